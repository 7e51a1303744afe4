// ORACLE SHIM (test infrastructure): the interface of ceres::LossFunction (ceres-solver 2.1.0, include/ceres/loss_function.h:
// `virtual void Evaluate(double sq_norm, double out[3]) const = 0;` with out = {rho, rho', rho''}) so that the reference's
// R/src/ndt_registration/ceres_loss_functions.cpp compiles unmodified.  Nothing of Ceres' own code is restated here.
#pragma once
#include <cmath>
#include <cstdlib>
#include <memory>
namespace ceres {
class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};
}  // namespace ceres
