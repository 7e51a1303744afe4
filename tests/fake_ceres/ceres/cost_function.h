// Minimal stand-in for Ceres Solver 2.1.0's include/ceres/cost_function.h, written from its documented public interface, used only by
// tests/test_abi.py to prove that include/randt_host.hpp compiles against "a real ceres" (the __has_include branch) and that
// randt::NdtCostFunction overrides the pure virtual with the right signature.  Not used by the product.
#ifndef RANDT_TEST_FAKE_CERES_COST_FUNCTION_H_
#define RANDT_TEST_FAKE_CERES_COST_FUNCTION_H_
#include <cstdint>
#include <vector>
namespace ceres {
class CostFunction {
 public:
  CostFunction() : num_residuals_(0) {}
  CostFunction(const CostFunction&) = delete;
  void operator=(const CostFunction&) = delete;
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int32_t>& parameter_block_sizes() const { return parameter_block_sizes_; }
  int num_residuals() const { return num_residuals_; }

 protected:
  std::vector<int32_t>* mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
  void set_num_residuals(int num_residuals) { num_residuals_ = num_residuals; }

 private:
  std::vector<int32_t> parameter_block_sizes_;
  int num_residuals_;
};
}  // namespace ceres
#endif
