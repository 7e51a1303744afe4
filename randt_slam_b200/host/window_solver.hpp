// The joint window problem of Matcher::estimateTransformCeres (R/src/ndt_registration/ndt_matcher.cpp:322-424) on the host: a small
// ceres-shaped problem (parameter blocks with constness and the SE(2) manifold, autodiff-style host factors, one device term holding
// every NDT residual block of the window) and ceres 2.1.0's trust-region Levenberg-Marquardt loop over its normal equations.
// Internal to librandt_host.so (see include/randt_host.hpp for the public surface).
#pragma once
#include <cstdint>
#include <functional>
#include <vector>

#include "../../include/randt_host.hpp"

namespace randt {
namespace window {

// forward-mode dual number over N seed slots (what ceres::AutoDiffCostFunction evaluates its functors on)
template <int N>
struct Dual {
  double v;
  double d[N];
  Dual() : v(0.0) { for (int i = 0; i < N; ++i) d[i] = 0.0; }
  Dual(double x) : v(x) { for (int i = 0; i < N; ++i) d[i] = 0.0; }   // NOLINT: constants enter expressions implicitly, as in ceres
};
constexpr int kSeeds = 20;   // two states x (pose 4 + lin_vel 2 + rot_vel 1 + lin_acc 2 + imu_bias 1)
using D20 = Dual<kSeeds>;

struct ParameterBlock {
  double* user = nullptr;   // where the caller keeps the values (written back after the solve)
  int size = 0;             // ambient size
  int tangent = 0;          // tangent size (3 for an SE(2) pose, else = size)
  bool se2 = false;         // Sophus::Manifold<SE2>
  bool constant = false;
  int x_off = -1, t_off = -1;   // offsets in the reduced ambient / tangent vectors (free blocks only)
};

// a host factor: up to 10 parameter blocks -> nres residuals, differentiated with dual numbers.  `eval` receives the blocks' values as
// dual numbers already seeded (slot of block b, component c = seed_off[b] + c; constant blocks carry no seed).
struct HostFactor {
  std::vector<int> blocks;
  int nres = 0;
  std::function<void(const D20* const* params, D20* residuals)> eval;
};

// every NDT residual block of the window: segment s of `problem` hangs off the parameter blocks `seg_blocks[s]` (one {4} pose block, or
// {2}, {1} = pos, rot in the vector parametrisation), in the order of the variant's parameter vector
struct NdtTerm {
  randt_ctx* ctx = nullptr;
  const randt_problem* problem = nullptr;
  int variant = RANDT_VAR_SE2_INTENSITY;
  int np = 4;
  std::vector<std::vector<int>> seg_blocks;
  // caller-owned staging for one evaluation: [S][np] poses and [S][RANDT_FUSED_STRIDE] records.  Pinned (randt_host_alloc) memory lets the
  // poses go up without a staging copy and K3 store its records straight into host memory.
  double* poses = nullptr;
  double* records = nullptr;
};

class JointProblem {
 public:
  int addParameterBlock(double* values, int size, bool se2_manifold);   // idempotent on `values`
  void setParameterBlockConstant(int id) { blocks_[id].constant = true; }
  void addFactor(HostFactor f) { factors_.push_back(std::move(f)); }
  void setNdtTerm(const NdtTerm& t) { ndt_ = t; }
  void finalize();
  int numAmbient() const { return n_amb_; }
  int numTangent() const { return n_tan_; }
  void gather(double* x) const;
  void scatter(const double* x) const;
  void plus(const double* x, const double* delta, double* x_plus) const;
  // cost, and with want_jac the tangent gradient g [nt] and J^T J H [nt x nt], at x.  loss == nullptr: raw NDT residuals (max_raw gets
  // their maximum).  Returns false when the device call fails or the cost is not finite.
  bool evaluate(const double* x, const randt_loss* loss, bool want_jac, double* cost, double* g, double* H, double* max_raw = nullptr);
  void evaluateHostFactors(const double* x, bool want_jac, double* cost, double* g, double* H) const;
  int evaluations() const { return n_evals_; }
  bool deviceFailed() const { return device_failed_; }   // a C-ABI call failed during evaluate (as opposed to a non-finite cost)
  uint32_t ndtBlocks() const { return ndt_blocks_; }

 private:
  const double* blockValues(int id, const double* x) const { const ParameterBlock& b = blocks_[id]; return b.constant ? b.user : x + b.x_off; }
  std::vector<ParameterBlock> blocks_;
  std::vector<HostFactor> factors_;
  NdtTerm ndt_;
  int n_amb_ = 0, n_tan_ = 0, n_evals_ = 0;
  bool device_failed_ = false;
  uint32_t ndt_blocks_ = 0;
};

struct MinimizerSummary {
  double initial_cost = 0, final_cost = 0;
  int num_iterations = 0;
  int termination = 0;   // 0 convergence, 1 iteration limit, 2 failure
};
// ceres::Solve with TRUST_REGION / LEVENBERG_MARQUARDT on the problem's normal equations; x in / out
MinimizerSummary minimize(JointProblem& problem, const randt_loss& loss, const randt_solver_options& options, double* x);

// the reference's functors (R/include/ndt_registration/ceres_residuals.h), same argument lists
template <typename T>
void MotionModelFactorSE2(double dt, const double* sqrtI, const T* old_pose, const T* old_lin_vel, const T* old_rot_vel, const T* old_lin_acc,
                          const T* new_pose, const T* new_lin_vel, const T* new_rot_vel, const T* new_lin_acc, T* residuals);
template <typename T>
void MotionModelFactor(double dt, const double* sqrtI, const T* old_pos, const T* old_rot, const T* old_lin_vel, const T* old_rot_vel,
                       const T* old_lin_acc, const T* new_pos, const T* new_rot, const T* new_lin_vel, const T* new_rot_vel, const T* new_lin_acc,
                       T* residuals);
template <typename T>
void RotationalResidualSE2(double rot, double weight, double dt, double bias_weight, const T* pose_old, const T* pose_new, const T* bias_old,
                           const T* bias_new, T* residuals);
template <typename T>
void RotationalResidual(double rot, double weight, double dt, double bias_weight, const T* rot_old, const T* rot_new, const T* bias_old,
                        const T* bias_new, T* residuals);

}  // namespace window
}  // namespace randt
