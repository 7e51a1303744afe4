// ORACLE — TEST INFRASTRUCTURE ONLY.  Full-reference fixture generator for the host factors of Matcher::estimateTransformCeres' joint
// window problem.
//
// NOT built in the development image (Eigen, Ceres, Sophus are absent there).  On a machine that has the reference's dependencies it
// evaluates the REFERENCE'S OWN functors, unmodified — MotionModelFactorSE2, MotionModelFactor, RotationalResidualSE2, RotationalResidual
// of ros/ndt_radar_slam/include/ndt_registration/ceres_residuals.h — through ceres::AutoDiffCostFunction with exactly the template
// arguments Matcher::addMotionModelFactor / addImuFactor use (ndt_matcher.cpp:60-110, 144-181), on the seeded state pairs of
// tests/golden/ref_full_window_inputs.txt, and writes tests/golden/ref_full_window_outputs.txt: per case the residuals and the ambient
// Jacobian over the 2 x 10 parameter slots (slot of state side s: 10 s + pose 0-3 | pos 0-1, rot 2 | lin_vel 4-5 | rot_vel 6 | lin_acc
// 7-8 | imu_bias 9).  tests/test_ref_full_fixtures.py holds the CPU oracle (oracle/window_oracle.h) and the product's host factors
// (randt_slam_b200/host/window_solver.cpp) to that file when it is present.
//
// Build + run (see oracle/ref_full/README.md):  make -C oracle/ref_full window REF=/path/to/RaNDT-SLAM && oracle/ref_full/gen_window_fixtures
#include <ceres/ceres.h>
#include <sophus/se2.hpp>

#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <ndt_registration/ceres_residuals.h>

namespace {
void expect(std::istream& in, const std::string& key) {
  std::string k; in >> k;
  if (k != key) { std::cerr << "input format: expected '" << key << "', got '" << k << "'\n"; std::exit(2); }
}
// state14: cos, sin, tx, ty, pos_x, pos_y, rot, vx, vy, omega, ax, ay, imu_bias, stamp
struct S14 { double v[14]; };
}  // namespace

int main(int argc, char** argv) {
  const std::string in_path = argc > 1 ? argv[1] : "tests/golden/ref_full_window_inputs.txt";
  const std::string out_path = argc > 2 ? argv[2] : "tests/golden/ref_full_window_outputs.txt";
  std::ifstream in(in_path);
  if (!in) { std::cerr << "cannot open " << in_path << "\n"; return 2; }
  FILE* out = std::fopen(out_path.c_str(), "w");
  if (!out) { std::cerr << "cannot write " << out_path << "\n"; return 2; }
  expect(in, "sqrtI");
  Eigen::Matrix<double, 8, 8> sqrtI;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) in >> sqrtI(i, j);
  double weight_imu, weight_bias;
  expect(in, "imu_weights"); in >> weight_imu >> weight_bias;
  size_t n;
  expect(in, "cases"); in >> n;
  std::fprintf(out, "cases %zu\n", n);
  for (size_t c = 0; c < n; ++c) {
    S14 a, b; double imu_rot;
    expect(in, "case");
    for (double& x : a.v) in >> x;
    for (double& x : b.v) in >> x;
    in >> imu_rot;
    const double dt = b.v[13] - a.v[13];
    // ---- kind 0 / 1 x manifold 1 / 0, in the order: motion SE2, motion vector, imu SE2, imu vector
    for (int kind = 0; kind < 2; ++kind)
      for (int manifold = 1; manifold >= 0; --manifold) {
        double res[8] = {0};
        double jac[8][20] = {{0}};
        const int nres = kind == 0 ? 8 : 2;
        bool ok = false;
        if (kind == 0 && manifold) {
          ceres::AutoDiffCostFunction<MotionModelFactorSE2, 8, 4, 2, 1, 2, 4, 2, 1, 2> f(new MotionModelFactorSE2(dt, sqrtI));
          const double* par[8] = {a.v, a.v + 7, a.v + 9, a.v + 10, b.v, b.v + 7, b.v + 9, b.v + 10};
          double J0[8 * 4], J1[8 * 2], J2[8], J3[8 * 2], J4[8 * 4], J5[8 * 2], J6[8], J7[8 * 2];
          double* jp[8] = {J0, J1, J2, J3, J4, J5, J6, J7};
          ok = f.Evaluate(par, res, jp);
          const int slot[8] = {0, 4, 6, 7, 10, 14, 16, 17}, size[8] = {4, 2, 1, 2, 4, 2, 1, 2};
          for (int bl = 0; bl < 8; ++bl) for (int r = 0; r < 8; ++r) for (int k = 0; k < size[bl]; ++k) jac[r][slot[bl] + k] = jp[bl][r * size[bl] + k];
        } else if (kind == 0) {
          ceres::AutoDiffCostFunction<MotionModelFactor, 8, 2, 1, 2, 1, 2, 2, 1, 2, 1, 2> f(new MotionModelFactor(dt, sqrtI));
          const double* par[10] = {a.v + 4, a.v + 6, a.v + 7, a.v + 9, a.v + 10, b.v + 4, b.v + 6, b.v + 7, b.v + 9, b.v + 10};
          double J[10][16];
          double* jp[10];
          for (int i = 0; i < 10; ++i) jp[i] = J[i];
          ok = f.Evaluate(par, res, jp);
          const int slot[10] = {0, 2, 4, 6, 7, 10, 12, 14, 16, 17}, size[10] = {2, 1, 2, 1, 2, 2, 1, 2, 1, 2};
          for (int bl = 0; bl < 10; ++bl) for (int r = 0; r < 8; ++r) for (int k = 0; k < size[bl]; ++k) jac[r][slot[bl] + k] = jp[bl][r * size[bl] + k];
        } else if (manifold) {
          ceres::AutoDiffCostFunction<RotationalResidualSE2, 2, 4, 4, 1, 1> f(new RotationalResidualSE2(imu_rot, weight_imu, dt, weight_bias));
          const double* par[4] = {a.v, b.v, a.v + 12, b.v + 12};
          double J0[8], J1[8], J2[2], J3[2];
          double* jp[4] = {J0, J1, J2, J3};
          ok = f.Evaluate(par, res, jp);
          const int slot[4] = {0, 10, 9, 19}, size[4] = {4, 4, 1, 1};
          for (int bl = 0; bl < 4; ++bl) for (int r = 0; r < 2; ++r) for (int k = 0; k < size[bl]; ++k) jac[r][slot[bl] + k] = jp[bl][r * size[bl] + k];
        } else {
          ceres::AutoDiffCostFunction<RotationalResidual, 2, 1, 1, 1, 1> f(new RotationalResidual(imu_rot, weight_imu, dt, weight_bias));
          const double* par[4] = {a.v + 6, b.v + 6, a.v + 12, b.v + 12};
          double J0[2], J1[2], J2[2], J3[2];
          double* jp[4] = {J0, J1, J2, J3};
          ok = f.Evaluate(par, res, jp);
          const int slot[4] = {2, 12, 9, 19};
          for (int bl = 0; bl < 4; ++bl) for (int r = 0; r < 2; ++r) jac[r][slot[bl]] = jp[bl][r];
        }
        std::fprintf(out, "block %d %d %d %d\n", kind, manifold, nres, ok ? 1 : 0);
        for (int r = 0; r < nres; ++r) {
          std::fprintf(out, "%.17g", res[r]);
          for (int k = 0; k < 20; ++k) std::fprintf(out, " %.17g", jac[r][k]);
          std::fprintf(out, "\n");
        }
      }
  }
  std::fclose(out);
  std::cout << "wrote " << out_path << "\n";
  return 0;
}
