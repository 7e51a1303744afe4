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
// A second section of the input file ("solves") holds whole window problems — states, frozen NDT pair lists with their cell tables, IMU
// constraints, parameters — which are solved the way Matcher::estimateTransformCeres does (ndt_matcher.cpp:322-397): parameter blocks of
// addMotionParameterBlock / addImuParameterBlock, the motion / IMU factors above, one AutoDiffCostFunction<NDTFrameToMapIntensityFactorResidualSE2,
// 1, 4> (or <NDTFrameToMapIntensityFactorResidual, 1, 2, 1>) per pair behind a LossFunctionWrapper, Sophus::Manifold<SE2>, DENSE_QR /
// LEVENBERG_MARQUARDT, the GNC loop around the real ceres::Solve.  Output: the window states after the solve, solves, iterations, final cost.
//
// Build + run (see oracle/ref_full/README.md):  make -C oracle/ref_full window REF=/path/to/RaNDT-SLAM && oracle/ref_full/gen_window_fixtures
#include <ceres/ceres.h>
#include <sophus/se2.hpp>
#include <sophus/ceres_manifold.hpp>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <ndt_registration/ceres_residuals.h>
#include <ndt_registration/ceres_loss_functions.h>

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

  // ---- whole window problems, solved as Matcher::estimateTransformCeres solves them --------------------------------------------------
  std::string key;
  if (in >> key) {
    if (key != "solves") { std::cerr << "input format: expected 'solves', got '" << key << "'\n"; return 2; }
    size_t n_solves; in >> n_solves;
    std::fprintf(out, "solves %zu\n", n_solves);
    struct RState { Sophus::SE2d pose; Eigen::Vector2d pos; double rot; Eigen::Vector2d lin_vel; double rot_vel; Eigen::Vector2d lin_acc; double imu_bias; double stamp; };
    struct Cell { double v[12]; };
    auto mean3 = [](const Cell& c) { return Eigen::Vector3d(c.v[0], c.v[1], c.v[2]); };
    auto cov3 = [](const Cell& c) { Eigen::Matrix3d m; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m(i, j) = c.v[3 + 3 * i + j]; return m; };
    for (size_t g = 0; g < n_solves; ++g) {
      expect(in, "solve");
      size_t W, nm, nf, P;
      in >> W >> nm >> nf >> P;
      double par[16];
      expect(in, "params"); for (double& x : par) in >> x;
      const int k = (int)par[0], gnc_steps = (int)par[1], max_iteration = (int)par[2];
      const double loss_scale = par[3], convexity = par[4], divisor = par[5], ndt_weight = par[6];
      const bool manifold_mode = par[7] != 0.0, constant_velocity = par[8] != 0.0, use_imu = par[9] != 0.0;
      const double w_imu = par[10], w_bias = par[11];
      Eigen::Matrix<double, 8, 8> sI;
      expect(in, "sqrtI"); for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) in >> sI(i, j);
      double tol[3]; expect(in, "tolerances"); in >> tol[0] >> tol[1] >> tol[2];
      size_t n_cells_total; expect(in, "n_cells"); in >> n_cells_total;
      std::vector<RState, Eigen::aligned_allocator<RState>> X(W + 1);
      expect(in, "states");
      for (RState& s : X) {
        double v[14]; for (double& x : v) in >> x;
        std::copy(v, v + 4, s.pose.data());
        s.pos = Eigen::Vector2d(v[4], v[5]); s.rot = v[6]; s.lin_vel = Eigen::Vector2d(v[7], v[8]); s.rot_vel = v[9];
        s.lin_acc = Eigen::Vector2d(v[10], v[11]); s.imu_bias = v[12]; s.stamp = v[13];
      }
      std::vector<double> imu(W); expect(in, "imu"); for (double& x : imu) in >> x;
      std::vector<Cell> cm(nm), cf(nf);
      expect(in, "cells_m"); for (Cell& c : cm) for (double& x : c.v) in >> x;
      expect(in, "cells_f"); for (Cell& c : cf) for (double& x : c.v) in >> x;
      std::vector<size_t> im(P), jf(P), seg_off(W + 1);
      expect(in, "pair_m"); for (size_t& x : im) in >> x;
      expect(in, "pair_f"); for (size_t& x : jf) in >> x;
      expect(in, "seg_off"); for (size_t& x : seg_off) in >> x;

      ceres::Problem problem;
      ceres::Manifold* manifold = manifold_mode ? new Sophus::Manifold<Sophus::SE2>() : nullptr;
      ceres::LossFunctionWrapper* current_loss = new ceres::LossFunctionWrapper(nullptr, ceres::TAKE_OWNERSHIP);
      auto add_motion_blocks = [&](RState& S, bool set_constant) {      // Matcher::addMotionParameterBlock
        if (manifold) problem.AddParameterBlock(S.pose.data(), 4, manifold);
        else { problem.AddParameterBlock(&S.pos[0], 2); problem.AddParameterBlock(&S.rot, 1); }
        problem.AddParameterBlock(&S.lin_vel[0], 2);
        problem.AddParameterBlock(&S.rot_vel, 1);
        problem.AddParameterBlock(&S.lin_acc[0], 2);
        if (constant_velocity) problem.SetParameterBlockConstant(&S.lin_acc[0]);
        if (set_constant) {
          if (manifold) problem.SetParameterBlockConstant(S.pose.data());
          else { problem.SetParameterBlockConstant(&S.pos[0]); problem.SetParameterBlockConstant(&S.rot); }
        }
      };
      add_motion_blocks(X[0], true);
      if (use_imu) { problem.AddParameterBlock(&X[0].imu_bias, 1); problem.SetParameterBlockConstant(&X[0].imu_bias); }
      std::vector<ceres::ResidualBlockId> ndt_residuals;
      for (size_t j = 1; j <= W; ++j) {
        add_motion_blocks(X[j], false);
        RState &A = X[j - 1], &B = X[j];
        const double dt = B.stamp - A.stamp;
        if (manifold)
          problem.AddResidualBlock(new ceres::AutoDiffCostFunction<MotionModelFactorSE2, 8, 4, 2, 1, 2, 4, 2, 1, 2>(new MotionModelFactorSE2(dt, sI)), nullptr,
                                   A.pose.data(), &A.lin_vel[0], &A.rot_vel, &A.lin_acc[0], B.pose.data(), &B.lin_vel[0], &B.rot_vel, &B.lin_acc[0]);
        else
          problem.AddResidualBlock(new ceres::AutoDiffCostFunction<MotionModelFactor, 8, 2, 1, 2, 1, 2, 2, 1, 2, 1, 2>(new MotionModelFactor(dt, sI)), nullptr,
                                   &A.pos[0], &A.rot, &A.lin_vel[0], &A.rot_vel, &A.lin_acc[0], &B.pos[0], &B.rot, &B.lin_vel[0], &B.rot_vel, &B.lin_acc[0]);
        if (use_imu) {
          problem.AddParameterBlock(&B.imu_bias, 1);
          if (manifold)
            problem.AddResidualBlock(new ceres::AutoDiffCostFunction<RotationalResidualSE2, 2, 4, 4, 1, 1>(new RotationalResidualSE2(imu[j - 1], w_imu, dt, w_bias)), nullptr,
                                     A.pose.data(), B.pose.data(), &A.imu_bias, &B.imu_bias);
          else
            problem.AddResidualBlock(new ceres::AutoDiffCostFunction<RotationalResidual, 2, 1, 1, 1, 1>(new RotationalResidual(imu[j - 1], w_imu, dt, w_bias)), nullptr,
                                     &A.rot, &B.rot, &A.imu_bias, &B.imu_bias);
        }
        for (size_t p = seg_off[j - 1]; p < seg_off[j]; ++p) {
          if (manifold)
            ndt_residuals.push_back(problem.AddResidualBlock(
                new ceres::AutoDiffCostFunction<NDTFrameToMapIntensityFactorResidualSE2, 1, 4>(
                    new NDTFrameToMapIntensityFactorResidualSE2(mean3(cm[im[p]]), cov3(cm[im[p]]), mean3(cf[jf[p]]), cov3(cf[jf[p]]))),
                current_loss, B.pose.data()));
          else
            ndt_residuals.push_back(problem.AddResidualBlock(
                new ceres::AutoDiffCostFunction<NDTFrameToMapIntensityFactorResidual, 1, 2, 1>(
                    new NDTFrameToMapIntensityFactorResidual(mean3(cm[im[p]]), cov3(cm[im[p]]), mean3(cf[jf[p]]), cov3(cf[jf[p]]))),
                current_loss, &B.pos[0], &B.rot));
        }
      }
      ceres::Solver::Options options;
      options.max_num_iterations = max_iteration;
      options.linear_solver_type = ceres::DENSE_QR;
      options.trust_region_strategy_type = ceres::LEVENBERG_MARQUARDT;
      options.num_threads = 1;
      if (tol[0] > 0.0) options.function_tolerance = tol[0];
      if (tol[1] > 0.0) options.parameter_tolerance = tol[1];
      if (tol[2] > 0.0) options.gradient_tolerance = tol[2];
      ceres::Problem::EvaluateOptions eo; eo.residual_blocks = ndt_residuals; eo.apply_loss_function = false;
      std::vector<double> raw;
      problem.Evaluate(eo, nullptr, &raw, nullptr, nullptr);
      const double max_residual = *std::max_element(raw.begin(), raw.end());
      double gnc_mu = 2.0 * std::pow(max_residual, 2) / std::pow(loss_scale, 2);
      gnc_mu = std::min(gnc_mu, std::pow(divisor, gnc_steps - 1));
      const double mu_first = gnc_mu;
      ceres::Solver::Summary summary;
      int solves = 0, iterations = 0;
      do {
        gnc_mu = std::max(gnc_mu, 1.0);
        current_loss->Reset(new ceres::ScaledLoss(new ceres::BarronLoss(loss_scale, convexity, gnc_mu), ndt_weight / static_cast<double>(n_cells_total * k), ceres::TAKE_OWNERSHIP),
                            ceres::TAKE_OWNERSHIP);
        ceres::Solve(options, &problem, &summary);
        ++solves; iterations += (int)summary.iterations.size();
        gnc_mu /= divisor;
      } while (gnc_mu > 1.0 / std::sqrt(divisor));
      std::fprintf(out, "solve %zu %d %d %.17g %.17g %.17g\n", W, solves, iterations, summary.final_cost, mu_first, max_residual);
      for (const RState& s : X)
        std::fprintf(out, "%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", s.pose.data()[0], s.pose.data()[1], s.pose.data()[2],
                     s.pose.data()[3], s.pos[0], s.pos[1], s.rot, s.lin_vel[0], s.lin_vel[1], s.rot_vel, s.lin_acc[0], s.lin_acc[1], s.imu_bias, s.stamp);
    }
  }
  std::fclose(out);
  std::cout << "wrote " << out_path << "\n";
  return 0;
}
