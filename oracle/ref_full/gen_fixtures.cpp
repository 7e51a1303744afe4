// ORACLE — TEST INFRASTRUCTURE ONLY.  Full-reference fixture generator.
//
// This program is NOT built in the development image (Eigen, Ceres, Sophus are absent there).  On a machine that has the reference's
// dependencies (the reference's own Dockerfile: Eigen 3.3.7, Ceres 2.1.0, Sophus 1.22.10) it evaluates the REFERENCE'S OWN headers —
// ros/ndt_radar_slam/include/ndt_registration/{ceres_residuals.h, ceres_loss_functions.h} and src/ndt_registration/ceres_loss_functions.cpp,
// unmodified — on the seeded inputs of tests/golden/ref_full_inputs.txt and writes tests/golden/ref_full_outputs.txt:
//   * r and the autodiff Jacobian row of all four NDT functors (ceres::AutoDiffCostFunction, exactly what Matcher::addNDTFactor builds,
//     ndt_matcher.cpp:217-284) for every (pair, pose);
//   * BarronLoss::Evaluate (rho, rho', rho'') for the listed settings;
//   * Matcher::estimateLoopConstraint's solve (ndt_matcher.cpp:426-493) restated around real ceres::Problem / ceres::Solve on the given
//     pair lists, in the raw-ambient mode the reference ends up in and on the manifold: final pose, score, iterations.
// tests/test_ref_full_fixtures.py holds the CPU oracle (and, under -m gpu, the CUDA path) to that file when it is present.
//
// Build + run (see oracle/ref_full/README.md):  make -C oracle/ref_full REF=/path/to/RaNDT-SLAM && oracle/ref_full/gen_fixtures
#include <ceres/ceres.h>
#include <sophus/se2.hpp>
#include <sophus/ceres_manifold.hpp>

#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <ndt_registration/ceres_residuals.h>
#include <ndt_registration/ceres_loss_functions.h>

namespace {

struct Cell { double v[12]; };
Eigen::Vector3d mean3(const Cell& c) { return Eigen::Vector3d(c.v[0], c.v[1], c.v[2]); }
Eigen::Matrix3d cov3(const Cell& c) { Eigen::Matrix3d m; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m(i, j) = c.v[3 + 3 * i + j]; return m; }
Eigen::Vector2d mean2(const Cell& c) { return Eigen::Vector2d(c.v[0], c.v[1]); }
Eigen::Matrix2d cov2(const Cell& c) { Eigen::Matrix2d m; m << c.v[3], c.v[4], c.v[6], c.v[7]; return m; }

void expect(std::istream& in, const std::string& key) {
  std::string k; in >> k;
  if (k != key) { std::cerr << "input format: expected '" << key << "', got '" << k << "'\n"; std::exit(2); }
}
std::vector<double> read_n(std::istream& in, size_t n) { std::vector<double> v(n); for (double& x : v) in >> x; return v; }

ceres::CostFunction* make_block(int variant, const Cell& m, const Cell& f) {
  switch (variant) {
    case 0: return new ceres::AutoDiffCostFunction<NDTFrameToMapIntensityFactorResidualSE2, 1, 4>(new NDTFrameToMapIntensityFactorResidualSE2(mean3(m), cov3(m), mean3(f), cov3(f)));
    case 1: return new ceres::AutoDiffCostFunction<NDTFrameToMapFactorResidualSE2, 1, 4>(new NDTFrameToMapFactorResidualSE2(mean2(m), cov2(m), mean2(f), cov2(f)));
    case 2: return new ceres::AutoDiffCostFunction<NDTFrameToMapIntensityFactorResidual, 1, 2, 1>(new NDTFrameToMapIntensityFactorResidual(mean3(m), cov3(m), mean3(f), cov3(f)));
    default: return new ceres::AutoDiffCostFunction<NDTFrameToMapFactorResidual, 1, 2, 1>(new NDTFrameToMapFactorResidual(mean2(m), cov2(m), mean2(f), cov2(f)));
  }
}

}  // namespace

int main(int argc, char** argv) {
  const std::string in_path = argc > 1 ? argv[1] : "tests/golden/ref_full_inputs.txt";
  const std::string out_path = argc > 2 ? argv[2] : "tests/golden/ref_full_outputs.txt";
  std::ifstream in(in_path);
  if (!in) { std::cerr << "cannot open " << in_path << "\n"; return 2; }
  FILE* out = std::fopen(out_path.c_str(), "w");
  if (!out) { std::cerr << "cannot write " << out_path << "\n"; return 2; }
  size_t n;
  expect(in, "pairs"); in >> n;
  std::vector<Cell> cm(n), cf(n);
  for (size_t i = 0; i < n; ++i) { for (double& x : cm[i].v) in >> x; for (double& x : cf[i].v) in >> x; }
  size_t n4, n3;
  expect(in, "poses4"); in >> n4; std::vector<double> p4 = read_n(in, 4 * n4);
  expect(in, "poses3"); in >> n3; std::vector<double> p3 = read_n(in, 3 * n3);
  // ---- functors: r and dr/dparams through ceres' autodiff, variant by variant ----
  for (int variant = 0; variant < 4; ++variant) {
    const size_t np = variant <= 1 ? n4 : n3;
    std::fprintf(out, "functor %d %zu %zu\n", variant, np, n);
    for (size_t s = 0; s < np; ++s)
      for (size_t i = 0; i < n; ++i) {
        ceres::CostFunction* cfun = make_block(variant, cm[i], cf[i]);
        double r = 0.0, J[4] = {0, 0, 0, 0};
        bool ok;
        if (variant <= 1) {
          const double* par[1] = {&p4[4 * s]};
          double* jac[1] = {J};
          ok = cfun->Evaluate(par, &r, jac);
        } else {
          const double* par[2] = {&p3[3 * s], &p3[3 * s + 2]};
          double* jac[2] = {J, J + 2};
          ok = cfun->Evaluate(par, &r, jac);
        }
        std::fprintf(out, "%d %.17g %.17g %.17g %.17g %.17g\n", ok ? 1 : 0, r, J[0], J[1], J[2], variant <= 1 ? J[3] : 0.0);
        delete cfun;
      }
  }
  // ---- Barron loss ----
  size_t ns, nl;
  expect(in, "loss_s"); in >> ns; std::vector<double> ss = read_n(in, ns);
  expect(in, "loss_settings"); in >> nl; std::vector<double> ls = read_n(in, 3 * nl);
  std::fprintf(out, "loss %zu %zu\n", nl, ns);
  for (size_t l = 0; l < nl; ++l) {
    ceres::BarronLoss loss(ls[3 * l], ls[3 * l + 1], ls[3 * l + 2]);
    for (double s : ss) { double rho[3]; loss.Evaluate(s, rho); std::fprintf(out, "%.17g %.17g %.17g\n", rho[0], rho[1], rho[2]); }
  }
  // ---- estimateLoopConstraint on given pair lists (ndt_matcher.cpp:426-493), raw ambient block and manifold ----
  expect(in, "solver"); std::vector<double> sv = read_n(in, 7);
  const double loss_function_scale = sv[0], loop_scale = sv[1], convexity = sv[2], divisor = sv[3];
  const int max_gnc_steps = (int)sv[4], max_iteration = (int)sv[6];
  size_t nr;
  expect(in, "registrations"); in >> nr;
  std::fprintf(out, "registrations %zu\n", nr);
  for (size_t g = 0; g < nr; ++g) {
    size_t nm, nf, P;
    expect(in, "registration"); in >> nm >> nf >> P;
    std::vector<Cell> rm(nm), rf(nf);
    for (Cell& c : rm) for (double& x : c.v) in >> x;
    for (Cell& c : rf) for (double& x : c.v) in >> x;
    std::vector<size_t> im(P), jf(P);
    for (size_t& x : im) in >> x;
    for (size_t& x : jf) in >> x;
    std::vector<double> pose0 = read_n(in, 4);
    for (int on_manifold = 0; on_manifold < 2; ++on_manifold) {
      Sophus::SE2d trans;
      std::copy(pose0.begin(), pose0.end(), trans.data());
      ceres::Problem problem;
      // estimateLoopConstraint attaches the manifold to trans.data() but the residual blocks to a COPY (two_representation_state.pose): the
      // copy is optimised as four raw parameters.  on_manifold == 1 is the behaviour the code intends.
      Sophus::SE2d state = trans;
      if (on_manifold) problem.AddParameterBlock(state.data(), 4, new Sophus::Manifold<Sophus::SE2>());
      ceres::LossFunctionWrapper* current_loss = new ceres::LossFunctionWrapper(nullptr, ceres::TAKE_OWNERSHIP);
      std::vector<ceres::ResidualBlockId> ids;
      for (size_t i = 0; i < P; ++i) ids.push_back(problem.AddResidualBlock(make_block(0, rm[im[i]], rf[jf[i]]), current_loss, state.data()));
      ceres::Solver::Options options;
      options.max_num_iterations = max_iteration;
      options.linear_solver_type = ceres::DENSE_QR;
      options.trust_region_strategy_type = ceres::LEVENBERG_MARQUARDT;
      options.num_threads = 1;
      ceres::Problem::EvaluateOptions eo; eo.residual_blocks = ids; eo.apply_loss_function = false;
      std::vector<double> raw;
      problem.Evaluate(eo, nullptr, &raw, nullptr, nullptr);
      const double max_residual = *std::max_element(raw.begin(), raw.end());
      double gnc_mu = 2.0 * std::pow(max_residual, 2) / std::pow(loss_function_scale, 2);
      gnc_mu = std::min(gnc_mu, std::pow(divisor, max_gnc_steps - 1));
      const double mu_first = gnc_mu;
      ceres::Solver::Summary summary;
      int solves = 0, iterations = 0;
      do {
        gnc_mu = std::max(gnc_mu, 1.0);
        current_loss->Reset(new ceres::ScaledLoss(new ceres::BarronLoss(loop_scale, convexity, gnc_mu), 1, ceres::TAKE_OWNERSHIP), ceres::TAKE_OWNERSHIP);
        ceres::Solve(options, &problem, &summary);
        ++solves; iterations += (int)summary.iterations.size();
        gnc_mu /= divisor;
      } while (gnc_mu > 1.0 / std::sqrt(divisor));
      std::fprintf(out, "%d %.17g %.17g %.17g %.17g %.17g %d %d %.17g\n", on_manifold, state.data()[0], state.data()[1], state.data()[2], state.data()[3],
                   summary.final_cost / summary.num_residual_blocks, solves, iterations, mu_first);
    }
  }
  std::fclose(out);
  std::cout << "wrote " << out_path << "\n";
  return 0;
}
