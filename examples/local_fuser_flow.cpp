// How the reference's per-scan flow (R/src/local_fuser/local_fuser.cpp:99-190) and its loop-closure refinement (:335-340) read on top of
// include/randt_host.hpp.  Compile-checked by tests/test_abi.py; link with -lrandt_host -lrandt_gpu to run it on a
// B200: `g++ -std=c++17 -Iinclude examples/local_fuser_flow.cpp -Lrandt_slam_b200 -lrandt_host -lrandt_gpu -Wl,-rpath,randt_slam_b200`.
#include <cmath>
#include <cstdio>
#include <random>
#include <deque>
#include <vector>

#include "randt_host.hpp"

// a toy "scan": points on a few walls seen from pose (x, y, theta), (x, y, unused, intensity) per point, sensor frame
static std::vector<float> toy_scan(double x, double y, double th, unsigned seed) {
  std::mt19937 rng(seed);
  std::normal_distribution<float> noise(0.f, 0.03f), inten(90.f, 10.f);
  std::vector<float> pts;
  const double c = std::cos(-th), s = std::sin(-th);
  for (int wall = 0; wall < 12; ++wall) {
    const double wx = 8.0 * std::cos(wall * 0.52), wy = 8.0 * std::sin(wall * 0.52), dx = -std::sin(wall * 0.52), dy = std::cos(wall * 0.52);
    for (int i = 0; i < 160; ++i) {
      const double px = wx + dx * (i - 80) * 0.02 - x, py = wy + dy * (i - 80) * 0.02 - y;
      pts.push_back((float)(c * px - s * py) + noise(rng));
      pts.push_back((float)(s * px + c * py) + noise(rng));
      pts.push_back(0.f);
      pts.push_back(inten(rng));
    }
  }
  return pts;
}

int main() {
  try {
    randt::Context gpu(0);                                   // one CUDA stream; the scan callback and the loop-closure timer each own one
    randt::NDTMapParameters map_params;                      // field names of ndt_slam_parameters.h; defaults = parameters_indoor.yaml
    randt::NDTMatcherParameters matcher_params;
    randt::Matcher matcher(gpu);
    matcher.initialize(matcher_params);

    // first scan becomes the submap (LocalFuser::initialize path)
    randt::Map submap(gpu, map_params);
    submap.addClusters(toy_scan(0.0, 0.0, 0.0, 1));

    randt::SE2d pose;                                        // identity; Sophus::SE2d storage order (cos, sin, tx, ty)
    for (int k = 1; k <= 4; ++k) {
      // processScan: clusters -> NDT of the scan (K1), registration against the submap (K2 + K3 + K4), keyframe insertion
      randt::Map scan(gpu, map_params);
      scan.addClusters(toy_scan(0.15 * k, 0.02 * k, 0.01 * k, 10 + k));
      const bool accepted = matcher.estimateTransformNDT(pose, {&submap}, scan);
      std::printf("scan %d: %s  x %.3f  y %.3f  theta %.4f  (%zu cells)\n", k, accepted ? "ok" : "rejected", pose.v[2], pose.v[3], pose.angle(),
                  scan.get_n_cells());
      if (accepted && k % 2 == 0) {                          // insertion_step
        scan.transformMap(pose);
        submap.mergeMapCell(scan);
      }
    }

    // the same scans through the reference's own odometry entry points: predictTransform + estimateTransformCeres over the smoothing
    // window (NDT blocks of the window states + motion-model factors), local_fuser.cpp:125-138
    {
      std::vector<randt::State> trajectory(1);               // the submap's first state: identity, zero velocities, stamp 0
      std::deque<randt::Map> f_maps, map_window;
      f_maps.emplace_back(gpu, map_params);
      f_maps.back().addClusters(toy_scan(0.0, 0.0, 0.0, 1));
      randt::SE2d current_transform;
      for (int k = 1; k <= 4; ++k) {
        const double stamp = 0.25 * k;
        matcher.predictTransform(/*yaw=*/0.0, stamp, trajectory);
        map_window.emplace_back(gpu, map_params);
        map_window.back().addClusters(toy_scan(0.15 * k, 0.02 * k, 0.01 * k, 10 + k));
        matcher.estimateTransformCeres(current_transform, trajectory, 0.0, stamp, f_maps, map_window);
        const randt::WindowSummary& ws = matcher.lastWindowSummary();
        std::printf("window %d: x %.3f  y %.3f  theta %.4f  v %.3f  (%d states, %d LM iterations, %d K3 launches)\n", k, current_transform.v[2],
                    current_transform.v[3], current_transform.angle(), trajectory.back().lin_vel[0], ws.n_free_states, ws.total_iterations, ws.evaluations);
        if ((int)map_window.size() >= matcher.parameters().smoothing_steps) map_window.pop_front();
      }
    }

    // loop-closure refinement + verification against an old submap (here: the same place seen again)
    randt::Map old_map(gpu, map_params), new_map(gpu, map_params);
    old_map.addClusters(toy_scan(0.0, 0.0, 0.0, 2));
    new_map.addClusters(toy_scan(0.3, -0.1, 0.02, 3));
    randt::SE2d loop(0.0, 0.2, 0.0);
    const double score = matcher.estimateLoopConstraint(loop, old_map, new_map, /*gnc_steps=*/3, /*use_intensity=*/true, /*scale=*/1.0);
    new_map.transformMap(loop);
    const double cs = old_map.calculateCSDivergence(new_map)[0];
    std::printf("loop: score %.4f  CS divergence %.4f  x %.3f  y %.3f  theta %.4f\n", score, cs, loop.v[2], loop.v[3], loop.angle());

    // what NDTSlam::createVisualizationMsg publishes
    std::vector<double> mean, cov;
    submap.exportNormalDistributions(mean, cov);
    std::printf("submap: %zu cells exported (ndt_msgs Mean + Covariance layout)\n", mean.size() / 3);
  } catch (const randt::Error& e) {
    std::fprintf(stderr, "randt error %d: %s\n", e.code, e.what());
    return 1;
  }
  return 0;
}
