"""Parameter presets mirroring the reference's shipped YAML configs, with the derived values the
reference computes in NDTSlam::readParameters (R/src/ndt_slam/ndt_slam.cpp:397-712).

R/ = /root/reference/ros/ndt_radar_slam/.  Sources: R/config/parameters_{indoor,outdoor,mixed,oxford}.yaml and
R/config/ndt_radar_slam_base_parameters.yaml.  Derived values:
  * map size in cells: `size_x /= resolution` on an int (truncates)           ndt_slam.cpp:653-654
  * n_clusters = int((2*max_range/resolution)^2)                               ndt_slam.cpp:691
  * loop-closure gnc steps / scale default to the matcher's                    ndt_slam.cpp:573-575,584-586
"""
import math
from dataclasses import dataclass, replace


@dataclass(frozen=True)
class NdtParams:
    name: str
    # ndt_map
    resolution: float
    max_neighbor_linf_distance: float
    min_points_per_cell: int
    map_size_m_x: int = 50          # base yaml (metres, before the /= resolution)
    map_size_m_y: int = 50
    # radar_preprocessor
    min_range: float = 0.6
    max_range: float = 12.0
    min_intensity: float = 6.0
    beam_distance_increment_threshold: float = 0.04
    # ndt_matcher
    gnc_control_parameter_divisor: float = 1.3
    gnc_steps: int = 3
    loss_function_convexity: float = -2.0
    loss_function_scale: float = 1.5
    n_results_nn_lookup: int = 4
    ndt_weight: float = 50000.0
    smoothing_steps: int = 3
    max_iteration: int = 200
    lookup_distribution: bool = True
    use_intensity_as_dimension: bool = True
    optimize_on_manifold: bool = True
    covariance_scaling_factor: float = 25.0
    pose_reject_translation: float = 2.0
    pose_reject_rotation: float = 2.0
    # local_fuser
    loop_closure_gnc_steps: int = 2
    loop_closure_scale: float = 1.5
    csm_cost_threshold: float = 0.82

    # ---- derived exactly as the reference derives them ----
    @property
    def size_x(self) -> int:
        return int(self.map_size_m_x / self.resolution)

    @property
    def size_y(self) -> int:
        return int(self.map_size_m_y / self.resolution)

    @property
    def n_clusters(self) -> int:
        return int(math.pow(2.0 * self.max_range / self.resolution, 2))

    @property
    def grid_row_size(self) -> int:
        return int(math.sqrt(self.n_clusters))

    @property
    def r_stop(self) -> int:
        """window-radius loop bound int(max_linf / res)  (R/src/ndt_representation/ndt_map.cpp:143)"""
        return int(self.max_neighbor_linf_distance / self.resolution)


INDOOR = NdtParams("indoor", resolution=0.5, max_neighbor_linf_distance=4.0, min_points_per_cell=5)
OUTDOOR = NdtParams("outdoor", resolution=1.2, max_neighbor_linf_distance=4.0, min_points_per_cell=3, max_range=16.0,
                    gnc_control_parameter_divisor=1.1, gnc_steps=3, loss_function_convexity=-1.0, loss_function_scale=2.0,
                    loop_closure_gnc_steps=1, loop_closure_scale=2.0)
MIXED = NdtParams("mixed", resolution=1.0, max_neighbor_linf_distance=4.0, min_points_per_cell=3, max_range=16.0,
                  gnc_control_parameter_divisor=1.1, gnc_steps=3, loss_function_convexity=-1.5, loss_function_scale=2.0,
                  loop_closure_gnc_steps=1, loop_closure_scale=2.0)
OXFORD = NdtParams("oxford", resolution=3.5, max_neighbor_linf_distance=10.0, min_points_per_cell=10, map_size_m_x=400,
                   map_size_m_y=400, min_range=2.0, max_range=100.0, min_intensity=70.0, beam_distance_increment_threshold=0.12,
                   gnc_control_parameter_divisor=1.1, gnc_steps=2, loss_function_convexity=-2.0, loss_function_scale=1.0,
                   n_results_nn_lookup=2, ndt_weight=5000.0, covariance_scaling_factor=0.01, pose_reject_translation=5.0,
                   loop_closure_gnc_steps=10, loop_closure_scale=0.5)

PRESETS = {p.name: p for p in (INDOOR, OUTDOOR, MIXED, OXFORD)}

# BASELINE config C1: "single synthetic 2D scan pair, 32x32 NDT grid (~200 cells)".  mixed.yaml is the shipped
# config whose cluster grid is exactly 32x32 (n_clusters = (2*16/1.0)^2 = 1024).
C1 = replace(MIXED, name="c1_32x32")
# BASELINE config C3 (beyond-reference SE(3) wording; run here with reference SE(2) semantics):
# 0.5 m cells, 200x200 map, outdoor loss parameters, k = 4.
C3 = replace(OUTDOOR, name="c3_dense", resolution=0.5, map_size_m_x=100, map_size_m_y=100, max_range=50.0,
             max_neighbor_linf_distance=2.0)
