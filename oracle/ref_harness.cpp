// ORACLE / _ref HARNESS — TEST INFRASTRUCTURE ONLY.
//
// C entry points over the reference's OWN classes, compiled unmodified from /root/reference (see oracle/Makefile, target `ref`):
//   R/src/ndt_registration/ceres_loss_functions.cpp      ceres::BarronLoss / WelschLoss            (a11)
//   R/src/radar_preprocessing/grid.cpp                    Grid::cluster                             (a1)
//   R/src/radar_preprocessing/radar_preprocessor.cpp      filterScan, ClusterGenerator::labelClouds (f1, a2)
//   R/src/ndt_representation/ndt_cell.cpp                 Cell::addPointCloud/updateCell/transformCell/mahalanobis*, operator+=  (a3-a6)
//   R/src/ndt_representation/ndt_map.cpp                  Map::insertCluster/transformMap/mergeMapCell/getClosestCells/calculateCSDivergence (a3, a5-a7, f2)
// against the shim headers in oracle/shim/ (Eigen, PCL, ceres::LossFunction, ROS message types: none of them is in this image).
// This file only marshals flat arrays in and out and strings the reference's calls together the way the reference's own callers do;
// each such place cites the caller it mirrors.  R/ = /root/reference/ros/ndt_radar_slam/.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>
#include <iostream>
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <complex>
#include <sstream>
#include <math.h>
// the shim headers and every standard header above are pulled in BEFORE the access override below, so that only the reference's own
// class definitions see it
#include <Eigen/Core>
#include <Eigen/Eigenvalues>
#include <boost/shared_ptr.hpp>
#include <pcl/common/centroid.h>
#include <pcl/common/transforms.h>
#include <pcl/impl/point_types.hpp>
#include <pcl/ml/kmeans.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl_conversions/pcl_conversions.h>
#include <pcl_ros/point_cloud.h>
#include <sensor_msgs/PointCloud.h>
#include <ceres/loss_function.h>

// RadarPreprocessor::filterScan is private; the harness calls it directly (the class is compiled from the unmodified source, only this
// translation unit sees the members as public)
#define private public
#include "radar_preprocessing/radar_preprocessor.h"
#include "ndt_representation/ndt_map.h"
#undef private
#include "ndt_registration/ceres_loss_functions.h"

using rc::navigation::ndt::Cell;
using rc::navigation::ndt::Map;

namespace {

// Sophus 1.22.10 SE2d::cast<float>().matrix() (sophus/se2.hpp, so2.hpp): the unit complex is cast to float and RE-NORMALISED by the SO2
// constructor (length = hypot(re, im); complex /= length), the translation is cast; matrix() = [[re, -im, tx], [im, re, ty], [0, 0, 1]].
// Sophus is not in this image: these four lines are restated.  Used exactly where the reference writes
// `Eigen::Affine2f(x.cast<float>().matrix())` (ndt_matcher.cpp:208, local_fuser.cpp:175,280,338).
Eigen::Affine2f affine_from_se2d(const double* pose) {
  float re = static_cast<float>(pose[0]), im = static_cast<float>(pose[1]);
  const float length = std::hypot(re, im);
  re /= length; im /= length;
  Eigen::Matrix3f m;
  m << re, -im, static_cast<float>(pose[2]), im, re, static_cast<float>(pose[3]), 0.f, 0.f, 1.f;
  return Eigen::Affine2f(m);
}

rc::navigation::ndt::NDTMapParameters map_params(int size_x, int size_y, double res, double max_linf, int min_points) {
  rc::navigation::ndt::NDTMapParameters p;
  p.size_x = size_x; p.size_y = size_y; p.resolution = res; p.ogm_resolution = res; p.ogm_threshold = 0.0;
  p.max_neighbour_manhattan_distance = max_linf; p.min_points_per_cell = min_points; p.visualize_ogm = false;
  p.ndt_cell_parameters.use_pndt = false;                 // off in every shipped config
  p.ndt_cell_parameters.beam_cov = Eigen::Matrix3f::Zero();
  return p;
}

pcl::PointCloud<pcl::PointXYZI>::Ptr cloud_from(const float* pts4, uint32_t n) {
  pcl::PointCloud<pcl::PointXYZI>::Ptr c(new pcl::PointCloud<pcl::PointXYZI>());
  c->reserve(n);
  for (uint32_t i = 0; i < n; ++i) c->push_back(pcl::PointXYZI(pts4[4 * i], pts4[4 * i + 1], pts4[4 * i + 2], pts4[4 * i + 3]));
  return c;
}

struct RefMap { Map map; };

}  // namespace

extern "C" {

int ref_version() { return 1; }

// ---- a11: the reference's loss classes -------------------------------------------------------------------------
// with_mu != 0: BarronLoss(a, alpha, mu) / WelschLoss(a, mu) (the GNC constructors, ndt_matcher.cpp:391,479); else the 2 / 1 argument forms
void ref_barron_evaluate(double a, double alpha, double mu, int with_mu, const double* s, int n, double* rho3) {
  if (with_mu) { ceres::BarronLoss l(a, alpha, mu); for (int i = 0; i < n; ++i) l.Evaluate(s[i], rho3 + 3 * i); }
  else { ceres::BarronLoss l(a, alpha); for (int i = 0; i < n; ++i) l.Evaluate(s[i], rho3 + 3 * i); }
}
void ref_welsch_evaluate(double a, double mu, int with_mu, const double* s, int n, double* rho3) {
  if (with_mu) { ceres::WelschLoss l(a, mu); for (int i = 0; i < n; ++i) l.Evaluate(s[i], rho3 + 3 * i); }
  else { ceres::WelschLoss l(a); for (int i = 0; i < n; ++i) l.Evaluate(s[i], rho3 + 3 * i); }
}

// ---- a1: Grid::cluster -------------------------------------------------------------------------------------------
void ref_grid_cluster(const float* pts4, uint32_t n, uint64_t n_clusters, double max_range, int32_t* labels) {
  rc::navigation::ndt::Grid g;
  g.setMaxRange(max_range);            // double -> float as RadarPreprocessor::initialize does (radar_preprocessor.cpp:27)
  std::vector<int> l(n);
  g.cluster(static_cast<size_t>(n_clusters), cloud_from(pts4, n), l);
  for (uint32_t i = 0; i < n; ++i) labels[i] = l[i];
}

// ---- a1-a4: filtered scan -> NDT map, as RadarPreprocessor::processScan (radar_preprocessor.cpp:36-38) followed by
// HierarchicalMap::addClusters (ndt_hierarchical_map.cpp:28-31) do.  Returns NULL (and *err = 1) when Map::insertCluster throws
// (vector::at for a cell mean outside the map).
void* ref_map_from_scan(const float* pts4, uint32_t n, int n_clusters, double max_range, int min_points, int size_x, int size_y, double res,
                        double max_linf, int* err) {
  if (err) *err = 0;
  std::unique_ptr<RefMap> m(new RefMap());
  m->map.initialize(map_params(size_x, size_y, res, max_linf, min_points), 0.0, 0.0);
  if (n == 0) return m.release();
  rc::navigation::ndt::Grid g;
  g.setMaxRange(max_range);
  pcl::PointCloud<pcl::PointXYZI>::Ptr cloud = cloud_from(pts4, n);
  std::vector<std::pair<double, double>> polar(n, std::make_pair(0.0, 0.0));   // pNDT only
  std::vector<int> labels(n);
  g.cluster(n_clusters, cloud, labels);
  std::vector<pcl::PointCloud<pcl::PointXYZI>> clusters;
  std::vector<std::vector<std::pair<double, double>>> polar_points;
  g.labelClouds(cloud, polar, labels, clusters, polar_points);
  try {
    for (size_t i = 0; i < clusters.size(); ++i) m->map.insertCluster(clusters[i], polar_points[i]);
  } catch (const std::out_of_range&) {
    if (err) *err = 1;
    return nullptr;
  }
  return m.release();
}
void* ref_map_empty(int min_points, int size_x, int size_y, double res, double max_linf) {
  RefMap* m = new RefMap();
  m->map.initialize(map_params(size_x, size_y, res, max_linf, min_points), 0.0, 0.0);
  return m;
}
void* ref_map_clone(const void* h) { return new RefMap(*static_cast<const RefMap*>(h)); }
void ref_map_free(void* h) { delete static_cast<RefMap*>(h); }
uint32_t ref_map_n_cells(const void* h) { return static_cast<const RefMap*>(h)->map.get_n_cells(); }
uint32_t ref_map_n_slots(const void* h) { return static_cast<uint32_t>(static_cast<const RefMap*>(h)->map.getGridIndizes().size()); }
// cells [n][12] = mean (x, y, intensity) + row-major 3x3 covariance; npts [n]; slot [n_slots]; any may be NULL
void ref_map_get(const void* h, float* cells, uint32_t* npts, int32_t* slot) {
  const Map& map = static_cast<const RefMap*>(h)->map;
  const std::vector<Cell> cs = map.getCells();
  for (size_t i = 0; i < cs.size(); ++i) {
    const Eigen::Vector3f mu = cs[i].getIntensityMean();
    const Eigen::Matrix3f cv = cs[i].getIntensityCov();
    if (cells) {
      for (int k = 0; k < 3; ++k) cells[12 * i + k] = mu(k);
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cells[12 * i + 3 + 3 * r + c] = cv(r, c);
    }
    if (npts) npts[i] = static_cast<uint32_t>(cs[i].getNumCells());
  }
  if (slot) { const std::vector<int> g = map.getGridIndizes(); for (size_t i = 0; i < g.size(); ++i) slot[i] = g[i]; }
}
// a5: Map::transformMap(Eigen::Affine2f(pose.cast<float>().matrix()))  (local_fuser.cpp:338)
void ref_map_transform(void* h, const double* pose) { static_cast<RefMap*>(h)->map.transformMap(affine_from_se2d(pose)); }
// the float affine the reference would hand to transformMap for this pose: (cos, sin, tx, ty) of its matrix
void ref_affine_from_se2d(const double* pose, float* csxy) {
  const Eigen::Affine2f a = affine_from_se2d(pose);
  csxy[0] = a.matrix()(0, 0); csxy[1] = a.matrix()(1, 0); csxy[2] = a.matrix()(0, 2); csxy[3] = a.matrix()(1, 2);
}
// Transform::rotation() of the 3-D lift Cell::transformCell builds (ndt_cell.cpp:118-122) for a float affine (cos, sin, tx, ty): row-major 3x3
void ref_rotation_of_affine(const float* csxy, float* rot9) {
  Eigen::Affine3f t3 = Eigen::Affine3f::Identity();
  t3.matrix()(0, 0) = csxy[0]; t3.matrix()(0, 1) = -csxy[1]; t3.matrix()(1, 0) = csxy[1]; t3.matrix()(1, 1) = csxy[0];
  const Eigen::Matrix3f r = t3.rotation();
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) rot9[3 * i + j] = r(i, j);
}
// a6: Map::mergeMapCell
int ref_map_merge(void* fixed, const void* moving) {
  try { static_cast<RefMap*>(fixed)->map.mergeMapCell(static_cast<const RefMap*>(moving)->map); }
  catch (const std::out_of_range&) { return 1; }
  return 0;
}
// a7 as Matcher::addNDTFactor drives it (ndt_matcher.cpp:200-214): metric 0 = lookup_mahalanobis (copy of moving cell i, transformCell by
// the float-cast initial guess, getClosestCells(cell)); metric 1 = Euclidean lookup of the transformed mean (SE2f * mean_xy).
// Returns the number of neighbours written to idx (<= k), -1 when the reference throws.
int ref_closest_cells(const void* fixed, const void* moving, uint32_t i, const double* pose, int k, int metric, uint64_t* idx) {
  const Map& F = static_cast<const RefMap*>(fixed)->map;
  const Map& M = static_cast<const RefMap*>(moving)->map;
  std::vector<size_t> indizes;
  try {
    if (metric == 0) {
      Cell query_cell = M.getCells()[i];
      const Eigen::Affine2f affine_trans = affine_from_se2d(pose);
      query_cell.transformCell(affine_trans);
      F.getClosestCells(query_cell, k, indizes);
    } else {
      // Sophus SE2f * Vector2f = so2 * p + translation, so2 * p = (re x - im y, im x + re y)   (sophus/so2.hpp operator*, se2.hpp operator*)
      const Eigen::Affine2f a = affine_from_se2d(pose);
      const float re = a.matrix()(0, 0), im = a.matrix()(1, 0);
      const Eigen::Vector2f mu = M.getCells()[i].getMean();
      Eigen::Vector2f q(re * mu(0) - im * mu(1), im * mu(0) + re * mu(1));
      q(0) = q(0) + a.matrix()(0, 2); q(1) = q(1) + a.matrix()(1, 2);
      F.getClosestCells(q, k, indizes);
    }
  } catch (const std::out_of_range&) { return -1; }
  for (size_t j = 0; j < indizes.size(); ++j) idx[j] = indizes[j];
  return static_cast<int>(indizes.size());
}
// f2: Map::calculateCSDivergence.  The reference never initialises its three accumulators (ndt_map.cpp:43-46): the value returned is the
// reference's arithmetic on top of whatever those stack slots held.  Kept for completeness; tests do not compare against it.
double ref_cs_divergence(void* fixed, const void* moving) { return static_cast<RefMap*>(fixed)->map.calculateCSDivergence(static_cast<const RefMap*>(moving)->map); }

// ---- f1: RadarPreprocessor::filterScan ---------------------------------------------------------------------------------------
// raw4 [n_az * n_bins][4] (x, y, z, intensity) azimuth-major = the organised cloud (height = n_az, width = n_bins);
// sensor_to_base: row-major 3x4 of initial_transform_radar_baselink.  out4 / polar receive the filtered points in the base frame and
// their (angle, range); max_det (may be NULL) the (angle, range, intensity) of each emitted peak.  Returns the number of kept points,
// -1 when the reference throws (pcl at()).
int ref_filter_scan(const float* raw4, uint32_t n_az, uint32_t n_bins, double min_range, double max_range, double min_intensity,
                    double beam_distance_increment_threshold, const float* sensor_to_base, float* out4, uint32_t cap, double* polar,
                    double* max_det, uint32_t* n_max_det) {
  rc::navigation::ndt::RadarPreprocessorParameters p;
  p.n_clusters = 1; p.min_intensity = min_intensity; p.min_range = min_range; p.max_range = max_range; p.cluster_all_points = false;
  p.beam_distance_increment_threshold = beam_distance_increment_threshold; p.min_points_per_cell = 0; p.sensor_frame = "radar"; p.base_frame = "base_link";
  Eigen::Affine3f T = Eigen::Affine3f::Identity();
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) T.matrix()(r, c) = sensor_to_base[4 * r + c];
  rc::navigation::ndt::RadarPreprocessor pre;
  std::streambuf* old = std::cout.rdbuf(nullptr);      // initialize() prints a banner
  pre.initialize(rc::navigation::ndt::ClusteringType::Grid, p, T);
  std::cout.rdbuf(old);
  boost::shared_ptr<sensor_msgs::PointCloud2> msg(new sensor_msgs::PointCloud2());
  msg->height = n_az; msg->width = n_bins;
  msg->xyzi.assign(raw4, raw4 + static_cast<size_t>(n_az) * n_bins * 4);
  pcl::PointCloud<pcl::PointXYZI>::Ptr filtered(new pcl::PointCloud<pcl::PointXYZI>());
  std::vector<std::pair<double, double>> polar_point;
  std::vector<std::tuple<double, double, double>> max_detections;
  try { pre.filterScan(msg, filtered, polar_point, max_detections); }
  catch (const std::out_of_range&) { return -1; }
  const uint32_t n = static_cast<uint32_t>(filtered->size());
  for (uint32_t i = 0; i < n && i < cap; ++i) {
    out4[4 * i] = filtered->at(i).x; out4[4 * i + 1] = filtered->at(i).y; out4[4 * i + 2] = filtered->at(i).z; out4[4 * i + 3] = filtered->at(i).intensity;
    if (polar) { polar[2 * i] = polar_point[i].first; polar[2 * i + 1] = polar_point[i].second; }
  }
  if (n_max_det) *n_max_det = static_cast<uint32_t>(max_detections.size());
  if (max_det) for (size_t i = 0; i < max_detections.size(); ++i) {
    max_det[3 * i] = std::get<0>(max_detections[i]); max_det[3 * i + 1] = std::get<1>(max_detections[i]); max_det[3 * i + 2] = std::get<2>(max_detections[i]);
  }
  return static_cast<int>(n);
}

}  // extern "C"
