// ORACLE SHIM (test infrastructure): pcl::transformPointCloud(cloud_in, cloud_out, Eigen::Affine3f) as PCL 1.10 evaluates it
// for dense clouds (common/impl/transforms.hpp, detail::Transformer<float>, SSE2 path): with c[k] = column k of the 4x4 matrix,
//   tgt = x*c[0] + (y*c[1] + (z*c[2] + c[3]))   per component, all four fields of the point copied first (copy_all_fields).
#pragma once
#include <Eigen/Core>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
namespace pcl {
template <typename PointT>
void transformPointCloud(const PointCloud<PointT>& cloud_in, PointCloud<PointT>& cloud_out, const Eigen::Affine3f& transform, bool copy_all_fields = true) {
  (void)copy_all_fields;
  if (&cloud_in != &cloud_out) {
    cloud_out.header = cloud_in.header; cloud_out.is_dense = cloud_in.is_dense; cloud_out.width = cloud_in.width; cloud_out.height = cloud_in.height;
    cloud_out.points.reserve(cloud_in.points.size());
    cloud_out.points.assign(cloud_in.points.begin(), cloud_in.points.end());
  }
  const Eigen::Matrix4f& tf = transform.matrix();
  for (std::size_t i = 0; i < cloud_out.points.size(); ++i) {
    const float sx = cloud_in.points[i].data[0], sy = cloud_in.points[i].data[1], sz = cloud_in.points[i].data[2];
    float t[4];
    for (int r = 0; r < 4; ++r) t[r] = sx * tf(r, 0) + (sy * tf(r, 1) + (sz * tf(r, 2) + tf(r, 3)));
    for (int r = 0; r < 4; ++r) cloud_out.points[i].data[r] = t[r];
  }
}
}  // namespace pcl
