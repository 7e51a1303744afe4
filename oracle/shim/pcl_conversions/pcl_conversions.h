// ORACLE SHIM (test infrastructure): pcl::fromROSMsg / pcl_conversions::toPCL for the shim message type (field copy, no arithmetic).
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <sensor_msgs/PointCloud.h>
namespace pcl_conversions {
inline void toPCL(const std_msgs::Header& h, pcl::PCLHeader& out) { out.seq = h.seq; out.stamp = static_cast<std::uint64_t>(h.stamp * 1e6); out.frame_id = h.frame_id; }
}  // namespace pcl_conversions
namespace pcl {
inline void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<PointXYZI>& cloud) {
  const std::size_t n = static_cast<std::size_t>(msg.height) * msg.width;
  cloud.points.resize(n); cloud.width = msg.width; cloud.height = msg.height; cloud.is_dense = true;
  for (std::size_t i = 0; i < n; ++i) cloud.points[i] = PointXYZI(msg.xyzi[4 * i], msg.xyzi[4 * i + 1], msg.xyzi[4 * i + 2], msg.xyzi[4 * i + 3]);
}
}  // namespace pcl
