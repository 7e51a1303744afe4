// ORACLE SHIM (test infrastructure): just enough of std_msgs::Header / sensor_msgs::PointCloud2 for RadarPreprocessor's signatures.
// The message payload is carried as packed (x, y, z, intensity) float32 records — what the Oxford radar driver's PointCloud2 holds
// for the fields pcl::fromROSMsg extracts into PointXYZI.
#pragma once
#include <boost/shared_ptr.hpp>
#include <cstdint>
#include <string>
#include <vector>
namespace std_msgs { struct Header { std::uint32_t seq = 0; double stamp = 0.0; std::string frame_id; }; }
namespace sensor_msgs {
struct PointCloud2 {
  typedef boost::shared_ptr<PointCloud2> Ptr;
  typedef boost::shared_ptr<const PointCloud2> ConstPtr;
  std_msgs::Header header;
  std::uint32_t height = 0, width = 0;
  std::vector<float> xyzi;   // height * width * 4
};
}  // namespace sensor_msgs
