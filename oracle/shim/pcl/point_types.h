// ORACLE SHIM (test infrastructure): pcl::PointXYZI with PCL 1.10's layout (x, y, z, pad = 1 | intensity + 3 pad; 32 bytes).
#pragma once
#include <cstdint>
namespace pcl {
struct alignas(16) PointXYZI {
  union { float data[4]; struct { float x, y, z; }; };
  union { struct { float intensity; }; float data_c[4]; };
  PointXYZI() : x(0.f), y(0.f), z(0.f) { data[3] = 1.0f; intensity = 0.f; data_c[1] = data_c[2] = data_c[3] = 0.f; }
  PointXYZI(float x_, float y_, float z_, float i_) : x(x_), y(y_), z(z_) { data[3] = 1.0f; intensity = i_; data_c[1] = data_c[2] = data_c[3] = 0.f; }
};
}  // namespace pcl
