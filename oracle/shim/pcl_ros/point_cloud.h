// ORACLE SHIM (test infrastructure)
#pragma once
#include <pcl/point_cloud.h>
