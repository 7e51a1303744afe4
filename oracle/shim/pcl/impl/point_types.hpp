// ORACLE SHIM (test infrastructure)
#pragma once
#include <pcl/point_types.h>
