// ORACLE SHIM (test infrastructure): the reference includes pcl/common/centroid.h but calls nothing from it on this path
#pragma once
