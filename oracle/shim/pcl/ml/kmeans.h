// ORACLE SHIM (test infrastructure): the reference includes pcl/ml/kmeans.h but the k-means clustering was removed (ClusteringType has Grid only)
#pragma once
