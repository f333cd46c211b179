// oracle/shim — pcl::PointCloud<T> as far as include/Scancontext.cpp uses it (`.points`).  TEST INFRASTRUCTURE.
#pragma once
#include <vector>
namespace pcl { template <class T> struct PointCloud { std::vector<T> points; }; }
