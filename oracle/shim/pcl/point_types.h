// oracle/shim — the one PCL type include/Scancontext.cpp reads (x, y, z of pcl::PointXYZI).  TEST INFRASTRUCTURE.
#pragma once
namespace pcl { struct PointXYZI { float x, y, z, intensity; }; }
