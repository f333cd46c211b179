// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).  The point block travels as raw bytes of the PCL point type it was made from
// (point_step = sizeof(point)); pcl_conversions copies it back when the sizes agree.
#pragma once
#include <std_msgs/Header.h>
namespace sensor_msgs {
struct PointField { std::string name; uint32_t offset = 0; uint8_t datatype = 0; uint32_t count = 1; };
struct PointCloud2 {
    std_msgs::Header header; uint32_t height = 1, width = 0; std::vector<PointField> fields; bool is_bigendian = false;
    uint32_t point_step = 0, row_step = 0; std::vector<uint8_t> data; bool is_dense = true;
    typedef std::shared_ptr<PointCloud2 const> ConstPtr;
};
typedef std::shared_ptr<PointCloud2 const> PointCloud2ConstPtr;
}
