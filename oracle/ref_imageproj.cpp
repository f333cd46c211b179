// ref_imageproj.cpp — extern "C" surface over the REFERENCE's own ImageProjection node: /root/reference/src/imageProjection.cpp is compiled UNCHANGED,
// from where it lies (textually included below — the class has no header; nothing is copied into the repo), together with lib/common_lib.cpp, against
// the header stand-ins of oracle/shim_ros/.  Output: oracle/_ref/libliorf_ref_imageproj.so (git-ignored, travels to the GPU box).  Recipe: oracle/Makefile.
//
// TEST INFRASTRUCTURE.  The node is driven the way ROS drives it — only through its public handlers: IMU samples go in through imuHandler (:149), scans
// through cloudHandler (:191), and the liorf/cloud_info message it publishes (:608-613) is read from the stand-in's outbox.  What this pins:
// cachePointCloud's queueing and time stamps (:206-329), deskewInfo's IMU gate (:331-348), imuDeskewInfo (:350-409), findRotation (:493-518),
// deskewPoint's first-kept-point reference (:536-566) and projectPointCloud's filters (:568-598) are the reference's own code.  Not pinned: the arithmetic
// of pcl::getTransformation and Eigen's Affine3f inverse / product (stand-ins forward to oracle/liorf_oracle.hpp) and tf's quaternion → RPY.
#include <cmath>
#include <cstring>
#include <iostream>
#include <sstream>
#define main liorf_ref_imageproj_main_unused
#include "imageProjection.cpp"
#undef main

namespace {
struct Quiet { std::streambuf* old; std::ostringstream sink; Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {} ~Quiet() { std::cout.rdbuf(old); } };
const char* kInfoTopic = "liorf/deskew/cloud_info";
}

extern "C" {

void refip_param_num(const char* name, double v) { ros::shim::params().num[name] = v; }
void refip_param_str(const char* name, const char* v) { ros::shim::params().str[name] = v; }

void* refip_create() {
    Quiet q;
    if (!ros::shim::params().str.count("liorf/sensor")) ros::shim::params().str["liorf/sensor"] = "velodyne";
    if (!common_lib_) common_lib_ = std::make_shared<CommonLib::common_lib>("mapping");                 // main() :612
    return new ImageProjection();
}
void refip_destroy(void* h) { delete (ImageProjection*)h; }

// one sensor_msgs/Imu through imuHandler: stamp, angular velocity, orientation quaternion (x, y, z, w)
void refip_imu(void* h, double stamp, const double* gyro_xyz, const double* quat_xyzw) {
    auto m = std::make_shared<sensor_msgs::Imu>();
    m->header.stamp.fromSec(stamp);
    m->angular_velocity.x = gyro_xyz[0]; m->angular_velocity.y = gyro_xyz[1]; m->angular_velocity.z = gyro_xyz[2];
    if (quat_xyzw) { m->orientation.x = quat_xyzw[0]; m->orientation.y = quat_xyzw[1]; m->orientation.z = quat_xyzw[2]; m->orientation.w = quat_xyzw[3]; }
    else m->orientation.w = 1.0;
    ((ImageProjection*)h)->imuHandler(m);
}

// one sensor_msgs/PointCloud2 through cloudHandler.  raw = n records {x, y, z, intensity f32, ring u16, pad u16, time f32} (24 bytes, the payload of
// VelodynePointXYZIRT); has_time = whether the message carries a "time" field (deskewFlag, :313-326).  Returns how many cloud_info messages the node has
// published so far (the node holds two scans back: the scan handled by this call is the one pushed two calls earlier, :209-215).
long refip_cloud(void* h, double stamp, const void* raw, int n, int has_time) {
    struct Raw { float x, y, z, i; uint16_t ring, pad; float time; };
    const Raw* r = (const Raw*)raw;
    pcl::PointCloud<PointXYZIRT> c; c.points.resize((size_t)n); c.width = (uint32_t)n;
    for (int k = 0; k < n; ++k) { PointXYZIRT p; p.x = r[k].x; p.y = r[k].y; p.z = r[k].z; p.intensity = r[k].i; p.ring = r[k].ring; p.time = r[k].time; c.points[k] = p; }
    auto m = std::make_shared<sensor_msgs::PointCloud2>();
    pcl::toROSMsg(c, *m);
    m->header.stamp.fromSec(stamp);
    const char* names[] = {"x", "y", "z", "intensity", "ring", "time"};
    for (int k = 0; k < (has_time ? 6 : 5); ++k) { sensor_msgs::PointField f; f.name = names[k]; m->fields.push_back(f); }
    ((ImageProjection*)h)->cloudHandler(m);
    return ros::shim::outcount<liorf::cloud_info>()[kInfoTopic];
}

// the last published cloud_info: header stamp, imuAvailable / odomAvailable, imu{Roll,Pitch,Yaw}Init, and the deskewed cloud (x, y, z, intensity); returns its size
int refip_last_info(double* stamp, int* avail2, float* rpy3, float* xyzi_out, int cap) {
    auto& box = ros::shim::outbox<liorf::cloud_info>();
    auto it = box.find(kInfoTopic);
    if (it == box.end()) return -1;
    const liorf::cloud_info& ci = it->second;
    if (stamp) *stamp = ci.header.stamp.toSec();
    if (avail2) { avail2[0] = (int)ci.imuAvailable; avail2[1] = (int)ci.odomAvailable; }
    if (rpy3) { rpy3[0] = ci.imuRollInit; rpy3[1] = ci.imuPitchInit; rpy3[2] = ci.imuYawInit; }
    pcl::PointCloud<PointType> c; pcl::fromROSMsg(ci.cloud_deskewed, c);
    const int n = (int)c.points.size();
    for (int k = 0; k < n && k < cap; ++k) { xyzi_out[4 * k] = c.points[k].x; xyzi_out[4 * k + 1] = c.points[k].y; xyzi_out[4 * k + 2] = c.points[k].z; xyzi_out[4 * k + 3] = c.points[k].intensity; }
    return n;
}

}  // extern "C"
