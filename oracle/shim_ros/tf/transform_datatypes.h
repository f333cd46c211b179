// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).  tf's quaternion arithmetic (double) = the oracle's restatement
// (oracle/liorf_oracle.hpp: MapOptScalarState::setRPY / slerp / getRPY).  Third-party arithmetic: NOT pinned (tf is not in this image).
#pragma once
#include <geometry_msgs/geometry.h>
#include "liorf_oracle.hpp"
namespace tf {
typedef liorf_oracle::MapOptScalarState::Q OQ;
class Quaternion {
public:
    Quaternion() : q_{0, 0, 0, 1} {}
    Quaternion(double x, double y, double z, double w) : q_{x, y, z, w} {}
    explicit Quaternion(const OQ& q) : q_(q) {}
    void setRPY(double roll, double pitch, double yaw) { q_ = liorf_oracle::MapOptScalarState::setRPY(roll, pitch, yaw); }
    Quaternion slerp(const Quaternion& o, double t) const { return Quaternion(liorf_oracle::MapOptScalarState::slerp(q_, o.q_, t)); }
    double x() const { return q_.x; } double y() const { return q_.y; } double z() const { return q_.z; } double w() const { return q_.w; }
    const OQ& raw() const { return q_; }
private:
    OQ q_;
};
class Matrix3x3 {
public:
    explicit Matrix3x3(const Quaternion& q) : q_(q) {}
    void getRPY(double& roll, double& pitch, double& yaw) const { liorf_oracle::MapOptScalarState::getRPY(q_.raw(), roll, pitch, yaw); }
private:
    Quaternion q_;
};
struct Vector3 { double x_, y_, z_; Vector3(double x = 0, double y = 0, double z = 0) : x_(x), y_(y), z_(z) {} };
struct Transform { Quaternion q; Vector3 v; Transform() {} Transform(const Quaternion& q_, const Vector3& v_) : q(q_), v(v_) {} };
struct StampedTransform : Transform { StampedTransform() {} StampedTransform(const Transform& t, const ros::Time&, const std::string&, const std::string&) : Transform(t) {} };
inline Quaternion createQuaternionFromRPY(double roll, double pitch, double yaw) { Quaternion q; q.setRPY(roll, pitch, yaw); return q; }
inline geometry_msgs::Quaternion createQuaternionMsgFromRollPitchYaw(double roll, double pitch, double yaw) {
    Quaternion q = createQuaternionFromRPY(roll, pitch, yaw); geometry_msgs::Quaternion m; m.x = q.x(); m.y = q.y(); m.z = q.z(); m.w = q.w(); return m;
}
inline void quaternionMsgToTF(const geometry_msgs::Quaternion& m, Quaternion& q) { q = Quaternion(m.x, m.y, m.z, m.w); }
}
