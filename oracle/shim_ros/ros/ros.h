// oracle/shim_ros — TEST INFRASTRUCTURE.  Stand-ins for the ROS / PCL / Eigen / OpenCV / tf / GTSAM names the reference's two nodes touch, so that
// /root/reference/src/mapOptmization.cpp and src/imageProjection.cpp compile UNCHANGED into oracle/_ref/ (none of those libraries is in this image)
// and can be driven message by message: handlers are called directly, published messages are kept in an outbox.
// What this pins: the reference's own control flow, indexing, thresholds, float / double mixes, member state and call order.
// What it does NOT pin: the arithmetic inside the third-party calls (VoxelGrid, kd-tree, ColPivHouseholderQR, cv::solve / eigen / inv / gemm,
// getTransformation, Affine3f inverse / product, tf quaternions) — those forward to oracle/liorf_oracle.hpp's restatements.
//
// ros/ros.h: time, parameters (a global store the harness fills before constructing a node), publishers that keep the last message per topic.
#pragma once
#include <unistd.h>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }

namespace ros {

struct Time {
    double t = 0.0;
    Time() {}
    explicit Time(double s) : t(s) {}
    double toSec() const { return t; }
    Time& fromSec(double s) { t = s; return *this; }
    static Time now() { return Time(); }
};
struct Duration { explicit Duration(double = 0) {} void sleep() {} };
struct Rate { explicit Rate(double) {} void sleep() {} };
struct TransportHints { TransportHints& tcpNoDelay(bool = true) { return *this; } };

namespace shim {
struct Params { std::map<std::string, double> num; std::map<std::string, std::string> str; std::map<std::string, std::vector<double>> vec; };
inline Params& params() { static Params p; return p; }
inline std::map<std::string, int>& subscribers() { static std::map<std::string, int> s; return s; }
template <class M> std::map<std::string, M>& outbox() { static std::map<std::string, M> b; return b; }        // last message per topic
template <class M> std::map<std::string, long>& outcount() { static std::map<std::string, long> b; return b; }
inline void assign(std::string& var, const std::string& name, const std::string& def) { auto it = params().str.find(name); var = it == params().str.end() ? def : it->second; }
inline void assign(std::vector<double>& var, const std::string& name, const std::vector<double>& def) { auto it = params().vec.find(name); var = it == params().vec.end() ? def : it->second; }
template <class T> void assign(T& var, const std::string& name, const T& def) { auto it = params().num.find(name); var = it == params().num.end() ? def : (T)it->second; }
}  // namespace shim

struct Publisher {
    std::string topic;
    template <class M> void publish(const M& m) const { shim::outbox<M>()[topic] = m; ++shim::outcount<M>()[topic]; }
    int getNumSubscribers() const { auto it = shim::subscribers().find(topic); return it == shim::subscribers().end() ? 0 : it->second; }
};
struct Subscriber {};
struct ServiceServer {};

struct NodeHandle {
    template <class T> void param(const std::string& name, T& var, const T& def) const { shim::assign(var, name, def); }
    template <class M, class C> Subscriber subscribe(const std::string&, uint32_t, void (C::*)(const std::shared_ptr<M const>&), C*, const TransportHints& = TransportHints()) { return Subscriber(); }
    template <class M> Publisher advertise(const std::string& topic, uint32_t, bool = false) { Publisher p; p.topic = topic; return p; }
    template <class C, class Req, class Res> ServiceServer advertiseService(const std::string&, bool (C::*)(Req&, Res&), C*) { return ServiceServer(); }
};

inline void init(int&, char**, const std::string&) {}
inline bool ok() { return false; }
inline void shutdown() {}
inline void spin() {}
struct MultiThreadedSpinner { explicit MultiThreadedSpinner(int = 0) {} void spin() {} };

}  // namespace ros

#define ROS_INFO(...) do {} while (0)
#define ROS_WARN(...) do {} while (0)
#define ROS_ERROR(...) do {} while (0)
#define ROS_DEBUG(...) do {} while (0)
#define ROS_INFO_STREAM(x) do {} while (0)
#define ROS_WARN_STREAM(x) do {} while (0)
#define ROS_ERROR_STREAM(x) do {} while (0)
#define ROS_DEBUG_STREAM(x) do {} while (0)
