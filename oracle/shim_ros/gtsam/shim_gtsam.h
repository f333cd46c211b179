// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).  GTSAM is outside the path (the factor graph stays on the CPU and is out of
// scope).  This stand-in keeps saveKeyFramesAndFactor compiling and gives it the answer iSAM2 gives for an odometry chain without loop / GPS factors:
// the estimate of a newly inserted pose is its initial value.  Pose3 / Rot3 carry (roll, pitch, yaw, x, y, z) verbatim, so "the optimised pose is the
// key pose" — exactly what liorf_process_frame and the oracle pipeline do.  Nothing here optimises anything.
#pragma once
#include <map>
#include <memory>
#include <vector>
#include <Eigen/Dense>
namespace gtsam {
typedef unsigned long Key;
struct Point3 { double x_, y_, z_; Point3(double x = 0, double y = 0, double z = 0) : x_(x), y_(y), z_(z) {} double x() const { return x_; } double y() const { return y_; } double z() const { return z_; } };
struct Rot3 {
    double r_ = 0, p_ = 0, y_ = 0;
    static Rot3 RzRyRx(double roll, double pitch, double yaw) { Rot3 r; r.r_ = roll; r.p_ = pitch; r.y_ = yaw; return r; }
    double roll() const { return r_; } double pitch() const { return p_; } double yaw() const { return y_; }
};
struct Pose3 {
    Rot3 R; Point3 t;
    Pose3() {}
    Pose3(const Rot3& R_, const Point3& t_) : R(R_), t(t_) {}
    const Point3& translation() const { return t; }
    const Rot3& rotation() const { return R; }
    Pose3 between(const Pose3&) const { return Pose3(); }          // only ever stored into factors, which this stand-in ignores
};
struct Vector {
    std::vector<double> v; size_t k = 0;
    explicit Vector(int n = 0) : v((size_t)n, 0.0) {}
    struct Comma { Vector* o; Comma& operator,(double x) { if (o->k < o->v.size()) o->v[o->k++] = x; return *this; } Vector finished() const { return *o; } };
    Comma operator<<(double x) { k = 0; if (!v.empty()) v[k++] = x; return Comma{this}; }
};
namespace noiseModel {
struct Base { typedef std::shared_ptr<Base> shared_ptr; virtual ~Base() {} };
struct Diagonal : Base { typedef std::shared_ptr<Base> shared_ptr; static shared_ptr Variances(const Vector&) { return std::make_shared<Diagonal>(); } };
namespace mEstimator { struct Cauchy { typedef std::shared_ptr<Cauchy> shared_ptr; static shared_ptr Create(double) { return std::make_shared<Cauchy>(); } }; }
struct Robust : Base { static Base::shared_ptr Create(const mEstimator::Cauchy::shared_ptr&, const Base::shared_ptr& n) { return n; } };
}  // namespace noiseModel
typedef noiseModel::Base::shared_ptr SharedNoiseModel;
template <class T> struct PriorFactor { PriorFactor(Key, const T&, const SharedNoiseModel&) {} };
template <class T> struct BetweenFactor { BetweenFactor(Key, Key, const T&, const SharedNoiseModel&) {} };
struct GPSFactor { GPSFactor(Key, const Point3&, const SharedNoiseModel&) {} };
struct NonlinearFactorGraph { template <class F> void add(const F&) {} void resize(size_t) {} };
struct Values {
    std::map<Key, Pose3> m;
    void insert(Key k, const Pose3& p) { m[k] = p; }
    template <class T> const T& at(Key k) const { return m.at(k); }
    size_t size() const { return m.size(); }
    void clear() { m.clear(); }
};
struct ISAM2Params { double relinearizeThreshold = 0.1; int relinearizeSkip = 1; };
struct ISAM2 {
    Values est;
    explicit ISAM2(const ISAM2Params& = ISAM2Params()) {}
    void update(const NonlinearFactorGraph&, const Values& initial) { for (auto& kv : initial.m) est.m[kv.first] = kv.second; }
    void update() {}
    Values calculateEstimate() const { return est; }
    Eigen::MatrixXd marginalCovariance(Key) const { return Eigen::MatrixXd::Zero(6, 6); }
};
namespace symbol_shorthand { inline Key X(unsigned long j) { return j; } inline Key V(unsigned long j) { return j; } inline Key B(unsigned long j) { return j; } inline Key G(unsigned long j) { return j; } }
}  // namespace gtsam
