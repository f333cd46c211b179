// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).  cv::Mat (CV_32F only) as LMOptimization uses it.  The arithmetic forwards to the oracle's
// restatement of OpenCV's small-matrix kernels, which tests/golden/cv2_lm6.npz pins against OpenCV 4.13 (cv_qr_solve6, cv_jacobi6, cv_lu_invert6;
// products with a double accumulator over ascending k, as cv::gemm does for small float matrices).
#pragma once
#include <cstring>
#include <stdexcept>
#include <vector>
#include "liorf_oracle.hpp"
#define CV_32F 5
namespace cv {
struct Scalar { double v[4]; static Scalar all(double x) { Scalar s; s.v[0] = s.v[1] = s.v[2] = s.v[3] = x; return s; } };
enum { DECOMP_LU = 0, DECOMP_SVD = 1, DECOMP_EIG = 2, DECOMP_CHOLESKY = 3, DECOMP_QR = 4 };
class Mat {
public:
    int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c, int type, const Scalar& s) : rows(r), cols(c), d_((size_t)r * c, (float)s.v[0]) { if (type != CV_32F) throw std::runtime_error("shim_ros cv::Mat: CV_32F only"); }
    template <class T> T& at(int i, int j) { return d_[(size_t)i * cols + j]; }
    template <class T> const T& at(int i, int j) const { return d_[(size_t)i * cols + j]; }
    void copyTo(Mat& o) const { o = *this; }
    float* ptr() { return d_.data(); }
    const float* ptr() const { return d_.data(); }
    Mat inv() const { if (rows != 6 || cols != 6) throw std::runtime_error("shim_ros cv::Mat::inv: 6x6 only"); Mat o(6, 6, CV_32F, Scalar::all(0)); liorf_oracle::cv_lu_invert6(ptr(), o.ptr()); return o; }
private:
    std::vector<float> d_;
};
inline Mat operator*(const Mat& a, const Mat& b) {
    Mat o(a.rows, b.cols, CV_32F, Scalar::all(0));
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < b.cols; ++j) { double s = 0; for (int k = 0; k < a.cols; ++k) s += (double)a.at<float>(i, k) * (double)b.at<float>(k, j); o.at<float>(i, j) = (float)s; }
    return o;
}
inline void transpose(const Mat& a, Mat& o) { Mat t(a.cols, a.rows, CV_32F, Scalar::all(0)); for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) t.at<float>(j, i) = a.at<float>(i, j); o = t; }
inline bool solve(const Mat& A, const Mat& B, Mat& X, int flags) {
    if (flags != DECOMP_QR || A.rows != 6 || A.cols != 6 || B.rows != 6 || B.cols != 1) throw std::runtime_error("shim_ros cv::solve: 6x6 DECOMP_QR only");
    X = Mat(6, 1, CV_32F, Scalar::all(0));
    return liorf_oracle::cv_qr_solve6(A.ptr(), B.ptr(), X.ptr());
}
inline bool eigen(const Mat& A, Mat& E, Mat& V) {           // eigenvalues descending as a 6x1 column (the caller's 1x6 buffer is replaced, as cv::eigen does), rows of V = eigenvectors
    if (A.rows != 6 || A.cols != 6) throw std::runtime_error("shim_ros cv::eigen: 6x6 only");
    E = Mat(6, 1, CV_32F, Scalar::all(0)); V = Mat(6, 6, CV_32F, Scalar::all(0));
    liorf_oracle::cv_jacobi6(A.ptr(), E.ptr(), V.ptr());
    return true;
}
}  // namespace cv
