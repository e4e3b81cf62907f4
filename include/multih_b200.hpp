// ============================================================================
// multih_b200.hpp — header-only C++ shim over the C ABI (multih_b200.h) with the
// reference's class and method names (MultiH/MultiH/MultiH.h:20-149), so that an
// `ApplyMultiH`-style caller (MultiH/MultiH/main.cpp:232-297) compiles unchanged
// apart from passing the fundamental matrix.
//
//   #define MULTIH_B200_WITH_OPENCV   -> Process() takes the reference's own types
//                                        (std::vector<cv::Point2d>, std::vector<cv::Mat>)
//   otherwise                          -> Process() takes flat double arrays
//
// Differences from the reference class, all deliberate:
//  * F is an input: GetFundamentalMatrixAndRefineData (MultiH.cpp:770-848) is upstream of the accelerated path;
//  * GetDestinationPoints returns the destination points (the reference returns src: MultiH.h:64);
//  * GetHomography(idx) stays 1-based (MultiH.h:69); labels: -1 = outlier, 0..K-1 (MultiH.h:61-62);
//  * Process() runs what the reference's Process() runs after its F-RANSAC: the per-correspondence refinement / affine
//    consistency filter (MultiH.cpp:807-838), the alternating optimisation and HomographyCompatibilityCheck (:76-86);
//    correspondences the filter drops disappear from the getters, as in the reference (which only keeps the survivors);
//  * no LM polish after the linear fits (its callbacks read out of bounds, Homography_RefineHAFCallback.h:148-151) and no
//    HandleDegenerateCase (cv::findHomography): with K <= 1 the single homography's inliers are returned.
// ============================================================================
#pragma once
#include <array>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "multih_b200.h"
#ifdef MULTIH_B200_WITH_OPENCV
#include <opencv2/core.hpp>
#endif

namespace multih_b200 {

class MultiH {
 public:
  // MultiH.h:49-53 (same defaults: MultiH.h:7-10)
  explicit MultiH(double thr_fund_mat = 3.0, double thr_hom = 2.5, double locality = 0.002, double lambda = 0.5,
                  int minimum_inlier_number = 0, int device = 0) {
    mh_default_params(&params_);
    params_.thr_fundamental = thr_fund_mat;
    params_.thr_homography = thr_hom;
    params_.locality = locality;
    params_.lambda = lambda;
    params_.min_inliers = minimum_inlier_number;
    params_.prefilter = 1;            // MultiH.cpp:807-838
    params_.compatibility_check = 1;  // MultiH.cpp:76-86
    const mh_status st = mh_create(&params_, device, &ctx_);
    if (st != MH_OK) throw std::runtime_error("multih_b200: no usable CUDA device (there is no CPU fallback)");
  }
  ~MultiH() { Release(); }
  MultiH(const MultiH&) = delete;
  MultiH& operator=(const MultiH&) = delete;
  void Release() { if (ctx_) { mh_destroy(ctx_); ctx_ = nullptr; } }  // MultiH.cpp:28-30

  // MultiH::Process (MultiH.cpp:32-98).  pts: N x (x1 y1 x2 y2), affines: N x (a11 a12 a21 a22), F row-major with
  // x2^T F x1 = 0.  Returns false on the reference's own refusal (fewer than 8 correspondences, MultiH.cpp:44-50).
  bool Process(const double* pts, const double* affines, const double F[9], int n) {
    std::printf("[Multi-H] Processing has been started.\n");  // MultiH.cpp:34
    src_.resize(n); dst_.resize(n); aff_.assign(affines, affines + 4 * (size_t)n);
    for (int i = 0; i < n; ++i) { src_[i] = {pts[4 * i], pts[4 * i + 1]}; dst_[i] = {pts[4 * i + 2], pts[4 * i + 3]}; }
    labeling_.assign(n, -1);
    homographies_.assign(9 * (size_t)kMaxClusters, 0.0);
    int32_t k = 0;
    const mh_status st = mh_process(ctx_, pts, affines, F, n, labeling_.data(), homographies_.data(), kMaxClusters, &k);
    if (st != MH_OK) {
      std::fprintf(stderr, "%s\n", mh_last_error(ctx_));
      labeling_.clear();
      cluster_number_ = 0;
      return false;
    }
    cluster_number_ = k;
    // the reference's getters only know the correspondences that survived its filter (label -2 = dropped here)
    size_t m = 0;
    for (int i = 0; i < n; ++i)
      if (labeling_[i] > -2) {
        labeling_[m] = labeling_[i]; src_[m] = src_[i]; dst_[m] = dst_[i];
        for (int k4 = 0; k4 < 4; ++k4) aff_[4 * m + k4] = aff_[4 * (size_t)i + k4];
        ++m;
      }
    labeling_.resize(m); src_.resize(m); dst_.resize(m); aff_.resize(4 * m);
    return true;
  }
#ifdef MULTIH_B200_WITH_OPENCV
  bool Process(const std::vector<cv::Point2d>& src, const std::vector<cv::Point2d>& dst, const std::vector<cv::Mat>& affines,
               const cv::Mat& F) {
    if (dst.size() != src.size() || affines.size() != src.size()) { std::fprintf(stderr, "Error: Features are not set!\n"); return false; }
    std::vector<double> p(4 * src.size()), a(4 * src.size());
    for (size_t i = 0; i < src.size(); ++i) {
      p[4 * i] = src[i].x; p[4 * i + 1] = src[i].y; p[4 * i + 2] = dst[i].x; p[4 * i + 3] = dst[i].y;
      for (int r = 0; r < 2; ++r) for (int c = 0; c < 2; ++c) a[4 * i + 2 * r + c] = affines[i].at<double>(r, c);
    }
    double f[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) f[3 * r + c] = F.at<double>(r, c);
    return Process(p.data(), a.data(), f, (int)src.size());
  }
  cv::Mat GetHomography(int idx) const {  // 1-based, MultiH.h:69
    cv::Mat H(3, 3, CV_64F);
    for (int k = 0; k < 9; ++k) H.at<double>(k / 3, k % 3) = homographies_[9 * (size_t)(idx - 1) + k];
    return H;
  }
#else
  std::array<double, 9> GetHomography(int idx) const {  // 1-based, MultiH.h:69
    std::array<double, 9> H;
    for (int k = 0; k < 9; ++k) H[k] = homographies_[9 * (size_t)(idx - 1) + k];
    return H;
  }
#endif
  int GetLabel(int idx) const { return labeling_[idx]; }                                  // MultiH.h:61
  void GetLabels(std::vector<int>& out) const { out.assign(labeling_.begin(), labeling_.end()); }  // MultiH.h:62
  void GetSourcePoints(std::vector<std::array<double, 2>>& out) const { out = src_; }     // MultiH.h:63
  void GetDestinationPoints(std::vector<std::array<double, 2>>& out) const { out = dst_; }  // MultiH.h:64 (fixed)
  void GetAffinities(std::vector<double>& out) const { out = aff_; }                      // MultiH.h:65
  int GetPointNumber() const { return (int)labeling_.size(); }                            // MultiH.h:66
  int GetClusterNumber() const { return cluster_number_; }                                // MultiH.h:67
  int GetIterationNumber() const { return mh_get_iterations(ctx_); }                      // MultiH.h:68
  double GetEnergy() const { return mh_get_energy(ctx_); }                                // MultiH.h:74
  double GetHomographyThreshold() const { return params_.thr_homography; }                // MultiH.h:75
  mh_ctx* context() { return ctx_; }

 private:
  static constexpr int kMaxClusters = 4096;
  mh_params params_{};
  mh_ctx* ctx_ = nullptr;
  std::vector<int32_t> labeling_;
  std::vector<double> homographies_, aff_;
  std::vector<std::array<double, 2>> src_, dst_;
  int cluster_number_ = 0;
};

}  // namespace multih_b200
