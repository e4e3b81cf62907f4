// ============================================================================
// ref_multih_wrapper.cpp — TEST INFRASTRUCTURE.  Compiles the REFERENCE's own hot-path
// sources — MultiH/MultiH/MultiH.cpp with MultiH.h, moduls/mode_seeking/MeanShiftClustering.h,
// moduls/homographies/Homography_Refine{HAF,3PT}Callback.h and the alpha-expansion under
// moduls/alpha_expansion/ — UNMODIFIED and IN PLACE from /root/reference into
// oracle/_ref/libmultih_ref.so (oracle/Makefile).  Nothing of the reference is copied into
// this repository; this file only holds the glue:
//   * OpenCV 3.1 (absent from the tree and the image) is replaced by oracle/cvshim/mini_cv.hpp,
//     MSVC's <ppl.h> by a sequential parallel_for, the Windows headers by empty files
//     (generated under oracle/_ref/inc by the Makefile);
//   * the reference's Utilities.hpp is skipped through its include guard (file utilities, ZNCC
//     ...); its copy of cv::LMSolverImpl (Utilities.hpp:750-860) is extracted by the Makefile into
//     oracle/_ref/inc/lm_solver_extract.hpp, also unmodified;
//   * rand() is MSVC's generator (the reference is an MSVC program; RAND_MAX 32767), seeded per call;
//   * cv::findFundamentalMat returns the F handed to ref_multih_process with an all-inlier mask
//     (the RANSAC is upstream of the hot path: F is an input of the accelerated path);
//   * FlannBasedMatcher::radiusMatch is the exactly-defined stand-in of mini_cv.hpp.
// Exported: ref_multih_process = MultiH::Process() as the reference runs it, plus per-function
// entry points (HAF, 3PT, data cost, mean-shift) used to pin oracle/multih_oracle.cpp.
// ============================================================================
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <chrono>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <numeric>
#include <random>
#include <string>
#include <thread>
#include <vector>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mini_cv.hpp"

// ---- MSVC rand() -------------------------------------------------------------------------------------------------------------------
static unsigned int g_holdrand = 1u;
static int ref_msvc_rand() {
  g_holdrand = g_holdrand * 214013u + 2531011u;
  return (int)((g_holdrand >> 16) & 0x7fff);
}
#include <cstdlib>
#undef RAND_MAX
#define RAND_MAX 32767
#define rand ref_msvc_rand

// ---- the injected fundamental matrix -----------------------------------------------------------------------------------------------
static double g_F[9];
static int g_lm_mode = 1;   // 1 = the reference's LM solver, 0 = LM leaves its parameters untouched (the linear solutions)
static bool g_degenerate = false;
namespace cv {
inline Mat findFundamentalMat(InputArray p1, InputArray, int, double, double, std::vector<uchar>& mask) {
  Mat F(3, 3, CV_64F);
  for (int k = 0; k < 9; ++k) F.at<double>(k / 3, k % 3) = g_F[k];
  mask.assign((size_t)p1.getMat().rows, (uchar)1);
  return F;
}
inline Mat findHomography(InputArray p1, InputArray, int, double, std::vector<uchar>& mask) {   // HandleDegenerateCase: out of scope
  g_degenerate = true;
  mask.assign((size_t)p1.getMat().rows, (uchar)0);
  return Mat::eye(3, 3, CV_64F);
}
}  // namespace cv

#define __UTILITIES__   // skip the reference's Utilities.hpp; its LM solver comes from the extract below
namespace cv {
#include "lm_solver_extract.hpp"
}
// LM switched off: run() returns without touching the parameters
namespace cv {
class LMSolverSwitch : public LMSolverImpl {
 public:
  LMSolverSwitch(const Ptr<LMSolver::Callback>& cb, int iters) : LMSolverImpl(cb, iters) {}
  int run(InputOutputArray p) const { return g_lm_mode ? LMSolverImpl::run(p) : 0; }
};
}  // namespace cv
#define LMSolverImpl LMSolverSwitch

// MSVC accepts a class template that is named before its declaration inside a function template; g++ needs the declaration
template <typename T> class Homography_RefineHAFCallback;
template <typename T> class Homography_Refine3PTCallback;
#include "GCoptimization.cpp"
#include "LinkedBlockList.cpp"
#include "MultiH.cpp"
#undef LMSolverImpl
#undef rand

namespace {
struct RefMultiH : public MultiH {
  using MultiH::MultiH;
  using MultiH::affinities;
  using MultiH::cluster_homographies;
  using MultiH::dst_points;
  using MultiH::epipole_2;
  using MultiH::fundamental_matrix;
  using MultiH::fundamental_matrix_ptr;
  using MultiH::GetHomography3PT;
  using MultiH::GetHomographyHAF;
  using MultiH::GetHomographyHAFNonminimal;
  using MultiH::homographies;
  using MultiH::labeling;
  using MultiH::src_points;
  void set_F(const double* F) {   // what GetFundamentalMatrixAndRefineData leaves behind (MultiH.cpp:775-799)
    fundamental_matrix = cv::Mat(3, 3, CV_64F);
    for (int k = 0; k < 9; ++k) fundamental_matrix.at<double>(k / 3, k % 3) = F[k];
    fundamental_matrix_ptr = (double*)fundamental_matrix.data;
    cv::Mat Ft = fundamental_matrix.t();
    cv::Mat FFt = fundamental_matrix * Ft;
    cv::Mat ev, evec;
    cv::eigen(FFt, ev, evec);
    epipole_2 = evec.row(evec.rows - 1);
    epipole_2 = epipole_2 / epipole_2.at<double>(2);
  }
};
cv::Mat mat_from(const double* p, int r, int c) {
  cv::Mat m(r, c, CV_64F);
  std::memcpy(m.data, p, sizeof(double) * (size_t)r * c);
  return m;
}
}  // namespace

extern "C" {

// MultiH::Process (MultiH.cpp:32-98).  labels_out/pts_out/haf_out describe the correspondences that survive the reference's
// refinement filter, in input order (the reference forgets the others); *kept_out = their number.
int ref_multih_process(const double* pts, const double* aff, const double* F, int N, double thr_F, double thr_H, double locality,
                       double lambda, int min_inliers, unsigned rng_seed, int lm_mode, int32_t* labels_out, double* H_out, int Kmax,
                       int* K_out, double* pts_out /*N x 4*/, double* haf_out /*N x 9*/, int* kept_out, int* iterations,
                       double* energy, int* degenerate) {
  g_holdrand = rng_seed;
  g_lm_mode = lm_mode;
  g_degenerate = false;
  std::memcpy(g_F, F, sizeof(g_F));
  std::vector<cv::Point2d> src(N), dst(N);
  std::vector<cv::Mat> affs(N);
  for (int i = 0; i < N; ++i) {
    src[i] = cv::Point2d(pts[4 * i], pts[4 * i + 1]);
    dst[i] = cv::Point2d(pts[4 * i + 2], pts[4 * i + 3]);
    affs[i] = mat_from(aff + 4 * i, 2, 2);
  }
  RefMultiH mh(thr_F, thr_H, locality, lambda, min_inliers);
  const bool ok = mh.Process(src, dst, affs);
  if (!ok) return 1;
  const int M = (int)mh.src_points.size();
  *kept_out = M;
  for (int i = 0; i < M; ++i) {
    labels_out[i] = i < (int)mh.labeling.size() ? mh.labeling[i] : -1;
    if (pts_out) { pts_out[4 * i] = mh.src_points[i].x; pts_out[4 * i + 1] = mh.src_points[i].y; pts_out[4 * i + 2] = mh.dst_points[i].x; pts_out[4 * i + 3] = mh.dst_points[i].y; }
    if (haf_out && i < (int)mh.homographies.size()) std::memcpy(haf_out + 9 * (size_t)i, mh.homographies[i].data, sizeof(double) * 9);
  }
  const int K = (int)mh.cluster_homographies.size();
  *K_out = K;
  for (int k = 0; k < K && k < Kmax; ++k) {
    const cv::Mat Hc = mh.cluster_homographies[k].clone();   // (a continuous copy)
    std::memcpy(H_out + 9 * (size_t)k, Hc.data, sizeof(double) * 9);
  }
  if (iterations) *iterations = mh.GetIterationNumber();
  if (energy) *energy = mh.GetEnergy();
  if (degenerate) *degenerate = g_degenerate ? 1 : 0;
  return 0;
}

// GetHomographyHAF (MultiH.cpp:850-911) per correspondence
void ref_haf_hypotheses(const double* pts, const double* aff, const double* F, int N, double* H_out) {
  RefMultiH mh;
  mh.set_F(F);
  for (int i = 0; i < N; ++i) {
    cv::Mat H(3, 3, CV_64F);
    mh.GetHomographyHAF(aff[4 * i], aff[4 * i + 1], aff[4 * i + 2], aff[4 * i + 3], pts[4 * i], pts[4 * i + 1], pts[4 * i + 2], pts[4 * i + 3], H);
    std::memcpy(H_out + 9 * (size_t)i, H.data, sizeof(double) * 9);
  }
}

// GetHomography3PT (MultiH.cpp:995-1055), do_numerical_refinement as given
void ref_homography_3pt(const double* pts1, const double* pts2, int n, const double* F, int refine, double* H_out) {
  RefMultiH mh;
  mh.set_F(F);
  g_lm_mode = 1;
  cv::Mat H(3, 3, CV_64F);
  mh.GetHomography3PT(mat_from(pts1, n, 2), mat_from(pts2, n, 2), H, refine != 0);
  const cv::Mat Hc = H.clone();
  std::memcpy(H_out, Hc.data, sizeof(double) * 9);
}

// GetHomographyHAFNonminimal (MultiH.cpp:913-993)
void ref_haf_nonminimal(const double* pts, const double* aff, int n, const double* F, int refine, double* H_out) {
  RefMultiH mh;
  mh.set_F(F);
  g_lm_mode = 1;
  cv::Mat p1(n, 2, CV_64F), p2(n, 2, CV_64F), a(n, 4, CV_64F);
  for (int i = 0; i < n; ++i) {
    p1.at<double>(i, 0) = pts[4 * i]; p1.at<double>(i, 1) = pts[4 * i + 1];
    p2.at<double>(i, 0) = pts[4 * i + 2]; p2.at<double>(i, 1) = pts[4 * i + 3];
    for (int k = 0; k < 4; ++k) a.at<double>(i, k) = aff[4 * i + k];
  }
  cv::Mat H(3, 3, CV_64F);
  mh.GetHomographyHAFNonminimal(a, p1, p2, H, refine != 0);
  const cv::Mat Hc = H.clone();
  std::memcpy(H_out, Hc.data, sizeof(double) * 9);
}

// dataEnergy (MultiH.cpp:473-504) over all (site, label), label 0 = outlier: site-major [N][K + 1]
void ref_data_cost_dense(const double* pts, int N, const double* H, int K, double lambda, double thr, int32_t* out) {
  std::vector<cv::Point2d> src(N), dst(N);
  for (int i = 0; i < N; ++i) { src[i] = cv::Point2d(pts[4 * i], pts[4 * i + 1]); dst[i] = cv::Point2d(pts[4 * i + 2], pts[4 * i + 3]); }
  std::vector<cv::Mat> hs(K);
  for (int k = 0; k < K; ++k) hs[k] = mat_from(H + 9 * (size_t)k, 3, 3);
  MultiH::EnergyDataStruct e(&src, &dst, &hs, lambda, thr * thr);
  for (int i = 0; i < N; ++i)
    for (int l = 0; l <= K; ++l) out[(size_t)i * (K + 1) + l] = dataEnergy(i, l, &e);
}
int ref_smooth_cost(int l1, int l2, double lambda) {
  std::vector<cv::Point2d> none;
  std::vector<cv::Mat> hs;
  MultiH::EnergyDataStruct e(&none, &none, &hs, lambda, 1.0);
  return smoothnessEnergy(0, 1, l1, l2, &e);
}

// MeanShiftClustering<double>::Cluster (MeanShiftClustering.h:22-157); returns the number of clusters
int ref_meanshift(const double* data, int N, int D, double bw, unsigned* rng_state, double* centres /*N x D*/, int32_t* assign) {
  g_holdrand = *rng_state;
  cv::Mat X = mat_from(data, N, D), clusters;
  std::vector<std::vector<int>> members;
  MeanShiftClustering<double> ms;
  ms.Cluster(X, bw, clusters, members);
  *rng_state = g_holdrand;
  for (int i = 0; i < N; ++i) assign[i] = -1;
  for (int c = 0; c < (int)members.size(); ++c)
    for (int i : members[c]) assign[i] = c;
  const cv::Mat cc = clusters.clone();
  std::memcpy(centres, cc.data, sizeof(double) * (size_t)cc.rows * D);
  return cc.rows;
}

}  // extern "C"
