// Shared definitions of libmultih_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/multih_b200.h"

namespace mh {

// Pair geometry in normalised image coordinates, passed to kernels by value
// (kernel parameters live in the constant bank, so every use is a uniform load).
struct GeomF {
  float F[9];      // F' = T2^-T F T1^-1, row-major, x2'^T F' x1' = 0
  float ex, ey;    // epipole of image 2 in normalised coordinates
  float s1, t1x, t1y;  // T1 = [s1 0 t1x; 0 s1 t1y; 0 0 1]
  float s2, t2x, t2y;
};
struct GeomD {
  double F[9];
  double ex, ey;
  double s1, t1x, t1y;
  double s2, t2x, t2y;
};

// Data-cost constants of dataEnergy (MultiH.cpp:473-504) in normalised units.
struct CostParams {
  float T;         // truncated_sqr_threshold * s2^2      (d2' < T  <=> d2 < thr^2*81/16)
  float inv_T;     // 1 / T
  float lam;       // one_per_energy_lambda = 100 / lambda
  float thr2;      // sqr_threshold_homography * s2^2
  int32_t cost_outlier;   // round(lam * T_px)            (label 0)
  int32_t cost_far;       // 2 * round(lam * T_px)        (d2 >= T)
};

}  // namespace mh

struct mh_ctx {
  mh_params params;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  bool have_geom = false;
  mh::GeomD gd;      // normalised, FP64
  mh::GeomF gf;      // normalised, FP32
  double F_px[9];    // as given
  double e2_px[2];
  std::string err;
  int64_t launches = 0;
  uint32_t rng_state = 1u;  // state of the restated MSVC rand() (MS.h:54); mh_process re-seeds it from params.rng_seed
  // grow-only scratch arena (device) + pinned staging (host)
  void* scratch = nullptr;
  uint64_t scratch_bytes = 0;
  void* staging = nullptr;
  uint64_t staging_bytes = 0;
  void* pinned = nullptr;
  uint64_t pinned_bytes = 0;
  // mh_process's working set: grow-only device buffers that live as long as the context (no cudaMalloc in steady state)
  void* pbuf[24] = {};
  uint64_t pcap[24] = {};
  // results of the last mh_process
  double energy = 0.0;
  int32_t iterations = 0;
  double stage_ms[5] = {0, 0, 0, 0, 0};
  double alt_ms[5] = {0, 0, 0, 0, 0};   // inside the alternating optimisation: mean-shift, mode fit + inlier scan, data cost, graph cut, refit
  // multi-GPU (comm.cu): NCCL communicator over the ranks that shard the correspondences, its own stream, and the double-buffered
  // refit statistics whose all-reduce overlaps the next sharded pass
  void* comm = nullptr;
  int32_t comm_rank = 0, comm_world = 1;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t comm_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  double* comm_acc = nullptr;
  int32_t comm_acc_k = 0, comm_cur = 0, comm_pending = -1;
  int32_t pend_K = 0;
  void* pend_inliers = nullptr;
  void* pend_ref = nullptr;
  cudaStream_t aux_stream = nullptr;   // mh_step_sharded: K1 runs here, behind K2
  cudaEvent_t aux_ev[2] = {nullptr, nullptr};
};

namespace mh {

mh_status fail(mh_ctx* ctx, mh_status st, const std::string& msg);
mh_status check_cuda(mh_ctx* ctx, cudaError_t e, const char* what);
mh_status ensure_scratch(mh_ctx* ctx, uint64_t bytes);   // ctx->scratch  >= bytes (device)
mh_status ensure_staging(mh_ctx* ctx, uint64_t bytes);   // ctx->staging  >= bytes (device)
mh_status ensure_pinned(mh_ctx* ctx, uint64_t bytes);    // ctx->pinned   >= bytes (host, page-locked)
CostParams cost_params(const mh_ctx* ctx);
void hyp_px_to_norm(const mh_ctx* ctx, const double* H_px, float* out12);
void hyp_norm_to_px(const mh_ctx* ctx, const float* in12, double* H_px, bool divide_h33);
void epipole2_host(const double* F, double* e2);
void sym_eigen3_host(const double* A, double* w, double* Vrows);

#define MH_CUDA(ctx, call)                                                  \
  do {                                                                      \
    mh_status _st = mh::check_cuda((ctx), (call), #call);                   \
    if (_st != MH_OK) return _st;                                           \
  } while (0)
#define MH_LAUNCHED(ctx, name)                                              \
  do {                                                                      \
    ++(ctx)->launches;                                                      \
    mh_status _st = mh::check_cuda((ctx), cudaGetLastError(), name);        \
    if (_st != MH_OK) return _st;                                           \
  } while (0)
// Opt-in for more than 48 KB of dynamic shared memory.  The attribute belongs to the FUNCTION (per device), not to a launch:
// contexts on several host threads launch the same kernels with different sizes, so it is always set to the hardware maximum —
// an idempotent call that cannot race with another thread's launch.
constexpr int MH_MAX_SMEM_PER_BLOCK = 227 * 1024;
template <typename Kernel>
inline cudaError_t mh_allow_max_smem(Kernel kernel) {
  cudaFuncAttributes a;
  const cudaError_t e = cudaFuncGetAttributes(&a, kernel);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MH_MAX_SMEM_PER_BLOCK - (int)a.sharedSizeBytes);
}

#define MH_TRY(expr)                                                        \
  do {                                                                      \
    mh_status _st = (expr);                                                 \
    if (_st != MH_OK) return _st;                                           \
  } while (0)

// ---- kernel launchers (defined in the k*.cu files) -------------------------
mh_status launch_normalize_points(mh_ctx*, const double* d_pts_raw, const double* d_aff_raw, int64_t N, float4* d_pts,
                                  float4* d_aff);
// Trailing FP64 pointers select the PRECISE path used by mh_process: raw FP64 pixel correspondences in, FP64 pixel-space
// homographies out (next to the FP32 normalised ones), so that the alternating optimisation tracks the FP64 reference.
mh_status launch_haf(mh_ctx*, const float4* d_pts, const float4* d_aff, int64_t N, float* d_hyp, int precision,
                     const double* d_pts64 = nullptr, const double* d_aff64 = nullptr, double* d_hyp64 = nullptr);
mh_status launch_cost_dense(mh_ctx*, const float4* d_pts, int64_t N, const float* d_hyp, int K, void* d_cost,
                            int elem_bytes);
mh_status launch_residuals(mh_ctx*, const float4* d_pts, int64_t N, const float* d_hyp, int K, float* d_d2);
mh_status launch_cost_fused(mh_ctx*, const float4* d_pts, int64_t N, const float* d_hyp, int K, int kmax,
                            uint32_t* d_list, int32_t* d_list_count, unsigned long long* d_best,
                            int32_t* d_inlier_count);
mh_status launch_inlier_stats(mh_ctx*, const float4* d_pts, int64_t N, const float* d_hyp, int K,
                              double* d_scatter /*K x 6, normalised coords*/);
mh_status launch_inliers_of(mh_ctx*, const float4* d_pts, int64_t N, const float* d_hyp_one, int idx, int32_t* d_labels);
mh_status launch_features10(mh_ctx*, const float* d_hyp, const float4* d_pts, int64_t N, double* d_feat,
                            const double* d_hyp64 = nullptr, const double* d_pts64 = nullptr);
mh_status launch_features6(mh_ctx*, const float* d_hyp, int K, double* d_feat, const double* d_hyp64 = nullptr);
mh_status launch_meanshift(mh_ctx*, const double* d_feat, int N, int D, double bw, int metric, uint32_t* rng_state,
                           double* d_centres, int max_c, int32_t* d_assign, int* C_out, int64_t* stats);
mh_status launch_refit_haf(mh_ctx*, const float4* d_pts, const float4* d_aff, const int32_t* d_labels, int64_t N, int K,
                           float* d_hyp, int32_t* d_count, const double* d_pts64 = nullptr,
                           const double* d_aff64 = nullptr, double* d_hyp64 = nullptr);
mh_status launch_refit_haf_accumulate(mh_ctx*, const float4* d_pts, const float4* d_aff, const int32_t* d_labels,
                                      int64_t N, int K, double* d_acc, const double* d_pts64 = nullptr,
                                      const double* d_aff64 = nullptr);
mh_status launch_refit_haf_solve(mh_ctx*, const double* d_acc, int K, float* d_hyp, int32_t* d_count,
                                 double* d_hyp64 = nullptr);
mh_status launch_labels_from_best(mh_ctx*, const unsigned long long* d_best, int64_t N, int32_t* d_labels);
mh_status launch_pack_inlier_counts(mh_ctx*, int32_t* d_cnt, int K, double* d_acc, int unpack);
mh_status launch_refit_3pt(mh_ctx*, const float4* d_pts, const int32_t* d_assign, int64_t N, int C, float* d_hyp,
                           int32_t* d_keep, const double* d_pts64 = nullptr, double* d_hyp64 = nullptr);
mh_status launch_prefilter(mh_ctx*, const double* d_pts64, const double* d_aff64, const double F[9], int64_t N,
                           double* d_pts_out, double* d_aff_out, int32_t* d_keep, int64_t* M_out);
mh_status launch_modes_to_hyp(mh_ctx*, const double* d_modes, int C, float* d_hyp, double* d_hyp64 = nullptr);
// FP64 members of the K2 family for the precise path (small N x K only): dataEnergy, inlier scan, single-H inliers
mh_status launch_cost_dense64(mh_ctx*, const double* d_pts64, int64_t N, const double* d_hyp64, int K, int32_t* d_cost);
mh_status launch_cost_list64(mh_ctx*, const double* d_pts64, int64_t N, const double* d_hyp64, int K, int kmax, uint32_t* d_list,
                             int32_t* d_count);
mh_status launch_inlier_stats64(mh_ctx*, const double* d_pts64, int64_t N, const double* d_hyp64, int K,
                                double* d_scatter /*K x 6, pixel coords*/);
mh_status launch_inliers_of64(mh_ctx*, const double* d_pts64, int64_t N, const double* d_hyp64_one, int idx,
                              int32_t* d_labels);

}  // namespace mh
