// ============================================================================
// pipeline.cu — mh_process: the control flow of MultiH::Process
// (MultiH/MultiH/MultiH.cpp:42-98) from ComputeLocalHomographies on, i.e.
// EstablishStablePointSets (:604-694) and ClusterMergingAndLabeling (:224-312)
// with MergingStep (:352-471) and LabelingStep (:513-602), driving the K1-K4
// kernels and the host alpha-expansion.  F is an input (north star), so
// GetFundamentalMatrixAndRefineData (:770-848) is the caller's business, and the
// post-processing HomographyCompatibilityCheck / HandleDegenerateCase
// (:100-222, :719-741) is out of scope.  The reference's
// Levenberg-Marquardt polish after each linear fit is not reproduced (its
// callbacks read out of bounds): every fit returns the
// reference's own linear solution (do_numerical_refinement = false).
// ============================================================================
#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace mh {
mh_status alpha_expansion(const int32_t* cost, int N, int L, int potts, const int64_t* offsets, const int32_t* adj,
                          const int32_t* init, int max_cycles, int32_t* lab, int64_t* energy_out);
int64_t radius_neighbourhood(const double* pts, int N, double radius, int max_neighbours, int64_t* offsets, int32_t* adj);
}  // namespace mh

using namespace mh;

namespace {

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  mh_status alloc(mh_ctx* ctx, uint64_t bytes) {
    if (p) { cudaFree(p); p = nullptr; }
    return check_cuda(ctx, cudaMalloc(&p, std::max<uint64_t>(bytes, 256)), "cudaMalloc");
  }
  template <typename T> T* as() { return (T*)p; }
};

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

mh_status mh_neighbourhood(mh_ctx* ctx, const double* pts, int32_t N, double radius, int32_t max_neighbours,
                           int64_t* offsets, int32_t* adj, int64_t* total_out) {
  // host-only: ctx may be NULL
  if (N < 0 || (N && !pts) || !(radius >= 0)) return ctx ? fail(ctx, MH_EINVAL, "mh_neighbourhood: bad arguments") : MH_EINVAL;
  const int64_t t = radius_neighbourhood(pts, N, radius, max_neighbours, offsets, adj);
  if (total_out) *total_out = t;
  return MH_OK;
}

mh_status mh_alpha_expansion(mh_ctx* ctx, const int32_t* cost, int32_t N, int32_t L, int32_t potts, const int64_t* offsets,
                             const int32_t* adj, const int32_t* init, int32_t max_cycles, int32_t* labels,
                             int64_t* energy) {
  if (!cost || !labels || N <= 0 || L < 1) return ctx ? fail(ctx, MH_EINVAL, "mh_alpha_expansion: bad arguments") : MH_EINVAL;
  const mh_status st = alpha_expansion(cost, N, L, potts, offsets, adj, init, max_cycles, labels, energy);
  if (st != MH_OK && ctx) ctx->err = "mh_alpha_expansion: invalid labels or sizes";
  return st;
}

mh_status mh_process(mh_ctx* ctx, const double* pts, const double* aff, const double F[9], int32_t N, int32_t* labels_out,
                     double* H_out, int32_t Kmax, int32_t* K_out) {
  if (!ctx) return MH_EINVAL;
  if (!pts || !aff || !F || !labels_out || !K_out) return fail(ctx, MH_EINVAL, "mh_process: null argument");
  if (N < 8) return fail(ctx, MH_EINVAL, "Error: Features are not set! (MultiH.cpp:44-50: fewer than 8 correspondences)");
  *K_out = 0;
  const mh_params& P = ctx->params;
  const double t_start = now_ms();
  MH_TRY(mh_set_geometry(ctx, F, nullptr, nullptr, pts, N));

  DevBuf b_pts, b_aff, b_hyp_pt, b_feat, b_centres, b_assign, b_hyp, b_keep, b_cost, b_labels, b_modes_hyp, b_feat6;
  MH_TRY(b_pts.alloc(ctx, sizeof(float4) * (uint64_t)N));
  MH_TRY(b_aff.alloc(ctx, sizeof(float4) * (uint64_t)N));
  MH_TRY(mh_upload_correspondences(ctx, pts, aff, N, b_pts.p, b_aff.p));

  // ---- ComputeLocalHomographies (MultiH.cpp:65, 696-717) ------------------------------
  double t0 = now_ms();
  MH_TRY(b_hyp_pt.alloc(ctx, sizeof(float) * 12 * (uint64_t)N));
  MH_TRY(mh_haf_hypotheses(ctx, b_pts.p, b_aff.p, N, b_hyp_pt.p, 0));
  MH_TRY(mh_sync(ctx));
  ctx->stage_ms[0] = now_ms() - t0;

  // ---- EstablishStablePointSets (MultiH.cpp:71, 604-694) ------------------------------
  t0 = now_ms();
  MH_TRY(b_feat.alloc(ctx, sizeof(double) * 10 * (uint64_t)N));
  MH_TRY(mh_features10(ctx, b_hyp_pt.p, b_pts.p, N, b_feat.p));
  MH_TRY(b_centres.alloc(ctx, sizeof(double) * 10 * (uint64_t)N));
  MH_TRY(b_assign.alloc(ctx, sizeof(int32_t) * (uint64_t)N));
  int32_t C = 0;
  MH_TRY(mh_meanshift(ctx, b_feat.p, N, 10, P.thr_homography, b_centres.p, N, b_assign.p, &C, nullptr));
  std::vector<float> hyp;  // current cluster homographies, K x 12 (normalised space)
  int K = 0;
  if (C > 0) {
    MH_TRY(b_hyp.alloc(ctx, sizeof(float) * 12 * (uint64_t)C));
    MH_TRY(b_keep.alloc(ctx, sizeof(int32_t) * (uint64_t)C));
    MH_TRY(mh_refit_3pt(ctx, b_pts.p, b_assign.p, N, C, b_hyp.p, b_keep.p));
    std::vector<float> all(12 * (size_t)C);
    std::vector<int32_t> keep(C);
    MH_TRY(mh_memcpy_d2h(ctx, all.data(), b_hyp.p, sizeof(float) * 12 * (size_t)C));
    MH_TRY(mh_memcpy_d2h(ctx, keep.data(), b_keep.p, sizeof(int32_t) * (size_t)C));
    for (int c = 0; c < C; ++c)
      if (keep[c]) { hyp.insert(hyp.end(), all.begin() + 12 * (size_t)c, all.begin() + 12 * (size_t)(c + 1)); ++K; }
  }
  ctx->stage_ms[1] = now_ms() - t0;

  // ---- neighbourhood (MultiH.cpp:231-258) ----------------------------------------------
  t0 = now_ms();
  std::vector<int64_t> offsets((size_t)N + 1);
  int64_t total = 0;
  MH_TRY(mh_neighbourhood(ctx, pts, N, 1.0 / P.locality, P.max_neighbours, offsets.data(), nullptr, &total));
  std::vector<int32_t> adj((size_t)std::max<int64_t>(total, 1));
  MH_TRY(mh_neighbourhood(ctx, pts, N, 1.0 / P.locality, P.max_neighbours, offsets.data(), adj.data(), &total));
  ctx->stage_ms[2] = now_ms() - t0;

  // ---- alternating optimisation (MultiH.cpp:260-311) ------------------------------------
  t0 = now_ms();
  std::vector<int32_t> labeling(N, -1), init(N), gc_labels(N);
  std::vector<int32_t> cost;
  MH_TRY(b_labels.alloc(ctx, sizeof(int32_t) * (uint64_t)N));
  double lastEnergy = (double)INT32_MAX;
  int not_changed_number = 0, iteration_number = 0;
  ctx->energy = 0.0;
  const int potts = (int)std::round(100.0 * P.lambda);  // smoothnessEnergy, MultiH.cpp:506-511
  while (iteration_number++ < P.max_iterations) {
    bool changed = false;
    // -- MergingStep (MultiH.cpp:352-471)
    if (K > 0) {
      MH_TRY(b_hyp.alloc(ctx, sizeof(float) * 12 * (uint64_t)K));
      MH_TRY(mh_memcpy_h2d(ctx, b_hyp.p, hyp.data(), sizeof(float) * 12 * (size_t)K));
      MH_TRY(b_feat6.alloc(ctx, sizeof(double) * 6 * (uint64_t)K));
      MH_TRY(mh_features6(ctx, b_hyp.p, K, b_feat6.p));
      MH_TRY(b_centres.alloc(ctx, sizeof(double) * 6 * (uint64_t)K));
      MH_TRY(b_assign.alloc(ctx, sizeof(int32_t) * (uint64_t)K));
      int32_t Cm = 0;
      MH_TRY(mh_meanshift(ctx, b_feat6.p, K, 6, P.thr_homography, b_centres.p, K, b_assign.p, &Cm, nullptr));
      std::vector<float> merged;
      int Kn = 0;
      if (Cm > 0) {
        MH_TRY(b_modes_hyp.alloc(ctx, sizeof(float) * 12 * (uint64_t)Cm));
        MH_TRY(mh_modes_to_hypotheses(ctx, b_centres.p, Cm, b_modes_hyp.p));
        std::vector<int32_t> keep(Cm);
        MH_TRY(mh_inlier_stats(ctx, b_pts.p, N, b_modes_hyp.p, Cm, nullptr, nullptr, keep.data()));
        std::vector<float> all(12 * (size_t)Cm);
        MH_TRY(mh_memcpy_d2h(ctx, all.data(), b_modes_hyp.p, sizeof(float) * 12 * (size_t)Cm));
        for (int c = 0; c < Cm; ++c)
          if (keep[c]) { merged.insert(merged.end(), all.begin() + 12 * (size_t)c, all.begin() + 12 * (size_t)(c + 1)); ++Kn; }
      }
      changed = Kn != K;  // MultiH.cpp:468
      if (changed) { hyp.swap(merged); K = Kn; }
    }
    if (changed) not_changed_number = 0; else ++not_changed_number;

    if (K == 1) {  // MultiH.cpp:280-285
      MH_TRY(b_hyp.alloc(ctx, sizeof(float) * 12));
      MH_TRY(mh_memcpy_h2d(ctx, b_hyp.p, hyp.data(), sizeof(float) * 12));
      MH_TRY(mh_memcpy_h2d(ctx, b_labels.p, labeling.data(), sizeof(int32_t) * (size_t)N));
      MH_TRY(mh_inliers_of_homography(ctx, b_pts.p, N, b_hyp.p, 0, b_labels.p));
      MH_TRY(mh_memcpy_d2h(ctx, labeling.data(), b_labels.p, sizeof(int32_t) * (size_t)N));
      break;
    } else if (K == 0)
      break;

    // -- LabelingStep (MultiH.cpp:513-602)
    const int L = K + 1;
    MH_TRY(b_hyp.alloc(ctx, sizeof(float) * 12 * (uint64_t)K));
    MH_TRY(mh_memcpy_h2d(ctx, b_hyp.p, hyp.data(), sizeof(float) * 12 * (size_t)K));
    MH_TRY(b_cost.alloc(ctx, sizeof(int32_t) * (uint64_t)N * L));
    MH_TRY(mh_data_cost_dense(ctx, b_pts.p, N, b_hyp.p, K, b_cost.p, 4));
    cost.resize((size_t)N * L);
    MH_TRY(mh_memcpy_d2h(ctx, cost.data(), b_cost.p, sizeof(int32_t) * (size_t)N * L));
    const int32_t* init_ptr = nullptr;
    if (!changed) {  // warm start (MultiH.cpp:525-529)
      for (int i = 0; i < N; ++i) init[i] = std::min(std::max(labeling[i] + 1, 0), L - 1);
      init_ptr = init.data();
    }
    int64_t e64 = 0;
    MH_TRY(mh_alpha_expansion(ctx, cost.data(), N, L, potts, offsets.data(), adj.data(), init_ptr, P.max_gc_cycles,
                              gc_labels.data(), &e64));
    const double energy = (double)e64;
    for (int i = 0; i < N; ++i) labeling[i] = gc_labels[i] - 1;  // MultiH.cpp:547-568
    MH_TRY(mh_memcpy_h2d(ctx, b_labels.p, labeling.data(), sizeof(int32_t) * (size_t)N));
    MH_TRY(mh_refit_haf(ctx, b_pts.p, b_aff.p, b_labels.p, N, K, b_hyp.p, nullptr));  // MultiH.cpp:587-599
    MH_TRY(mh_memcpy_d2h(ctx, hyp.data(), b_hyp.p, sizeof(float) * 12 * (size_t)K));

    if ((!changed && std::fabs(lastEnergy - energy) < P.convergence) || not_changed_number > 10) {  // MultiH.cpp:295
      ctx->energy = energy;
      break;
    }
    lastEnergy = energy;
  }
  ctx->iterations = iteration_number - 1;  // MultiH.cpp:311
  ctx->stage_ms[3] = now_ms() - t0;

  std::memcpy(labels_out, labeling.data(), sizeof(int32_t) * (size_t)N);
  *K_out = K;
  if (H_out) {
    for (int k = 0; k < std::min(K, (int)Kmax); ++k) mh::hyp_norm_to_px(ctx, hyp.data() + 12 * (size_t)k, H_out + 9 * (size_t)k, false);
  }
  ctx->stage_ms[4] = now_ms() - t_start;
  return MH_OK;
}

}  // extern "C"
