// ============================================================================
// pipeline.cu — mh_process: the control flow of MultiH::Process
// (MultiH/MultiH/MultiH.cpp:42-98) from ComputeLocalHomographies on, i.e.
// EstablishStablePointSets (:604-694) and ClusterMergingAndLabeling (:224-312)
// with MergingStep (:352-471) and LabelingStep (:513-602), driving the K1-K4
// kernels and the host alpha-expansion.  F is an input (north star), so
// GetFundamentalMatrixAndRefineData (:770-848) is the caller's business, and the
// HandleDegenerateCase (:719-741, OpenCV's RANSAC findHomography) stays with the caller;
// HomographyCompatibilityCheck (:100-222) runs at the end when params.compatibility_check is set.  The reference's
// Levenberg-Marquardt polish after each linear fit is not reproduced (its
// callbacks read out of bounds): every fit returns the
// reference's own linear solution (do_numerical_refinement = false).
// ============================================================================
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace mh {
mh_status alpha_expansion(const int32_t* cost, int N, int L, int potts, const int64_t* offsets, const int32_t* adj,
                          const int32_t* init, int max_cycles, int32_t* lab, int64_t* energy_out);
int64_t radius_neighbourhood(const double* pts, int N, double radius, int max_neighbours, int64_t* offsets, int32_t* adj);
mh_status neighbourhood_device(mh_ctx* ctx, const double* pts, int N, double radius, int max_neighbours, int64_t* offsets,
                               int32_t* adj, int64_t* total_out);   // k5_neighbourhood.cu
mh_status launch_compat_trials(mh_ctx* ctx, const double* d_pts64, const int32_t* d_members, const int32_t* d_moff,
                               const int32_t* d_samples, int T, int trials, int max_n, double* d_out);   // k4_refit.cu
int g_nb_backend = 0;   // 0 = auto (device when there is a context and the list length is bounded), 1 = host, 2 = device
}  // namespace mh

using namespace mh;

namespace {

struct DevBuf {  // grow-only device buffer owned by the context (slot of mh_ctx::pbuf): mh_process allocates on the first pair
                 // of a size class and re-uses the buffers afterwards — no cudaMalloc/cudaFree in steady state
  void*& p;
  uint64_t& cap;
  DevBuf(mh_ctx* ctx, int slot) : p(ctx->pbuf[slot]), cap(ctx->pcap[slot]) {}
  mh_status alloc(mh_ctx* ctx, uint64_t bytes) {
    bytes = std::max<uint64_t>(bytes, 256);
    if (p && cap >= bytes) return MH_OK;
    if (p) { cudaStreamSynchronize(ctx->stream); cudaFree(p); p = nullptr; cap = 0; }
    const uint64_t want = bytes + bytes / 4;
    mh_status st = check_cuda(ctx, cudaMalloc(&p, want), "cudaMalloc");
    if (st == MH_OK) cap = want;
    return st;
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
  // ctx may be NULL: host search.  With a context and a bounded list length (the reference's effective 31) the search runs on
  // the GPU (K5, k5_neighbourhood.cu) — same set, bit for bit.
  if (N < 0 || (N && !pts) || !(radius >= 0)) return ctx ? fail(ctx, MH_EINVAL, "mh_neighbourhood: bad arguments") : MH_EINVAL;
  const bool can_device = ctx && max_neighbours >= 1 && max_neighbours <= 64;
  if (g_nb_backend == 2 && !can_device) return ctx ? fail(ctx, MH_EINVAL, "mh_neighbourhood: device backend needs a context and 1 <= max_neighbours <= 64") : MH_EINVAL;
  if (can_device && (g_nb_backend == 2 || (g_nb_backend == 0 && N >= 256)))
    return neighbourhood_device(ctx, pts, N, radius, max_neighbours, offsets, adj, total_out);
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

// Sparse-with-default form of the same call (cf. GCO's setDataCost(SparseDataCost), GCoptimization.h:220-224): per site up to
// kmax entries (label << 16 | cost), label 1..L-1, as cost_list64_kernel / the pipeline's labelling step produce them; label 0
// (the outlier label) costs cost_label0 everywhere and every label not listed costs cost_default.  A site with more than kmax
// entries (count > kmax) is an error: the lists would be truncated.  The solver indexes a dense matrix, which is built here (in the
// caller-provided or a thread-local buffer), so the result is exactly mh_alpha_expansion's on the expanded costs.
mh_status mh_alpha_expansion_sparse(mh_ctx* ctx, const uint32_t* lists, const int32_t* counts, int32_t kmax, int32_t N, int32_t L,
                                    int32_t cost_label0, int32_t cost_default, int32_t potts, const int64_t* offsets,
                                    const int32_t* adj, const int32_t* init, int32_t max_cycles, int32_t* labels, int64_t* energy) {
  if (!lists || !counts || !labels || N <= 0 || L < 1 || kmax < 1)
    return ctx ? fail(ctx, MH_EINVAL, "mh_alpha_expansion_sparse: bad arguments") : MH_EINVAL;
  static thread_local std::vector<int32_t> dense;
  dense.resize((size_t)N * L);
  for (int i = 0; i < N; ++i) {
    if (counts[i] < 0 || counts[i] > kmax)
      return ctx ? fail(ctx, MH_EINVAL, "mh_alpha_expansion_sparse: a site has more entries than kmax") : MH_EINVAL;
    int32_t* row = dense.data() + (size_t)i * L;
    row[0] = cost_label0;
    std::fill(row + 1, row + L, cost_default);
    const uint32_t* e = lists + (size_t)i * kmax;
    for (int k = 0; k < counts[i]; ++k) {
      const uint32_t l = e[k] >> 16;
      if (l == 0 || l >= (uint32_t)L) return ctx ? fail(ctx, MH_EINVAL, "mh_alpha_expansion_sparse: label out of range") : MH_EINVAL;
      row[l] = (int32_t)(e[k] & 0xffffu);
    }
  }
  return mh_alpha_expansion(ctx, dense.data(), N, L, potts, offsets, adj, init, max_cycles, labels, energy);
}

// HomographyCompatibilityCheck (MultiH.cpp:100-222) in three parts.
//   mh_compat_plan   (host)  the sampling: rand() draws without replacement from a point vector whose order evolves from trial
//                            to trial (:142-154, :183-194) — sequential index bookkeeping, replayed exactly;
//   compat_trial_kernel (GPU) the 501 x clusters three-point fits, transfer errors and their order statistics;
//   mh_compat_decide (host)  the reference's "median" per trial INCLUDING the three stale entries of its distance buffer
//                            (see oracle/multih_oracle.cpp orc_compatibility_check), median over the trials, the decision and
//                            the relabelling (:200-230).
// The two host halves take plain arrays (no context) so that the CPU tests can drive them with oracle-computed statistics.
mh_status mh_compat_plan(const int32_t* labels, int32_t N, int32_t K, int32_t min_inliers, uint32_t* rng_state,
                         int32_t* tested /*[K]*/, int32_t* T_out, int32_t* members /*[N]*/, int32_t* moff /*[K+1]*/,
                         int32_t* samples /*[K][501][3]*/, int32_t* removed /*[K]*/) {
  if (!labels || !rng_state || !tested || !T_out || !members || !moff || !samples || !removed || N < 0 || K < 0) return MH_EINVAL;
  const int trials = MH_COMPAT_TRIALS;
  uint32_t hold = *rng_state;
  auto msvc_rand = [&]() { hold = hold * 214013u + 2531011u; return (int)((hold >> 16) & 0x7fff); };
  std::vector<std::vector<int32_t>> per(K);                     // members in index order (MultiH.cpp:106-114)
  for (int i = 0; i < N; ++i)
    if (labels[i] > -1 && labels[i] < K) per[labels[i]].push_back(i);
  int T = 0;
  size_t nm = 0, ns = 0;
  moff[0] = 0;
  for (int c = 0; c < K; ++c) {
    const int n = (int)per[c].size();
    removed[c] = 0;
    if (n >= std::max(min_inliers, 4)) {
      tested[T] = c;
      std::memcpy(members + nm, per[c].data(), sizeof(int32_t) * (size_t)n);
      nm += (size_t)n;
      moff[++T] = (int32_t)nm;
      std::vector<int32_t> v = per[c];                          // the evolving point vector, as indices
      for (int t = 0; t < trials; ++t) {
        int32_t pick[3];
        for (int j = 0; j < 3; ++j) {
          const int idx = (int)((double)(v.size() - 1) * ((double)msvc_rand() / 32767.0));   // MultiH.cpp:145
          pick[j] = v[idx];
          v.erase(v.begin() + idx);
        }
        samples[ns++] = pick[0]; samples[ns++] = pick[1]; samples[ns++] = pick[2];
        v.resize(n);
        for (int j = 0; j < 3; ++j) v[n - j - 1] = pick[j];     // MultiH.cpp:183-194
      }
    } else if (n < min_inliers) {
      removed[c] = 1;                                           // MultiH.cpp:207-208
    }
  }
  *T_out = T;
  *rng_state = hold;
  return MH_OK;
}

mh_status mh_compat_decide(const int32_t* tested, int32_t T, const int32_t* moff, const double* stats /*[T][501][8]*/,
                           double thr_homography, int32_t* removed /*[K] in/out*/, int32_t N, int32_t* labels, double* H,
                           int32_t* K_inout, double* medians_out /*[K] or NULL*/) {
  if (!K_inout || (T > 0 && (!tested || !moff || !stats)) || !removed || (N > 0 && !labels) || !H) return MH_EINVAL;
  const int trials = MH_COMPAT_TRIALS, K = *K_inout;
  const double limit = thr_homography * thr_homography * 81.0 / 16.0;
  if (medians_out)
    for (int c = 0; c < K; ++c) medians_out[c] = std::nan("");
  std::vector<double> distances(trials);
  for (int k = 0; k < T; ++k) {
    const int c = tested[k], n = moff[k + 1] - moff[k] - 3, m = n / 2, lo = std::max(0, m - 3);
    double stale[3] = {0.0, 0.0, 0.0};                          // the buffer starts zeroed (MultiH.cpp:140)
    // q-th smallest of (the trial's n errors) U (3 stale entries), from the window [lo, min(n-1, m+1)] of the sorted errors:
    // the lo entries below the window all rank below q, so it is the (q - lo)-th smallest of window U stale
    auto kth = [&](const double* w, int q) {
      double u[8];
      int cnt = 0;
      for (int i = 0; i < 5; ++i)
        if (lo + i <= std::min(n - 1, q)) u[cnt++] = w[i];
      for (int i = 0; i < 3; ++i) u[cnt++] = stale[i];
      std::sort(u, u + cnt);
      return u[q - lo];
    };
    for (int t = 0; t < trials; ++t) {
      const double* w = stats + 8 * ((size_t)k * trials + t);
      distances[t] = n % 2 ? kth(w, m) : 0.5 * (kth(w, m) + kth(w, m + 1));   // MultiH.cpp:178
      double u[6] = {w[5], w[6], w[7], stale[0], stale[1], stale[2]};         // next trial's stale entries: the buffer's top 3
      std::sort(u, u + 6);
      stale[0] = u[3]; stale[1] = u[4]; stale[2] = u[5];
    }
    std::sort(distances.begin(), distances.end());
    const double median = distances[trials / 2];                // trials is odd (MultiH.cpp:200)
    if (medians_out) medians_out[c] = median;
    removed[c] = median > limit;                                // MultiH.cpp:202
  }
  int Kn = K;
  for (int c = K - 1; c >= 0; --c)                              // MultiH.cpp:216-230
    if (removed[c]) {
      for (int j = 0; j < N; ++j) {
        if (labels[j] == c) labels[j] = -1;
        else if (labels[j] > c) --labels[j];
      }
      for (int k = c; k + 1 < Kn; ++k) std::memcpy(H + 9 * (size_t)k, H + 9 * (size_t)(k + 1), sizeof(double) * 9);
      --Kn;
    }
  *K_inout = Kn;
  return MH_OK;
}

mh_status mh_compatibility_check(mh_ctx* ctx, const double* pts, int32_t N, int32_t* labels, double* H, int32_t* K_inout,
                                 double* medians_out) {
  if (!ctx) return MH_EINVAL;
  if (!pts || !labels || !H || !K_inout || N < 0) return fail(ctx, MH_EINVAL, "mh_compatibility_check: null argument");
  if (!ctx->have_geom) return fail(ctx, MH_EINVAL, "mh_compatibility_check: call mh_set_geometry first");
  const int K = *K_inout;
  if (K <= 0) return MH_OK;
  const int trials = MH_COMPAT_TRIALS;
  std::vector<int32_t> tested(K), members((size_t)std::max(N, 1)), moff((size_t)K + 1), samples((size_t)K * trials * 3), removed(K);
  int32_t T = 0;
  if (mh_compat_plan(labels, N, K, ctx->params.min_inliers, &ctx->rng_state, tested.data(), &T, members.data(), moff.data(),
                     samples.data(), removed.data()) != MH_OK)
    return fail(ctx, MH_EINVAL, "mh_compatibility_check: bad labels");
  std::vector<double> stat(8 * (size_t)T * trials);
  if (T > 0) {
    int max_n = 0;
    for (int k = 0; k < T; ++k) max_n = std::max(max_n, moff[k + 1] - moff[k] - 3);
    uint64_t off = 0;
    auto take = [&](uint64_t bytes) { uint64_t o = off; off = (off + bytes + 255) & ~uint64_t(255); return o; };
    const uint64_t o_pts = take(sizeof(double) * 4 * (uint64_t)N), o_mem = take(sizeof(int32_t) * (uint64_t)moff[T]),
                   o_off = take(sizeof(int32_t) * ((uint64_t)T + 1)), o_smp = take(sizeof(int32_t) * 3 * (uint64_t)T * trials),
                   o_out = take(sizeof(double) * 8 * (uint64_t)T * trials);
    MH_TRY(ensure_staging(ctx, off));
    char* base = (char*)ctx->staging;
    MH_TRY(mh_memcpy_h2d(ctx, base + o_pts, pts, sizeof(double) * 4 * (size_t)N));
    MH_TRY(mh_memcpy_h2d(ctx, base + o_mem, members.data(), sizeof(int32_t) * (size_t)moff[T]));
    MH_TRY(mh_memcpy_h2d(ctx, base + o_off, moff.data(), sizeof(int32_t) * ((size_t)T + 1)));
    MH_TRY(mh_memcpy_h2d(ctx, base + o_smp, samples.data(), sizeof(int32_t) * 3 * (size_t)T * trials));
    MH_TRY(launch_compat_trials(ctx, (const double*)(base + o_pts), (const int32_t*)(base + o_mem), (const int32_t*)(base + o_off),
                                (const int32_t*)(base + o_smp), T, trials, max_n, (double*)(base + o_out)));
    MH_TRY(mh_memcpy_d2h(ctx, stat.data(), base + o_out, sizeof(double) * stat.size()));
  }
  if (mh_compat_decide(tested.data(), T, moff.data(), stat.data(), ctx->params.thr_homography, removed.data(), N, labels, H,
                       K_inout, medians_out) != MH_OK)
    return fail(ctx, MH_EINVAL, "mh_compatibility_check: bad arguments");
  return MH_OK;
}

mh_status mh_process(mh_ctx* ctx, const double* pts, const double* aff, const double F[9], int32_t N, int32_t* labels_out,
                     double* H_out, int32_t Kmax, int32_t* K_out) {
  if (!ctx) return MH_EINVAL;
  if (!pts || !aff || !F || !labels_out || !K_out) return fail(ctx, MH_EINVAL, "mh_process: null argument");
  if (N < 8) return fail(ctx, MH_EINVAL, "Error: Features are not set! (MultiH.cpp:44-50: fewer than 8 correspondences)");
  *K_out = 0;
  const mh_params& P = ctx->params;
  const bool trace = std::getenv("MH_TRACE") != nullptr;  // the reference's LOG_TO_CONSOLE (MultiH.h:18, MultiH.cpp:292-293)
  // precise = 1 (default): FP64 pixel-space data path on the device, tracking the reference's arithmetic; 0: FP32
  // normalised data path (the throughput kernels) — same control flow, results equal up to FP32 rounding.
  const bool precise = P.precise_pipeline != 0;
  const double t_start = now_ms();
  ctx->rng_state = P.rng_seed;  // every Process() call starts from the same generator state, like a fresh run of the reference
  MH_TRY(mh_set_geometry(ctx, F, nullptr, nullptr, pts, N));

  int slot = 0;
  DevBuf b_pts(ctx, slot++), b_aff(ctx, slot++), b_pts64(ctx, slot++), b_aff64(ctx, slot++), b_raw_p(ctx, slot++),
      b_raw_a(ctx, slot++), b_keepmask(ctx, slot++), b_hyp_pt(ctx, slot++), b_hyp_pt64(ctx, slot++), b_feat(ctx, slot++),
      b_centres(ctx, slot++), b_assign(ctx, slot++), b_hyp(ctx, slot++), b_hyp64(ctx, slot++), b_keep(ctx, slot++),
      b_cost(ctx, slot++), b_labels(ctx, slot++), b_modes_hyp(ctx, slot++), b_modes_hyp64(ctx, slot++), b_feat6(ctx, slot++),
      b_scatter(ctx, slot++);
  static_assert(sizeof(((mh_ctx*)nullptr)->pbuf) / sizeof(void*) >= 21, "mh_ctx::pbuf has a slot per mh_process buffer");
  const int32_t N_in = N;
  std::vector<int32_t> keepmask;      // prefilter survivors (only with params.prefilter)
  std::vector<double> kept_pts_host;  // their refined coordinates, for the host neighbourhood search
  if (P.prefilter) {
    // ---- per-correspondence refinement of GetFundamentalMatrixAndRefineData (MultiH.cpp:807-838), K0 ------------
    MH_TRY(b_raw_p.alloc(ctx, sizeof(double) * 4 * (uint64_t)N));
    MH_TRY(b_raw_a.alloc(ctx, sizeof(double) * 4 * (uint64_t)N));
    MH_TRY(b_pts64.alloc(ctx, sizeof(double) * 4 * (uint64_t)N));
    MH_TRY(b_aff64.alloc(ctx, sizeof(double) * 4 * (uint64_t)N));
    MH_TRY(b_keepmask.alloc(ctx, sizeof(int32_t) * (uint64_t)N));
    MH_CUDA(ctx, cudaMemcpyAsync(b_raw_p.p, pts, sizeof(double) * 4 * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
    MH_CUDA(ctx, cudaMemcpyAsync(b_raw_a.p, aff, sizeof(double) * 4 * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
    int64_t M = 0;
    MH_TRY(launch_prefilter(ctx, b_raw_p.as<double>(), b_raw_a.as<double>(), F, N, b_pts64.as<double>(), b_aff64.as<double>(),
                            b_keepmask.as<int32_t>(), &M));
    keepmask.resize(N);
    MH_TRY(mh_memcpy_d2h(ctx, keepmask.data(), b_keepmask.p, sizeof(int32_t) * (size_t)N));
    if (trace) std::fprintf(stderr, "[Multi-H] %lld points kept from the initial %d after filtering.\n", (long long)M, N);  // MultiH.cpp:840
    if (M < 8) {  // MultiH.cpp:842-847: degenerate, not enough points remained
      for (int i = 0; i < N_in; ++i) labels_out[i] = keepmask[i] ? -1 : -2;
      return fail(ctx, MH_EDEGENERATE, "Degenerate case, not enough points remained (MultiH.cpp:842-847)");
    }
    N = (int32_t)M;
    kept_pts_host.resize(4 * (size_t)N);
    MH_TRY(mh_memcpy_d2h(ctx, kept_pts_host.data(), b_pts64.p, sizeof(double) * 4 * (size_t)N));
    pts = kept_pts_host.data();
    MH_TRY(mh_set_geometry(ctx, F, nullptr, nullptr, pts, N));  // normalisation from the refined set
  } else {
    MH_TRY(b_pts64.alloc(ctx, sizeof(double) * 4 * (uint64_t)N));
    MH_TRY(b_aff64.alloc(ctx, sizeof(double) * 4 * (uint64_t)N));
    MH_CUDA(ctx, cudaMemcpyAsync(b_pts64.p, pts, sizeof(double) * 4 * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
    MH_CUDA(ctx, cudaMemcpyAsync(b_aff64.p, aff, sizeof(double) * 4 * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
  }
  MH_TRY(b_pts.alloc(ctx, sizeof(float4) * (uint64_t)N));
  MH_TRY(b_aff.alloc(ctx, sizeof(float4) * (uint64_t)N));
  MH_TRY(launch_normalize_points(ctx, b_pts64.as<double>(), b_aff64.as<double>(), N, b_pts.as<float4>(), b_aff.as<float4>()));
  const double* pts64 = precise ? b_pts64.as<double>() : nullptr;
  const double* aff64 = precise ? b_aff64.as<double>() : nullptr;

  // ---- ComputeLocalHomographies (MultiH.cpp:65, 696-717) ------------------------------
  double t0 = now_ms();
  MH_TRY(b_hyp_pt.alloc(ctx, sizeof(float) * 12 * (uint64_t)N));
  if (precise) MH_TRY(b_hyp_pt64.alloc(ctx, sizeof(double) * 9 * (uint64_t)N));
  MH_TRY(launch_haf(ctx, b_pts.as<float4>(), b_aff.as<float4>(), N, b_hyp_pt.as<float>(), 0, pts64, aff64,
                    precise ? b_hyp_pt64.as<double>() : nullptr));
  MH_TRY(mh_sync(ctx));
  ctx->stage_ms[0] = now_ms() - t0;

  // ---- EstablishStablePointSets (MultiH.cpp:71, 604-694) ------------------------------
  t0 = now_ms();
  MH_TRY(b_feat.alloc(ctx, sizeof(double) * 10 * (uint64_t)N));
  MH_TRY(launch_features10(ctx, b_hyp_pt.as<float>(), b_pts.as<float4>(), N, b_feat.as<double>(),
                           precise ? b_hyp_pt64.as<double>() : nullptr, pts64));
  MH_TRY(b_centres.alloc(ctx, sizeof(double) * 10 * (uint64_t)N));
  MH_TRY(b_assign.alloc(ctx, sizeof(int32_t) * (uint64_t)N));
  int32_t C = 0;
  MH_TRY(mh_meanshift(ctx, b_feat.p, N, 10, P.thr_homography, b_centres.p, N, b_assign.p, &C, nullptr));
  // current cluster homographies: K x 12 FP32 normalised (hyp) and, on the precise path, K x 9 FP64 pixels (hyp64)
  std::vector<float> hyp;
  std::vector<double> hyp64;
  int K = 0;
  if (C > 0) {
    MH_TRY(b_hyp.alloc(ctx, sizeof(float) * 12 * (uint64_t)C));
    MH_TRY(b_hyp64.alloc(ctx, sizeof(double) * 9 * (uint64_t)C));
    MH_TRY(b_keep.alloc(ctx, sizeof(int32_t) * (uint64_t)C));
    MH_TRY(launch_refit_3pt(ctx, b_pts.as<float4>(), b_assign.as<int32_t>(), N, C, b_hyp.as<float>(), b_keep.as<int32_t>(),
                            pts64, precise ? b_hyp64.as<double>() : nullptr));
    std::vector<float> all(12 * (size_t)C);
    std::vector<double> all64(precise ? 9 * (size_t)C : 0);
    std::vector<int32_t> keep(C);
    MH_TRY(mh_memcpy_d2h(ctx, all.data(), b_hyp.p, sizeof(float) * 12 * (size_t)C));
    if (precise) MH_TRY(mh_memcpy_d2h(ctx, all64.data(), b_hyp64.p, sizeof(double) * 9 * (size_t)C));
    MH_TRY(mh_memcpy_d2h(ctx, keep.data(), b_keep.p, sizeof(int32_t) * (size_t)C));
    for (int c = 0; c < C; ++c)
      if (keep[c]) {
        hyp.insert(hyp.end(), all.begin() + 12 * (size_t)c, all.begin() + 12 * (size_t)(c + 1));
        if (precise) hyp64.insert(hyp64.end(), all64.begin() + 9 * (size_t)c, all64.begin() + 9 * (size_t)(c + 1));
        ++K;
      }
  }
  ctx->stage_ms[1] = now_ms() - t0;
  if (trace) std::fprintf(stderr, "[mh_process] N = %d: %d mean-shift centres, %d stable clusters\n", N, C, K);

  // ---- neighbourhood (MultiH.cpp:231-258) ----------------------------------------------
  t0 = now_ms();
  std::vector<int64_t> offsets((size_t)N + 1);
  int64_t total = 0;
  std::vector<int32_t> adj;
  if (P.max_neighbours > 0) {   // list lengths are bounded: one pass into a buffer of the bound
    adj.resize((size_t)N * (size_t)P.max_neighbours + 1);
    MH_TRY(mh_neighbourhood(ctx, pts, N, 1.0 / P.locality, P.max_neighbours, offsets.data(), adj.data(), &total));
  } else {                      // full ball: count, then fill
    MH_TRY(mh_neighbourhood(ctx, pts, N, 1.0 / P.locality, P.max_neighbours, offsets.data(), nullptr, &total));
    adj.resize((size_t)std::max<int64_t>(total, 1));
    MH_TRY(mh_neighbourhood(ctx, pts, N, 1.0 / P.locality, P.max_neighbours, offsets.data(), adj.data(), &total));
  }
  ctx->stage_ms[2] = now_ms() - t0;

  // uploads the current hypothesis set (both representations)
  auto upload_hyps = [&](int k) -> mh_status {
    MH_TRY(b_hyp.alloc(ctx, sizeof(float) * 12 * (uint64_t)std::max(k, 1)));
    MH_TRY(mh_memcpy_h2d(ctx, b_hyp.p, hyp.data(), sizeof(float) * 12 * (size_t)k));
    if (precise) {
      MH_TRY(b_hyp64.alloc(ctx, sizeof(double) * 9 * (uint64_t)std::max(k, 1)));
      MH_TRY(mh_memcpy_h2d(ctx, b_hyp64.p, hyp64.data(), sizeof(double) * 9 * (size_t)k));
    }
    return MH_OK;
  };

  // ---- alternating optimisation (MultiH.cpp:260-311) ------------------------------------
  t0 = now_ms();
  std::vector<int32_t> labeling(N, -1), init(N), gc_labels(N);
  std::vector<int32_t> cost;
  std::vector<uint32_t> sparse;   // per-site (label << 16 | cost) lists + counts of a labelling step
  MH_TRY(b_labels.alloc(ctx, sizeof(int32_t) * (uint64_t)N));
  double lastEnergy = (double)INT32_MAX;
  int not_changed_number = 0, iteration_number = 0;
  ctx->energy = 0.0;
  const int potts = (int)std::round(100.0 * P.lambda);  // smoothnessEnergy, MultiH.cpp:506-511
  for (double& v : ctx->alt_ms) v = 0.0;
  double tm = 0.0;
  while (iteration_number++ < P.max_iterations) {
    bool changed = false;
    // -- MergingStep (MultiH.cpp:352-471)
    if (K > 0) {
      tm = now_ms();
      MH_TRY(upload_hyps(K));
      MH_TRY(b_feat6.alloc(ctx, sizeof(double) * 6 * (uint64_t)K));
      MH_TRY(launch_features6(ctx, b_hyp.as<float>(), K, b_feat6.as<double>(), precise ? b_hyp64.as<double>() : nullptr));
      MH_TRY(b_centres.alloc(ctx, sizeof(double) * 6 * (uint64_t)K));
      MH_TRY(b_assign.alloc(ctx, sizeof(int32_t) * (uint64_t)K));
      int32_t Cm = 0;
      MH_TRY(mh_meanshift(ctx, b_feat6.p, K, 6, P.thr_homography, b_centres.p, K, b_assign.p, &Cm, nullptr));
      ctx->alt_ms[0] += now_ms() - tm;
      tm = now_ms();
      std::vector<float> merged;
      std::vector<double> merged64;
      int Kn = 0;
      if (Cm > 0) {
        MH_TRY(b_modes_hyp.alloc(ctx, sizeof(float) * 12 * (uint64_t)Cm));
        MH_TRY(b_modes_hyp64.alloc(ctx, sizeof(double) * 9 * (uint64_t)Cm));
        MH_TRY(launch_modes_to_hyp(ctx, b_centres.as<double>(), Cm, b_modes_hyp.as<float>(),
                                   precise ? b_modes_hyp64.as<double>() : nullptr));
        std::vector<int32_t> keep(Cm);
        if (precise) {  // inlier scan + straightness test (MultiH.cpp:430-463) in FP64 pixel coordinates
          MH_TRY(b_scatter.alloc(ctx, sizeof(double) * 6 * (uint64_t)Cm));
          MH_TRY(launch_inlier_stats64(ctx, pts64, N, b_modes_hyp64.as<double>(), Cm, b_scatter.as<double>()));
          std::vector<double> sc(6 * (size_t)Cm);
          MH_TRY(mh_memcpy_d2h(ctx, sc.data(), b_scatter.p, sizeof(double) * 6 * (size_t)Cm));
          for (int c = 0; c < Cm; ++c) {
            const double* q = sc.data() + 6 * (size_t)c;
            const double S[9] = {q[0], q[1], q[2], q[1], q[3], q[4], q[2], q[4], q[5]};
            double w[3], V[9];
            sym_eigen3_host(S, w, V);
            keep[c] = !(w[2] < P.straightness || q[5] < 3.0);
          }
        } else {
          MH_TRY(mh_inlier_stats(ctx, b_pts.p, N, b_modes_hyp.p, Cm, nullptr, nullptr, keep.data()));
        }
        std::vector<float> all(12 * (size_t)Cm);
        std::vector<double> all64(precise ? 9 * (size_t)Cm : 0);
        MH_TRY(mh_memcpy_d2h(ctx, all.data(), b_modes_hyp.p, sizeof(float) * 12 * (size_t)Cm));
        if (precise) MH_TRY(mh_memcpy_d2h(ctx, all64.data(), b_modes_hyp64.p, sizeof(double) * 9 * (size_t)Cm));
        for (int c = 0; c < Cm; ++c)
          if (keep[c]) {
            merged.insert(merged.end(), all.begin() + 12 * (size_t)c, all.begin() + 12 * (size_t)(c + 1));
            if (precise) merged64.insert(merged64.end(), all64.begin() + 9 * (size_t)c, all64.begin() + 9 * (size_t)(c + 1));
            ++Kn;
          }
      }
      changed = Kn != K;  // MultiH.cpp:468
      if (changed) { hyp.swap(merged); hyp64.swap(merged64); K = Kn; }
      ctx->alt_ms[1] += now_ms() - tm;
    }
    if (changed) not_changed_number = 0; else ++not_changed_number;

    if (K == 1) {  // MultiH.cpp:280-285: ComputeInliersOfHomography(0) (:743-768)
      // The reference leaves the labels of the previous labelling step in place here, but then (K <= 1, MultiH.cpp:88-94) clears
      // them all and falls back to cv::findHomography.  That fallback is the caller's; what is returned is the one homography
      // with its inliers: every other label is -1, so that labels < K_out.
      MH_TRY(upload_hyps(1));
      std::fill(labeling.begin(), labeling.end(), -1);
      MH_TRY(mh_memcpy_h2d(ctx, b_labels.p, labeling.data(), sizeof(int32_t) * (size_t)N));
      if (precise) MH_TRY(launch_inliers_of64(ctx, pts64, N, b_hyp64.as<double>(), 0, b_labels.as<int32_t>()));
      else MH_TRY(mh_inliers_of_homography(ctx, b_pts.p, N, b_hyp.p, 0, b_labels.p));
      MH_TRY(mh_memcpy_d2h(ctx, labeling.data(), b_labels.p, sizeof(int32_t) * (size_t)N));
      break;
    } else if (K == 0) {
      std::fill(labeling.begin(), labeling.end(), -1);
      break;
    }

    // -- LabelingStep (MultiH.cpp:513-602)
    const int L = K + 1;
    tm = now_ms();
    MH_TRY(upload_hyps(K));
    // The labelling step's data costs (dataEnergy, MultiH.cpp:473-504).  Precise path: sparse with default — the device lists
    // the few (label, cost) entries with d2 < T per site (cost_list64_kernel), every other entry is one of two constants, so
    // N x kmax words cross the bus instead of the N x (K + 1) matrix and go to the solver as they are
    // (mh_alpha_expansion_sparse).  A site with more than kmax entries in range (rare) makes this step fall back to the dense matrix.
    const int kmax = 32;
    const double lam_d = 100.0 / P.lambda, T_d = P.thr_homography * P.thr_homography * 81.0 / 16.0;
    const int c_out = (int)std::round(lam_d * T_d), c_far = 2 * c_out;
    bool sparse_ok = precise && K < 65535 && lam_d < 65535.0 && K > kmax;
    if (sparse_ok) {
      MH_TRY(b_cost.alloc(ctx, sizeof(uint32_t) * (uint64_t)N * kmax + sizeof(int32_t) * (uint64_t)N));
      uint32_t* d_list = b_cost.as<uint32_t>();
      int32_t* d_cnt = reinterpret_cast<int32_t*>(d_list + (size_t)N * kmax);
      MH_TRY(launch_cost_list64(ctx, pts64, N, b_hyp64.as<double>(), K, kmax, d_list, d_cnt));
      sparse.resize((size_t)N * kmax + (size_t)N);
      MH_TRY(mh_memcpy_d2h(ctx, sparse.data(), b_cost.p, sizeof(uint32_t) * sparse.size()));
      const uint32_t* cnt = sparse.data() + (size_t)N * kmax;
      for (int i = 0; i < N && sparse_ok; ++i) sparse_ok = cnt[i] <= (uint32_t)kmax;
    }
    if (!sparse_ok) {
      cost.resize((size_t)N * L);
      MH_TRY(b_cost.alloc(ctx, sizeof(int32_t) * (uint64_t)N * L));
      if (precise) MH_TRY(launch_cost_dense64(ctx, pts64, N, b_hyp64.as<double>(), K, b_cost.as<int32_t>()));
      else MH_TRY(mh_data_cost_dense(ctx, b_pts.p, N, b_hyp.p, K, b_cost.p, 4));
      MH_TRY(mh_memcpy_d2h(ctx, cost.data(), b_cost.p, sizeof(int32_t) * (size_t)N * L));
    }
    const int32_t* init_ptr = nullptr;
    if (!changed) {  // warm start (MultiH.cpp:525-529)
      for (int i = 0; i < N; ++i) init[i] = std::min(std::max(labeling[i] + 1, 0), L - 1);
      init_ptr = init.data();
    }
    int64_t e64 = 0;
    ctx->alt_ms[2] += now_ms() - tm;
    tm = now_ms();
    if (sparse_ok)   // the lists go to the solver as they came off the device: sparse with the two default costs
      MH_TRY(mh_alpha_expansion_sparse(ctx, sparse.data(), reinterpret_cast<const int32_t*>(sparse.data() + (size_t)N * kmax), kmax, N,
                                       L, c_out, c_far, potts, offsets.data(), adj.data(), init_ptr, P.max_gc_cycles,
                                       gc_labels.data(), &e64));
    else
      MH_TRY(mh_alpha_expansion(ctx, cost.data(), N, L, potts, offsets.data(), adj.data(), init_ptr, P.max_gc_cycles,
                                gc_labels.data(), &e64));
    ctx->alt_ms[3] += now_ms() - tm;
    tm = now_ms();
    const double energy = (double)e64;
    if (trace) std::fprintf(stderr, "[mh_process] iteration %d: K = %d changed = %d energy = %.0f\n", iteration_number, K, (int)changed, energy);
    for (int i = 0; i < N; ++i) labeling[i] = gc_labels[i] - 1;  // MultiH.cpp:547-568
    MH_TRY(mh_memcpy_h2d(ctx, b_labels.p, labeling.data(), sizeof(int32_t) * (size_t)N));
    MH_TRY(launch_refit_haf(ctx, b_pts.as<float4>(), b_aff.as<float4>(), b_labels.as<int32_t>(), N, K, b_hyp.as<float>(),
                            nullptr, pts64, aff64, precise ? b_hyp64.as<double>() : nullptr));  // MultiH.cpp:587-599
    MH_TRY(mh_memcpy_d2h(ctx, hyp.data(), b_hyp.p, sizeof(float) * 12 * (size_t)K));
    if (precise) MH_TRY(mh_memcpy_d2h(ctx, hyp64.data(), b_hyp64.p, sizeof(double) * 9 * (size_t)K));
    ctx->alt_ms[4] += now_ms() - tm;

    if ((!changed && std::fabs(lastEnergy - energy) < P.convergence) || not_changed_number > 10) {  // MultiH.cpp:295
      ctx->energy = energy;
      break;
    }
    lastEnergy = energy;
  }
  ctx->iterations = iteration_number - 1;  // MultiH.cpp:311
  ctx->stage_ms[3] = now_ms() - t0;

  // pixel-space homographies of the surviving clusters
  std::vector<double> Hpix(9 * (size_t)std::max(K, 1));
  for (int k = 0; k < K; ++k) {
    if (precise) std::memcpy(Hpix.data() + 9 * (size_t)k, hyp64.data() + 9 * (size_t)k, sizeof(double) * 9);
    else mh::hyp_norm_to_px(ctx, hyp.data() + 12 * (size_t)k, Hpix.data() + 9 * (size_t)k, false);
  }
  if (P.compatibility_check && K > 1) {  // MultiH.cpp:78-86 (HandleDegenerateCase, :88-94, stays with the caller)
    int32_t Kc = K;
    MH_TRY(mh_compatibility_check(ctx, pts, N, labeling.data(), Hpix.data(), &Kc, nullptr));
    K = Kc;
  }
  if (P.prefilter) {  // labels of the survivors in input order; -2 marks correspondences the pre-filter dropped
    int j = 0;
    for (int i = 0; i < N_in; ++i) labels_out[i] = keepmask[i] ? labeling[j++] : -2;
  } else {
    std::memcpy(labels_out, labeling.data(), sizeof(int32_t) * (size_t)N);
  }
  *K_out = K;
  if (H_out) std::memcpy(H_out, Hpix.data(), sizeof(double) * 9 * (size_t)std::min(K, (int)Kmax));
  ctx->stage_ms[4] = now_ms() - t_start;
  return MH_OK;
}

}  // extern "C"
