// ============================================================================
// K2 — N x K reprojection residual / data-cost kernels (sm_100a).
//
// Replaces the reference's lazily evaluated dataEnergy callback
// (MultiH/MultiH/MultiH.cpp:473-504, constants MultiH.h:41-44), the inlier
// scans of MergingStep (MultiH.cpp:430-458) and ComputeInliersOfHomography
// (MultiH.cpp:743-768).
//
// All kernels work in the context's normalised coordinates; squared residuals
// scale by s2^2, so the thresholds are pre-scaled (CostParams) and the cost
// round(lam * (1 - d2/T)) is scale-free.
//
// The throughput kernel (cost_fused) keeps 8 correspondences per thread in
// registers, streams hypothesis PAIRS from shared memory and evaluates two
// hypotheses per instruction with Blackwell's packed FFMA2 (fma.rn.f32x2):
// per residual 5 FFMA2 + 1 MUFU.RCP + 0.5 FMNMX3 issue slots.  Nothing of size
// N x K is written: hits (d2 < T) take a rare, divergent path that appends to
// per-site sparse lists / argmin / per-hypothesis inlier counters.
// ============================================================================
#include "common.cuh"
#include "k2_device.cuh"

namespace mh {

// ============================================================================
// Dense matrix: out[p*(K+1) + l], l = 0 outlier label (GCoptimization.h:339).
// Thread = one hypothesis (registers), loop over a shared-memory tile of
// correspondences; consecutive threads write consecutive labels => coalesced.
// HBM-bound on the store: 4 (or 2) B per residual.
// ============================================================================
constexpr int DENSE_THREADS = 256;
constexpr int DENSE_PT = 128;

template <typename OutT, bool RAW>
__global__ void __launch_bounds__(DENSE_THREADS) cost_dense_kernel(const float4* __restrict__ pts, long long N,
                                                                   const float* __restrict__ hyp, int K,
                                                                   OutT* __restrict__ out, CostParams cp,
                                                                   float inv_s2sq) {
  __shared__ float4 sp[DENSE_PT];
  const long long p0 = (long long)blockIdx.x * DENSE_PT;
  const int l = blockIdx.y * DENSE_THREADS + threadIdx.x;
  for (int i = threadIdx.x; i < DENSE_PT; i += DENSE_THREADS) {
    long long p = p0 + i;
    sp[i] = pts[p < N ? p : N - 1];
  }
  float h[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bool valid = l < K;
  if (valid) {
    const float4* hp = reinterpret_cast<const float4*>(hyp + (size_t)l * 12);
    const float4 a = hp[0], b = hp[1], c = hp[2];
    h[0] = a.x; h[1] = a.y; h[2] = a.z; h[3] = a.w; h[4] = b.x; h[5] = b.y; h[6] = b.z; h[7] = b.w; h[8] = c.x;
  }
  __syncthreads();
  const int np = (int)min((long long)DENSE_PT, N - p0);
  const long long stride = RAW ? K : (K + 1);
  if (!RAW && blockIdx.y == 0 && threadIdx.x < np)
    out[(p0 + threadIdx.x) * stride] = (OutT)cp.cost_outlier;
  if (!valid) return;
#pragma unroll 4
  for (int i = 0; i < np; ++i) {
    const float4 q = sp[i];
    const float d2 = residual(h, q.x, q.y, q.z, q.w);
    if (RAW)
      out[(p0 + i) * stride + l] = (OutT)(d2 * inv_s2sq);
    else
      out[(p0 + i) * stride + 1 + l] = (OutT)cost_of(d2, cp);
  }
}

mh_status launch_cost_dense_tiled(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, void* d_cost,
                                  int elem_bytes, int conv);   // k2_dense.cu
// 1 = cost_dense_tiled_kernel (TMA row stores; float->int through a denormal product, default), 2 / 3 = the same kernel with
// the 2^23-magic / F2I conversion, 0 = cost_dense_kernel (first generation, scalar stores)
int g_dense_variant = 1;

mh_status launch_cost_dense(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, void* d_cost,
                            int elem_bytes) {
  if (N <= 0 || K < 0) return MH_OK;
  if (elem_bytes != 2 && elem_bytes != 4) return fail(ctx, MH_EINVAL, "mh_data_cost_dense: elem_bytes must be 2 or 4");
  if (g_dense_variant >= 1) return launch_cost_dense_tiled(ctx, d_pts, N, d_hyp, K, d_cost, elem_bytes, g_dense_variant);
  const CostParams cp = cost_params(ctx);
  dim3 grid((unsigned)((N + DENSE_PT - 1) / DENSE_PT), (unsigned)std::max(1, (K + DENSE_THREADS - 1) / DENSE_THREADS));
  if (elem_bytes == 4)
    cost_dense_kernel<int32_t, false><<<grid, DENSE_THREADS, 0, ctx->stream>>>(d_pts, N, d_hyp, K, (int32_t*)d_cost, cp, 0.f);
  else if (elem_bytes == 2)
    cost_dense_kernel<int16_t, false><<<grid, DENSE_THREADS, 0, ctx->stream>>>(d_pts, N, d_hyp, K, (int16_t*)d_cost, cp, 0.f);
  else
    return fail(ctx, MH_EINVAL, "mh_data_cost_dense: elem_bytes must be 2 or 4");
  MH_LAUNCHED(ctx, "cost_dense_kernel");
  return MH_OK;
}

mh_status launch_residuals(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, float* d_d2) {
  if (N <= 0 || K <= 0) return MH_OK;
  const CostParams cp = cost_params(ctx);
  dim3 grid((unsigned)((N + DENSE_PT - 1) / DENSE_PT), (unsigned)((K + DENSE_THREADS - 1) / DENSE_THREADS));
  const float inv = (float)(1.0 / (ctx->gd.s2 * ctx->gd.s2));
  cost_dense_kernel<float, true><<<grid, DENSE_THREADS, 0, ctx->stream>>>(d_pts, N, d_hyp, K, d_d2, cp, inv);
  MH_LAUNCHED(ctx, "residual_kernel");
  return MH_OK;
}

// ============================================================================
// Fused throughput kernel.
// ============================================================================
constexpr int FUSED_THREADS = 256;
constexpr int FUSED_P = 8;                              // correspondences per thread
constexpr int FUSED_TILE = FUSED_THREADS * FUSED_P;     // 2048 correspondences per CTA
constexpr int FUSED_CHUNK_PAIRS = 256;                  // hypothesis pairs staged per chunk (512 hypotheses, 20 KB)

struct FusedOut {
  uint32_t* list;
  int32_t* list_count;
  u64* best;
  int32_t* inlier_count;
  int kmax;
};

__device__ __forceinline__ void fused_emit(const FusedOut& o, const CostParams& cp, long long pt, int label, float d2) {
  const int cost = cost_in_range(d2, cp);
  if (o.best) atomicMin(o.best + pt, ((u64)(uint32_t)cost << 32) | (uint32_t)label);
  if (o.list_count) {
    const int slot = atomicAdd(o.list_count + pt, 1);
    if (o.list && slot < o.kmax) o.list[pt * o.kmax + slot] = ((uint32_t)label << 8) | (uint32_t)min(cost, 255);
  }
  if (o.inlier_count && d2 < cp.thr2) atomicAdd(o.inlier_count + (label - 1), 1);
}

// Stage `npairs` hypothesis pairs starting at hypothesis index h0 into smem as
// 10 u64 per pair: {h_k(a), h_k(b)} for k = 0..8, + pad.  Out-of-range
// hypotheses become a "far" homography (residual ~1e36, never a hit).
__device__ __forceinline__ void stage_pairs(u64* sm, const float* __restrict__ hyp, int h0, int hend, int npairs) {
  for (int j = threadIdx.x; j < npairs; j += FUSED_THREADS) {
    float a[9], b[9];
    const int ia = h0 + 2 * j, ib = ia + 1;
    if (ia < hend) {
      const float4* p = reinterpret_cast<const float4*>(hyp + (size_t)ia * 12);
      const float4 u = p[0], v = p[1], w = p[2];
      a[0] = u.x; a[1] = u.y; a[2] = u.z; a[3] = u.w; a[4] = v.x; a[5] = v.y; a[6] = v.z; a[7] = v.w; a[8] = w.x;
    } else {
      a[0] = a[1] = a[3] = a[4] = a[6] = a[7] = 0.f; a[2] = a[5] = 1e18f; a[8] = 1.f;
    }
    if (ib < hend) {
      const float4* p = reinterpret_cast<const float4*>(hyp + (size_t)ib * 12);
      const float4 u = p[0], v = p[1], w = p[2];
      b[0] = u.x; b[1] = u.y; b[2] = u.z; b[3] = u.w; b[4] = v.x; b[5] = v.y; b[6] = v.z; b[7] = v.w; b[8] = w.x;
    } else {
      b[0] = b[1] = b[3] = b[4] = b[6] = b[7] = 0.f; b[2] = b[5] = 1e18f; b[8] = 1.f;
    }
    u64* d = sm + (size_t)j * 10;
#pragma unroll
    for (int k = 0; k < 9; ++k) d[k] = pk(a[k], b[k]);
    d[9] = 0ull;
  }
}

template <bool PACKED>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
cost_fused_kernel(const float4* __restrict__ pts, long long N, const float* __restrict__ hyp, int K, int k_per_block,
                  CostParams cp, FusedOut o) {
  __shared__ __align__(16) u64 sm[2][FUSED_CHUNK_PAIRS * 10];

  const long long tile0 = (long long)blockIdx.x * FUSED_TILE;
  const int kbeg = blockIdx.y * k_per_block;
  const int kend = min(K, kbeg + k_per_block);

  // correspondences -> registers (duplicated into both f32x2 lanes)
  u64 X[FUSED_P], Y[FUSED_P], NX2[FUSED_P], NY2[FUSED_P];
#pragma unroll
  for (int p = 0; p < FUSED_P; ++p) {
    long long idx = tile0 + (long long)p * FUSED_THREADS + threadIdx.x;
    const float4 q = pts[idx < N ? idx : N - 1];
    X[p] = pk(q.x, q.x); Y[p] = pk(q.y, q.y); NX2[p] = pk(-q.z, -q.z); NY2[p] = pk(-q.w, -q.w);
  }

  int buf = 0;
  stage_pairs(sm[0], hyp, kbeg, kend, min(FUSED_CHUNK_PAIRS, (kend - kbeg + 1) / 2));
  __syncthreads();

  for (int c0 = kbeg; c0 < kend; c0 += 2 * FUSED_CHUNK_PAIRS) {
    const int npairs = min(FUSED_CHUNK_PAIRS, (kend - c0 + 1) / 2);
    const int cn = c0 + 2 * FUSED_CHUNK_PAIRS;
    if (cn < kend) stage_pairs(sm[buf ^ 1], hyp, cn, kend, min(FUSED_CHUNK_PAIRS, (kend - cn + 1) / 2));

    const u64* s = sm[buf];
#pragma unroll 1
    for (int j = 0; j < npairs; ++j) {
      const ulonglong2* hp = reinterpret_cast<const ulonglong2*>(s + (size_t)j * 10);
      const ulonglong2 q0 = hp[0], q1 = hp[1], q2 = hp[2], q3 = hp[3], q4 = hp[4];
      const u64 H0 = q0.x, H1 = q0.y, H2 = q1.x, H3 = q1.y, H4 = q2.x, H5 = q2.y, H6 = q3.x, H7 = q3.y, H8 = q4.x;
      u64 D[FUSED_P];
      float m = 3.0e38f;
#pragma unroll
      for (int p = 0; p < FUSED_P; ++p) {
        if (PACKED) {
          const u64 sv = fma2(H6, X[p], fma2(H7, Y[p], H8));
          const u64 xn = fma2(H0, X[p], fma2(H1, Y[p], H2));
          const u64 yn = fma2(H3, X[p], fma2(H4, Y[p], H5));
          float sl, sh;
          upk(sv, sl, sh);
          const u64 r = pk(rcp_approx(sl), rcp_approx(sh));
          const u64 dx = fma2(xn, r, NX2[p]);
          const u64 dy = fma2(yn, r, NY2[p]);
          D[p] = fma2(dx, dx, mul2(dy, dy));
        } else {
          float ha[9], hb[9], x, y, nx2, ny2, t;
          upk(H0, ha[0], hb[0]); upk(H1, ha[1], hb[1]); upk(H2, ha[2], hb[2]);
          upk(H3, ha[3], hb[3]); upk(H4, ha[4], hb[4]); upk(H5, ha[5], hb[5]);
          upk(H6, ha[6], hb[6]); upk(H7, ha[7], hb[7]); upk(H8, ha[8], hb[8]);
          upk(X[p], x, t); upk(Y[p], y, t); upk(NX2[p], nx2, t); upk(NY2[p], ny2, t);
          D[p] = pk(residual(ha, x, y, -nx2, -ny2), residual(hb, x, y, -nx2, -ny2));
        }
        float da, db;
        upk(D[p], da, db);
        m = min3(m, da, db);
      }
      if (m < cp.T) {  // rare: some (correspondence, hypothesis) of this step is in range
        const int la = c0 + 2 * j + 1;  // 1-based label of lane a
#pragma unroll
        for (int p = 0; p < FUSED_P; ++p) {
          float da, db;
          upk(D[p], da, db);
          const long long idx = tile0 + (long long)p * FUSED_THREADS + threadIdx.x;
          if (idx < N) {
            if (da < cp.T) fused_emit(o, cp, idx, la, da);
            if (db < cp.T) fused_emit(o, cp, idx, la + 1, db);
          }
        }
      }
    }
    __syncthreads();
    buf ^= 1;
  }
}

// ----------------------------------------------------------------------------
// Fast path (no per-site lists): per-site data-term argmin + per-hypothesis inlier
// counts, built for HIGH hit rates (on the 200-plane scene ~7 % of all residuals
// are below the truncation threshold, so "emit on hit" is not a rare path).
//
//  * argmin: each thread keeps, per correspondence, the best (cost, label) so far
//    and the d2 interval [lo, T) a residual must fall into to IMPROVE on it (lower
//    integer cost; costs fall as d2 rises, MultiH.cpp:502).  The interval test is
//    folded into the last FMA (t = d2 - mid, |t| < half), so the common path pays one
//    FMNMX + one FSETP per hypothesis pair; the exact update (residual recomputed
//    with the dense kernel's instruction sequence => bit-identical costs and ties)
//    runs only on record-breaking candidates, ~ln(#hits) times per correspondence.
//  * inlier counts: one FSETP + one predicated shared-memory RED per residual into a
//    per-chunk counter array, flushed with one global RED per hypothesis and chunk.
// Issue budget per residual: 5 FFMA2 + 1 MUFU + ~0.3 LDS + 1 (argmin filter) + 2
// (inlier) ~ 9.5 slots against ~11 FMA-pipe cycles for the 5 FFMA2.
// ----------------------------------------------------------------------------
// v3 register tile: 4 correspondences per thread x 2 hypothesis pairs (4 hypotheses) per iteration.
constexpr int FAST_P = 4;  // <= 1024 correspondences per CTA => per-CTA inlier counts fit 16 bits

// stage_pairs for an arbitrary CTA size
template <int THREADS>
__device__ __forceinline__ void stage_pairs_t(u64* sm, const float* __restrict__ hyp, int h0, int hend, int npairs) {
  for (int j = threadIdx.x; j < npairs; j += THREADS) {
    float a[9], b[9];
    const int ia = h0 + 2 * j, ib = ia + 1;
    if (ia < hend) {
      const float4* p = reinterpret_cast<const float4*>(hyp + (size_t)ia * 12);
      const float4 u = p[0], v = p[1], w = p[2];
      a[0] = u.x; a[1] = u.y; a[2] = u.z; a[3] = u.w; a[4] = v.x; a[5] = v.y; a[6] = v.z; a[7] = v.w; a[8] = w.x;
    } else {
      a[0] = a[1] = a[3] = a[4] = a[6] = a[7] = 0.f; a[2] = a[5] = 1e18f; a[8] = 1.f;
    }
    if (ib < hend) {
      const float4* p = reinterpret_cast<const float4*>(hyp + (size_t)ib * 12);
      const float4 u = p[0], v = p[1], w = p[2];
      b[0] = u.x; b[1] = u.y; b[2] = u.z; b[3] = u.w; b[4] = v.x; b[5] = v.y; b[6] = v.z; b[7] = v.w; b[8] = w.x;
    } else {
      b[0] = b[1] = b[3] = b[4] = b[6] = b[7] = 0.f; b[2] = b[5] = 1e18f; b[8] = 1.f;
    }
    u64* d = sm + (size_t)j * 10;
#pragma unroll
    for (int k = 0; k < 9; ++k) d[k] = pk(a[k], b[k]);
    d[9] = 0ull;
  }
}

template <bool COUNT_INLIERS, int THREADS, int MINB, int CHUNK_PAIRS>
__global__ void __launch_bounds__(THREADS, MINB)
cost_argmin_kernel(const float4* __restrict__ pts, long long N, const float* __restrict__ hyp, int K, int k_per_block,
                   CostParams cp, FastOut o, int use_atomic_best) {
  __shared__ __align__(16) u64 sm[2][CHUNK_PAIRS * 10];
  __shared__ __align__(8) unsigned short s_cnt[2][2 * CHUNK_PAIRS];  // per-chunk inlier counters (<= 1024 each)

  const long long tile0 = (long long)blockIdx.x * (THREADS * FAST_P);
  const int kbeg = blockIdx.y * k_per_block;
  const int kend = min(K, kbeg + k_per_block);
  const int lane = threadIdx.x & 31;

  float X[FAST_P], Y[FAST_P], NX2[FAST_P], NY2[FAST_P], NEGMID[FAST_P], HALF[FAST_P], C[FAST_P];
  int BC[FAST_P], BL[FAST_P];
#pragma unroll
  for (int p = 0; p < FAST_P; ++p) {
    long long idx = tile0 + (long long)p * THREADS + threadIdx.x;
    const float4 q = pts[idx < N ? idx : N - 1];
    X[p] = q.x; Y[p] = q.y; NX2[p] = -q.z; NY2[p] = -q.w;
    BC[p] = cp.cost_outlier; BL[p] = 0;
    fast_thresholds(BC[p], cp, NEGMID[p], HALF[p]);
    C[p] = -NEGMID[p] - cp.thr2;  // (d2 - mid) + C = d2 - thr2: its sign bit is the inlier flag
    if (idx >= N) { HALF[p] = -1.f; C[p] = 3.0e38f; BC[p] = -1; }  // padding lanes: never an inlier, never a candidate
  }
  const u64 ONE2 = pk(1.f, 1.f);

  int buf = 0;
  // chunks are padded to an even number of pairs with "far" hypotheses so that the loop can take two pairs at a time
  auto pairs_in = [&](int c) { return min(CHUNK_PAIRS, (((kend - c + 1) / 2) + 1) & ~1); };
  stage_pairs_t<THREADS>(sm[0], hyp, kbeg, kend, pairs_in(kbeg));
  if (COUNT_INLIERS)
    for (int i = threadIdx.x; i < 4 * CHUNK_PAIRS; i += THREADS) (&s_cnt[0][0])[i] = 0;
  __syncthreads();

  for (int c0 = kbeg; c0 < kend; c0 += 2 * CHUNK_PAIRS) {
    const int npairs = pairs_in(c0);
    const int cn = c0 + 2 * CHUNK_PAIRS;
    if (cn < kend) stage_pairs_t<THREADS>(sm[buf ^ 1], hyp, cn, kend, pairs_in(cn));

    const ulonglong2* hp = reinterpret_cast<const ulonglong2*>(sm[buf]);
    unsigned cnt_addr = (unsigned)__cvta_generic_to_shared(s_cnt[buf]);
#pragma unroll 1
    for (int j = 0; j < npairs; j += 2, hp += 10, cnt_addr += 8) {
      u64 H[2][9];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const ulonglong2 q0 = hp[5 * q], q1 = hp[5 * q + 1], q2 = hp[5 * q + 2], q3 = hp[5 * q + 3], q4 = hp[5 * q + 4];
        H[q][0] = q0.x; H[q][1] = q0.y; H[q][2] = q1.x; H[q][3] = q1.y; H[q][4] = q2.x;
        H[q][5] = q2.y; H[q][6] = q3.x; H[q][7] = q3.y; H[q][8] = q4.x;
      }
      bool any = false;
      float M[2][FAST_P];
      unsigned mask[4] = {0u, 0u, 0u, 0u};  // sign bits of (d2 - thr2): one bit per correspondence, per hypothesis
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int p = 0; p < FAST_P; ++p) {
          const u64 xx = pk(X[p], X[p]), yy = pk(Y[p], Y[p]);
          const u64 sv = fma2(H[q][6], xx, fma2(H[q][7], yy, H[q][8]));
          const u64 xn = fma2(H[q][0], xx, fma2(H[q][1], yy, H[q][2]));
          const u64 yn = fma2(H[q][3], xx, fma2(H[q][4], yy, H[q][5]));
          float sl, sh;
          upk(sv, sl, sh);
          const u64 r = pk(rcp_approx(sl), rcp_approx(sh));
          const u64 dx = fma2(xn, r, pk(NX2[p], NX2[p]));
          const u64 dy = fma2(yn, r, pk(NY2[p], NY2[p]));
          const u64 t = fma2(dx, dx, fma2(dy, dy, pk(NEGMID[p], NEGMID[p])));  // d2 - mid
          float ta, tb;
          upk(t, ta, tb);
          if (COUNT_INLIERS) {
            float va, vb;
            upk(fma2(t, ONE2, pk(C[p], C[p])), va, vb);  // d2 - thr2
            mask[2 * q] = __funnelshift_l(__float_as_uint(va), mask[2 * q], 1);
            mask[2 * q + 1] = __funnelshift_l(__float_as_uint(vb), mask[2 * q + 1], 1);
          }
          M[q][p] = fminf(fabsf(ta), fabsf(tb));
          any = any || (M[q][p] < HALF[p]);
        }
      }
      if (COUNT_INLIERS) {
        // per-thread counts (<= 4) in byte fields -> one warp REDUX (<= 128 per field) -> one 64-bit shared RED of four
        // 16-bit counters by lane 0
        const unsigned bytes = __popc(mask[0]) | (__popc(mask[1]) << 8) | (__popc(mask[2]) << 16) | (__popc(mask[3]) << 24);
        const unsigned tot = __reduce_add_sync(0xffffffffu, bytes);
        const unsigned lo16 = __byte_perm(tot, 0u, 0x4140), hi16 = __byte_perm(tot, 0u, 0x4342);
        asm volatile("{ .reg .pred p; .reg .b64 v;\n\t"
                     "setp.eq.u32 p, %0, 0;\n\t"
                     "mov.b64 v, {%2, %3};\n\t"
                     "@p red.shared.add.u64 [%1], v; }"
                     :: "r"(lane), "r"(cnt_addr), "r"(lo16), "r"(hi16) : "memory");
      }
      if (any) {  // a candidate may lower some correspondence's best cost: exact update (rare after warm-up)
        float half_old[FAST_P];  // M[][] was measured against the thresholds in force before any update of this step
#pragma unroll
        for (int p = 0; p < FAST_P; ++p) half_old[p] = HALF[p];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float ha[9], hb[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) upk(H[q][k], ha[k], hb[k]);
          const int la = c0 + 2 * (j + q) + 1;  // 1-based label of lane a
#pragma unroll
          for (int p = 0; p < FAST_P; ++p) {
            if (M[q][p] < half_old[p]) {
              const float da = residual(ha, X[p], Y[p], -NX2[p], -NY2[p]);
              const float db = residual(hb, X[p], Y[p], -NX2[p], -NY2[p]);
              if (da < cp.T) { const int c = cost_in_range(da, cp); if (c < BC[p]) { BC[p] = c; BL[p] = la; } }
              if (db < cp.T) { const int c = cost_in_range(db, cp); if (c < BC[p]) { BC[p] = c; BL[p] = la + 1; } }
              fast_thresholds(BC[p], cp, NEGMID[p], HALF[p]);
              C[p] = -NEGMID[p] - cp.thr2;
            }
          }
        }
      }
    }
    __syncthreads();
    if (COUNT_INLIERS) {
      unsigned short* cnt = s_cnt[buf];
      for (int i = threadIdx.x; i < 2 * npairs; i += THREADS) {
        const int v = cnt[i];
        cnt[i] = 0;
        if (v && c0 + i < kend) atomicAdd(o.inlier_count + c0 + i, v);
      }
      // the zeroing above is ordered before the buffer's next use by the __syncthreads of the following iteration
    }
    buf ^= 1;
  }
  if (o.best) {
#pragma unroll
    for (int p = 0; p < FAST_P; ++p) {
      const long long idx = tile0 + (long long)p * THREADS + threadIdx.x;
      if (idx < N && BL[p] != 0) {
        const u64 v = ((u64)(uint32_t)BC[p] << 32) | (uint32_t)BL[p];
        if (use_atomic_best) atomicMin(o.best + idx, v);
        else o.best[idx] = v;
      }
    }
  }
}

#ifdef MH_TUNING   // superseded K2 generations v4 (transposed tile) and v5 (first mma.sync kernel): tuning builds only (make TUNING=1)
// ----------------------------------------------------------------------------
// v4 "transposed" register tile: each thread keeps HP hypothesis PAIRS in registers (18 regs per pair) and the CTA's
// correspondence tile lives in shared memory, read with warp-uniform (broadcast) LDS — per correspondence one LDS.128
// of (x, y, -x2, -y2) and one LDS.128 of its mutable argmin filter (-mid, half, mid - thr2).  Compared with v3 this
// makes the per-hypothesis inlier counters thread-private (no REDUX / shared atomics in the loop: sign bits are
// funnel-shifted into a 32-correspondence mask and POPCed once per 32 correspondences) and keeps every FFMA2 operand in
// ordinary registers (no R2UR traffic).  The per-correspondence argmin state is shared by the CTA instead: a candidate
// that passes the folded interval test recomputes its exact cost, the warp REDUX.MINs (cost << 16 | label) over its
// candidate lanes, and one lane atomicMins the shared best and publishes the tightened filter with a single STS.128.
// A reader may see a filter that is one update old — always the looser one, so no candidate is ever missed; the exact
// integer compare rejects the false positives.  Per residual: 5 FFMA2 + 0.5 FADD2 + 1 MUFU + 1 SHF + 0.5 FMNMX +
// 0.5 FSETP + 2/(2 HP) LDS.
// ----------------------------------------------------------------------------
// hypotheses [K][12] -> pair-interleaved [ceil(K/2)][10] 64-bit words {h_k(a), h_k(b)} (k = 0..8, + pad), so that a thread
// loads its register-resident pairs with 128-bit loads and every FFMA2 operand is born as an aligned register pair.
// Out-of-range hypotheses become "far" (residual ~1e36, never a hit).
__global__ void pack_hyp_pairs_kernel(const float* __restrict__ hyp, int K, int npairs, u64* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= npairs) return;
  float a[9], b[9];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    float* d = e ? b : a;
    const int ih = 2 * j + e;
    if (ih < K) {
      const float4* p = reinterpret_cast<const float4*>(hyp + (size_t)ih * 12);
      const float4 u = p[0], v = p[1], w = p[2];
      d[0] = u.x; d[1] = u.y; d[2] = u.z; d[3] = u.w; d[4] = v.x; d[5] = v.y; d[6] = v.z; d[7] = v.w; d[8] = w.x;
    } else {
      d[0] = d[1] = d[3] = d[4] = d[6] = d[7] = 0.f; d[2] = d[5] = 1e18f; d[8] = 1.f;
    }
  }
  u64* o = out + (size_t)j * 10;
#pragma unroll
  for (int k = 0; k < 9; ++k) o[k] = pk(a[k], b[k]);
  o[9] = 0ull;
}

template <bool COUNT_INLIERS, int THREADS, int MINB, int HP, int PT>
__global__ void __launch_bounds__(THREADS, MINB)
cost_argmin_t_kernel(const float4* __restrict__ pts, long long N, const u64* __restrict__ hyp_pairs, int K, int k_per_block,
                     CostParams cp, FastOut o, int use_atomic_best) {
  __shared__ __align__(16) float4 s_pt[PT];      // x, y, -x2, -y2
  __shared__ __align__(16) float4 s_flt[PT];     // -mid, half, mid - thr2, unused
  __shared__ unsigned s_best[PT];                // cost << 16 | label  (label 0 = outlier at the outlier cost)
  static_assert(PT % 32 == 0, "tile must be a multiple of 32 correspondences");
  constexpr int HYP_PER_THREAD = 2 * HP, HYP_PER_BLOCK = THREADS * HYP_PER_THREAD;

  const long long tile0 = (long long)blockIdx.x * PT;
  const int kbeg = blockIdx.y * k_per_block;
  const int kend = min(K, kbeg + k_per_block);
  const unsigned best_init = ((unsigned)min(cp.cost_outlier, 0xffff) << 16);

  for (int i = threadIdx.x; i < PT; i += THREADS) {
    const long long idx = tile0 + i;
    const float4 q = pts[idx < N ? idx : N - 1];
    float negmid, half;
    fast_thresholds(cp.cost_outlier, cp, negmid, half);
    float c = -negmid - cp.thr2;
    if (idx >= N) { half = -1.f; c = 3.0e38f; }  // padding: never an inlier, never a candidate
    s_pt[i] = make_float4(q.x, q.y, -q.z, -q.w);
    s_flt[i] = make_float4(negmid, half, c, 0.f);
    s_best[i] = best_init;
  }
  __syncthreads();
  const u64 ONE2 = pk(1.f, 1.f);

  for (int hb = kbeg; hb < kend; hb += HYP_PER_BLOCK) {
    // this thread's hypothesis pairs -> registers, as aligned 64-bit pairs straight from the packed array
    u64 H[HP][9];
    const int h0 = hb + threadIdx.x * HYP_PER_THREAD;
#pragma unroll
    for (int q = 0; q < HP; ++q) {
      const int jp = min((h0 >> 1) + q, (K + 1) / 2 - 1);  // clamp: pairs past the end are never counted (labels >= kend)
      const ulonglong2* hp = reinterpret_cast<const ulonglong2*>(hyp_pairs + (size_t)jp * 10);
      const ulonglong2 q0 = hp[0], q1 = hp[1], q2 = hp[2], q3 = hp[3], q4 = hp[4];
      H[q][0] = q0.x; H[q][1] = q0.y; H[q][2] = q1.x; H[q][3] = q1.y; H[q][4] = q2.x;
      H[q][5] = q2.y; H[q][6] = q3.x; H[q][7] = q3.y; H[q][8] = q4.x;
    }
    const bool live_pair0 = true;
    (void)live_pair0;
    int cnt[HYP_PER_THREAD];
#pragma unroll
    for (int h = 0; h < HYP_PER_THREAD; ++h) cnt[h] = 0;

#pragma unroll 1
    for (int i0 = 0; i0 < PT; i0 += 32) {
      unsigned mask[HYP_PER_THREAD];
#pragma unroll
      for (int h = 0; h < HYP_PER_THREAD; ++h) mask[h] = 0u;
#pragma unroll 2
      for (int ii = 0; ii < 32; ++ii) {
        const int i = i0 + ii;
        const float4 P = s_pt[i];
        const float4 Fl = s_flt[i];
        const u64 xx = pk(P.x, P.x), yy = pk(P.y, P.y), nx2 = pk(P.z, P.z), ny2 = pk(P.w, P.w);
        const u64 negmid = pk(Fl.x, Fl.x), cc = pk(Fl.z, Fl.z);
        bool any = false;
#pragma unroll
        for (int q = 0; q < HP; ++q) {
          const u64 sv = fma2(H[q][6], xx, fma2(H[q][7], yy, H[q][8]));
          const u64 xn = fma2(H[q][0], xx, fma2(H[q][1], yy, H[q][2]));
          const u64 yn = fma2(H[q][3], xx, fma2(H[q][4], yy, H[q][5]));
          float sl, sh;
          upk(sv, sl, sh);
          const u64 r = pk(rcp_approx(sl), rcp_approx(sh));
          const u64 dx = fma2(xn, r, nx2);
          const u64 dy = fma2(yn, r, ny2);
          const u64 t = fma2(dx, dx, fma2(dy, dy, negmid));  // d2 - mid
          float ta, tb;
          upk(t, ta, tb);
          if (COUNT_INLIERS) {
            float va, vb;
            upk(fma2(t, ONE2, cc), va, vb);  // d2 - thr2: the sign bit is the inlier flag
            mask[2 * q] = __funnelshift_l(__float_as_uint(va), mask[2 * q], 1);
            mask[2 * q + 1] = __funnelshift_l(__float_as_uint(vb), mask[2 * q + 1], 1);
          }
          any = any || (fminf(fabsf(ta), fabsf(tb)) < Fl.y);
        }
        if (any) {  // some lane holds a candidate for correspondence i: exact update
          unsigned mine = 0xffffffffu;
#pragma unroll
          for (int q = 0; q < HP; ++q) {
            float ha[9], hb2[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) upk(H[q][k], ha[k], hb2[k]);
            const float da = residual(ha, P.x, P.y, -P.z, -P.w);
            const float db = residual(hb2, P.x, P.y, -P.z, -P.w);
            const unsigned la = (unsigned)(h0 + 2 * q + 1);  // 1-based label of lane a
            if (da < cp.T && (int)la <= kend) mine = min(mine, ((unsigned)cost_in_range(da, cp) << 16) | la);
            if (db < cp.T && (int)la + 1 <= kend) mine = min(mine, ((unsigned)cost_in_range(db, cp) << 16) | (la + 1));
          }
          const unsigned act = __activemask();
          const unsigned wbest = __reduce_min_sync(act, mine);
          if (mine == wbest && wbest < s_best[i]) {  // at most one lane per distinct value; ties carry the same value
            const unsigned old = atomicMin(&s_best[i], wbest);
            const unsigned now = min(old, wbest);
            float nm, hf;
            // hypotheses arrive in no particular label order here, so an equal-cost candidate with a LOWER label must
            // still get through (GCO keeps the first label among equal costs): filter on cost <= best, not < best
            fast_thresholds((int)(now >> 16) + 1, cp, nm, hf);
            if (tile0 + i < N) s_flt[i] = make_float4(nm, hf, -nm - cp.thr2, 0.f);
          }
        }
      }
      if (COUNT_INLIERS) {
#pragma unroll
        for (int h = 0; h < HYP_PER_THREAD; ++h) cnt[h] += __popc(mask[h]);
      }
    }
    if (COUNT_INLIERS) {
#pragma unroll
      for (int h = 0; h < HYP_PER_THREAD; ++h)
        if (cnt[h] && h0 + h < kend) atomicAdd(o.inlier_count + h0 + h, cnt[h]);
    }
  }
  __syncthreads();
  if (o.best) {
    for (int i = threadIdx.x; i < PT; i += THREADS) {
      const long long idx = tile0 + i;
      const unsigned b = s_best[i];
      if (idx < N && (b & 0xffffu) != 0u) {
        const u64 v = ((u64)(b >> 16) << 32) | (u64)(b & 0xffffu);
        if (use_atomic_best) atomicMin(o.best + idx, v);
        else o.best[idx] = v;
      }
    }
  }
}

// ----------------------------------------------------------------------------
// v5: the 3x3 homography product on the tensor cores.  60 % of the FP32 work of a residual is the projective product
// (s, xn, yn) = H (x, y, 1)^T — a GEMM with inner dimension 3.  Here it runs as mma.sync.m16n8k8 TF32 with the 3xTF32
// split packed INTO the k = 8 inner dimension, so one MMA gives FP32-grade sums:
//     A[point][k]   = [ xhi, yhi, 1, xlo, xhi, yhi, ylo, 1 ]                (16 correspondences x 8)
//     B_s[k][hyp]   = [ h6hi, h7hi, h8hi, h6hi, h6lo, h7lo, h7hi, h8lo ]    (8 x 8 hypotheses; same pattern for xn, yn)
//     s = xhi h6hi + yhi h7hi + h8hi + xlo h6hi + xhi h6lo + yhi h7lo + ylo h7hi + h8lo     (only lo*lo terms dropped)
// Each thread then owns 2 correspondences (fragment rows g, g+8) x 2 hypotheses (columns 2t, 2t+1) per MMA triple and
// finishes them on the FP32 pipe with 4 FFMA2 + 1 FADD2 per pair (reciprocal, residual, folded interval / inlier tests
// exactly as in v3).  The tensor-core sums only feed the conservative argmin filter and the inlier counts; every
// candidate is re-evaluated with the dense kernel's FP32 instruction sequence, so (cost, label) stay bit-identical.
// Per-correspondence argmin state is replicated in the 4 lanes of a quad and kept coherent by a quad min-reduce in the
// (rare) update path; per-hypothesis inlier bits are funnel-shifted into masks and added to shared counters once per
// 8-hypothesis block.
// ----------------------------------------------------------------------------
constexpr int MMA_THREADS = 256;
constexpr int MMA_CH = 512;  // hypotheses staged per chunk (48 KB of split fragments)

template <bool COUNT_INLIERS, int MINB, int PB>
__global__ void __launch_bounds__(MMA_THREADS, MINB)
cost_argmin_mma_kernel(const float4* __restrict__ pts, long long N, const float* __restrict__ hyp, int K, int k_per_block,
                       CostParams cp, FastOut o, int use_atomic_best) {
  // dynamic shared memory (56 KB > the 48 KB static limit):
  //   sB   [MMA_CH][3][4] float2 : [hyp][s|xn|yn][t] = {B[t][hyp], B[t+4][hyp]} as tf32 bit patterns
  //   sCntW[warps][MMA_CH] u16   : per-warp inlier counts of the chunk (<= 16 PB each)
  extern __shared__ __align__(16) unsigned char mma_smem[];
  float2 (*sB)[3][4] = reinterpret_cast<float2 (*)[3][4]>(mma_smem);
  unsigned short (*sCntW)[MMA_CH] = reinterpret_cast<unsigned short (*)[MMA_CH]>(mma_smem + sizeof(float2) * MMA_CH * 12);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  constexpr int PTS_PER_WARP = 16 * PB, TILE = (MMA_THREADS / 32) * PTS_PER_WARP;
  const long long tile0 = (long long)blockIdx.x * TILE + (long long)warp * PTS_PER_WARP;
  const int kbeg = blockIdx.y * k_per_block;
  const int kend = min(K, kbeg + k_per_block);

  // ---- per-thread correspondence state: rows g and g+8 of each of the warp's PB 16-row blocks ----------------------
  unsigned A[PB][4];
  float X[PB][2], Y[PB][2], NX2[PB][2], NY2[PB][2], NEGMID[PB][2], HALF[PB][2], C[PB][2];
  unsigned BEST[PB][2];
  const unsigned best_init = ((unsigned)min(cp.cost_outlier, 0xffff) << 16);
#pragma unroll
  for (int pb = 0; pb < PB; ++pb)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const long long idx = tile0 + pb * 16 + g + 8 * r;
      const float4 q = pts[idx < N ? idx : N - 1];
      X[pb][r] = q.x; Y[pb][r] = q.y; NX2[pb][r] = -q.z; NY2[pb][r] = -q.w;
      BEST[pb][r] = best_init;
      fast_thresholds(cp.cost_outlier, cp, NEGMID[pb][r], HALF[pb][r]);
      C[pb][r] = -NEGMID[pb][r] - cp.thr2;
      if (idx >= N) { HALF[pb][r] = -1.f; C[pb][r] = 3.0e38f; }
      // A fragment entries of this row: k = t and k = t + 4 of [xhi, yhi, 1, xlo, xhi, yhi, ylo, 1]
      const unsigned xhi = to_tf32(q.x), yhi = to_tf32(q.y);
      const unsigned xlo = to_tf32(q.x - __uint_as_float(xhi)), ylo = to_tf32(q.y - __uint_as_float(yhi));
      const unsigned one = 0x3f800000u;
      A[pb][r] = t == 0 ? xhi : t == 1 ? yhi : t == 2 ? one : xlo;      // a0 (row g) / a1 (row g+8): k = t
      A[pb][2 + r] = t == 0 ? xhi : t == 1 ? yhi : t == 2 ? ylo : one;  // a2 / a3: k = t + 4
    }
  const u64 ONE2 = pk(1.f, 1.f);

  for (int c0 = kbeg; c0 < kend; c0 += MMA_CH) {
    // ---- stage a chunk of hypotheses as 3xTF32 split B fragments ---------------------------------------------------------
    __syncthreads();
    for (int j = threadIdx.x; j < MMA_CH; j += MMA_THREADS) {
      const int ih = c0 + j;
      float h[9];
      if (ih < kend) {
        const float4* p = reinterpret_cast<const float4*>(hyp + (size_t)ih * 12);
        const float4 u = p[0], v = p[1], w = p[2];
        h[0] = u.x; h[1] = u.y; h[2] = u.z; h[3] = u.w; h[4] = v.x; h[5] = v.y; h[6] = v.z; h[7] = v.w; h[8] = w.x;
      } else {
        h[0] = h[1] = h[3] = h[4] = h[6] = h[7] = 0.f; h[2] = h[5] = 1e18f; h[8] = 1.f;
      }
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const int base = m == 0 ? 6 : m == 1 ? 0 : 3;  // s uses (h6,h7,h8); xn (h0,h1,h2); yn (h3,h4,h5)
        const float ha = h[base], hb = h[base + 1], hc = h[base + 2];
        const float ahi = __uint_as_float(to_tf32(ha)), bhi = __uint_as_float(to_tf32(hb)), chi = __uint_as_float(to_tf32(hc));
        const float alo = __uint_as_float(to_tf32(ha - ahi)), blo = __uint_as_float(to_tf32(hb - bhi)),
                    clo = __uint_as_float(to_tf32(hc - chi));
        // B[k] = [ahi, bhi, chi, ahi, alo, blo, bhi, clo]; entry [t] = {B[t], B[t+4]}
        float4* d = reinterpret_cast<float4*>(&sB[j][m][0]);
        d[0] = make_float4(ahi, alo, bhi, blo);
        d[1] = make_float4(chi, bhi, ahi, clo);
      }
    }
    __syncthreads();
    const int nblk = (min(MMA_CH, kend - c0) + 7) >> 3;
    const float2* bp = &sB[g][0][t];                      // + hb * 8 hypotheses = hb * 96 float2
    unsigned short* cw = &sCntW[warp][2 * t];

#pragma unroll 1
    for (int hb = 0; hb < nblk; hb += 2) {                // two 8-hypothesis blocks per trip
      unsigned packed = 0u;                               // byte fields: inlier counts of (blk A: 2t, 2t+1), (blk B: 2t, 2t+1)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const float2 bs = bp[(hb + hh) * 96], bx = bp[(hb + hh) * 96 + 4], by = bp[(hb + hh) * 96 + 8];
        unsigned mask0 = 0u, mask1 = 0u;
        float T_[PB][2][2];
        bool cand = false;
#pragma unroll
        for (int pb = 0; pb < PB; ++pb) {
          float S[4], XN[4], YN[4];
          mma_tf32(S, A[pb], __float_as_uint(bs.x), __float_as_uint(bs.y));
          mma_tf32(XN, A[pb], __float_as_uint(bx.x), __float_as_uint(bx.y));
          mma_tf32(YN, A[pb], __float_as_uint(by.x), __float_as_uint(by.y));
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const u64 rr = pk(rcp_approx(S[2 * r]), rcp_approx(S[2 * r + 1]));
            const u64 dx = fma2(pk(XN[2 * r], XN[2 * r + 1]), rr, pk(NX2[pb][r], NX2[pb][r]));
            const u64 dy = fma2(pk(YN[2 * r], YN[2 * r + 1]), rr, pk(NY2[pb][r], NY2[pb][r]));
            const u64 tt = fma2(dx, dx, fma2(dy, dy, pk(NEGMID[pb][r], NEGMID[pb][r])));  // d2 - mid
            upk(tt, T_[pb][r][0], T_[pb][r][1]);
            if (COUNT_INLIERS) {
              float va, vb;
              upk(fma2(tt, ONE2, pk(C[pb][r], C[pb][r])), va, vb);  // d2 - thr2: sign bit = inlier
              mask0 = __funnelshift_l(__float_as_uint(va), mask0, 1);
              mask1 = __funnelshift_l(__float_as_uint(vb), mask1, 1);
            }
            cand = cand || (fminf(fabsf(T_[pb][r][0]), fabsf(T_[pb][r][1])) < HALF[pb][r]);
          }
        }
        if (COUNT_INLIERS) packed |= (__popc(mask0) | (__popc(mask1) << 8)) << (16 * hh);
        if (__any_sync(0xffffffffu, cand)) {  // warp-uniform: some lane holds a candidate in this 8-block -> exact update
#pragma unroll
          for (int pb = 0; pb < PB; ++pb)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              unsigned mine = 0xffffffffu;
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                if (fabsf(T_[pb][r][e]) < HALF[pb][r]) {
                  const int ih = c0 + (hb + hh) * 8 + 2 * t + e;  // 0-based hypothesis index
                  if (ih < kend) {
                    const float4* hp = reinterpret_cast<const float4*>(hyp + (size_t)ih * 12);
                    const float4 u = hp[0], v = hp[1], w = hp[2];
                    const float hx[9] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w, w.x};
                    const float d2 = residual(hx, X[pb][r], Y[pb][r], -NX2[pb][r], -NY2[pb][r]);
                    if (d2 < cp.T) mine = min(mine, ((unsigned)cost_in_range(d2, cp) << 16) | (unsigned)(ih + 1));
                  }
                }
              }
              mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, 1));  // the 4 lanes of a quad share rows g, g+8
              mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, 2));
              if (mine < BEST[pb][r]) {
                BEST[pb][r] = mine;
                // labels rise along the loop for a given correspondence: only strictly cheaper candidates matter
                fast_thresholds((int)(mine >> 16), cp, NEGMID[pb][r], HALF[pb][r]);
                C[pb][r] = -NEGMID[pb][r] - cp.thr2;
              }
            }
        }
      }
      if (COUNT_INLIERS) {
        // sum the byte fields over the 8 lanes that share t (<= 2 PB per lane, <= 16 PB per field): 3 shuffles, then
        // lanes g == 0 store this warp's counts of the two blocks (each (warp, hypothesis) is visited once per chunk)
        packed += __shfl_xor_sync(0xffffffffu, packed, 4);
        packed += __shfl_xor_sync(0xffffffffu, packed, 8);
        packed += __shfl_xor_sync(0xffffffffu, packed, 16);
        if (g == 0) {
          *reinterpret_cast<unsigned*>(cw + hb * 8) = (packed & 0xffu) | ((packed & 0xff00u) << 8);
          *reinterpret_cast<unsigned*>(cw + hb * 8 + 8) = ((packed >> 16) & 0xffu) | ((packed >> 8) & 0xff0000u);
        }
      }
    }
    if (COUNT_INLIERS) {
      __syncthreads();
      const int nh = min(MMA_CH, kend - c0);
      for (int j = threadIdx.x; j < nh; j += MMA_THREADS) {
        int v = 0;
#pragma unroll
        for (int w = 0; w < MMA_THREADS / 32; ++w) v += sCntW[w][j];
        if (v) atomicAdd(o.inlier_count + c0 + j, v);
      }
    }
  }
  if (o.best && t == 0) {  // the quad holds identical state: lane t == 0 writes
#pragma unroll
    for (int pb = 0; pb < PB; ++pb)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const long long idx = tile0 + pb * 16 + g + 8 * r;
        const unsigned b = BEST[pb][r];
        if (idx < N && (b & 0xffffu) != 0u) {
          const u64 v = ((u64)(b >> 16) << 32) | (u64)(b & 0xffffu);
          if (use_atomic_best) atomicMin(o.best + idx, v);
          else o.best[idx] = v;
        }
      }
  }
}
#endif  // MH_TUNING

__global__ void fused_init_kernel(long long N, int K, int32_t* list_count, u64* best, int32_t* inlier_count,
                                  u64 best_init) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    if (list_count) list_count[i] = 0;
    if (best) best[i] = best_init;
  }
  if (inlier_count && i < K) inlier_count[i] = 0;
}

mh_status launch_cost_argmin_tc(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, const CostParams& cp,
                                const FastOut& fo, int config);  // k2_mma.cu
mh_status launch_cost_list_tc(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, const CostParams& cp,
                              const FastOut& fo);  // k2_mma.cu
mh_status launch_cost_argmin_tmem(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, const CostParams& cp,
                                  const FastOut& fo, int config);  // k2_tmem.cu

int g_list_variant = 1;   // 1 = tensor-core list kernel (default), 0 = first-generation cost_fused_kernel
int g_fused_variant = 1;  // 1 = packed FFMA2 (default), 0 = scalar FFMA (A/B evidence)
int g_fast_config = 55;   // fast-path variant, see launch_cost_fused / launch_cost_argmin_tc; 55 = v7 tensor-core kernel, 4 warps x 5 CTAs/SM x 48 rows/warp (default); 5 = v3 FFMA2 kernel

mh_status launch_cost_fused(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, int kmax,
                            uint32_t* d_list, int32_t* d_list_count, unsigned long long* d_best,
                            int32_t* d_inlier_count) {
  if (N <= 0) return MH_OK;
  const CostParams cp = cost_params(ctx);
  if (cp.lam > 255.f && d_list) return fail(ctx, MH_EINVAL, "mh_data_cost_fused: 100/lambda must be <= 255 for packed lists");
  {
    const long long n = std::max<long long>(N, K);
    const u64 init = ((u64)(uint32_t)cp.cost_outlier << 32);  // label 0 at the outlier cost
    fused_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(N, K, d_list_count, d_best, d_inlier_count, init);
    MH_LAUNCHED(ctx, "fused_init_kernel");
  }
  if (K <= 0) return MH_OK;
  const unsigned tiles = (unsigned)((N + FUSED_TILE - 1) / FUSED_TILE);
  // split K across CTAs only when there are too few point tiles to fill the chip (~2 waves)
  int ksplit = 1;
  const int want = 2 * ctx->sm_count;
  if ((int)tiles < want) ksplit = std::min((K + 2 * FUSED_CHUNK_PAIRS - 1) / (2 * FUSED_CHUNK_PAIRS), (want + (int)tiles - 1) / (int)tiles);
  ksplit = std::max(1, ksplit);
  int k_per_block = (K + ksplit - 1) / ksplit;
  k_per_block = (k_per_block + 1) & ~1;  // even, so label pairs stay aligned
  ksplit = (K + k_per_block - 1) / k_per_block;
  dim3 grid(tiles, (unsigned)ksplit);
  if (d_list && d_list_count && kmax > 0 && g_fused_variant == 1 && g_list_variant == 1 && K < 65535 && cp.cost_outlier <= 0xffff) {
    // sparse lists from the tensor-core kernel (filter + deferred queue + append-drain); g_list_variant 0 = the first-generation
    // emit-on-every-hit kernel below (cross-check)
    FastOut fo{(u64*)d_best, d_inlier_count, d_list, d_list_count, kmax};
    if (!d_best) return fail(ctx, MH_EINVAL, "mh_data_cost_fused: the list kernel also produces the argmin: pass d_best");
    return launch_cost_list_tc(ctx, d_pts, N, d_hyp, K, cp, fo);
  }
  if (!d_list && !d_list_count && g_fused_variant == 1) {
    FastOut fo{(u64*)d_best, d_inlier_count};
    const bool cnt = d_inlier_count != nullptr;
    // g_fast_config selects (threads per CTA, CTAs per SM, hypothesis pairs per staged chunk)
    auto launch = [&](auto kernel_cnt, auto kernel_nocnt, int threads, int chunk_pairs) -> mh_status {
      const long long tile = (long long)threads * FAST_P;
      const unsigned tiles_f = (unsigned)((N + tile - 1) / tile);
      int ks = 1;
      if ((int)tiles_f < want) ks = std::min((K + 2 * chunk_pairs - 1) / (2 * chunk_pairs), (want + (int)tiles_f - 1) / (int)tiles_f);
      ks = std::max(1, ks);
      int kpb = (K + ks - 1) / ks;
      kpb = (kpb + 3) & ~3;  // multiple of 4: pair-of-pairs alignment
      ks = (K + kpb - 1) / kpb;
      dim3 gridf(tiles_f, (unsigned)ks);
      if (cnt) kernel_cnt<<<gridf, threads, 0, ctx->stream>>>(d_pts, N, d_hyp, K, kpb, cp, fo, ks > 1);
      else kernel_nocnt<<<gridf, threads, 0, ctx->stream>>>(d_pts, N, d_hyp, K, kpb, cp, fo, ks > 1);
      return MH_OK;
    };
#ifdef MH_TUNING
    auto launch_t = [&](auto kernel_cnt, auto kernel_nocnt, int threads, int hyp_per_block, int pt) -> mh_status {
      const int npairs = (K + 1) / 2;
      MH_TRY(ensure_scratch(ctx, sizeof(u64) * 10 * (uint64_t)npairs));
      u64* d_pairs = (u64*)ctx->scratch;
      pack_hyp_pairs_kernel<<<(unsigned)((npairs + 127) / 128), 128, 0, ctx->stream>>>(d_hyp, K, npairs, d_pairs);
      MH_LAUNCHED(ctx, "pack_hyp_pairs_kernel");
      const unsigned tiles_f = (unsigned)((N + pt - 1) / pt);
      int ks = 1;
      if ((int)tiles_f < want) ks = std::min((K + hyp_per_block - 1) / hyp_per_block, (want + (int)tiles_f - 1) / (int)tiles_f);
      ks = std::max(1, ks);
      int kpb = (K + ks - 1) / ks;
      kpb = ((kpb + hyp_per_block - 1) / hyp_per_block) * hyp_per_block;  // whole hypothesis blocks per CTA
      ks = (K + kpb - 1) / kpb;
      dim3 gridf(tiles_f, (unsigned)ks);
      if (cnt) kernel_cnt<<<gridf, threads, 0, ctx->stream>>>(d_pts, N, d_pairs, K, kpb, cp, fo, ks > 1);
      else kernel_nocnt<<<gridf, threads, 0, ctx->stream>>>(d_pts, N, d_pairs, K, kpb, cp, fo, ks > 1);
      return MH_OK;
    };
    auto launch_m = [&](auto kernel_cnt, auto kernel_nocnt, int pb) -> mh_status {
      const int tile = (MMA_THREADS / 32) * 16 * pb;
      const unsigned tiles_f = (unsigned)((N + tile - 1) / tile);
      int ks = 1;
      if ((int)tiles_f < want) ks = std::min((K + MMA_CH - 1) / MMA_CH, (want + (int)tiles_f - 1) / (int)tiles_f);
      ks = std::max(1, ks);
      int kpb = (K + ks - 1) / ks;
      kpb = ((kpb + MMA_CH - 1) / MMA_CH) * MMA_CH;
      ks = (K + kpb - 1) / kpb;
      dim3 gridf(tiles_f, (unsigned)ks);
      const size_t smem = sizeof(float2) * MMA_CH * 12 + sizeof(unsigned short) * (MMA_THREADS / 32) * MMA_CH;
      if (cnt) {
        MH_CUDA(ctx, cudaFuncSetAttribute(kernel_cnt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel_cnt<<<gridf, MMA_THREADS, smem, ctx->stream>>>(d_pts, N, d_hyp, K, kpb, cp, fo, ks > 1);
      } else {
        MH_CUDA(ctx, cudaFuncSetAttribute(kernel_nocnt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel_nocnt<<<gridf, MMA_THREADS, smem, ctx->stream>>>(d_pts, N, d_hyp, K, kpb, cp, fo, ks > 1);
      }
      return MH_OK;
    };
#endif
    if (K >= 65535 || cp.cost_outlier > 0xffff) g_fast_config = std::min(g_fast_config, 6);  // the tensor-core kernel packs (cost, label) in 32 bits
    if (g_fast_config >= 70 && g_fast_config < 100) return launch_cost_argmin_tmem(ctx, d_pts, N, d_hyp, K, cp, fo, g_fast_config);
    if (g_fast_config >= 30 && g_fast_config < 70) return launch_cost_argmin_tc(ctx, d_pts, N, d_hyp, K, cp, fo, g_fast_config);
    switch (g_fast_config) {
#ifdef MH_TUNING
      case 20: MH_TRY(launch_m(cost_argmin_mma_kernel<true, 2, 2>, cost_argmin_mma_kernel<false, 2, 2>, 2)); break;
      case 21: MH_TRY(launch_m(cost_argmin_mma_kernel<true, 3, 2>, cost_argmin_mma_kernel<false, 3, 2>, 2)); break;
      case 22: MH_TRY(launch_m(cost_argmin_mma_kernel<true, 2, 4>, cost_argmin_mma_kernel<false, 2, 4>, 4)); break;
      case 23: MH_TRY(launch_m(cost_argmin_mma_kernel<true, 3, 1>, cost_argmin_mma_kernel<false, 3, 1>, 1)); break;
      case 24: MH_TRY(launch_m(cost_argmin_mma_kernel<true, 4, 1>, cost_argmin_mma_kernel<false, 4, 1>, 1)); break;
      case 25: MH_TRY(launch_m(cost_argmin_mma_kernel<true, 3, 3>, cost_argmin_mma_kernel<false, 3, 3>, 3)); break;
      case 10: launch_t(cost_argmin_t_kernel<true, 128, 6, 2, 512>, cost_argmin_t_kernel<false, 128, 6, 2, 512>, 128, 512, 512); break;
      case 11: launch_t(cost_argmin_t_kernel<true, 128, 5, 2, 512>, cost_argmin_t_kernel<false, 128, 5, 2, 512>, 128, 512, 512); break;
      case 12: launch_t(cost_argmin_t_kernel<true, 128, 4, 4, 512>, cost_argmin_t_kernel<false, 128, 4, 4, 512>, 128, 1024, 512); break;
      case 13: launch_t(cost_argmin_t_kernel<true, 128, 3, 4, 512>, cost_argmin_t_kernel<false, 128, 3, 4, 512>, 128, 1024, 512); break;
      case 14: launch_t(cost_argmin_t_kernel<true, 256, 3, 2, 1024>, cost_argmin_t_kernel<false, 256, 3, 2, 1024>, 256, 1024, 1024); break;
      case 15: launch_t(cost_argmin_t_kernel<true, 128, 7, 2, 256>, cost_argmin_t_kernel<false, 128, 7, 2, 256>, 128, 512, 256); break;
#endif
#ifdef MH_TUNING
      case 0: launch(cost_argmin_kernel<true, 256, 3, 256>, cost_argmin_kernel<false, 256, 3, 256>, 256, 256); break;
      case 1: launch(cost_argmin_kernel<true, 256, 2, 256>, cost_argmin_kernel<false, 256, 2, 256>, 256, 256); break;
      case 3: launch(cost_argmin_kernel<true, 128, 5, 256>, cost_argmin_kernel<false, 128, 5, 256>, 128, 256); break;
      case 4: launch(cost_argmin_kernel<true, 128, 6, 128>, cost_argmin_kernel<false, 128, 6, 128>, 128, 128); break;
#endif
      case 5: launch(cost_argmin_kernel<true, 128, 7, 128>, cost_argmin_kernel<false, 128, 7, 128>, 128, 128); break;
      case 6: launch(cost_argmin_kernel<true, 128, 4, 256>, cost_argmin_kernel<false, 128, 4, 256>, 128, 256); break;
      default: return fail(ctx, MH_EINVAL, "unknown fast-path config (this build keeps 5, 6, 52 and 55; make TUNING=1 for the rest)");
    }
    MH_LAUNCHED(ctx, "cost_argmin_kernel");
    return MH_OK;
  }
  FusedOut o{d_list, d_list_count, (u64*)d_best, d_inlier_count, kmax};
  if (g_fused_variant)
    cost_fused_kernel<true><<<grid, FUSED_THREADS, 0, ctx->stream>>>(d_pts, N, d_hyp, K, k_per_block, cp, o);
  else
    cost_fused_kernel<false><<<grid, FUSED_THREADS, 0, ctx->stream>>>(d_pts, N, d_hyp, K, k_per_block, cp, o);
  MH_LAUNCHED(ctx, "cost_fused_kernel");
  return MH_OK;
}

// ============================================================================
// Inlier statistics (MergingStep, MultiH.cpp:430-458): per hypothesis the 6
// uniques of S = sum [x y 1]^T [x y 1] over sites with d2 < thr^2, accumulated
// in normalised coordinates (FP64 atomics, one per warp and hypothesis).
// ============================================================================
constexpr int STATS_THREADS = 256;

__global__ void __launch_bounds__(STATS_THREADS) inlier_stats_kernel(const float4* __restrict__ pts, long long N,
                                                                     const float* __restrict__ hyp, int K,
                                                                     double* __restrict__ scatter, CostParams cp) {
  extern __shared__ float sh[];  // K x 9
  for (int i = threadIdx.x; i < K * 9; i += STATS_THREADS) sh[i] = hyp[(size_t)(i / 9) * 12 + (i % 9)];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * STATS_THREADS + threadIdx.x;
  const bool live = idx < N;
  const float4 q = pts[live ? idx : N - 1];
  const int lane = threadIdx.x & 31;
  for (int k = 0; k < K; ++k) {
    const float d2 = residual(sh + 9 * k, q.x, q.y, q.z, q.w);
    const bool in = live && d2 < cp.thr2;
    const unsigned b = __ballot_sync(0xffffffffu, in);
    if (b == 0) continue;
    double v[5] = {in ? (double)q.x * q.x : 0.0, in ? (double)q.x * q.y : 0.0, in ? (double)q.x : 0.0,
                   in ? (double)q.y * q.y : 0.0, in ? (double)q.y : 0.0};
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int t = 0; t < 5; ++t) v[t] += __shfl_xor_sync(0xffffffffu, v[t], off);
    if (lane == 0) {
      double* o = scatter + 6 * (size_t)k;
#pragma unroll
      for (int t = 0; t < 5; ++t) atomicAdd(o + t, v[t]);
      atomicAdd(o + 5, (double)__popc(b));
    }
  }
}

mh_status launch_inlier_stats(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, double* d_scatter) {
  if (K <= 0) return MH_OK;
  MH_CUDA(ctx, cudaMemsetAsync(d_scatter, 0, sizeof(double) * 6 * (size_t)K, ctx->stream));
  if (N <= 0) return MH_OK;
  const CostParams cp = cost_params(ctx);
  // hypotheses are processed in groups that fit shared memory
  const int group = 1024;
  for (int k0 = 0; k0 < K; k0 += group) {
    const int kn = std::min(group, K - k0);
    inlier_stats_kernel<<<(unsigned)((N + STATS_THREADS - 1) / STATS_THREADS), STATS_THREADS, sizeof(float) * 9 * kn,
                          ctx->stream>>>(d_pts, N, d_hyp + (size_t)k0 * 12, kn, d_scatter + 6 * (size_t)k0, cp);
    MH_LAUNCHED(ctx, "inlier_stats_kernel");
  }
  return MH_OK;
}

__global__ void inliers_of_kernel(const float4* __restrict__ pts, long long N, const float* __restrict__ hyp, int idx,
                                  int32_t* __restrict__ labels, CostParams cp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float h[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) h[k] = hyp[k];
  const float4 q = pts[i];
  if (residual(h, q.x, q.y, q.z, q.w) < cp.thr2) labels[i] = idx;
}

mh_status launch_inliers_of(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp_one, int idx, int32_t* d_labels) {
  if (N <= 0) return MH_OK;
  inliers_of_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_pts, N, d_hyp_one, idx, d_labels, cost_params(ctx));
  MH_LAUNCHED(ctx, "inliers_of_kernel");
  return MH_OK;
}

// ============================================================================
// FP64 members of the family, for the PRECISE path of mh_process (N x K small): the alternating optimisation is
// chaotic — one +-1 cost can flip a cut, move a refit and change the next merge — so the pipeline evaluates
// dataEnergy / the inlier scans exactly as the reference does (FP64, pixel coordinates).  Not a throughput path.
// ============================================================================
__device__ __forceinline__ double residual64(const double* __restrict__ h, double x, double y, double x2, double y2) {
  const double s1 = h[6] * x + h[7] * y + h[8];                      // MultiH.cpp:491-498, operation for operation
  const double x1 = (h[0] * x + h[1] * y + h[2]) / s1;
  const double y1 = (h[3] * x + h[4] * y + h[5]) / s1;
  const double dx = x1 - x2, dy = y1 - y2;
  return dx * dx + dy * dy;
}

__global__ void cost_dense64_kernel(const double* __restrict__ pts, long long N, const double* __restrict__ hyp, int K,
                                    int32_t* __restrict__ out, double lam, double T, int cost_outlier, int cost_far) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long L = K + 1;
  if (e >= N * L) return;
  const long long p = e / L;
  const int l = (int)(e % L);
  if (l == 0) { out[e] = cost_outlier; return; }
  const double* q = pts + 4 * p;
  const double d = residual64(hyp + 9 * (size_t)(l - 1), q[0], q[1], q[2], q[3]);
  out[e] = (d < T) ? (int)round(lam * (1.0 - d / T)) : cost_far;     // MultiH.cpp:501-503 (C round(): half away)
}

mh_status launch_cost_dense64(mh_ctx* ctx, const double* d_pts64, int64_t N, const double* d_hyp64, int K, int32_t* d_cost) {
  if (N <= 0) return MH_OK;
  const double thr2 = ctx->params.thr_homography * ctx->params.thr_homography, T = thr2 * 81.0 / 16.0;
  const double lam = 100.0 / ctx->params.lambda;
  const int c0 = (int)std::round(lam * T);
  const long long total = (long long)N * (K + 1);
  cost_dense64_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_pts64, N, d_hyp64, K, d_cost, lam, T, c0, 2 * c0);
  MH_LAUNCHED(ctx, "cost_dense64_kernel");
  return MH_OK;
}

// Sparse-with-default form of the same costs for the labelling step: per site the (label, cost) entries with d2 < T, in
// ascending label order, and their number; every label not listed costs cost_far, label 0 costs cost_outlier.  One warp per
// site, lanes over the hypotheses, ballot-compacted — the entries are the dense kernel's values bit for bit (same expression).
// entry = label << 16 | cost (the caller checks that both fit).  count may exceed kmax: the site's list is then incomplete.
constexpr int LIST64_WARPS = 8;
__global__ void __launch_bounds__(32 * LIST64_WARPS) cost_list64_kernel(const double* __restrict__ pts, long long N,
                                                                        const double* __restrict__ hyp, int K, int kmax,
                                                                        uint32_t* __restrict__ list, int32_t* __restrict__ count,
                                                                        double lam, double T) {
  const long long p = (long long)blockIdx.x * LIST64_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= N) return;
  const double* q = pts + 4 * p;
  const double x = q[0], y = q[1], x2 = q[2], y2 = q[3];
  int n = 0;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int k = k0 + lane;
    double d = T;
    if (k < K) d = residual64(hyp + 9 * (size_t)k, x, y, x2, y2);
    const bool in = k < K && d < T;
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (in) {
      const int slot = n + __popc(m & ((1u << lane) - 1u));
      if (slot < kmax) list[(size_t)p * kmax + slot] = ((uint32_t)(k + 1) << 16) | (uint32_t)(int)round(lam * (1.0 - d / T));
    }
    n += __popc(m);
  }
  if (lane == 0) count[p] = n;
}

mh_status launch_cost_list64(mh_ctx* ctx, const double* d_pts64, int64_t N, const double* d_hyp64, int K, int kmax, uint32_t* d_list,
                             int32_t* d_count) {
  if (N <= 0) return MH_OK;
  const double thr2 = ctx->params.thr_homography * ctx->params.thr_homography, T = thr2 * 81.0 / 16.0;
  const double lam = 100.0 / ctx->params.lambda;
  cost_list64_kernel<<<(unsigned)((N + LIST64_WARPS - 1) / LIST64_WARPS), 32 * LIST64_WARPS, 0, ctx->stream>>>(d_pts64, N, d_hyp64, K, kmax,
                                                                                                          d_list, d_count, lam, T);
  MH_LAUNCHED(ctx, "cost_list64_kernel");
  return MH_OK;
}

__global__ void __launch_bounds__(STATS_THREADS) inlier_stats64_kernel(const double* __restrict__ pts, long long N,
                                                                       const double* __restrict__ hyp, int K,
                                                                       double* __restrict__ scatter, double thr2) {
  const long long idx = (long long)blockIdx.x * STATS_THREADS + threadIdx.x;
  const bool live = idx < N;
  const double* q = pts + 4 * (live ? idx : N - 1);
  const double x = q[0], y = q[1], x2 = q[2], y2 = q[3];
  const int lane = threadIdx.x & 31;
  for (int k = 0; k < K; ++k) {
    const double h6 = hyp[9 * (size_t)k + 6], h7 = hyp[9 * (size_t)k + 7], h8 = hyp[9 * (size_t)k + 8];
    const double s = h6 * x + h7 * y + h8;                          // MultiH.cpp:434-441
    const double x1 = (hyp[9 * (size_t)k] * x + hyp[9 * (size_t)k + 1] * y + hyp[9 * (size_t)k + 2]) / s;
    const double y1 = (hyp[9 * (size_t)k + 3] * x + hyp[9 * (size_t)k + 4] * y + hyp[9 * (size_t)k + 5]) / s;
    const double dx = x2 - x1, dy = y2 - y1;
    const bool in = live && (dx * dx + dy * dy < thr2);
    const unsigned b = __ballot_sync(0xffffffffu, in);
    if (b == 0) continue;
    double v[5] = {in ? x * x : 0.0, in ? x * y : 0.0, in ? x : 0.0, in ? y * y : 0.0, in ? y : 0.0};
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int t = 0; t < 5; ++t) v[t] += __shfl_xor_sync(0xffffffffu, v[t], off);
    if (lane == 0) {
      double* o = scatter + 6 * (size_t)k;
#pragma unroll
      for (int t = 0; t < 5; ++t) atomicAdd(o + t, v[t]);
      atomicAdd(o + 5, (double)__popc(b));
    }
  }
}

mh_status launch_inlier_stats64(mh_ctx* ctx, const double* d_pts64, int64_t N, const double* d_hyp64, int K,
                                double* d_scatter) {
  if (K <= 0) return MH_OK;
  MH_CUDA(ctx, cudaMemsetAsync(d_scatter, 0, sizeof(double) * 6 * (size_t)K, ctx->stream));
  if (N <= 0) return MH_OK;
  const double thr2 = ctx->params.thr_homography * ctx->params.thr_homography;
  inlier_stats64_kernel<<<(unsigned)((N + STATS_THREADS - 1) / STATS_THREADS), STATS_THREADS, 0, ctx->stream>>>(
      d_pts64, N, d_hyp64, K, d_scatter, thr2);
  MH_LAUNCHED(ctx, "inlier_stats64_kernel");
  return MH_OK;
}

__global__ void inliers_of64_kernel(const double* __restrict__ pts, long long N, const double* __restrict__ h, int idx,
                                    int32_t* __restrict__ labels, double thr2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double* q = pts + 4 * i;
  if (residual64(h, q[0], q[1], q[2], q[3]) < thr2) labels[i] = idx;   // MultiH.cpp:753-763
}

mh_status launch_inliers_of64(mh_ctx* ctx, const double* d_pts64, int64_t N, const double* d_hyp64_one, int idx,
                              int32_t* d_labels) {
  if (N <= 0) return MH_OK;
  const double thr2 = ctx->params.thr_homography * ctx->params.thr_homography;
  inliers_of64_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_pts64, N, d_hyp64_one, idx, d_labels, thr2);
  MH_LAUNCHED(ctx, "inliers_of64_kernel");
  return MH_OK;
}

}  // namespace mh
