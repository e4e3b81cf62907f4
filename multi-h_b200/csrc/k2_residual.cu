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

namespace mh {

typedef unsigned long long u64;

// ---- packed f32x2 helpers --------------------------------------------------
__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// squared reprojection residual of one correspondence under one homography
// (MultiH.cpp:491-498), normalised units; one reciprocal, two FMAs for the divide.
__device__ __forceinline__ float residual(const float h[9], float x, float y, float x2, float y2) {
  const float s = fmaf(h[6], x, fmaf(h[7], y, h[8]));
  const float xn = fmaf(h[0], x, fmaf(h[1], y, h[2]));
  const float yn = fmaf(h[3], x, fmaf(h[4], y, h[5]));
  const float r = rcp_approx(s);
  const float dx = fmaf(xn, r, -x2);
  const float dy = fmaf(yn, r, -y2);
  return fmaf(dx, dx, dy * dy);
}

// integer data cost of an in-range residual: round(lam * (1 - d2/T)), C round()
// on a non-negative value == floor(v + 0.5) (MultiH.cpp:501-502).
__device__ __forceinline__ int cost_in_range(float d2, const CostParams& cp) {
  const float v = fmaf(-cp.lam * cp.inv_T, d2, cp.lam);
  return (int)floorf(fmaxf(v, 0.f) + 0.5f);
}
__device__ __forceinline__ int cost_of(float d2, const CostParams& cp) {
  return (d2 < cp.T) ? cost_in_range(d2, cp) : cp.cost_far;  // NaN compares false -> far, as in the reference
}

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
  float h[9];
  const bool valid = l < K;
  {
    const float4* hp = reinterpret_cast<const float4*>(hyp + (size_t)(valid ? l : 0) * 12);
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

mh_status launch_cost_dense(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, void* d_cost,
                            int elem_bytes) {
  if (N <= 0 || K < 0) return MH_OK;
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

__global__ void fused_init_kernel(long long N, int K, int32_t* list_count, u64* best, int32_t* inlier_count,
                                  u64 best_init) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    if (list_count) list_count[i] = 0;
    if (best) best[i] = best_init;
  }
  if (inlier_count && i < K) inlier_count[i] = 0;
}

int g_fused_variant = 1;  // 1 = packed FFMA2 (default), 0 = scalar FFMA (A/B evidence)

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
  FusedOut o{d_list, d_list_count, (u64*)d_best, d_inlier_count, kmax};
  dim3 grid(tiles, (unsigned)ksplit);
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

}  // namespace mh
