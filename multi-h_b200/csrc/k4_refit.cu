// ============================================================================
// K4 — per-label refit (sm_100a).
//
// Replaces LabelingStep's serial per-label gather (MultiH/MultiH/MultiH.cpp:545-584)
// followed by GetHomographyHAFNonminimal (MultiH.cpp:913-990, linear solution)
// and EstablishStablePointSets' per-cluster GetHomography3PT (MultiH.cpp:664-688,
// :995-1055 + NormalizePoints, Homography_Refine3PTCallback.h:146-197).
//
// Pipeline (all on the context's stream, no host round trip):
//   1 label histogram   (warp-aggregated atomics)
//   2 exclusive scan    (one CTA)
//   3 scatter to CSR    (warp-aggregated cursors) — the device version of the
//                        reference's gather; also yields the member lists
//   4 HAF: segmented accumulation of SUM A_i^T A_i (10 uniques, FP64) with a
//          warp-shuffle segmented scan over the label-sorted order
//     3PT: one warp per cluster, three passes over its members (centroid, mean
//          distance, 3x3 normal equations) in FP64
//   5 batched register-resident eigen-solves (one thread / warp per label)
// The normal equations are accumulated in PIXEL coordinates in FP64, as the
// reference does, because the least-squares estimate is not invariant to
// normalisation (see k1_haf.cu).
// ============================================================================
#include "haf_device.cuh"

namespace mh {

// ---- 1..3: labels -> CSR ------------------------------------------------------
__global__ void label_count_kernel(const int32_t* __restrict__ labels, long long N, int K, int32_t* __restrict__ count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int l = (i < N) ? labels[i] : -1;
  const bool ok = l >= 0 && l < K;
  const unsigned act = __ballot_sync(0xffffffffu, ok);
  if (!ok) return;
  const unsigned peers = __match_any_sync(act, l);
  const int lane = threadIdx.x & 31;
  if (lane == __ffs(peers) - 1) atomicAdd(count + l, __popc(peers));
}

// offsets[0..K] = exclusive scan of count[0..K-1]; cursor[l] = offsets[l]
__global__ void scan_kernel(const int32_t* __restrict__ count, int K, int32_t* __restrict__ offsets,
                            int32_t* __restrict__ cursor) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < K; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = (i < K) ? count[i] : 0;
    int s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, d);
      if (lane >= d) s += t;
    }
    if (lane == 31) warp_sums[warp] = s;
    __syncthreads();
    if (warp == 0) {
      int w = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += t;
      }
      warp_sums[lane] = w;  // inclusive
    }
    __syncthreads();
    const int excl = carry + (warp ? warp_sums[warp - 1] : 0) + s - v;
    if (i < K) { offsets[i] = excl; cursor[i] = excl; }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[K] = carry;
}

__global__ void label_scatter_kernel(const int32_t* __restrict__ labels, long long N, int K, int32_t* __restrict__ cursor,
                                     int32_t* __restrict__ members, int32_t* __restrict__ sorted_label) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int l = (i < N) ? labels[i] : -1;
  const bool ok = l >= 0 && l < K;
  const unsigned act = __ballot_sync(0xffffffffu, ok);
  if (!ok) return;
  const unsigned peers = __match_any_sync(act, l);
  const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(cursor + l, __popc(peers));
  base = __shfl_sync(peers, base, leader);
  const int pos = base + __popc(peers & ((1u << lane) - 1u));
  members[pos] = (int32_t)i;
  if (sorted_label) sorted_label[pos] = l;
}

// Builds CSR (offsets[K+1], members[M], sorted_label[M]) in ctx->scratch; returns device pointers.
struct Csr {
  int32_t* count;
  int32_t* offsets;
  int32_t* cursor;
  int32_t* members;
  int32_t* sorted_label;
  double* acc;  // K x 12 doubles, zeroed
};

static mh_status build_csr(mh_ctx* ctx, const int32_t* d_labels, int64_t N, int K, Csr& c) {
  const uint64_t kpad = ((uint64_t)K + 64) & ~uint64_t(63);
  const uint64_t npad = ((uint64_t)N + 64) & ~uint64_t(63);
  const uint64_t bytes = sizeof(int32_t) * (3 * kpad + 2 * npad) + sizeof(double) * 12 * kpad;
  MH_TRY(ensure_scratch(ctx, bytes));
  char* p = (char*)ctx->scratch;
  c.acc = (double*)p; p += sizeof(double) * 12 * kpad;
  c.count = (int32_t*)p; p += sizeof(int32_t) * kpad;
  c.offsets = (int32_t*)p; p += sizeof(int32_t) * kpad;
  c.cursor = (int32_t*)p; p += sizeof(int32_t) * kpad;
  c.members = (int32_t*)p; p += sizeof(int32_t) * npad;
  c.sorted_label = (int32_t*)p;
  MH_CUDA(ctx, cudaMemsetAsync(c.acc, 0, sizeof(double) * 12 * kpad + sizeof(int32_t) * kpad, ctx->stream));  // acc + count
  if (N > 0) {
    label_count_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_labels, N, K, c.count);
    MH_LAUNCHED(ctx, "label_count_kernel");
  }
  scan_kernel<<<1, 1024, 0, ctx->stream>>>(c.count, K, c.offsets, c.cursor);
  MH_LAUNCHED(ctx, "scan_kernel");
  if (N > 0) {
    label_scatter_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_labels, N, K, c.cursor, c.members,
                                                                               c.sorted_label);
    MH_LAUNCHED(ctx, "label_scatter_kernel");
  }
  return MH_OK;
}

// ---- 4a: HAF normal equations, warp-shuffle segmented reduction -----------------
__global__ void __launch_bounds__(256) haf_segment_accumulate_kernel(const float4* __restrict__ pts,
                                                                     const float4* __restrict__ aff,
                                                                     const int32_t* __restrict__ members,
                                                                     const int32_t* __restrict__ sorted_label,
                                                                     const int32_t* __restrict__ offsets, int K,
                                                                     double* __restrict__ acc, HafGeom g,
                                                                     const double* __restrict__ pts64,
                                                                     const double* __restrict__ aff64) {
  const int M = offsets[K];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = j < M;
  const int l = live ? sorted_label[j] : -1;
  double v[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) v[k] = 0.0;
  if (live) {
    const int i = members[j];
    double x1, y1, x2, y2, a11, a12, a21, a22;
    load_point_px(pts, pts64, i, g, x1, y1, x2, y2);
    load_affine_px(aff, aff64, i, g, a11, a12, a21, a22);
    double Mx[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) Mx[r][c] = 0.0;
    haf_accumulate(x1, y1, x2, y2, a11, a12, a21, a22, g, Mx);
    int q = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = r; c < 4; ++c) v[q++] = Mx[r][c];
  }
  // segmented inclusive scan over the (sorted) labels of the warp
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int lo = __shfl_up_sync(0xffffffffu, l, d);
    const bool take = lane >= d && lo == l;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const double t = __shfl_up_sync(0xffffffffu, v[k], d);
      if (take) v[k] += t;
    }
  }
  const int lnext = __shfl_down_sync(0xffffffffu, l, 1);
  const bool tail = live && (lane == 31 || lnext != l);
  if (tail) {
    double* o = acc + 12 * (size_t)l;
#pragma unroll
    for (int k = 0; k < 10; ++k) atomicAdd(o + k, v[k]);
  }
}

// ---- 5a: batched 4x4 eigen-solves ------------------------------------------------
// acc layout per label: 10 uniques of SUM A^T A (row-major upper triangle), [10] = member count, [11] = pad — one
// (K,12) FP64 array so that a single all-reduce(sum) combines the partial statistics of correspondence shards.
__global__ void acc_count_kernel(const int32_t* __restrict__ count, int K, double* __restrict__ acc) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < K) { acc[12 * (size_t)l + 10] = (double)count[l]; acc[12 * (size_t)l + 11] = 0.0; }
}

__global__ void haf_solve_kernel(const double* __restrict__ acc, int K, float* __restrict__ hyp,
                                 int32_t* __restrict__ count_out, HafGeom g, double* __restrict__ hyp64) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= K) return;
  const double* a = acc + 12 * (size_t)l;
  const int n = (int)(a[10] + 0.5);
  if (count_out) count_out[l] = n;
  if (n == 0) return;  // MultiH.cpp:592-593: label without members keeps its homography
  double M[4][4];
  int q = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = r; c < 4; ++c) { M[r][c] = a[q]; M[c][r] = a[q]; ++q; }
  double v[4];
  smallest_eigvec4(M, v);
  haf_store(v, g, false, hyp + 12 * (size_t)l, hyp64 ? hyp64 + 9 * (size_t)l : nullptr);  // no division by h33 (MultiH.cpp:977-989)
}

// label_i = (best_i & 0xffffffff) - 1  (-1 = outlier): the argmin output of K2 as the label array K4 consumes
__global__ void labels_from_best_kernel(const unsigned long long* __restrict__ best, long long N, int32_t* __restrict__ labels) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) labels[i] = (int32_t)(best[i] & 0xffffffffull) - 1;
}
// acc[k][11] = inlier count of hypothesis k, so that ONE all-reduce carries the refit statistics and the inlier counts
__global__ void pack_inlier_counts_kernel(const int32_t* __restrict__ cnt, int K, double* __restrict__ acc) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) acc[12 * (size_t)k + 11] = (double)cnt[k];
}
__global__ void unpack_inlier_counts_kernel(const double* __restrict__ acc, int K, int32_t* __restrict__ cnt) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) cnt[k] = (int32_t)(acc[12 * (size_t)k + 11] + 0.5);
}

mh_status launch_labels_from_best(mh_ctx* ctx, const unsigned long long* d_best, int64_t N, int32_t* d_labels) {
  if (N <= 0) return MH_OK;
  labels_from_best_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_best, N, d_labels);
  MH_LAUNCHED(ctx, "labels_from_best_kernel");
  return MH_OK;
}
mh_status launch_pack_inlier_counts(mh_ctx* ctx, int32_t* d_cnt, int K, double* d_acc, int unpack) {
  if (K <= 0) return MH_OK;
  if (unpack) unpack_inlier_counts_kernel<<<(unsigned)((K + 255) / 256), 256, 0, ctx->stream>>>(d_acc, K, d_cnt);
  else pack_inlier_counts_kernel<<<(unsigned)((K + 255) / 256), 256, 0, ctx->stream>>>(d_cnt, K, d_acc);
  MH_LAUNCHED(ctx, "pack_inlier_counts_kernel");
  return MH_OK;
}

mh_status launch_refit_haf_accumulate(mh_ctx* ctx, const float4* d_pts, const float4* d_aff, const int32_t* d_labels,
                                      int64_t N, int K, double* d_acc, const double* d_pts64, const double* d_aff64) {
  if (K <= 0) return MH_OK;
  if (N > 0x7fffffff) return fail(ctx, MH_EINVAL, "mh_refit_haf: N must fit int32");
  Csr c;
  MH_TRY(build_csr(ctx, d_labels, N, K, c));
  MH_CUDA(ctx, cudaMemsetAsync(d_acc, 0, sizeof(double) * 12 * (size_t)K, ctx->stream));
  const HafGeom g = haf_geom(ctx);
  if (N > 0) {
    haf_segment_accumulate_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_pts, d_aff, c.members,
                                                                                        c.sorted_label, c.offsets, K, d_acc, g,
                                                                                        d_pts64, d_aff64);
    MH_LAUNCHED(ctx, "haf_segment_accumulate_kernel");
  }
  acc_count_kernel<<<(unsigned)((K + 255) / 256), 256, 0, ctx->stream>>>(c.count, K, d_acc);
  MH_LAUNCHED(ctx, "acc_count_kernel");
  return MH_OK;
}

mh_status launch_refit_haf_solve(mh_ctx* ctx, const double* d_acc, int K, float* d_hyp, int32_t* d_count, double* d_hyp64) {
  if (K <= 0) return MH_OK;
  haf_solve_kernel<<<(unsigned)((K + 63) / 64), 64, 0, ctx->stream>>>(d_acc, K, d_hyp, d_count, haf_geom(ctx), d_hyp64);
  MH_LAUNCHED(ctx, "haf_solve_kernel");
  return MH_OK;
}

mh_status launch_refit_haf(mh_ctx* ctx, const float4* d_pts, const float4* d_aff, const int32_t* d_labels, int64_t N,
                           int K, float* d_hyp, int32_t* d_count, const double* d_pts64, const double* d_aff64,
                           double* d_hyp64) {
  if (K <= 0) return MH_OK;
  // the statistics live behind the CSR arrays in the scratch arena (build_csr reserves K x 12 doubles at its head)
  Csr c;
  MH_TRY(build_csr(ctx, d_labels, N, K, c));
  double* acc = c.acc;
  const HafGeom g = haf_geom(ctx);
  if (N > 0) {
    haf_segment_accumulate_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_pts, d_aff, c.members,
                                                                                        c.sorted_label, c.offsets, K, acc, g,
                                                                                        d_pts64, d_aff64);
    MH_LAUNCHED(ctx, "haf_segment_accumulate_kernel");
  }
  acc_count_kernel<<<(unsigned)((K + 255) / 256), 256, 0, ctx->stream>>>(c.count, K, acc);
  MH_LAUNCHED(ctx, "acc_count_kernel");
  return launch_refit_haf_solve(ctx, acc, K, d_hyp, d_count, d_hyp64);
}

// ---- 3PT -------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void mat3_mul_d(const double* A, const double* B, double* C) {
  double T[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) T[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
#pragma unroll
  for (int i = 0; i < 9; ++i) C[i] = T[i];
}

// Everything of GetHomography3PT after the point sums: given the two similarity
// normalisations (s, mx, my per image) and the 3x3 normal equations, return the
// pixel-space H.  N (sym 6: 00 01 02 11 12 22) and r (3) are A^T A and A^T b in
// the per-cluster normalised coordinates.
struct Norm2 { double s1, mx1, my1, s2, mx2, my2; };

__device__ __forceinline__ void normalised_F_and_epipole(const HafGeom& g, const Norm2& n, double (&Fn)[9], double& ex,
                                                         double& ey) {
  // T1^-1 = [1/s 0 mx; 0 1/s my; 0 0 1] for T = [s 0 -mx s; 0 s -my s; 0 0 1]  (3PTcb.h:186-193)
  const double T1i[9] = {1.0 / n.s1, 0, n.mx1, 0, 1.0 / n.s1, n.my1, 0, 0, 1};
  const double T2it[9] = {1.0 / n.s2, 0, 0, 0, 1.0 / n.s2, 0, n.mx2, n.my2, 1};
  mat3_mul_d(T2it, g.F, Fn);
  mat3_mul_d(Fn, T1i, Fn);  // MultiH.cpp:1009
  double FFt[3][3], V[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) FFt[i][j] = Fn[i * 3] * Fn[j * 3] + Fn[i * 3 + 1] * Fn[j * 3 + 1] + Fn[i * 3 + 2] * Fn[j * 3 + 2];
  jacobi_sym<3>(FFt, V);  // MultiH.cpp:1013-1017: eigenvector of the smallest eigenvalue, / z
  int best = 0;
  double lo = FFt[0][0];
#pragma unroll
  for (int i = 1; i < 3; ++i)
    if (FFt[i][i] < lo) { lo = FFt[i][i]; best = i; }
  const double vx = best == 0 ? V[0][0] : best == 1 ? V[0][1] : V[0][2];
  const double vy = best == 0 ? V[1][0] : best == 1 ? V[1][1] : V[1][2];
  const double vz = best == 0 ? V[2][0] : best == 1 ? V[2][1] : V[2][2];
  ex = vx / vz;
  ey = vy / vz;
}

// h3 = pinv(A) b through the eigen-decomposition of A^T A: singular values w = sqrt(lambda), dropped when
// w <= 2*DBL_EPSILON*sum(w) (OpenCV SVBkSb threshold used by Mat::inv(DECOMP_SVD), MultiH.cpp:1038).
__device__ __forceinline__ void pinv_normal3(const double (&Nq)[6], const double (&r)[3], double (&h)[3]) {
  double A[3][3] = {{Nq[0], Nq[1], Nq[2]}, {Nq[1], Nq[3], Nq[4]}, {Nq[2], Nq[4], Nq[5]}}, V[3][3];
  jacobi_sym<3>(A, V);
  double w[3], sumw = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) { w[j] = sqrt(fmax(A[j][j], 0.0)); sumw += w[j]; }
  const double thr = 2.0 * 2.220446049250313e-16 * sumw;
  h[0] = h[1] = h[2] = 0.0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (w[j] <= thr) continue;
    const double coef = (V[0][j] * r[0] + V[1][j] * r[1] + V[2][j] * r[2]) / (w[j] * w[j]);
    h[0] += V[0][j] * coef; h[1] += V[1][j] * coef; h[2] += V[2][j] * coef;
  }
}

__device__ __forceinline__ void assemble_3pt_and_store(const double (&h3)[3], const double (&Fn)[9], double ex, double ey,
                                                       const Norm2& n, const HafGeom& g, float* out,
                                                       double* out64 = nullptr) {
  // MultiH.cpp:1040-1050 (lambda == 1), then H = T2^-1 Hn T1 (MultiH.cpp:1054)
  double Hn[9];
  Hn[6] = h3[0]; Hn[7] = h3[1]; Hn[8] = h3[2];
  Hn[3] = ey * Hn[6] - Fn[0]; Hn[4] = ey * Hn[7] - Fn[1]; Hn[5] = ey * Hn[8] - Fn[2];
  Hn[0] = ex * Hn[6] + Fn[3]; Hn[1] = ex * Hn[7] + Fn[4]; Hn[2] = ex * Hn[8] + Fn[5];
  const double T1[9] = {n.s1, 0, -n.mx1 * n.s1, 0, n.s1, -n.my1 * n.s1, 0, 0, 1};
  const double T2i[9] = {1.0 / n.s2, 0, n.mx2, 0, 1.0 / n.s2, n.my2, 0, 0, 1};
  double H[9];
  mat3_mul_d(T2i, Hn, H);
  mat3_mul_d(H, T1, H);
  if (out64) {
#pragma unroll
    for (int k = 0; k < 9; ++k) out64[k] = H[k];
  }
  // pixel H -> context-normalised FP32 (free scale): H' = Tc2 H Tc1^-1
  const double C2[9] = {g.s2, 0, g.t2x, 0, g.s2, g.t2y, 0, 0, 1};
  const double C1i[9] = {1.0 / g.s1, 0, -g.t1x / g.s1, 0, 1.0 / g.s1, -g.t1y / g.s1, 0, 0, 1};
  mat3_mul_d(C2, H, H);
  mat3_mul_d(H, C1i, H);
  double m = 0.0;
#pragma unroll
  for (int k = 0; k < 9; ++k) m = fmax(m, fabs(H[k]));
  const double sc = m > 0.0 ? 1.0 / m : 1.0;
  float4* o = reinterpret_cast<float4*>(out);
  o[0] = make_float4((float)(H[0] * sc), (float)(H[1] * sc), (float)(H[2] * sc), (float)(H[3] * sc));
  o[1] = make_float4((float)(H[4] * sc), (float)(H[5] * sc), (float)(H[6] * sc), (float)(H[7] * sc));
  o[2] = make_float4((float)(H[8] * sc), 0.f, 0.f, 0.f);
}

// accumulate the two 3PT rows of one normalised pair (MultiH.cpp:1026-1035)
__device__ __forceinline__ void rows_3pt(double x1, double y1, double x2, double y2, double ex, double ey,
                                         const double (&Fn)[9], double (&Nq)[6], double (&r)[3]) {
  const double a0 = ex * x1 - x2 * x1, a1 = ex * y1 - x2 * y1, a2 = ex - x2;
  const double b0 = ey * x1 - y2 * x1, b1 = ey * y1 - y2 * y1, b2 = ey - y2;
  const double ra = -(x1 * Fn[3] + y1 * Fn[4] + Fn[5]);
  const double rb = (x1 * Fn[0] + y1 * Fn[1] + Fn[2]);
  Nq[0] += a0 * a0 + b0 * b0; Nq[1] += a0 * a1 + b0 * b1; Nq[2] += a0 * a2 + b0 * b2;
  Nq[3] += a1 * a1 + b1 * b1; Nq[4] += a1 * a2 + b1 * b2; Nq[5] += a2 * a2 + b2 * b2;
  r[0] += a0 * ra + b0 * rb; r[1] += a1 * ra + b1 * rb; r[2] += a2 * ra + b2 * rb;
}

// ---- the reference's LM polish of a 3PT fit ----------------------------------------------------------------------------------
// RefineHomography3PT + Homography_Refine3PTCallback::compute (Homography_Refine3PTCallback.h:7-58, 81-143) driven by the
// reference's copy of cv::LMSolverImpl::run (Utilities.hpp:762-869; 1000 iterations, epsx = epsf = FLT_EPSILON), step for step:
// parameters (h31, h32, h33) of H in normalised coordinates, residuals (x2 - x', y2 - y'), the callback's own approximate
// Jacobian rows e_x s (x1, y1, 1) and e_y s (x1, y1, 1) — they ignore the derivative of the projective division and carry the
// opposite sign of d(err)/dh, so most trial steps are rejected; whatever the iteration accepts is what the reference keeps, and
// so does this.  (The HAF polish, RefineHomographyHAF, rebinds a local header at Homography_RefineHAFCallback.h:58 and never
// writes its result back — there is nothing to reproduce.)  WARP: the members of a cluster are spread over the lanes and every
// lane follows the same control flow on the warp-summed statistics; otherwise one thread walks all n points.
struct Lm3Stats { double S, rmax, A[6], v[3]; };
template <bool WARP, typename Fetch>
__device__ __forceinline__ void lm3pt_eval(int n, Fetch fetch, const double (&Fn)[9], double ex, double ey, const double (&h)[3],
                                           bool jac, int lane, Lm3Stats& o) {
  o.S = 0; o.rmax = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) o.A[k] = 0;
  o.v[0] = o.v[1] = o.v[2] = 0;
  for (int i = WARP ? lane : 0; i < n; i += WARP ? 32 : 1) {
    double x1, y1, x2, y2;
    fetch(i, x1, y1, x2, y2);
    double s = h[0] * x1 + h[1] * y1 + h[2];
    s = fabs(s) > 2.220446049250313e-16 ? 1. / s : 0;
    const double h21 = ey * h[0] - Fn[0], h22 = ey * h[1] - Fn[1], h23 = ey * h[2] - Fn[2];
    const double h11 = ex * h[0] + Fn[3], h12 = ex * h[1] + Fn[4], h13 = ex * h[2] + Fn[5];
    const double e0 = x2 - (h11 * x1 + h12 * y1 + h13) * s, e1 = y2 - (h21 * x1 + h22 * y1 + h23) * s;
    o.S += e0 * e0 + e1 * e1;
    o.rmax = fmax(o.rmax, fmax(fabs(e0), fabs(e1)));
    if (jac) {   // J^T J and J^T r, row by row as mulTransposed / gemm accumulate them
      const double j0[3] = {ex * s * x1, ex * s * y1, ex * s}, j1[3] = {ey * s * x1, ey * s * y1, ey * s};
      o.A[0] += j0[0] * j0[0]; o.A[1] += j0[0] * j0[1]; o.A[2] += j0[0] * j0[2];
      o.A[3] += j0[1] * j0[1]; o.A[4] += j0[1] * j0[2]; o.A[5] += j0[2] * j0[2];
      o.v[0] += j0[0] * e0; o.v[1] += j0[1] * e0; o.v[2] += j0[2] * e0;
      o.A[0] += j1[0] * j1[0]; o.A[1] += j1[0] * j1[1]; o.A[2] += j1[0] * j1[2];
      o.A[3] += j1[1] * j1[1]; o.A[4] += j1[1] * j1[2]; o.A[5] += j1[2] * j1[2];
      o.v[0] += j1[0] * e1; o.v[1] += j1[1] * e1; o.v[2] += j1[2] * e1;
    }
  }
  if (WARP) {
    o.S = warp_sum(o.S);
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) o.rmax = fmax(o.rmax, __shfl_xor_sync(0xffffffffu, o.rmax, of));
    if (jac) {
#pragma unroll
      for (int k = 0; k < 6; ++k) o.A[k] = warp_sum(o.A[k]);
#pragma unroll
      for (int k = 0; k < 3; ++k) o.v[k] = warp_sum(o.v[k]);
    }
  }
}
// pseudo-inverse of a symmetric 3x3 through its eigen-decomposition (cv::solve / cv::invert with DECOMP_EIG: eigenvalues
// <= 2 DBL_EPSILON sum |w| are dropped)
__device__ __forceinline__ void sym3_pinv(const double (&Au)[6], double (&inv)[3][3]) {
  double A[3][3] = {{Au[0], Au[1], Au[2]}, {Au[1], Au[3], Au[4]}, {Au[2], Au[4], Au[5]}}, V[3][3];
  jacobi_sym<3>(A, V);
  const double thr = 2.0 * 2.220446049250313e-16 * (fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]));
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) inv[i][j] = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (!(fabs(A[k][k]) > thr)) continue;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) inv[i][j] += V[i][k] * V[j][k] / A[k][k];
  }
}
template <bool WARP, typename Fetch>
__device__ __forceinline__ void lm3pt_refine(int n, Fetch fetch, const double (&Fn)[9], double ex, double ey, double (&x)[3], int lane) {
  const double eps = 1.1920928955078125e-07, deps = 2.220446049250313e-16, Rlo = 0.25, Rhi = 0.75;
  Lm3Stats cur, trial;
  lm3pt_eval<WARP>(n, fetch, Fn, ex, ey, x, true, lane, cur);
  const double D[3] = {cur.A[0], cur.A[3], cur.A[5]};
  double lambda = 1, lc = 0.75;
  for (int iter = 0;;) {
    const double Ap[6] = {cur.A[0] + lambda * D[0], cur.A[1], cur.A[2], cur.A[3] + lambda * D[1], cur.A[4], cur.A[5] + lambda * D[2]};
    double inv[3][3], d[3], xd[3];
    sym3_pinv(Ap, inv);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      d[i] = inv[i][0] * cur.v[0] + inv[i][1] * cur.v[1] + inv[i][2] * cur.v[2];
      xd[i] = x[i] - d[i];
    }
    lm3pt_eval<WARP>(n, fetch, Fn, ex, ey, xd, false, lane, trial);
    const double Ad[3] = {cur.A[0] * d[0] + cur.A[1] * d[1] + cur.A[2] * d[2], cur.A[1] * d[0] + cur.A[3] * d[1] + cur.A[4] * d[2],
                          cur.A[2] * d[0] + cur.A[4] * d[1] + cur.A[5] * d[2]};
    const double dS = d[0] * (2 * cur.v[0] - Ad[0]) + d[1] * (2 * cur.v[1] - Ad[1]) + d[2] * (2 * cur.v[2] - Ad[2]);
    const double R = (cur.S - trial.S) / (fabs(dS) > deps ? dS : 1);
    if (R > Rhi) {
      lambda *= 0.5;
      if (lambda < lc) lambda = 0;
    } else if (R < Rlo) {
      const double t = d[0] * cur.v[0] + d[1] * cur.v[1] + d[2] * cur.v[2];
      double nu = (trial.S - cur.S) / (fabs(t) > deps ? t : 1) + 2;
      nu = fmin(fmax(nu, 2.), 10.);
      if (lambda == 0) {
        sym3_pinv(cur.A, inv);
        const double maxval = fmax(deps, fmax(fabs(inv[0][0]), fmax(fabs(inv[1][1]), fabs(inv[2][2]))));
        lambda = lc = 1. / maxval;
        nu *= 0.5;
      }
      lambda *= nu;
    }
    if (trial.S < cur.S) {
      x[0] = xd[0]; x[1] = xd[1]; x[2] = xd[2];
      lm3pt_eval<WARP>(n, fetch, Fn, ex, ey, x, true, lane, cur);
    }
    ++iter;
    if (!(iter < 1000 && fmax(fabs(d[0]), fmax(fabs(d[1]), fabs(d[2]))) >= eps && cur.rmax >= eps)) break;
  }
}

// one warp per cluster
__global__ void __launch_bounds__(128) cluster_3pt_kernel(const float4* __restrict__ pts,
                                                          const int32_t* __restrict__ members,
                                                          const int32_t* __restrict__ offsets, int C,
                                                          float* __restrict__ hyp, int32_t* __restrict__ keep, HafGeom g,
                                                          const double* __restrict__ pts64, double* __restrict__ hyp64, int lm) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  const int beg = offsets[c], end = offsets[c + 1], n = end - beg;
  if (n < 3) {  // MultiH.cpp:667
    if (lane == 0) keep[c] = 0;
    return;
  }
  // pass 1: centroids (3PTcb.h:166-168)
  double sx1 = 0, sy1 = 0, sx2 = 0, sy2 = 0;
  for (int j = beg + lane; j < end; j += 32) {
    double px1, py1, px2, py2;
    load_point_px(pts, pts64, members[j], g, px1, py1, px2, py2);
    sx1 += px1; sy1 += py1; sx2 += px2; sy2 += py2;
  }
  Norm2 nm;
  nm.mx1 = warp_sum(sx1) / n; nm.my1 = warp_sum(sy1) / n; nm.mx2 = warp_sum(sx2) / n; nm.my2 = warp_sum(sy2) / n;
  // pass 2: mean distance to the centroid (3PTcb.h:171-179)
  double d1 = 0, d2 = 0;
  for (int j = beg + lane; j < end; j += 32) {
    double px1, py1, px2, py2;
    load_point_px(pts, pts64, members[j], g, px1, py1, px2, py2);
    const double x1 = px1 - nm.mx1, y1 = py1 - nm.my1, x2 = px2 - nm.mx2, y2 = py2 - nm.my2;
    d1 += sqrt(x1 * x1 + y1 * y1);
    d2 += sqrt(x2 * x2 + y2 * y2);
  }
  nm.s1 = sqrt(2.0) / (warp_sum(d1) / n);
  nm.s2 = sqrt(2.0) / (warp_sum(d2) / n);
  double Fn[9], ex, ey;
  normalised_F_and_epipole(g, nm, Fn, ex, ey);
  // pass 3: normal equations
  double Nq[6] = {0, 0, 0, 0, 0, 0}, r[3] = {0, 0, 0};
  for (int j = beg + lane; j < end; j += 32) {
    double px1, py1, px2, py2;
    load_point_px(pts, pts64, members[j], g, px1, py1, px2, py2);
    const double x1 = (px1 - nm.mx1) * nm.s1, y1 = (py1 - nm.my1) * nm.s1;
    const double x2 = (px2 - nm.mx2) * nm.s2, y2 = (py2 - nm.my2) * nm.s2;
    rows_3pt(x1, y1, x2, y2, ex, ey, Fn, Nq, r);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) Nq[k] = warp_sum(Nq[k]);
#pragma unroll
  for (int k = 0; k < 3; ++k) r[k] = warp_sum(r[k]);
  double h3[3];
  pinv_normal3(Nq, r, h3);   // every lane holds the same sums
  if (lm) {   // MultiH.cpp:1052-1053
    auto fetch = [&](int i, double& x1, double& y1, double& x2, double& y2) {
      double px1, py1, px2, py2;
      load_point_px(pts, pts64, members[beg + i], g, px1, py1, px2, py2);
      x1 = (px1 - nm.mx1) * nm.s1; y1 = (py1 - nm.my1) * nm.s1;
      x2 = (px2 - nm.mx2) * nm.s2; y2 = (py2 - nm.my2) * nm.s2;
    };
    lm3pt_refine<true>(n, fetch, Fn, ex, ey, h3, lane);
  }
  if (lane == 0) {
    assemble_3pt_and_store(h3, Fn, ex, ey, nm, g, hyp + 12 * (size_t)c, hyp64 ? hyp64 + 9 * (size_t)c : nullptr);
    keep[c] = 1;
  }
}

mh_status launch_refit_3pt(mh_ctx* ctx, const float4* d_pts, const int32_t* d_assign, int64_t N, int C, float* d_hyp,
                           int32_t* d_keep, const double* d_pts64, double* d_hyp64) {
  if (C <= 0) return MH_OK;
  if (N > 0x7fffffff) return fail(ctx, MH_EINVAL, "mh_refit_3pt: N must fit int32");
  Csr c;
  MH_TRY(build_csr(ctx, d_assign, N, C, c));
  const unsigned blocks = (unsigned)(((uint64_t)C * 32 + 127) / 128);
  cluster_3pt_kernel<<<blocks, 128, 0, ctx->stream>>>(d_pts, c.members, c.offsets, C, d_hyp, d_keep, haf_geom(ctx), d_pts64,
                                                      d_hyp64, ctx->params.lm_refine);
  MH_LAUNCHED(ctx, "cluster_3pt_kernel");
  return MH_OK;
}

// MergingStep: mode (6-D feature, pixel units) -> homography by 3PT on (0,0),(1,0),(0,1) (MultiH.cpp:408-427).
__global__ void modes_to_hyp_kernel(const double* __restrict__ modes, int C, float* __restrict__ hyp, HafGeom g,
                                    double* __restrict__ hyp64, int lm) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double* m = modes + 6 * (size_t)c;
  const double p1[3][2] = {{0, 0}, {1, 0}, {0, 1}};
  const double p2[3][2] = {{m[0], m[1]}, {m[2], m[3]}, {m[4], m[5]}};
  Norm2 nm;
  nm.mx1 = 1.0 / 3.0; nm.my1 = 1.0 / 3.0;
  nm.mx2 = (p2[0][0] + p2[1][0] + p2[2][0]) * (1.0 / 3.0);
  nm.my2 = (p2[0][1] + p2[1][1] + p2[2][1]) * (1.0 / 3.0);
  double d1 = 0, d2 = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    d1 += sqrt((p1[i][0] - nm.mx1) * (p1[i][0] - nm.mx1) + (p1[i][1] - nm.my1) * (p1[i][1] - nm.my1));
    d2 += sqrt((p2[i][0] - nm.mx2) * (p2[i][0] - nm.mx2) + (p2[i][1] - nm.my2) * (p2[i][1] - nm.my2));
  }
  nm.s1 = sqrt(2.0) / (d1 / 3.0);
  nm.s2 = sqrt(2.0) / (d2 / 3.0);
  double Fn[9], ex, ey;
  normalised_F_and_epipole(g, nm, Fn, ex, ey);
  double Nq[6] = {0, 0, 0, 0, 0, 0}, r[3] = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < 3; ++i)
    rows_3pt((p1[i][0] - nm.mx1) * nm.s1, (p1[i][1] - nm.my1) * nm.s1, (p2[i][0] - nm.mx2) * nm.s2,
             (p2[i][1] - nm.my2) * nm.s2, ex, ey, Fn, Nq, r);
  double h3[3];
  pinv_normal3(Nq, r, h3);
  if (lm) {   // MultiH.cpp:427 calls GetHomography3PT with its default do_numerical_refinement = true
    auto fetch = [&](int i, double& x1, double& y1, double& x2, double& y2) {
      x1 = (p1[i][0] - nm.mx1) * nm.s1; y1 = (p1[i][1] - nm.my1) * nm.s1;
      x2 = (p2[i][0] - nm.mx2) * nm.s2; y2 = (p2[i][1] - nm.my2) * nm.s2;
    };
    lm3pt_refine<false>(3, fetch, Fn, ex, ey, h3, 0);
  }
  assemble_3pt_and_store(h3, Fn, ex, ey, nm, g, hyp + 12 * (size_t)c, hyp64 ? hyp64 + 9 * (size_t)c : nullptr);
}

mh_status launch_modes_to_hyp(mh_ctx* ctx, const double* d_modes, int C, float* d_hyp, double* d_hyp64) {
  if (C <= 0) return MH_OK;
  modes_to_hyp_kernel<<<(unsigned)((C + 63) / 64), 64, 0, ctx->stream>>>(d_modes, C, d_hyp, haf_geom(ctx), d_hyp64, ctx->params.lm_refine);
  MH_LAUNCHED(ctx, "modes_to_hyp_kernel");
  return MH_OK;
}

// ---- HomographyCompatibilityCheck (MultiH.cpp:100-222), the data-parallel part ----------------------------------------
// One CTA per (tested cluster, trial): thread 0 fits GetHomography3PT (no refinement, :157) to the trial's three sampled
// members — the same device functions as modes_to_hyp_kernel, a general first point set — then all threads evaluate the squared
// transfer error of the cluster's other members (:162-175) into shared memory, sort it (bitonic, FP64) and write the order
// statistics the host needs to replay the reference's median with its three stale buffer entries (pipeline.cu): sorted values
// [max(0, m-3) .. min(n-1, m+1)] around the median index m = n/2 (5 slots, +inf when absent) and the three largest (-inf when
// absent).  The sampling itself — sequential rand() draws without replacement from an evolving point vector — is index
// bookkeeping and stays on the host.
constexpr int COMPAT_THREADS = 256;
__global__ void __launch_bounds__(COMPAT_THREADS)
compat_trial_kernel(const double* __restrict__ pts64, const int32_t* __restrict__ members, const int32_t* __restrict__ moff,
                    const int32_t* __restrict__ samples /*[T][trials][3]*/, int trials, int P /*pow2 >= max n*/, HafGeom g,
                    double* __restrict__ out /*[T][trials][8]*/) {
  extern __shared__ double compat_sd[];
  __shared__ double sH[9];
  const int c = blockIdx.y, t = blockIdx.x, tid = threadIdx.x;
  const int beg = moff[c], end = moff[c + 1], n = end - beg - 3;
  if (n > P) return;   // more members than fit the shared-memory sort: compat_trial_big_kernel's share
  const int32_t* sm = samples + ((size_t)c * trials + t) * 3;
  const int s0 = sm[0], s1 = sm[1], s2 = sm[2];
  if (tid == 0) {
    double p1[3][2], p2[3][2];
    const int si[3] = {s0, s1, s2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double* q = pts64 + 4 * (size_t)si[k];
      p1[k][0] = q[0]; p1[k][1] = q[1]; p2[k][0] = q[2]; p2[k][1] = q[3];
    }
    Norm2 nm;   // NormalizePoints (3PTcb.h:146-197) on the three pairs
    nm.mx1 = (p1[0][0] + p1[1][0] + p1[2][0]) / 3.0; nm.my1 = (p1[0][1] + p1[1][1] + p1[2][1]) / 3.0;
    nm.mx2 = (p2[0][0] + p2[1][0] + p2[2][0]) / 3.0; nm.my2 = (p2[0][1] + p2[1][1] + p2[2][1]) / 3.0;
    double d1 = 0, d2 = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      d1 += sqrt((p1[k][0] - nm.mx1) * (p1[k][0] - nm.mx1) + (p1[k][1] - nm.my1) * (p1[k][1] - nm.my1));
      d2 += sqrt((p2[k][0] - nm.mx2) * (p2[k][0] - nm.mx2) + (p2[k][1] - nm.my2) * (p2[k][1] - nm.my2));
    }
    nm.s1 = sqrt(2.0) / (d1 / 3.0);
    nm.s2 = sqrt(2.0) / (d2 / 3.0);
    double Fn[9], ex, ey;
    normalised_F_and_epipole(g, nm, Fn, ex, ey);
    double Nq[6] = {0, 0, 0, 0, 0, 0}, r[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 3; ++k)
      rows_3pt((p1[k][0] - nm.mx1) * nm.s1, (p1[k][1] - nm.my1) * nm.s1, (p2[k][0] - nm.mx2) * nm.s2,
               (p2[k][1] - nm.my2) * nm.s2, ex, ey, Fn, Nq, r);
    double h3[3], H[9];
    pinv_normal3(Nq, r, h3);
    __align__(16) float unused[12];
    assemble_3pt_and_store(h3, Fn, ex, ey, nm, g, unused, H);
#pragma unroll
    for (int k = 0; k < 9; ++k) sH[k] = H[k];
  }
  __syncthreads();
  double h[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) h[k] = sH[k];
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  // members are listed in ascending index order, so a member's slot among the non-sampled ones is its rank minus the number
  // of sampled indices below it
  for (int j = beg + tid; j < end; j += COMPAT_THREADS) {
    const int idx = members[j];
    if (idx == s0 || idx == s1 || idx == s2) continue;
    const int pos = (j - beg) - (s0 < idx) - (s1 < idx) - (s2 < idx);
    const double* q = pts64 + 4 * (size_t)idx;
    const double s = h[6] * q[0] + h[7] * q[1] + h[8];
    const double x1 = (h[0] * q[0] + h[1] * q[1] + h[2]) / s, y1 = (h[3] * q[0] + h[4] * q[1] + h[5]) / s;
    const double dx = q[2] - x1, dy = q[3] - y1;
    double d = dx * dx + dy * dy;
    if (!(d == d)) d = inf;   // NaN (a member on the homography's horizon line): sorts last
    compat_sd[pos] = d;
  }
  for (int j = n + tid; j < P; j += COMPAT_THREADS) compat_sd[j] = inf;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += COMPAT_THREADS) {
        const int l = i ^ j;
        if (l > i) {
          const double a = compat_sd[i], b = compat_sd[l];
          if ((a > b) == ((i & k) == 0)) { compat_sd[i] = b; compat_sd[l] = a; }
        }
      }
      __syncthreads();
    }
  if (tid == 0) {
    double* o = out + ((size_t)c * trials + t) * 8;
    const int m = n / 2, lo = max(0, m - 3), hi = min(n - 1, m + 1);
    for (int k = 0; k < 5; ++k) o[k] = (lo + k <= hi) ? compat_sd[lo + k] : inf;
    for (int k = 0; k < 3; ++k) o[5 + k] = (n - 1 - k >= 0) ? compat_sd[n - 1 - k] : -inf;
  }
}

// The same statistics for clusters too large to sort in shared memory (the reference has no size limit): nothing is stored —
// every pass recomputes the transfer errors.  The value at sorted index lo = n/2 - 3 is found by an 8-bit radix select over the
// errors' bit patterns (non-negative doubles order like unsigned integers; 8 passes), one more pass collects what else is
// needed: how many errors lie below / at that value, the five smallest above it and the three largest of all.
constexpr int COMPAT_SMEM_MAX_N = 16384;
__global__ void __launch_bounds__(COMPAT_THREADS)
compat_trial_big_kernel(const double* __restrict__ pts64, const int32_t* __restrict__ members, const int32_t* __restrict__ moff,
                        const int32_t* __restrict__ samples /*[T][trials][3]*/, int trials, HafGeom g,
                        double* __restrict__ out /*[T][trials][8]*/) {
  __shared__ double sH[9];
  __shared__ unsigned hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ unsigned s_rank;
  __shared__ unsigned long long s_small[COMPAT_THREADS][5], s_large[COMPAT_THREADS][3];
  __shared__ unsigned s_less, s_eq;
  const int c = blockIdx.y, t = blockIdx.x, tid = threadIdx.x;
  const int beg = moff[c], end = moff[c + 1], n = end - beg - 3;
  if (n <= COMPAT_SMEM_MAX_N) return;   // compat_trial_kernel's share
  const int32_t* sm = samples + ((size_t)c * trials + t) * 3;
  const int s0 = sm[0], s1 = sm[1], s2 = sm[2];
  if (tid == 0) {
    double p1[3][2], p2[3][2];
    const int si[3] = {s0, s1, s2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double* q = pts64 + 4 * (size_t)si[k];
      p1[k][0] = q[0]; p1[k][1] = q[1]; p2[k][0] = q[2]; p2[k][1] = q[3];
    }
    Norm2 nm;   // NormalizePoints (3PTcb.h:146-197) on the three pairs
    nm.mx1 = (p1[0][0] + p1[1][0] + p1[2][0]) / 3.0; nm.my1 = (p1[0][1] + p1[1][1] + p1[2][1]) / 3.0;
    nm.mx2 = (p2[0][0] + p2[1][0] + p2[2][0]) / 3.0; nm.my2 = (p2[0][1] + p2[1][1] + p2[2][1]) / 3.0;
    double d1 = 0, d2 = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      d1 += sqrt((p1[k][0] - nm.mx1) * (p1[k][0] - nm.mx1) + (p1[k][1] - nm.my1) * (p1[k][1] - nm.my1));
      d2 += sqrt((p2[k][0] - nm.mx2) * (p2[k][0] - nm.mx2) + (p2[k][1] - nm.my2) * (p2[k][1] - nm.my2));
    }
    nm.s1 = sqrt(2.0) / (d1 / 3.0);
    nm.s2 = sqrt(2.0) / (d2 / 3.0);
    double Fn[9], ex, ey;
    normalised_F_and_epipole(g, nm, Fn, ex, ey);
    double Nq[6] = {0, 0, 0, 0, 0, 0}, r[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 3; ++k)
      rows_3pt((p1[k][0] - nm.mx1) * nm.s1, (p1[k][1] - nm.my1) * nm.s1, (p2[k][0] - nm.mx2) * nm.s2,
               (p2[k][1] - nm.my2) * nm.s2, ex, ey, Fn, Nq, r);
    double h3[3], H[9];
    pinv_normal3(Nq, r, h3);
    __align__(16) float unused[12];
    assemble_3pt_and_store(h3, Fn, ex, ey, nm, g, unused, H);
#pragma unroll
    for (int k = 0; k < 9; ++k) sH[k] = H[k];
    s_prefix = 0ull;
    s_rank = (unsigned)max(0, n / 2 - 3);
  }
  __syncthreads();
  double h[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) h[k] = sH[k];
  const unsigned long long inf_bits = 0x7ff0000000000000ull;
  auto key_of = [&](int j, bool& live) {   // bit pattern of member j's squared transfer error (the same arithmetic as above)
    const int idx = members[j];
    live = !(idx == s0 || idx == s1 || idx == s2);
    const double* q = pts64 + 4 * (size_t)idx;
    const double s = h[6] * q[0] + h[7] * q[1] + h[8];
    const double x1 = (h[0] * q[0] + h[1] * q[1] + h[2]) / s, y1 = (h[3] * q[0] + h[4] * q[1] + h[5]) / s;
    const double dx = q[2] - x1, dy = q[3] - y1;
    const double d = dx * dx + dy * dy;
    return d == d ? (unsigned long long)__double_as_longlong(d) : inf_bits;   // NaN sorts last, as +inf
  };
  for (int shift = 56; shift >= 0; shift -= 8) {
    hist[tid] = 0;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    for (int j = beg + tid; j < end; j += COMPAT_THREADS) {
      bool live;
      const unsigned long long k = key_of(j, live);
      if (live && (shift == 56 || (k >> (shift + 8)) == (prefix >> (shift + 8)))) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned r = s_rank, b = 0;
      while (r >= hist[b]) { r -= hist[b]; ++b; }
      s_rank = r;
      s_prefix = prefix | ((unsigned long long)b << shift);
    }
    __syncthreads();
  }
  const unsigned long long klo = s_prefix;   // the error at sorted index max(0, n/2 - 3)
  unsigned long long small[5] = {~0ull, ~0ull, ~0ull, ~0ull, ~0ull}, large[3] = {0ull, 0ull, 0ull};
  unsigned less = 0, eq = 0, nlarge = 0;
  for (int j = beg + tid; j < end; j += COMPAT_THREADS) {
    bool live;
    const unsigned long long k = key_of(j, live);
    if (!live) continue;
    less += k < klo;
    eq += k == klo;
    if (k > klo && k < small[4]) {   // keep the five smallest above klo, ascending, duplicates included
      small[4] = k;
#pragma unroll
      for (int a = 4; a > 0; --a)
        if (small[a] < small[a - 1]) { const unsigned long long x = small[a]; small[a] = small[a - 1]; small[a - 1] = x; }
    }
    if (nlarge < 3 || k > large[2]) {   // and the three largest of all, descending
      large[2] = k;
      nlarge = min(nlarge + 1, 3u);
#pragma unroll
      for (int a = 2; a > 0; --a)
        if (large[a] > large[a - 1]) { const unsigned long long x = large[a]; large[a] = large[a - 1]; large[a - 1] = x; }
    }
  }
  if (tid == 0) { s_less = 0; s_eq = 0; }
  __syncthreads();
  atomicAdd(&s_less, less);
  atomicAdd(&s_eq, eq);
#pragma unroll
  for (int a = 0; a < 5; ++a) s_small[tid][a] = small[a];
#pragma unroll
  for (int a = 0; a < 3; ++a) s_large[tid][a] = large[a];
  __syncthreads();
  if (tid == 0) {
    unsigned long long S5[5] = {~0ull, ~0ull, ~0ull, ~0ull, ~0ull}, L3[3] = {0ull, 0ull, 0ull};
    for (int w = 0; w < COMPAT_THREADS; ++w) {
      for (int a = 0; a < 5; ++a) {
        const unsigned long long k = s_small[w][a];
        if (k < S5[4]) {
          S5[4] = k;
          for (int e = 4; e > 0; --e)
            if (S5[e] < S5[e - 1]) { const unsigned long long x = S5[e]; S5[e] = S5[e - 1]; S5[e - 1] = x; }
        }
      }
      for (int a = 0; a < 3; ++a) {   // every thread saw >= 3 live members (n > COMPAT_SMEM_MAX_N), so all slots are real
        const unsigned long long k = s_large[w][a];
        if (k > L3[2]) {
          L3[2] = k;
          for (int e = 2; e > 0; --e)
            if (L3[e] > L3[e - 1]) { const unsigned long long x = L3[e]; L3[e] = L3[e - 1]; L3[e - 1] = x; }
        }
      }
    }
    double* o = out + ((size_t)c * trials + t) * 8;
    const int m = n / 2, lo = max(0, m - 3), hi = min(n - 1, m + 1);
    for (int k = 0; k < 5; ++k) {
      const unsigned rank = (unsigned)(lo + k);
      unsigned long long v = inf_bits;
      if (lo + k <= hi) v = rank < s_less + s_eq ? klo : S5[rank - s_less - s_eq];
      o[k] = __longlong_as_double((long long)v);
    }
    for (int k = 0; k < 3; ++k) o[5 + k] = __longlong_as_double((long long)L3[k]);
  }
}

mh_status launch_compat_trials(mh_ctx* ctx, const double* d_pts64, const int32_t* d_members, const int32_t* d_moff,
                               const int32_t* d_samples, int T, int trials, int max_n, double* d_out) {
  if (T <= 0) return MH_OK;
  int P = 2;
  while (P < std::min(max_n, COMPAT_SMEM_MAX_N)) P <<= 1;
  const size_t smem = (size_t)P * sizeof(double);
  MH_CUDA(ctx, mh_allow_max_smem(compat_trial_kernel));
  compat_trial_kernel<<<dim3((unsigned)trials, (unsigned)T), COMPAT_THREADS, smem, ctx->stream>>>(
      d_pts64, d_members, d_moff, d_samples, trials, P, haf_geom(ctx), d_out);
  MH_LAUNCHED(ctx, "compat_trial_kernel");
  if (max_n > COMPAT_SMEM_MAX_N) {
    compat_trial_big_kernel<<<dim3((unsigned)trials, (unsigned)T), COMPAT_THREADS, 0, ctx->stream>>>(
        d_pts64, d_members, d_moff, d_samples, trials, haf_geom(ctx), d_out);
    MH_LAUNCHED(ctx, "compat_trial_big_kernel");
  }
  return MH_OK;
}

}  // namespace mh
