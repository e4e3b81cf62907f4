// ============================================================================
// K1 — per-correspondence HAF homography hypotheses + feature vectors (sm_100a).
//
// Replaces MultiH::ComputeLocalHomographies (MultiH/MultiH/MultiH.cpp:696-717)
// -> MultiH::GetHomographyHAF (MultiH.cpp:850-911) and the feature builders of
// EstablishStablePointSets (MultiH.cpp:612-646) / MergingStep (:359-389).
//
// One thread per correspondence; the 6x4 system, its 4x4 normal matrix, the
// cyclic-Jacobi eigen-solve and the 3x3 assembly all live in registers.
// The least-squares problem is NOT invariant to coordinate normalisation (6
// equations, 3 dof), so to return the reference's estimate the solve is done in
// PIXEL coordinates in FP64 exactly as the reference does (A^T A, smallest
// eigenvector): FP32 on the pixel-space system is off by 0.14 px (p99) in the
// mean-shift features, see DESIGN.md.  B200 sustains ~37 TFLOP/s FP64, so the
// ~3 kflop solve costs ~0.3 ms at 4M correspondences.
// ============================================================================
#include "haf_device.cuh"

namespace mh {

// raw FP64 pixel correspondences -> normalised float4 (and A' = s2/s1 A)
__global__ void normalize_points_kernel(const double* __restrict__ pts_raw, const double* __restrict__ aff_raw,
                                        long long N, float4* __restrict__ pts, float4* __restrict__ aff, HafGeom g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (pts_raw && pts) {
    const double2* p = reinterpret_cast<const double2*>(pts_raw + 4 * i);
    const double2 a = p[0], b = p[1];
    pts[i] = make_float4((float)(a.x * g.s1 + g.t1x), (float)(a.y * g.s1 + g.t1y), (float)(b.x * g.s2 + g.t2x),
                         (float)(b.y * g.s2 + g.t2y));
  }
  if (aff_raw && aff) {
    const double2* q = reinterpret_cast<const double2*>(aff_raw + 4 * i);
    const double2 c = q[0], d = q[1];
    const double r = g.s2 / g.s1;
    aff[i] = make_float4((float)(c.x * r), (float)(c.y * r), (float)(d.x * r), (float)(d.y * r));
  }
}

mh_status launch_normalize_points(mh_ctx* ctx, const double* d_pts_raw, const double* d_aff_raw, int64_t N,
                                  float4* d_pts, float4* d_aff) {
  if (N <= 0) return MH_OK;
  normalize_points_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_pts_raw, d_aff_raw, N, d_pts, d_aff,
                                                                                haf_geom(ctx));
  MH_LAUNCHED(ctx, "normalize_points_kernel");
  return MH_OK;
}

__global__ void __launch_bounds__(128) haf_kernel(const float4* __restrict__ pts, const float4* __restrict__ aff,
                                                  const double* __restrict__ pts64, const double* __restrict__ aff64,
                                                  long long N, float* __restrict__ hyp, double* __restrict__ hyp64,
                                                  HafGeom g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double x1, y1, x2, y2, a11, a12, a21, a22;
  load_point_px(pts, pts64, i, g, x1, y1, x2, y2);
  load_affine_px(aff, aff64, i, g, a11, a12, a21, a22);
  double M[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) M[r][c] = 0.0;
  haf_accumulate(x1, y1, x2, y2, a11, a12, a21, a22, g, M);
#pragma unroll
  for (int r = 1; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < r; ++c) M[r][c] = M[c][r];
  double v[4];
  smallest_eigvec4(M, v);
  haf_store(v, g, true, hyp + 12 * i, hyp64 ? hyp64 + 9 * i : nullptr);
}

mh_status launch_haf(mh_ctx* ctx, const float4* d_pts, const float4* d_aff, int64_t N, float* d_hyp, int precision,
                     const double* d_pts64, const double* d_aff64, double* d_hyp64) {
  (void)precision;
  if (N <= 0) return MH_OK;
  haf_kernel<<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>(d_pts, d_aff, d_pts64, d_aff64, N, d_hyp, d_hyp64,
                                                                   haf_geom(ctx));
  MH_LAUNCHED(ctx, "haf_kernel");
  return MH_OK;
}

__global__ void features_kernel(const float* __restrict__ hyp, const double* __restrict__ hyp64,
                                const float4* __restrict__ pts, const double* __restrict__ pts64, long long N, int D,
                                double locality, double* __restrict__ feat, HafGeom g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double h[9];
  load_hyp_px(hyp, hyp64, i, g, h);
  const double s1 = h[8], x1 = h[2] / s1, y1 = h[5] / s1;
  const double s2 = h[6] + h[8], x2 = (h[0] + h[2]) / s2, y2 = (h[3] + h[5]) / s2;
  const double s3 = h[7] + h[8], x3 = (h[1] + h[2]) / s3, y3 = (h[4] + h[5]) / s3;
  double* f = feat + (size_t)D * i;
  if (D == 10) {  // MultiH.cpp:635-644
    double px1, py1, px2, py2;
    load_point_px(pts, pts64, i, g, px1, py1, px2, py2);
    f[0] = x1; f[1] = x2; f[2] = x3; f[3] = y1; f[4] = y2; f[5] = y3;
    f[6] = px1 * locality; f[7] = py1 * locality; f[8] = px2 * locality; f[9] = py2 * locality;
  } else {        // MultiH.cpp:382-387
    f[0] = x1; f[1] = y1; f[2] = x2; f[3] = y2; f[4] = x3; f[5] = y3;
  }
}

mh_status launch_features10(mh_ctx* ctx, const float* d_hyp, const float4* d_pts, int64_t N, double* d_feat,
                            const double* d_hyp64, const double* d_pts64) {
  if (N <= 0) return MH_OK;
  features_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_hyp, d_hyp64, d_pts, d_pts64, N, 10,
                                                                        ctx->params.locality, d_feat, haf_geom(ctx));
  MH_LAUNCHED(ctx, "features10_kernel");
  return MH_OK;
}

mh_status launch_features6(mh_ctx* ctx, const float* d_hyp, int K, double* d_feat, const double* d_hyp64) {
  if (K <= 0) return MH_OK;
  features_kernel<<<(unsigned)((K + 255) / 256), 256, 0, ctx->stream>>>(d_hyp, d_hyp64, nullptr, nullptr, K, 6, 0.0, d_feat,
                                                                        haf_geom(ctx));
  MH_LAUNCHED(ctx, "features6_kernel");
  return MH_OK;
}

}  // namespace mh
