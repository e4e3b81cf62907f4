// Device helpers shared by K1 (k1_haf.cu) and K4 (k4_refit.cu): pixel-space HAF rows, register-resident
// 4x4 Jacobi eigen-solve, homography assembly and normalised<->pixel conversion.
#pragma once
#include "common.cuh"

namespace mh {

struct HafGeom {
  double F[9];     // pixel-space F (x2^T F x1 = 0)
  double ex, ey;   // pixel-space epipole of image 2
  double s1, t1x, t1y, s2, t2x, t2y;
};

inline HafGeom haf_geom(const mh_ctx* ctx) {
  HafGeom g;
  for (int i = 0; i < 9; ++i) g.F[i] = ctx->F_px[i];
  g.ex = ctx->e2_px[0]; g.ey = ctx->e2_px[1];
  g.s1 = ctx->gd.s1; g.t1x = ctx->gd.t1x; g.t1y = ctx->gd.t1y;
  g.s2 = ctx->gd.s2; g.t2x = ctx->gd.t2x; g.t2y = ctx->gd.t2y;
  return g;
}

// ---------------------------------------------------------------------------
// Register-resident symmetric 4x4 eigen-solve (cyclic Jacobi, FP64).  M holds the
// full symmetric matrix; on return the diagonal holds the eigenvalues and the
// columns of V the eigenvectors.  Fixed sweep count with a convergence test that
// is uniform work for all threads of a warp.
// ---------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void jacobi_sym(double (&M)[N][N], double (&V)[N][N]) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
#pragma unroll 1
  for (int sweep = 0; sweep < 24; ++sweep) {
    double off = 0.0, diag = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      diag += M[i][i] * M[i][i];
#pragma unroll
      for (int j = i + 1; j < N; ++j) off += M[i][j] * M[i][j];
    }
    if (!(off > 1e-34 * diag) || off <= 1e-300) break;
#pragma unroll
    for (int p = 0; p < N - 1; ++p)
#pragma unroll
      for (int q = p + 1; q < N; ++q) {
        const double apq = M[p][q];
        if (apq != 0.0) {
          const double theta = (M[q][q] - M[p][p]) / (2.0 * apq);
          const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double c = rsqrt(t * t + 1.0), s = t * c;
#pragma unroll
          for (int k = 0; k < N; ++k) {
            const double akp = M[k][p], akq = M[k][q];
            M[k][p] = c * akp - s * akq;
            M[k][q] = s * akp + c * akq;
          }
#pragma unroll
          for (int k = 0; k < N; ++k) {
            const double apk = M[p][k], aqk = M[q][k];
            M[p][k] = c * apk - s * aqk;
            M[q][k] = s * apk + c * aqk;
          }
#pragma unroll
          for (int k = 0; k < N; ++k) {
            const double vkp = V[k][p], vkq = V[k][q];
            V[k][p] = c * vkp - s * vkq;
            V[k][q] = s * vkp + c * vkq;
          }
        }
      }
  }
}

__device__ __forceinline__ void jacobi4(double (&M)[4][4], double (&V)[4][4]) { jacobi_sym<4>(M, V); }

// v = eigenvector of the smallest eigenvalue (EVec.row(3), MultiH.cpp:895)
__device__ __forceinline__ void smallest_eigvec4(double (&M)[4][4], double (&v)[4]) {
  double V[4][4];
  jacobi4(M, V);
  int best = 0;
  double lo = M[0][0];
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (M[i][i] < lo) { lo = M[i][i]; best = i; }
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = best == 0 ? V[k][0] : best == 1 ? V[k][1] : best == 2 ? V[k][2] : V[k][3];
}

// accumulate A_i^T A_i of one correspondence (rows of MultiH.cpp:859-887) into M (upper triangle used)
__device__ __forceinline__ void haf_accumulate(double x1, double y1, double x2, double y2, double a11, double a12,
                                               double a21, double a22, const HafGeom& g, double (&M)[4][4]) {
  const double ex = g.ex, ey = g.ey;
  double R[6][4];
  R[0][0] = a11 * x1 + x2 - ex; R[0][1] = a11 * y1;           R[0][2] = a11; R[0][3] = -g.F[3];
  R[1][0] = a12 * x1;           R[1][1] = a12 * y1 + x2 - ex; R[1][2] = a12; R[1][3] = -g.F[4];
  R[2][0] = a21 * x1 + y2 - ey; R[2][1] = a21 * y1;           R[2][2] = a21; R[2][3] = g.F[0];
  R[3][0] = a22 * x1;           R[3][1] = a22 * y1 + y2 - ey; R[3][2] = a22; R[3][3] = g.F[1];
  R[4][0] = ex * x1 - x2 * x1;  R[4][1] = ex * y1 - x2 * y1;  R[4][2] = ex - x2;
  R[4][3] = x1 * g.F[3] + y1 * g.F[4] + g.F[5];
  R[5][0] = ey * x1 - y2 * x1;  R[5][1] = ey * y1 - y2 * y1;  R[5][2] = ey - y2;
  R[5][3] = -(x1 * g.F[0] + y1 * g.F[1] + g.F[2]);
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = r; c < 4; ++c) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) s = fma(R[k][r], R[k][c], s);
      M[r][c] += s;
    }
}

// H_px from v (MultiH.cpp:899-909), then H' = T2 H_px T1^-1 scaled by `scale`, stored as 12 floats
__device__ __forceinline__ void haf_store(const double (&v)[4], const HafGeom& g, bool divide_h33, float* out,
                                          double* out64 = nullptr) {
  double H[9];
  H[6] = v[0]; H[7] = v[1]; H[8] = v[2];
  const double lam = v[3];
  H[3] = g.ey * H[6] - lam * g.F[0]; H[4] = g.ey * H[7] - lam * g.F[1]; H[5] = g.ey * H[8] - lam * g.F[2];
  H[0] = g.ex * H[6] + lam * g.F[3]; H[1] = g.ex * H[7] + lam * g.F[4]; H[2] = g.ex * H[8] + lam * g.F[5];
  double sc = 1.0;
  if (divide_h33) sc = 1.0 / H[8];  // MultiH.cpp:910
  else {
    // free scale: make the largest entry of the normalised matrix O(1) for FP32 storage
    double m = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) m = fmax(m, fabs(H[k]));
    sc = m > 0.0 ? 1.0 / m : 1.0;
  }
  if (out64) {  // precise path: the pixel-space FP64 homography exactly as the reference holds it
    const double s64 = divide_h33 ? sc : 1.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) out64[k] = H[k] * s64;
  }
  // G = H_px T1^-1 : T1^-1 = [1/s1 0 -t1x/s1; 0 1/s1 -t1y/s1; 0 0 1]
  const double is1 = 1.0 / g.s1;
  double G[9];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    G[3 * r + 0] = H[3 * r + 0] * is1;
    G[3 * r + 1] = H[3 * r + 1] * is1;
    G[3 * r + 2] = H[3 * r + 2] - (H[3 * r + 0] * g.t1x + H[3 * r + 1] * g.t1y) * is1;
  }
  // H' = T2 G : rows 0,1 scaled by s2 plus t2 * row 2
  float4* o = reinterpret_cast<float4*>(out);
  const double h0 = (g.s2 * G[0] + g.t2x * G[6]) * sc, h1 = (g.s2 * G[1] + g.t2x * G[7]) * sc,
               h2 = (g.s2 * G[2] + g.t2x * G[8]) * sc;
  const double h3 = (g.s2 * G[3] + g.t2y * G[6]) * sc, h4 = (g.s2 * G[4] + g.t2y * G[7]) * sc,
               h5 = (g.s2 * G[5] + g.t2y * G[8]) * sc;
  o[0] = make_float4((float)h0, (float)h1, (float)h2, (float)h3);
  o[1] = make_float4((float)h4, (float)h5, (float)(G[6] * sc), (float)(G[7] * sc));
  o[2] = make_float4((float)(G[8] * sc), 0.f, 0.f, 0.f);
}

// ---------------------------------------------------------------------------
// Feature vectors in PIXEL units from normalised FP32 hypotheses.
// H_px = T2^-1 H' T1; images of (0,0), (1,0), (0,1).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void hyp_to_pixel(const float* __restrict__ hp, const HafGeom& g, double (&H)[9]) {
  const float4* q = reinterpret_cast<const float4*>(hp);
  const float4 a = q[0], b = q[1], c = q[2];
  const double h[9] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x};
  // G = H' T1 : T1 = [s1 0 t1x; 0 s1 t1y; 0 0 1]
  double G[9];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    G[3 * r + 0] = h[3 * r + 0] * g.s1;
    G[3 * r + 1] = h[3 * r + 1] * g.s1;
    G[3 * r + 2] = h[3 * r + 0] * g.t1x + h[3 * r + 1] * g.t1y + h[3 * r + 2];
  }
  // H = T2^-1 G : T2^-1 = [1/s2 0 -t2x/s2; 0 1/s2 -t2y/s2; 0 0 1]
  const double is2 = 1.0 / g.s2;
#pragma unroll
  for (int c2 = 0; c2 < 3; ++c2) {
    H[c2] = (G[c2] - g.t2x * G[6 + c2]) * is2;
    H[3 + c2] = (G[3 + c2] - g.t2y * G[6 + c2]) * is2;
    H[6 + c2] = G[6 + c2];
  }
}

// ---- point / hypothesis sources: FP32 normalised device-space, or (precise path) raw FP64 pixels ------------------
__device__ __forceinline__ void load_point_px(const float4* __restrict__ pts, const double* __restrict__ pts64, long long i,
                                              const HafGeom& g, double& x1, double& y1, double& x2, double& y2) {
  if (pts64) {
    const double2* p = reinterpret_cast<const double2*>(pts64 + 4 * i);
    const double2 a = p[0], b = p[1];
    x1 = a.x; y1 = a.y; x2 = b.x; y2 = b.y;
  } else {
    const float4 p = pts[i];
    x1 = ((double)p.x - g.t1x) / g.s1; y1 = ((double)p.y - g.t1y) / g.s1;
    x2 = ((double)p.z - g.t2x) / g.s2; y2 = ((double)p.w - g.t2y) / g.s2;
  }
}
__device__ __forceinline__ void load_affine_px(const float4* __restrict__ aff, const double* __restrict__ aff64, long long i,
                                               const HafGeom& g, double& a11, double& a12, double& a21, double& a22) {
  if (aff64) {
    const double2* p = reinterpret_cast<const double2*>(aff64 + 4 * i);
    const double2 a = p[0], b = p[1];
    a11 = a.x; a12 = a.y; a21 = b.x; a22 = b.y;
  } else {
    const float4 a = aff[i];
    const double ra = g.s1 / g.s2;
    a11 = a.x * ra; a12 = a.y * ra; a21 = a.z * ra; a22 = a.w * ra;
  }
}
__device__ __forceinline__ void load_hyp_px(const float* __restrict__ hyp, const double* __restrict__ hyp64, long long i,
                                            const HafGeom& g, double (&H)[9]) {
  if (hyp64) {
#pragma unroll
    for (int k = 0; k < 9; ++k) H[k] = hyp64[9 * i + k];
  } else {
    hyp_to_pixel(hyp + 12 * i, g, H);
  }
}

}  // namespace mh
