// ============================================================================
// K0 — per-correspondence pre-filter (the step directly
// before K1): the refinement loop of MultiH::GetFundamentalMatrixAndRefineData
// (MultiH/MultiH/MultiH.cpp:786-838) with F given —
//   OptimalTriangulation (:1116-1188): Hartley-Sturm correction of the point pair; degree-6 polynomial, roots by
//     Durand-Kerner in registers (restating cv::solvePoly: start values (1+i)^k, leading coefficients <= DBL_EPSILON
//     trimmed), real-root test |imag| <= 1e-10, cost compared with the asymptotic value;
//   GetAffineConsistency (:1057-1090) + GetBetaScale (:1092-1114): drop when distanceError > 1 (:826);
//   GetOptimalAffineTransformation (:1190-1223): the 6x6 KKT system solved in closed form (G^T G = beta^2 I).
// One thread per correspondence, FP64 (B200's FP64 pipe), then an ORDER-PRESERVING compaction (block counts -> scan ->
// scatter), because the reference push_back()s the survivors in input order.  The serial CPU loop becomes ~25 kflop of
// independent FP64 work per correspondence.
// ============================================================================
#include "common.cuh"

namespace mh {

struct PreGeom {
  double F[9];
  double R1[9], R2[9];
};

struct Cx { double re, im; };
__device__ __forceinline__ Cx cmul(Cx a, Cx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ Cx cdiv(Cx a, Cx b) {
  const double d = 1.0 / (b.re * b.re + b.im * b.im);
  return {(a.re * b.re + a.im * b.im) * d, (a.im * b.re - a.re * b.im) * d};
}
__device__ __forceinline__ void m3mul(const double* A, const double* B, double* C) {
  double T[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) T[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
#pragma unroll
  for (int i = 0; i < 9; ++i) C[i] = T[i];
}
__device__ __forceinline__ void m3inv(const double* m, double* o) {
  const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g, det = 1.0 / (a * A + b * B + c * C);
  o[0] = A * det; o[1] = -(b * i - c * h) * det; o[2] = (b * f - c * e) * det;
  o[3] = B * det; o[4] = (a * i - c * g) * det; o[5] = -(a * f - c * d) * det;
  o[6] = C * det; o[7] = -(a * h - b * g) * det; o[8] = (a * e - b * d) * det;
}

// Durand-Kerner on a real polynomial of degree <= 6 (coefficients ascending); returns the degree actually solved
__device__ __forceinline__ int solve_poly_dk(const double (&c)[7], Cx (&roots)[6]) {
  int n = 6;
  for (; n > 1; --n)
    if (fabs(c[n]) > 2.220446049250313e-16) break;
  Cx p{1, 0};
  const Cx r{1, 1};
#pragma unroll
  for (int i = 0; i < 6; ++i) { roots[i] = p; p = cmul(p, r); }
  for (int iter = 0; iter < 300 * n; ++iter) {
    double maxDiff = 0, maxRoot = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (i < n) {
        p = roots[i];
        Cx num{c[n], 0}, den{c[n], 0};
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          if (j < n) {
            num = cmul(num, p);
            num.re += c[n - j - 1];
            if (j != i) {
              const Cx d{p.re - roots[j].re, p.im - roots[j].im};
              if (d.re != 0 || d.im != 0) den = cmul(den, d);
            }
          }
        }
        num = cdiv(num, den);
        roots[i] = {p.re - num.re, p.im - num.im};
        maxDiff = fmax(maxDiff, hypot(num.re, num.im));
        maxRoot = fmax(maxRoot, hypot(roots[i].re, roots[i].im));
      }
    }
    if (maxDiff <= 1e-16 * fmax(maxRoot, 1e-300)) break;
  }
  return n;
}

__device__ __forceinline__ double beta_scale(const double* F, const double* p1, const double* p2) {  // MultiH.cpp:1092-1114
  const double l1x = F[0] * p2[0] + F[3] * p2[1] + F[6], l1y = F[1] * p2[0] + F[4] * p2[1] + F[7],
               l1z = F[2] * p2[0] + F[5] * p2[1] + F[8];
  const double l2x = F[0] * p1[0] + F[1] * p1[1] + F[2], l2y = F[3] * p1[0] + F[4] * p1[1] + F[5];
  const double xn = p1[0] + 1.0, yn = -(l1x * xn + l1z) / l1y;
  double dx = xn - p1[0], dy = yn - p1[1];
  const double nn = sqrt(dx * dx + dy * dy);
  dx /= nn; dy /= nn;
  return fabs(sqrt(l2x * l2x + l2y * l2y) /
              ((-F[0] * dy + F[1] * dx) * p2[0] + (-F[3] * dy + F[4] * dx) * p2[1] - F[6] * dy + F[7] * dx));
}

__global__ void __launch_bounds__(128) prefilter_kernel(const double* __restrict__ pts, const double* __restrict__ aff,
                                                        long long N, double* __restrict__ out_pts,
                                                        double* __restrict__ out_aff, int32_t* __restrict__ keep, PreGeom g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  keep[i] = 0;
  const double2* pp = reinterpret_cast<const double2*>(pts + 4 * i);
  const double2 pa = pp[0], pb = pp[1];
  const double p1[3] = {pa.x, pa.y, 1.0}, p2[3] = {pb.x, pb.y, 1.0};
  const double* F = g.F;
  // ---- OptimalTriangulation (MultiH.cpp:1116-1188)
  const double T1i[9] = {1, 0, p1[0], 0, 1, p1[1], 0, 0, 1};    // T1.inv()
  const double T2ti[9] = {1, 0, 0, 0, 1, 0, p2[0], p2[1], 1};   // T2.t().inv()
  double F2[9], F3[9], R1t[9];
  m3mul(T2ti, F, F2); m3mul(F2, T1i, F2);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) R1t[r * 3 + c] = g.R1[c * 3 + r];
  m3mul(g.R2, F2, F3); m3mul(F3, R1t, F3);
  const double a = F3[4], b = F3[5], c = F3[7], d = F3[8];
  const double f1 = 1.0, f2 = 1.0;  // epipoles are divided by z (MultiH.cpp:793, 799)
  const double f14 = f1 * f1 * f1 * f1, adbc = a * d - b * c;
  double co[7];
  co[6] = -a * c * f14 * adbc;
  co[5] = (a * a + f2 * f2 * c * c) * (a * a + f2 * f2 * c * c) - (a * d + b * c) * f14 * adbc;
  co[4] = 2 * (a * a + f2 * f2 * c * c) * (2 * a * b + 2 * c * d * f2 * f2) - d * b * f14 * adbc - 2 * a * c * f1 * f1 * adbc;
  co[3] = (2 * a * b + 2 * c * d * f2 * f2) * (2 * a * b + 2 * c * d * f2 * f2) +
          2 * (a * a + f2 * f2 * c * c) * (b * b + f2 * f2 * d * d) - 2 * f1 * f1 * adbc * (a * d + b * c);
  co[2] = 2 * (2 * a * b + 2 * c * d * f2 * f2) * (b * b + f2 * f2 * d * d) - 2 * (f1 * f1 * a * d - f1 * f1 * b * c) * b * d -
          a * c * adbc;
  co[1] = (b * b + f2 * f2 * d * d) * (b * b + f2 * f2 * d * d) - (a * d + b * c) * adbc;
  co[0] = -adbc * b * d;
  Cx roots[6];
  const int n = solve_poly_dk(co, roots);
  double bestS = 2147483647.0, bestT = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if (k < n && fabs(roots[k].im) <= 1e-10) {
      const double t = roots[k].re;
      const double val = t * t / (1 + f1 * f1 * t * t) +
                         ((c * t + d) * (c * t + d)) / ((a * t + b) * (a * t + b) + f2 * f2 * ((c * t + d) * (c * t + d)));
      if (val < bestS) { bestS = val; bestT = t; }
    }
  const double valInf = 1 / (f1 * f1) + (c * c) / (a * a + f2 * f2 * c * c);
  if (valInf < bestS) return;  // MultiH.cpp:1170-1175 -> error -> not pushed (:816-817)
  const double q1[3] = {0, bestT, 1};
  const double l0 = F3[0] * q1[0] + F3[1] * q1[1] + F3[2], l1 = F3[3] * q1[0] + F3[4] * q1[1] + F3[5],
               l2 = F3[6] * q1[0] + F3[7] * q1[1] + F3[8];
  double q2[3] = {-l0 * l2, -l1 * l2, l0 * l0 + l1 * l1};
  q2[0] /= q2[2]; q2[1] /= q2[2]; q2[2] = 1.0;
  const double T1[9] = {1, 0, -p1[0], 0, 1, -p1[1], 0, 0, 1}, T2[9] = {1, 0, -p2[0], 0, 1, -p2[1], 0, 0, 1};
  double M1[9], M2[9], M1i[9], M2i[9];
  m3mul(g.R1, T1, M1); m3mul(g.R2, T2, M2); m3inv(M1, M1i); m3inv(M2, M2i);
  const double cpt[3] = {M1i[0] * q1[0] + M1i[1] * q1[1] + M1i[2], M1i[3] * q1[0] + M1i[4] * q1[1] + M1i[5], 1.0};
  const double dpt[3] = {M2i[0] * q2[0] + M2i[1] * q2[1] + M2i[2], M2i[3] * q2[0] + M2i[4] * q2[1] + M2i[5], 1.0};
  // ---- GetAffineConsistency (MultiH.cpp:1057-1090): distanceError
  const double2* ap = reinterpret_cast<const double2*>(aff + 4 * i);
  const double2 a0 = ap[0], a1 = ap[1];
  const double A[4] = {a0.x, a0.y, a1.x, a1.y};
  const double l1v[3] = {F[0] * dpt[0] + F[3] * dpt[1] + F[6], F[1] * dpt[0] + F[4] * dpt[1] + F[7], F[2] * dpt[0] + F[5] * dpt[1] + F[8]};
  const double l2v[3] = {F[0] * cpt[0] + F[1] * cpt[1] + F[2], F[3] * cpt[0] + F[4] * cpt[1] + F[5], F[6] * cpt[0] + F[7] * cpt[1] + F[8]};
  double n1[2] = {l1v[0] / l1v[2], l1v[1] / l1v[2]}, n2[2] = {l2v[0] / l2v[2], l2v[1] / l2v[2]};
  double nn = sqrt(n1[0] * n1[0] + n1[1] * n1[1]); n1[0] /= nn; n1[1] /= nn;
  nn = sqrt(n2[0] * n2[0] + n2[1] * n2[1]); n2[0] /= nn; n2[1] /= nn;
  const double beta = beta_scale(F, cpt, dpt);
  const double det = A[0] * A[3] - A[1] * A[2];
  const double r1x = (A[3] * n1[0] - A[2] * n1[1]) / det, r1y = (-A[1] * n1[0] + A[0] * n1[1]) / det;  // A^-T n1
  const double ex = r1x - beta * n2[0], ey = r1y - beta * n2[1];
  if (sqrt(ex * ex + ey * ey) > 1.0) return;  // MultiH.cpp:826
  // ---- GetOptimalAffineTransformation (MultiH.cpp:1190-1223)
  if (n1[0] * n2[0] + n1[1] * n2[1] < 0) { n2[0] = -n2[0]; n2[1] = -n2[1]; }
  const double bx = beta * n2[0], by = beta * n2[1];
  const double gta0 = -bx * A[0] - by * A[2], gta1 = -bx * A[1] - by * A[3], s2 = bx * bx + by * by;
  const double mu0 = (gta0 + n1[0]) / s2, mu1 = (gta1 + n1[1]) / s2;
  double2* op = reinterpret_cast<double2*>(out_pts + 4 * i);
  double2* oa = reinterpret_cast<double2*>(out_aff + 4 * i);
  op[0] = make_double2(cpt[0], cpt[1]); op[1] = make_double2(dpt[0], dpt[1]);
  oa[0] = make_double2(A[0] + bx * mu0, A[1] + bx * mu1); oa[1] = make_double2(A[2] + by * mu0, A[3] + by * mu1);
  keep[i] = 1;
}

// ---- order-preserving compaction ------------------------------------------------------------------------------------
constexpr int CB = 1024;
__global__ void __launch_bounds__(CB) block_count_kernel(const int32_t* __restrict__ keep, long long N, int32_t* __restrict__ counts) {
  const long long i = (long long)blockIdx.x * CB + threadIdx.x;
  const int c = __syncthreads_count(i < N && keep[i]);
  if (threadIdx.x == 0) counts[blockIdx.x] = c;
}
__global__ void __launch_bounds__(CB) compact_kernel(const int32_t* __restrict__ keep, long long N,
                                                     const int32_t* __restrict__ block_off, const double* __restrict__ in_pts,
                                                     const double* __restrict__ in_aff, double* __restrict__ out_pts,
                                                     double* __restrict__ out_aff) {
  __shared__ int warp_off[CB / 32];
  const long long i = (long long)blockIdx.x * CB + threadIdx.x;
  const bool k = i < N && keep[i];
  const unsigned b = __ballot_sync(0xffffffffu, k);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_off[warp] = __popc(b);
  __syncthreads();
  if (warp == 0) {
    int v = warp_off[lane], s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, d); if (lane >= d) s += t; }
    warp_off[lane] = s - v;
  }
  __syncthreads();
  if (!k) return;
  const long long pos = (long long)block_off[blockIdx.x] + warp_off[warp] + __popc(b & ((1u << lane) - 1u));
  const double2* sp = reinterpret_cast<const double2*>(in_pts + 4 * i);
  const double2* sa = reinterpret_cast<const double2*>(in_aff + 4 * i);
  double2* dp = reinterpret_cast<double2*>(out_pts + 4 * pos);
  double2* da = reinterpret_cast<double2*>(out_aff + 4 * pos);
  dp[0] = sp[0]; dp[1] = sp[1]; da[0] = sa[0]; da[1] = sa[1];
}
__global__ void scan_small_kernel(const int32_t* __restrict__ count, int n, int32_t* __restrict__ offsets) {
  // exclusive scan by one CTA (n = number of 1024-blocks: 4096 for 4M correspondences); offsets[n] = total
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = (i < n) ? count[i] : 0;
    int s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, d); if (lane >= d) s += t; }
    if (lane == 31) warp_sums[warp] = s;
    __syncthreads();
    if (warp == 0) {
      int w = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += t; }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const int excl = carry + (warp ? warp_sums[warp - 1] : 0) + s - v;
    if (i < n) offsets[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = carry;
}

// d_pts64 / d_aff64: N x 4 FP64 pixels (device).  Outputs compacted in input order; *M_out = survivors.
mh_status launch_prefilter(mh_ctx* ctx, const double* d_pts64, const double* d_aff64, const double F[9], int64_t N,
                           double* d_pts_out, double* d_aff_out, int32_t* d_keep, int64_t* M_out) {
  *M_out = 0;
  if (N <= 0) return MH_OK;
  PreGeom g;
  double e1[2], e2[2], Ft[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Ft[i * 3 + j] = F[j * 3 + i];
  epipole2_host(F, e2);    // MultiH.cpp:789-793
  epipole2_host(Ft, e1);   // MultiH.cpp:795-799
  for (int i = 0; i < 9; ++i) g.F[i] = F[i];
  const double R1[9] = {e1[0], e1[1], 0, -e1[1], e1[0], 0, 0, 0, 1};    // MultiH.cpp:801
  const double R2[9] = {-e2[0], -e2[1], 0, e2[1], -e2[0], 0, 0, 0, 1};  // MultiH.cpp:802
  for (int i = 0; i < 9; ++i) { g.R1[i] = R1[i]; g.R2[i] = R2[i]; }
  const int nblk = (int)((N + CB - 1) / CB);
  const uint64_t rows = sizeof(double) * 4 * (uint64_t)N;
  MH_TRY(ensure_scratch(ctx, 2 * rows + sizeof(int32_t) * (2 * (uint64_t)nblk + 8)));
  double* tmp_pts = (double*)ctx->scratch;
  double* tmp_aff = tmp_pts + 4 * N;
  int32_t* counts = (int32_t*)(tmp_aff + 4 * N);
  int32_t* offs = counts + nblk;
  prefilter_kernel<<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>(d_pts64, d_aff64, N, tmp_pts, tmp_aff, d_keep, g);
  MH_LAUNCHED(ctx, "prefilter_kernel");
  block_count_kernel<<<nblk, CB, 0, ctx->stream>>>(d_keep, N, counts);
  MH_LAUNCHED(ctx, "block_count_kernel");
  scan_small_kernel<<<1, 1024, 0, ctx->stream>>>(counts, nblk, offs);
  MH_LAUNCHED(ctx, "scan_small_kernel");
  compact_kernel<<<nblk, CB, 0, ctx->stream>>>(d_keep, N, offs, tmp_pts, tmp_aff, d_pts_out, d_aff_out);
  MH_LAUNCHED(ctx, "compact_kernel");
  int32_t total = 0;
  MH_CUDA(ctx, cudaMemcpyAsync(&total, offs + nblk, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *M_out = total;
  return MH_OK;
}

}  // namespace mh
