// ============================================================================
// capi.cu — the C ABI of libmultih_b200.so (include/multih_b200.h): context,
// geometry, host<->device-space conversion and thin argument-checking wrappers
// around the kernel launchers.  No CPU fallback exists anywhere in this file:
// every compute entry point launches a CUDA kernel or fails.
// ============================================================================
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace mh {

mh_status fail(mh_ctx* ctx, mh_status st, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return st;
}

mh_status check_cuda(mh_ctx* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return MH_OK;
  return fail(ctx, e == cudaErrorMemoryAllocation ? MH_ENOMEM : MH_ECUDA,
              std::string(what) + ": " + cudaGetErrorString(e));
}

static mh_status grow(mh_ctx* ctx, void** p, uint64_t* have, uint64_t want, bool host) {
  if (*have >= want) return MH_OK;
  uint64_t n = std::max<uint64_t>(want, *have + *have / 2);
  n = (n + 255) & ~uint64_t(255);
  if (*p) {
    MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MH_CUDA(ctx, host ? cudaFreeHost(*p) : cudaFree(*p));
    *p = nullptr; *have = 0;
  }
  MH_CUDA(ctx, host ? cudaMallocHost(p, n) : cudaMalloc(p, n));
  *have = n;
  return MH_OK;
}
mh_status ensure_scratch(mh_ctx* ctx, uint64_t bytes) { return grow(ctx, &ctx->scratch, &ctx->scratch_bytes, bytes, false); }
mh_status ensure_staging(mh_ctx* ctx, uint64_t bytes) { return grow(ctx, &ctx->staging, &ctx->staging_bytes, bytes, false); }
mh_status ensure_pinned(mh_ctx* ctx, uint64_t bytes) { return grow(ctx, &ctx->pinned, &ctx->pinned_bytes, bytes, true); }

// dataEnergy constants (MultiH.h:41-44, MultiH.cpp:478-503) in normalised units
CostParams cost_params(const mh_ctx* ctx) {
  CostParams cp;
  const double thr2_px = ctx->params.thr_homography * ctx->params.thr_homography;
  const double T_px = thr2_px * 81.0 / 16.0;
  const double lam = 100.0 / ctx->params.lambda;
  const double s2sq = ctx->gd.s2 * ctx->gd.s2;
  cp.T = (float)(T_px * s2sq);
  cp.inv_T = (float)(1.0 / (T_px * s2sq));
  cp.lam = (float)lam;
  cp.thr2 = (float)(thr2_px * s2sq);
  cp.cost_outlier = (int32_t)std::round(lam * T_px);
  cp.cost_far = 2 * (int32_t)std::round(lam * T_px);
  return cp;
}

// cv::eigen-equivalent symmetric 3x3 Jacobi (host, FP64): eigenvalues descending, eigenvectors in rows
static void sym_eigen3(const double* Ain, double* w, double* Vrows) {
  double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  std::memcpy(A, Ain, sizeof(A));
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5], diag = A[0] * A[0] + A[4] * A[4] + A[8] * A[8];
    if (off <= 1e-300 || off <= 1e-34 * diag) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = A[p * 3 + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) { double a = A[k * 3 + p], b = A[k * 3 + q]; A[k * 3 + p] = c * a - s * b; A[k * 3 + q] = s * a + c * b; }
        for (int k = 0; k < 3; ++k) { double a = A[p * 3 + k], b = A[q * 3 + k]; A[p * 3 + k] = c * a - s * b; A[q * 3 + k] = s * a + c * b; }
        for (int k = 0; k < 3; ++k) { double a = V[k * 3 + p], b = V[k * 3 + q]; V[k * 3 + p] = c * a - s * b; V[k * 3 + q] = s * a + c * b; }
      }
  }
  int o[3] = {0, 1, 2};
  std::sort(o, o + 3, [&](int a, int b) { return A[a * 3 + a] > A[b * 3 + b]; });
  for (int r = 0; r < 3; ++r) {
    w[r] = A[o[r] * 3 + o[r]];
    for (int k = 0; k < 3; ++k) Vrows[r * 3 + k] = V[k * 3 + o[r]];
  }
}

static void mat3_mul(const double* A, const double* B, double* C) {
  double T[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
  std::memcpy(C, T, sizeof(T));
}

// epipole in image 2: last row of eigen(F F^T) / z (MultiH.cpp:789-793)
void epipole2_host(const double* F, double* e) {
  double FFt[9], w[3], V[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) FFt[i * 3 + j] = F[i * 3] * F[j * 3] + F[i * 3 + 1] * F[j * 3 + 1] + F[i * 3 + 2] * F[j * 3 + 2];
  sym_eigen3(FFt, w, V);
  e[0] = V[6] / V[8];
  e[1] = V[7] / V[8];
}

void hyp_px_to_norm(const mh_ctx* ctx, const double* H, float* out12) {
  const GeomD& g = ctx->gd;
  const double T2[9] = {g.s2, 0, g.t2x, 0, g.s2, g.t2y, 0, 0, 1};
  const double T1i[9] = {1.0 / g.s1, 0, -g.t1x / g.s1, 0, 1.0 / g.s1, -g.t1y / g.s1, 0, 0, 1};
  double M[9];
  mat3_mul(T2, H, M);
  mat3_mul(M, T1i, M);
  double m = 0;
  for (int k = 0; k < 9; ++k) m = std::max(m, std::fabs(M[k]));
  const double sc = m > 0 ? 1.0 / m : 1.0;  // homographies are scale-free: keep FP32 entries O(1)
  for (int k = 0; k < 9; ++k) out12[k] = (float)(M[k] * sc);
  out12[9] = out12[10] = out12[11] = 0.f;
}

void hyp_norm_to_px(const mh_ctx* ctx, const float* in12, double* H, bool divide_h33) {
  const GeomD& g = ctx->gd;
  const double T1[9] = {g.s1, 0, g.t1x, 0, g.s1, g.t1y, 0, 0, 1};
  const double T2i[9] = {1.0 / g.s2, 0, -g.t2x / g.s2, 0, 1.0 / g.s2, -g.t2y / g.s2, 0, 0, 1};
  double M[9];
  for (int k = 0; k < 9; ++k) M[k] = in12[k];
  mat3_mul(T2i, M, M);
  mat3_mul(M, T1, M);
  const double sc = (divide_h33 && M[8] != 0.0) ? 1.0 / M[8] : 1.0;
  for (int k = 0; k < 9; ++k) H[k] = M[k] * sc;
}

void sym_eigen3_host(const double* A, double* w, double* Vrows) { sym_eigen3(A, w, Vrows); }

}  // namespace mh

using namespace mh;

extern "C" void mh_step_release(mh_ctx* ctx);   // comm.cu (internal)

extern "C" {

const char* mh_version(void) { return "multih_b200 0.1 (sm_100a)"; }

void mh_default_params(mh_params* p) {
  if (!p) return;
  p->thr_fundamental = 2.6;  // main.cpp:56
  p->thr_homography = 2.2;   // main.cpp:57
  p->locality = 0.005;       // main.cpp:58
  p->lambda = 0.5;           // main.cpp:59
  p->min_inliers = 20;       // main.cpp:55
  p->straightness = 0.005;   // MultiH.h:13
  p->max_iterations = 500;   // MultiH.h:14
  p->convergence = 1e-5;     // MultiH.h:15
  p->meanshift_metric = 0;
  p->rng_seed = 1u;
  p->max_gc_cycles = 1000;   // MultiH.cpp:543
  p->max_neighbours = 31;    // FLANN default SearchParams: checks = 32 (query included)
  p->precise_pipeline = 1;
  p->prefilter = 0;            // inputs are taken as already refined; the MultiH class shims switch it on (Process() takes raw rows)
  p->lm_refine = 1;            // the 3PT fits are LM-polished as in the reference (MultiH.cpp:1052-1053)
  p->compatibility_check = 1;  // Process() always ends with HomographyCompatibilityCheck when K > 1 (MultiH.cpp:76-86)
}

mh_status mh_create(const mh_params* params, int device, mh_ctx** out) {
  if (!out) return MH_EINVAL;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return MH_ECUDA;  // no CPU fallback
  mh_ctx* ctx = new mh_ctx();
  if (params) ctx->params = *params; else mh_default_params(&ctx->params);
  ctx->device = device;
  ctx->rng_state = ctx->params.rng_seed;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return MH_ECUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return MH_ECUDA; }
  if (prop.major < 10) { delete ctx; return MH_ECUDA; }  // sm_100a binary only
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return MH_ECUDA; }
  ctx->own_stream = true;
  *out = ctx;
  return MH_OK;
}

void mh_destroy(mh_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  mh_comm_destroy(ctx);
  mh_step_release(ctx);
  if (ctx->comm_acc) cudaFree(ctx->comm_acc);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->staging) cudaFree(ctx->staging);
  for (void* b : ctx->pbuf)
    if (b) cudaFree(b);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* mh_last_error(const mh_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

mh_status mh_set_stream(mh_ctx* ctx, void* s) {
  if (!ctx) return MH_EINVAL;
  MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)s;
  ctx->own_stream = false;
  return MH_OK;
}

mh_status mh_sync(mh_ctx* ctx) {
  if (!ctx) return MH_EINVAL;
  MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MH_OK;
}

mh_status mh_alloc(mh_ctx* ctx, uint64_t bytes, void** p) {
  if (!ctx || !p) return MH_EINVAL;
  MH_CUDA(ctx, cudaSetDevice(ctx->device));
  MH_CUDA(ctx, cudaMalloc(p, std::max<uint64_t>(bytes, 16)));
  return MH_OK;
}
mh_status mh_free(mh_ctx* ctx, void* p) {
  if (!ctx) return MH_EINVAL;
  if (p) MH_CUDA(ctx, cudaFree(p));
  return MH_OK;
}
mh_status mh_host_alloc(mh_ctx* ctx, uint64_t bytes, void** p) {
  if (!ctx || !p) return MH_EINVAL;
  MH_CUDA(ctx, cudaMallocHost(p, std::max<uint64_t>(bytes, 16)));
  return MH_OK;
}
mh_status mh_host_free(mh_ctx* ctx, void* p) {
  if (!ctx) return MH_EINVAL;
  if (p) MH_CUDA(ctx, cudaFreeHost(p));
  return MH_OK;
}
mh_status mh_memcpy_d2h(mh_ctx* ctx, void* host, const void* d, uint64_t bytes) {
  if (!ctx || (bytes && (!host || !d))) return MH_EINVAL;
  MH_CUDA(ctx, cudaMemcpyAsync(host, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MH_OK;
}
mh_status mh_memcpy_h2d(mh_ctx* ctx, void* d, const void* host, uint64_t bytes) {
  if (!ctx || (bytes && (!host || !d))) return MH_EINVAL;
  MH_CUDA(ctx, cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MH_OK;
}
int64_t mh_kernel_launches(const mh_ctx* ctx) { return ctx ? ctx->launches : 0; }

mh_status mh_set_geometry(mh_ctx* ctx, const double F[9], const double n1[3], const double n2[3], const double* pts,
                          int64_t N) {
  if (!ctx || !F) return MH_EINVAL;
  double fn = 0;
  for (int i = 0; i < 9; ++i) fn += F[i] * F[i];
  if (!(std::sqrt(fn) >= 1e-5)) return fail(ctx, MH_EDEGENERATE, "fundamental matrix norm < 1e-5 (MultiH.cpp:779)");
  double a[3], b[3];
  if (n1 && n2) {
    std::memcpy(a, n1, sizeof(a));
    std::memcpy(b, n2, sizeof(b));
  } else {
    if (!pts || N <= 0) return fail(ctx, MH_EINVAL, "mh_set_geometry: need norm1/norm2 or points");
    // Hartley-style similarity from a strided sample (<= 8192 correspondences): deterministic, O(1) host work
    const int64_t step = std::max<int64_t>(1, N / 8192);
    double m[4] = {0, 0, 0, 0};
    int64_t n = 0;
    for (int64_t i = 0; i < N; i += step, ++n)
      for (int k = 0; k < 4; ++k) m[k] += pts[4 * i + k];
    for (int k = 0; k < 4; ++k) m[k] /= (double)n;
    double d1 = 0, d2 = 0;
    for (int64_t i = 0; i < N; i += step) {
      d1 += std::hypot(pts[4 * i] - m[0], pts[4 * i + 1] - m[1]);
      d2 += std::hypot(pts[4 * i + 2] - m[2], pts[4 * i + 3] - m[3]);
    }
    d1 /= (double)n; d2 /= (double)n;
    const double s1 = d1 > 0 ? std::sqrt(2.0) / d1 : 1.0, s2 = d2 > 0 ? std::sqrt(2.0) / d2 : 1.0;
    a[0] = s1; a[1] = -m[0] * s1; a[2] = -m[1] * s1;
    b[0] = s2; b[1] = -m[2] * s2; b[2] = -m[3] * s2;
  }
  if (!(a[0] > 0) || !(b[0] > 0)) return fail(ctx, MH_EINVAL, "mh_set_geometry: scales must be positive");
  std::memcpy(ctx->F_px, F, sizeof(double) * 9);
  epipole2_host(F, ctx->e2_px);
  GeomD& g = ctx->gd;
  g.s1 = a[0]; g.t1x = a[1]; g.t1y = a[2];
  g.s2 = b[0]; g.t2x = b[1]; g.t2y = b[2];
  // F' = T2^-T F T1^-1
  const double T1i[9] = {1.0 / g.s1, 0, -g.t1x / g.s1, 0, 1.0 / g.s1, -g.t1y / g.s1, 0, 0, 1};
  const double T2it[9] = {1.0 / g.s2, 0, 0, 0, 1.0 / g.s2, 0, -g.t2x / g.s2, -g.t2y / g.s2, 1};
  double Fn[9];
  mat3_mul(T2it, F, Fn);
  mat3_mul(Fn, T1i, Fn);
  double nf = 0;
  for (int i = 0; i < 9; ++i) nf = std::max(nf, std::fabs(Fn[i]));
  for (int i = 0; i < 9; ++i) g.F[i] = Fn[i] / nf;
  g.ex = ctx->e2_px[0] * g.s2 + g.t2x;
  g.ey = ctx->e2_px[1] * g.s2 + g.t2y;
  GeomF& f = ctx->gf;
  for (int i = 0; i < 9; ++i) f.F[i] = (float)g.F[i];
  f.ex = (float)g.ex; f.ey = (float)g.ey;
  f.s1 = (float)g.s1; f.t1x = (float)g.t1x; f.t1y = (float)g.t1y;
  f.s2 = (float)g.s2; f.t2x = (float)g.t2x; f.t2y = (float)g.t2y;
  ctx->have_geom = true;
  return MH_OK;
}

mh_status mh_get_geometry(const mh_ctx* ctx, double F[9], double e2[2], double n1[3], double n2[3]) {
  if (!ctx || !ctx->have_geom) return MH_EINVAL;
  if (F) std::memcpy(F, ctx->F_px, sizeof(double) * 9);
  if (e2) { e2[0] = ctx->e2_px[0]; e2[1] = ctx->e2_px[1]; }
  if (n1) { n1[0] = ctx->gd.s1; n1[1] = ctx->gd.t1x; n1[2] = ctx->gd.t1y; }
  if (n2) { n2[0] = ctx->gd.s2; n2[1] = ctx->gd.t2x; n2[2] = ctx->gd.t2y; }
  return MH_OK;
}

#define NEED_GEOM(ctx)                                                                         \
  do {                                                                                         \
    if (!(ctx)) return MH_EINVAL;                                                              \
    if (!(ctx)->have_geom) return fail((ctx), MH_EINVAL, "call mh_set_geometry first");        \
  } while (0)

mh_status mh_upload_correspondences(mh_ctx* ctx, const double* pts, const double* aff, int64_t N, void* d_pts,
                                    void* d_aff) {
  NEED_GEOM(ctx);
  // either half may be omitted (both pointers of a half NULL), so that a caller can overlap the two uploads with compute
  if (N < 0 || (!pts != !d_pts) || (!aff != !d_aff) || (N && !pts && !aff))
    return fail(ctx, MH_EINVAL, "mh_upload_correspondences: bad arguments");
  if (N == 0) return MH_OK;
  const uint64_t bytes = sizeof(double) * 4 * (uint64_t)N;
  MH_TRY(ensure_staging(ctx, 2 * bytes));
  double* raw_p = (double*)ctx->staging;
  double* raw_a = raw_p + 4 * N;
  if (pts) MH_CUDA(ctx, cudaMemcpyAsync(raw_p, pts, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (aff) MH_CUDA(ctx, cudaMemcpyAsync(raw_a, aff, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return launch_normalize_points(ctx, pts ? raw_p : nullptr, aff ? raw_a : nullptr, N, (float4*)d_pts, (float4*)d_aff);
}

mh_status mh_hypotheses_from_host(mh_ctx* ctx, const double* H, int32_t K, void* d_hyp) {
  NEED_GEOM(ctx);
  if (K < 0 || (K && (!H || !d_hyp))) return fail(ctx, MH_EINVAL, "mh_hypotheses_from_host: bad arguments");
  if (K == 0) return MH_OK;
  MH_TRY(ensure_pinned(ctx, sizeof(float) * 12 * (uint64_t)K));
  MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // pinned buffer may still be in flight
  float* h = (float*)ctx->pinned;
  for (int k = 0; k < K; ++k) hyp_px_to_norm(ctx, H + 9 * (size_t)k, h + 12 * (size_t)k);
  MH_CUDA(ctx, cudaMemcpyAsync(d_hyp, h, sizeof(float) * 12 * (size_t)K, cudaMemcpyHostToDevice, ctx->stream));
  return MH_OK;
}

mh_status mh_hypotheses_to_host(mh_ctx* ctx, const void* d_hyp, int32_t K, double* H, int32_t divide) {
  NEED_GEOM(ctx);
  if (K < 0 || (K && (!H || !d_hyp))) return fail(ctx, MH_EINVAL, "mh_hypotheses_to_host: bad arguments");
  if (K == 0) return MH_OK;
  MH_TRY(ensure_pinned(ctx, sizeof(float) * 12 * (uint64_t)K));
  float* h = (float*)ctx->pinned;
  MH_CUDA(ctx, cudaMemcpyAsync(h, d_hyp, sizeof(float) * 12 * (size_t)K, cudaMemcpyDeviceToHost, ctx->stream));
  MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < K; ++k) hyp_norm_to_px(ctx, h + 12 * (size_t)k, H + 9 * (size_t)k, divide != 0);
  return MH_OK;
}

mh_status mh_prefilter(mh_ctx* ctx, const double* pts, const double* aff, const double F[9], int64_t N, double* pts_out,
                       double* aff_out, uint8_t* keep_out, int64_t* M_out) {
  if (!ctx) return MH_EINVAL;
  if (!F || !M_out || N < 0 || (N && (!pts || !aff || !pts_out || !aff_out))) return fail(ctx, MH_EINVAL, "mh_prefilter: bad arguments");
  *M_out = 0;
  if (N == 0) return MH_OK;
  const uint64_t rows = sizeof(double) * 4 * (uint64_t)N;
  MH_TRY(ensure_staging(ctx, 4 * rows + sizeof(int32_t) * (uint64_t)N));
  double* d_p = (double*)ctx->staging;
  double* d_a = d_p + 4 * N;
  double* d_po = d_a + 4 * N;
  double* d_ao = d_po + 4 * N;
  int32_t* d_keep = (int32_t*)(d_ao + 4 * N);
  MH_CUDA(ctx, cudaMemcpyAsync(d_p, pts, rows, cudaMemcpyHostToDevice, ctx->stream));
  MH_CUDA(ctx, cudaMemcpyAsync(d_a, aff, rows, cudaMemcpyHostToDevice, ctx->stream));
  int64_t M = 0;
  MH_TRY(launch_prefilter(ctx, d_p, d_a, F, N, d_po, d_ao, d_keep, &M));
  MH_CUDA(ctx, cudaMemcpyAsync(pts_out, d_po, sizeof(double) * 4 * (size_t)M, cudaMemcpyDeviceToHost, ctx->stream));
  MH_CUDA(ctx, cudaMemcpyAsync(aff_out, d_ao, sizeof(double) * 4 * (size_t)M, cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<int32_t> k32;
  if (keep_out) {
    k32.resize(N);
    MH_CUDA(ctx, cudaMemcpyAsync(k32.data(), d_keep, sizeof(int32_t) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
  }
  MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (keep_out) for (int64_t i = 0; i < N; ++i) keep_out[i] = (uint8_t)(k32[i] != 0);
  *M_out = M;
  return MH_OK;
}

mh_status mh_prefilter_device(mh_ctx* ctx, const void* d_pts64, const void* d_aff64, const double F[9], int64_t N,
                              void* d_pts64_out, void* d_aff64_out, void* d_keep, int64_t* M_out) {
  if (!ctx) return MH_EINVAL;
  if (!F || !M_out || N < 0 || (N && (!d_pts64 || !d_aff64 || !d_pts64_out || !d_aff64_out || !d_keep)))
    return fail(ctx, MH_EINVAL, "mh_prefilter_device: bad arguments");
  return launch_prefilter(ctx, (const double*)d_pts64, (const double*)d_aff64, F, N, (double*)d_pts64_out, (double*)d_aff64_out,
                          (int32_t*)d_keep, M_out);
}

mh_status mh_haf_hypotheses(mh_ctx* ctx, const void* d_pts, const void* d_aff, int64_t N, void* d_hyp, int32_t prec) {
  NEED_GEOM(ctx);
  if (N < 0 || (N && (!d_pts || !d_aff || !d_hyp))) return fail(ctx, MH_EINVAL, "mh_haf_hypotheses: bad arguments");
  return launch_haf(ctx, (const float4*)d_pts, (const float4*)d_aff, N, (float*)d_hyp, prec);
}

mh_status mh_data_cost_dense(mh_ctx* ctx, const void* d_pts, int64_t N, const void* d_hyp, int32_t K, void* d_cost,
                             int32_t elem_bytes) {
  NEED_GEOM(ctx);
  if (N < 0 || K < 0 || (N && (!d_pts || !d_cost)) || (K && !d_hyp)) return fail(ctx, MH_EINVAL, "mh_data_cost_dense: bad arguments");
  if (elem_bytes == 2 && cost_params(ctx).cost_far > 32767) return fail(ctx, MH_EINVAL, "mh_data_cost_dense: costs do not fit int16");
  return launch_cost_dense(ctx, (const float4*)d_pts, N, (const float*)d_hyp, K, d_cost, elem_bytes);
}

mh_status mh_residuals(mh_ctx* ctx, const void* d_pts, int64_t N, const void* d_hyp, int32_t K, void* d_d2) {
  NEED_GEOM(ctx);
  if (N < 0 || K < 0 || (N && K && (!d_pts || !d_hyp || !d_d2))) return fail(ctx, MH_EINVAL, "mh_residuals: bad arguments");
  return launch_residuals(ctx, (const float4*)d_pts, N, (const float*)d_hyp, K, (float*)d_d2);
}

mh_status mh_data_cost_fused(mh_ctx* ctx, const void* d_pts, int64_t N, const void* d_hyp, int32_t K, int32_t kmax,
                             void* d_list, void* d_list_count, void* d_best, void* d_inlier_count) {
  NEED_GEOM(ctx);
  if (N < 0 || K < 0 || kmax < 0 || (N && !d_pts) || (K && !d_hyp) || (d_list && !d_list_count))
    return fail(ctx, MH_EINVAL, "mh_data_cost_fused: bad arguments");
  return launch_cost_fused(ctx, (const float4*)d_pts, N, (const float*)d_hyp, K, kmax, (uint32_t*)d_list,
                           (int32_t*)d_list_count, (unsigned long long*)d_best, (int32_t*)d_inlier_count);
}

mh_status mh_inlier_stats(mh_ctx* ctx, const void* d_pts, int64_t N, const void* d_hyp, int32_t K, double* scatter,
                          double* lambda_min, int32_t* keep) {
  NEED_GEOM(ctx);
  if (N < 0 || K < 0 || (N && !d_pts) || (K && !d_hyp)) return fail(ctx, MH_EINVAL, "mh_inlier_stats: bad arguments");
  if (K == 0) return MH_OK;
  MH_TRY(ensure_scratch(ctx, sizeof(double) * 6 * (uint64_t)K));
  MH_TRY(ensure_pinned(ctx, sizeof(double) * 6 * (uint64_t)K));
  double* d_sc = (double*)ctx->scratch;
  MH_TRY(launch_inlier_stats(ctx, (const float4*)d_pts, N, (const float*)d_hyp, K, d_sc));
  double* h = (double*)ctx->pinned;
  MH_CUDA(ctx, cudaMemcpyAsync(h, d_sc, sizeof(double) * 6 * (size_t)K, cudaMemcpyDeviceToHost, ctx->stream));
  MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // normalised -> pixel: p = A p', A = T1^-1 ; S = A S' A^T
  const GeomD& g = ctx->gd;
  const double A[9] = {1.0 / g.s1, 0, -g.t1x / g.s1, 0, 1.0 / g.s1, -g.t1y / g.s1, 0, 0, 1};
  double At[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) At[i * 3 + j] = A[j * 3 + i];
  for (int k = 0; k < K; ++k) {
    const double* s = h + 6 * (size_t)k;
    const double Sn[9] = {s[0], s[1], s[2], s[1], s[3], s[4], s[2], s[4], s[5]};
    double S[9];
    mat3_mul(A, Sn, S);
    mat3_mul(S, At, S);
    if (scatter) {
      double* o = scatter + 6 * (size_t)k;
      o[0] = S[0]; o[1] = S[1]; o[2] = S[2]; o[3] = S[4]; o[4] = S[5]; o[5] = S[8];
    }
    if (lambda_min || keep) {
      double w[3], V[9];
      sym_eigen3(S, w, V);  // MultiH.cpp:458-459
      if (lambda_min) lambda_min[k] = w[2];
      if (keep) keep[k] = !(w[2] < ctx->params.straightness || s[5] < 3.0);  // MultiH.cpp:462
    }
  }
  return MH_OK;
}

mh_status mh_inliers_of_homography(mh_ctx* ctx, const void* d_pts, int64_t N, const void* d_hyp_one, int32_t idx,
                                   void* d_labels) {
  NEED_GEOM(ctx);
  if (N < 0 || (N && (!d_pts || !d_hyp_one || !d_labels))) return fail(ctx, MH_EINVAL, "mh_inliers_of_homography: bad arguments");
  return launch_inliers_of(ctx, (const float4*)d_pts, N, (const float*)d_hyp_one, idx, (int32_t*)d_labels);
}

mh_status mh_features10(mh_ctx* ctx, const void* d_hyp, const void* d_pts, int64_t N, void* d_feat) {
  NEED_GEOM(ctx);
  if (N < 0 || (N && (!d_hyp || !d_pts || !d_feat))) return fail(ctx, MH_EINVAL, "mh_features10: bad arguments");
  return launch_features10(ctx, (const float*)d_hyp, (const float4*)d_pts, N, (double*)d_feat);
}

mh_status mh_features6(mh_ctx* ctx, const void* d_hyp, int32_t K, void* d_feat) {
  NEED_GEOM(ctx);
  if (K < 0 || (K && (!d_hyp || !d_feat))) return fail(ctx, MH_EINVAL, "mh_features6: bad arguments");
  return launch_features6(ctx, (const float*)d_hyp, K, (double*)d_feat);
}

mh_status mh_meanshift(mh_ctx* ctx, const void* d_feat, int32_t N, int32_t D, double bw, void* d_centres, int32_t max_c,
                       void* d_assign, int32_t* C_out, int64_t* stats) {
  if (!ctx) return MH_EINVAL;
  if (N < 0 || D <= 0 || D > 16 || !(bw > 0) || !C_out || (N && (!d_feat || !d_centres || !d_assign)))
    return fail(ctx, MH_EINVAL, "mh_meanshift: bad arguments");
  *C_out = 0;
  if (N == 0) return MH_OK;
  return launch_meanshift(ctx, (const double*)d_feat, N, D, bw, ctx->params.meanshift_metric, &ctx->rng_state,
                          (double*)d_centres, max_c, (int32_t*)d_assign, C_out, stats);
}

mh_status mh_refit_haf(mh_ctx* ctx, const void* d_pts, const void* d_aff, const void* d_labels, int64_t N, int32_t K,
                       void* d_hyp, void* d_count) {
  NEED_GEOM(ctx);
  if (N < 0 || K < 0 || (N && (!d_pts || !d_aff || !d_labels)) || (K && !d_hyp)) return fail(ctx, MH_EINVAL, "mh_refit_haf: bad arguments");
  return launch_refit_haf(ctx, (const float4*)d_pts, (const float4*)d_aff, (const int32_t*)d_labels, N, K, (float*)d_hyp,
                          (int32_t*)d_count);
}

mh_status mh_refit_haf_accumulate(mh_ctx* ctx, const void* d_pts, const void* d_aff, const void* d_labels, int64_t N,
                                  int32_t K, void* d_acc) {
  NEED_GEOM(ctx);
  if (N < 0 || K < 0 || (N && (!d_pts || !d_aff || !d_labels)) || (K && !d_acc)) return fail(ctx, MH_EINVAL, "mh_refit_haf_accumulate: bad arguments");
  return launch_refit_haf_accumulate(ctx, (const float4*)d_pts, (const float4*)d_aff, (const int32_t*)d_labels, N, K,
                                     (double*)d_acc);
}

mh_status mh_refit_haf_solve(mh_ctx* ctx, const void* d_acc, int32_t K, void* d_hyp, void* d_count) {
  NEED_GEOM(ctx);
  if (K < 0 || (K && (!d_acc || !d_hyp))) return fail(ctx, MH_EINVAL, "mh_refit_haf_solve: bad arguments");
  return launch_refit_haf_solve(ctx, (const double*)d_acc, K, (float*)d_hyp, (int32_t*)d_count);
}

mh_status mh_labels_from_best(mh_ctx* ctx, const void* d_best, int64_t N, void* d_labels) {
  if (!ctx) return MH_EINVAL;
  if (N < 0 || (N && (!d_best || !d_labels))) return fail(ctx, MH_EINVAL, "mh_labels_from_best: bad arguments");
  return launch_labels_from_best(ctx, (const unsigned long long*)d_best, N, (int32_t*)d_labels);
}

mh_status mh_pack_inlier_counts(mh_ctx* ctx, void* d_inlier_count, int32_t K, void* d_acc, int32_t unpack) {
  if (!ctx) return MH_EINVAL;
  if (K < 0 || (K && (!d_inlier_count || !d_acc))) return fail(ctx, MH_EINVAL, "mh_pack_inlier_counts: bad arguments");
  return launch_pack_inlier_counts(ctx, (int32_t*)d_inlier_count, K, (double*)d_acc, unpack);
}

mh_status mh_refit_3pt(mh_ctx* ctx, const void* d_pts, const void* d_assign, int64_t N, int32_t C, void* d_hyp,
                       void* d_keep) {
  NEED_GEOM(ctx);
  if (N < 0 || C < 0 || (N && (!d_pts || !d_assign)) || (C && (!d_hyp || !d_keep))) return fail(ctx, MH_EINVAL, "mh_refit_3pt: bad arguments");
  return launch_refit_3pt(ctx, (const float4*)d_pts, (const int32_t*)d_assign, N, C, (float*)d_hyp, (int32_t*)d_keep);
}

mh_status mh_modes_to_hypotheses(mh_ctx* ctx, const void* d_modes, int32_t C, void* d_hyp) {
  NEED_GEOM(ctx);
  if (C < 0 || (C && (!d_modes || !d_hyp))) return fail(ctx, MH_EINVAL, "mh_modes_to_hypotheses: bad arguments");
  return launch_modes_to_hyp(ctx, (const double*)d_modes, C, (float*)d_hyp);
}

mh_status mh_set_rng_state(mh_ctx* ctx, uint32_t state) {
  if (!ctx) return MH_EINVAL;
  ctx->rng_state = state;
  return MH_OK;
}
uint32_t mh_get_rng_state(const mh_ctx* ctx) { return ctx ? ctx->rng_state : 0u; }

double mh_get_energy(const mh_ctx* ctx) { return ctx ? ctx->energy : 0.0; }
int32_t mh_get_iterations(const mh_ctx* ctx) { return ctx ? ctx->iterations : 0; }
mh_status mh_get_stage_ms(const mh_ctx* ctx, double ms[5]) {
  if (!ctx || !ms) return MH_EINVAL;
  for (int i = 0; i < 5; ++i) ms[i] = ctx->stage_ms[i];
  return MH_OK;
}
mh_status mh_diag_get_alternating_ms(const mh_ctx* ctx, double ms[5]) {
  if (!ctx || !ms) return MH_EINVAL;
  for (int i = 0; i < 5; ++i) ms[i] = ctx->alt_ms[i];
  return MH_OK;
}

}  // extern "C"
