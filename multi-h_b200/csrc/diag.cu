// Diagnostics: an FP32 FMA-pipe peak probe used by bench.py as the roofline denominator of the FP32-bound residual
// kernel (MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks).  variant 0 = scalar FFMA, 1 = packed FFMA2.
#include "common.cuh"

namespace mh {

template <bool PACKED>
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float a, float b) {
  constexpr int R = 16;
  if (PACKED) {
    unsigned long long acc[R], av, bv;
    asm("mov.b64 %0, {%1, %2};" : "=l"(av) : "f"(a), "f"(a * 1.0001f));
    asm("mov.b64 %0, {%1, %2};" : "=l"(bv) : "f"(b), "f"(b * 0.9999f));
#pragma unroll
    for (int r = 0; r < R; ++r) asm("mov.b64 %0, {%1, %2};" : "=l"(acc[r]) : "f"((float)(threadIdx.x + r)), "f"((float)r));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int r = 0; r < R; ++r) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[r]) : "l"(av), "l"(bv));
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[r]));
      s += lo + hi;
    }
    if (s == 123.456f) out[0] = s;
  } else {
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = (float)(threadIdx.x + r);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int r = 0; r < R; ++r) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[r]) : "f"(a), "f"(b));
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) s += acc[r];
    if (s == 123.456f) out[0] = s;
  }
}

}  // namespace mh

extern "C" mh_status mh_diag_fp32_peak(mh_ctx* ctx, int32_t variant, int32_t iters, double* tflops_out, double* ms_out) {
  using namespace mh;
  if (!ctx || !tflops_out || iters <= 0) return MH_EINVAL;
  MH_TRY(ensure_scratch(ctx, 256));
  const int blocks = ctx->sm_count * 8, threads = 256;
  cudaEvent_t e0, e1;
  MH_CUDA(ctx, cudaEventCreate(&e0));
  MH_CUDA(ctx, cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {  // first pass warms up
    MH_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    if (variant) fma_peak_kernel<true><<<blocks, threads, 0, ctx->stream>>>((float*)ctx->scratch, iters, 0.999f, 0.001f);
    else fma_peak_kernel<false><<<blocks, threads, 0, ctx->stream>>>((float*)ctx->scratch, iters, 0.999f, 0.001f);
    MH_LAUNCHED(ctx, "fma_peak_kernel");
    MH_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    MH_CUDA(ctx, cudaEventSynchronize(e1));
  }
  float ms = 0.f;
  MH_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double flops = (double)blocks * threads * 16.0 * (double)iters * 2.0 * (variant ? 2.0 : 1.0);
  *tflops_out = flops / (ms * 1e-3) / 1e12;
  if (ms_out) *ms_out = ms;
  return MH_OK;
}

namespace mh { extern int g_fused_variant; extern int g_fast_config; extern int g_dense_variant; extern int g_nb_backend; }
// 1 = packed FFMA2 inner loop (default), 0 = scalar FFMA (kept for A/B evidence)
extern "C" mh_status mh_diag_set_fused_variant(mh_ctx*, int32_t v) { mh::g_fused_variant = v ? 1 : 0; return MH_OK; }
// launch shape of the K2 fast path (threads/CTA x CTAs/SM): 0 = 256x3, 1 = 256x2, 2 = 256x4, 3 = 128x5, 4 = 128x6, 5 = 128x7 (default), 6 = 128x4
extern "C" mh_status mh_diag_set_fast_config(mh_ctx*, int32_t v) { mh::g_fast_config = v; return MH_OK; }
extern "C" int32_t mh_diag_get_fast_config(mh_ctx*) { return mh::g_fast_config; }
extern "C" mh_status mh_diag_set_dense_variant(mh_ctx*, int32_t v) { mh::g_dense_variant = v; return MH_OK; }
// mh_neighbourhood backend: 0 = auto, 1 = host grid search, 2 = K5 on the device (needs a context)
extern "C" mh_status mh_diag_set_neighbourhood_backend(mh_ctx*, int32_t v) { mh::g_nb_backend = v; return MH_OK; }

// legacy-path tensor probe: mma.sync.m16n8k8 TF32 issue rate (used to judge whether the 3x3 homography product of K2
// could move off the FP32 pipe); returns dense TFLOP/s (2*16*8*8 flop per warp instruction)
namespace mh {
__global__ void __launch_bounds__(256) mma_tf32_peak_kernel(float* out, int iters) {
  float c[8][4];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) c[r][k] = (float)(threadIdx.x + r + k);
  const unsigned a0 = 0x3f800000u + threadIdx.x, a1 = 0x3f900000u, a2 = 0x3fa00000u, a3 = 0x3fb00000u, b0 = 0x3f000000u + threadIdx.x, b1 = 0x3f100000u;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[r][0]), "+f"(c[r][1]), "+f"(c[r][2]), "+f"(c[r][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < 8; ++r) s += c[r][0] + c[r][1] + c[r][2] + c[r][3];
  if (s == 123.456f) out[0] = s;
}
}  // namespace mh

extern "C" mh_status mh_diag_mma_tf32_peak(mh_ctx* ctx, int32_t iters, double* tflops_out, double* mma_per_clk_per_sm) {
  using namespace mh;
  if (!ctx || !tflops_out || iters <= 0) return MH_EINVAL;
  MH_TRY(ensure_scratch(ctx, 256));
  const int blocks = ctx->sm_count * 4, threads = 256;
  cudaEvent_t e0, e1;
  MH_CUDA(ctx, cudaEventCreate(&e0));
  MH_CUDA(ctx, cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    MH_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    mma_tf32_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((float*)ctx->scratch, iters);
    MH_LAUNCHED(ctx, "mma_tf32_peak_kernel");
    MH_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    MH_CUDA(ctx, cudaEventSynchronize(e1));
  }
  float ms = 0.f;
  MH_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  const double n_mma = (double)blocks * (threads / 32) * 8.0 * iters;
  *tflops_out = n_mma * 2.0 * 16 * 8 * 8 / (ms * 1e-3) / 1e12;
  if (mma_per_clk_per_sm) *mma_per_clk_per_sm = n_mma / (ms * 1e-3) / 1.965e9 / ctx->sm_count;
  return MH_OK;
}
