// Device helpers shared by the K2 kernel files (k2_residual.cu, k2_mma.cu): packed f32x2 arithmetic, the residual /
// data-cost formulas of dataEnergy (MultiH/MultiH/MultiH.cpp:473-504) and the TF32 tensor-core primitives.
#pragma once
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

struct FastOut {
  u64* best;             // [N] packed (cost << 32 | label), pre-initialised to (cost_outlier << 32 | 0)
  int32_t* inlier_count; // [K]
  // list-producing members only (zero otherwise): per site up to kmax entries (label << 8 | cost) with d2 < T, and their number
  uint32_t* list;        // [N][kmax]
  int32_t* list_count;   // [N], pre-initialised to 0
  int kmax;
};

__device__ __forceinline__ void fast_thresholds(int best_cost, const CostParams& cp, float& negmid, float& half) {
  // a residual improves on best_cost iff cost(d2) <= best_cost - 1  <=>  d2 > T * (lam + 0.5 - best_cost) / lam
  float lo = cp.T * (cp.lam + 0.5f - (float)best_cost) * (1.0f / cp.lam);
  // conservative by 1e-3 (relative): the filter sees residuals whose rounding differs from the exact FP32 sequence (folded
  // constants; in the tensor-core kernels 3xTF32 products of hypotheses with |h_i| <= 4 |h_8|, worth <~ 2e-4); false
  // positives are rejected by the exact update
  lo = fmaxf(lo, 0.f) * 0.999f - 1e-12f;
  const float hi = cp.T * 1.001f;
  negmid = -0.5f * (lo + hi);
  half = 0.5f * (hi - lo);
}

__device__ __forceinline__ unsigned to_tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}


}  // namespace mh
