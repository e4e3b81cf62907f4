// ============================================================================
// K2 dense — the materialised N x (K+1) data-cost matrix (dataEnergy,
// MultiH/MultiH/MultiH.cpp:473-504; layout = GCoptimization.h:336-343: site-major,
// column 0 = outlier label).  This is the HBM-bound member of the K2 family: 4 B (int32)
// or 2 B (int16) per residual leave the chip, nothing else does.
//
// cost_dense_tiled_kernel: a CTA owns a 512-column chunk of the matrix for its whole life —
// two hypotheses per thread, held as packed f32x2 register pairs, arithmetic in FFMA2 —
// and walks 16-row tiles of correspondences.  Costs are staged in shared memory
// row by row; every row segment is then written by the TMA engine
// (cp.async.bulk.global.shared::cta, one bulk copy per row issued by one thread) as
// full 16-byte-aligned lines, with at most 15 bytes of head/tail per row stored by
// hand.  Rows of an odd-length matrix (K+1 is odd for the usual even K) start at
// every 4-byte phase, so each staged row is shifted by its own (address mod 16) to
// keep shared and global alignment equal.  The store side is its own warp: compute
// warps and store warp hand the two staging buffers back and forth through named
// barriers (bar.arrive / bar.sync), so nobody who computes ever waits on the TMA.
//
// The arithmetic is instruction-for-instruction the sequence of residual() /
// cost_in_range() in k2_device.cuh (packed f32x2 ops are two independent IEEE
// operations), so the matrix is bit-identical to cost_dense_kernel's and to what
// the fused kernels' exact update evaluates.
// ============================================================================
#include "k2_device.cuh"

namespace mh {

constexpr int DT_COMPUTE = 256;              // compute threads (8 warps, 2 matrix columns each)
constexpr int DT_THREADS = DT_COMPUTE + 32;  // + one store warp
constexpr int DT_P = 16;                     // correspondences (matrix rows) per tile
constexpr int DT_KC = 512;                   // matrix columns per CTA
constexpr int DT_STAGES = 2;

__device__ __forceinline__ unsigned dt_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// named barriers 1.. : full[b] = 1 + b (compute warps arrive, store warp syncs), empty[b] = 1 + STAGES + b (the reverse)
__device__ __forceinline__ void dt_bar_sync(int id) { asm volatile("bar.sync %0, %1;" :: "r"(id), "n"(DT_THREADS) : "memory"); }
__device__ __forceinline__ void dt_bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "n"(DT_THREADS) : "memory"); }

template <typename OutT>
__global__ void __launch_bounds__(DT_THREADS, 3)
cost_dense_tiled_kernel(const float4* __restrict__ pts, long long N, const float* __restrict__ hyp, int K,
                        OutT* __restrict__ out, CostParams cp) {
  constexpr int EPV = 16 / (int)sizeof(OutT);   // elements per 16-byte vector
  constexpr int PITCH = DT_KC + EPV;            // staged row pitch (multiple of EPV)
  extern __shared__ __align__(128) unsigned char dt_smem[];
  OutT* stage = reinterpret_cast<OutT*>(dt_smem);   // [DT_STAGES][DT_P][PITCH]

  const int tid = threadIdx.x;
  const long long L = (long long)K + 1;
  const int c0 = blockIdx.y * DT_KC;
  const int ncols = (int)min((long long)DT_KC, L - c0);
  const long long NB = (N + DT_P - 1) / DT_P;
  const unsigned out_phase = (unsigned)((reinterpret_cast<unsigned long long>(out) / sizeof(OutT)) % EPV);
  const unsigned lstep = (unsigned)(L % EPV);
  // phase of a row's first staged element = (global element index of (row, c0)) mod EPV
  auto phase_of = [&](long long p0) {
    return (unsigned)((out_phase + (unsigned long long)p0 * (unsigned long long)L + (unsigned)c0) % EPV);
  };

  if (tid >= DT_COMPUTE) {
    // ================= store warp: one staged row per lane -> one bulk copy (+ <= 15 B head / tail by hand) ===========
    const int r = tid - DT_COMPUTE;
    int it = 0;
    for (long long pb = blockIdx.x; pb < NB; pb += gridDim.x, ++it) {
      const int b = it % DT_STAGES;
      dt_bar_sync(1 + b);   // every compute thread has staged its columns of this tile (and fenced them for the async proxy)
      const long long p = pb * DT_P + r;
      if (r < DT_P && p < N && ncols > 0) {
        const unsigned a = (phase_of(pb * DT_P) + (unsigned)r * lstep) & (EPV - 1);
        const OutT* srow = stage + ((size_t)b * DT_P + r) * PITCH + a;   // staged element of column c0
        OutT* grow = out + p * L + c0;
        int head = (int)((EPV - a) & (EPV - 1));
        if (head > ncols) head = ncols;
        const int body = ((ncols - head) / EPV) * EPV;
        if (body > 0)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       :: "l"(grow + head), "r"(dt_smem_u32(srow + head)), "r"((unsigned)(body * sizeof(OutT))) : "memory");
        for (int j = 0; j < head; ++j) grow[j] = srow[j];
        for (int j = head + body; j < ncols; ++j) grow[j] = srow[j];
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the copies have READ the staged rows
      __syncwarp();
      dt_bar_arrive(1 + DT_STAGES + b);                                 // buffer b may be overwritten
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    return;
  }

  // ================= compute warps: columns c0 + tid and c0 + tid + 256 (hypothesis = column - 1) =====================
  u64 H[9];
  float Ta, Tb;        // in-range threshold, -1 for the outlier column and for columns past the matrix
  int fara, farb;      // value stored when not in range: cost_far, or the outlier cost for column 0
  {
    float ha[9], hb[9];
    auto load = [&](long long c, float (&h)[9], float& T, int& far) {
      const bool is_hyp = c >= 1 && c < L;
      T = is_hyp ? cp.T : -1.f;
      far = (c == 0) ? cp.cost_outlier : cp.cost_far;
#pragma unroll
      for (int k = 0; k < 9; ++k) h[k] = (k == 8) ? 1.f : 0.f;
      if (is_hyp) {
        const float4* hp = reinterpret_cast<const float4*>(hyp + (size_t)(c - 1) * 12);
        const float4 u = __ldg(hp), v = __ldg(hp + 1), w = __ldg(hp + 2);
        h[0] = u.x; h[1] = u.y; h[2] = u.z; h[3] = u.w; h[4] = v.x; h[5] = v.y; h[6] = v.z; h[7] = v.w; h[8] = w.x;
      }
    };
    load((long long)c0 + tid, ha, Ta, fara);
    load((long long)c0 + tid + 256, hb, Tb, farb);
#pragma unroll
    for (int k = 0; k < 9; ++k) H[k] = pk(ha[k], hb[k]);
  }
  const bool warp_active = (tid & ~31) < ncols;   // warps whose columns all lie past the chunk only take part in barriers
  const float kslope = -cp.lam * cp.inv_T;        // the same expression as cost_in_range()
  const u64 KS2 = pk(kslope, kslope), LAM2 = pk(cp.lam, cp.lam), HALF2 = pk(0.5f, 0.5f);

  int it = 0;
  for (long long pb = blockIdx.x; pb < NB; pb += gridDim.x, ++it) {
    const int b = it % DT_STAGES;
    if (it >= DT_STAGES) dt_bar_sync(1 + DT_STAGES + b);   // the store warp is done with this buffer's previous tile
    OutT* st = stage + (size_t)b * DT_P * PITCH;
    const long long p0 = pb * DT_P;
    if (warp_active) {
      unsigned off = phase_of(p0);
#pragma unroll 4
      for (int r = 0; r < DT_P; ++r) {
        const float4 q = __ldg(pts + min(p0 + r, N - 1));   // warp-uniform address: one broadcast L1 hit
        const u64 xx = pk(q.x, q.x), yy = pk(q.y, q.y);
        const u64 sv = fma2(H[6], xx, fma2(H[7], yy, H[8]));
        const u64 xn = fma2(H[0], xx, fma2(H[1], yy, H[2]));
        const u64 yn = fma2(H[3], xx, fma2(H[4], yy, H[5]));
        float sl, sh;
        upk(sv, sl, sh);
        const u64 rr = pk(rcp_approx(sl), rcp_approx(sh));
        const u64 dx = fma2(xn, rr, pk(-q.z, -q.z));
        const u64 dy = fma2(yn, rr, pk(-q.w, -q.w));
        const u64 d2 = fma2(dx, dx, mul2(dy, dy));
        // cost_in_range(): floor(max(fma(-lam/T, d2, lam), 0) + 0.5); in range the max() never binds on the result
        // (v > -1e-4 there, and floor(v + 0.5) = 0 either way)
        const u64 v2 = fma2(KS2, d2, LAM2);
        u64 w2;
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(w2) : "l"(v2), "l"(HALF2));
        float da, db, wa, wb;
        upk(d2, da, db);
        upk(w2, wa, wb);
        const int ca = (da < Ta) ? (int)floorf(wa) : fara;   // NaN compares false -> far, as in the reference
        const int cb = (db < Tb) ? (int)floorf(wb) : farb;
        OutT* row = st + r * PITCH + off + tid;
        row[0] = (OutT)ca;
        row[256] = (OutT)cb;
        off = (off + lstep) & (EPV - 1);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged costs -> visible to the TMA engine
    dt_bar_arrive(1 + b);
  }
}

template <typename OutT>
static mh_status launch_tiled(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, OutT* d_cost,
                              const CostParams& cp) {
  constexpr int EPV = 16 / (int)sizeof(OutT);
  const size_t smem = (size_t)DT_STAGES * DT_P * (DT_KC + EPV) * sizeof(OutT);
  const int nchunks = (int)(((long long)K + 1 + DT_KC - 1) / DT_KC);
  const long long NB = (N + DT_P - 1) / DT_P;
  auto kern = cost_dense_tiled_kernel<OutT>;
  MH_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  MH_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, DT_THREADS, smem));
  const long long resident = (long long)std::max(1, occ) * ctx->sm_count;
  const unsigned gx = (unsigned)std::max<long long>(1, std::min<long long>(NB, resident / nchunks));
  kern<<<dim3(gx, (unsigned)nchunks), DT_THREADS, smem, ctx->stream>>>(d_pts, N, d_hyp, K, d_cost, cp);
  MH_LAUNCHED(ctx, "cost_dense_tiled_kernel");
  return MH_OK;
}

mh_status launch_cost_dense_tiled(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, void* d_cost,
                                  int elem_bytes) {
  const CostParams cp = cost_params(ctx);
  if ((reinterpret_cast<unsigned long long>(d_cost) % (unsigned)elem_bytes) != 0)
    return fail(ctx, MH_EINVAL, "mh_data_cost_dense: output pointer must be aligned to its element size");
  if (elem_bytes == 4) return launch_tiled<int32_t>(ctx, d_pts, N, d_hyp, K, (int32_t*)d_cost, cp);
  return launch_tiled<int16_t>(ctx, d_pts, N, d_hyp, K, (int16_t*)d_cost, cp);
}

}  // namespace mh
