// ============================================================================
// K2 v6 — data-term argmin + per-hypothesis inlier counts with the 3x3 projective
// product on the tensor cores (dataEnergy, MultiH/MultiH/MultiH.cpp:473-504; inlier
// scan :430-443).  Successor of v5 (cost_argmin_mma_kernel in k2_residual.cu).
//
//  * (s, xn, yn) = H (x, y, 1)^T runs as mma.sync.m16n8k8 TF32 with the 3xTF32 split
//    packed INTO k = 8 (one MMA = one FP32-grade 3-term dot product for 16
//    correspondences x 8 hypotheses).  The split B fragments are prepared ONCE per
//    launch by split_hyp_tf32_kernel and streamed into shared memory by the TMA engine
//    (cp.async.bulk + mbarrier, double-buffered chunks) — no thread touches the staging.
//  * each thread finishes 2 correspondences x 2 hypotheses per MMA triple on the FP32
//    pipe: 2 MUFU.RCP + 4 FFMA2 + 1 FADD2 per hypothesis pair, the argmin interval test
//    and the inlier test folded into the last FMA exactly as in v3.
//  * software pipeline: the MMAs of hypothesis block b+1 are issued before the FP32
//    epilogue of block b.
//  * the exact update of a record-breaking candidate is entered per (row-slot), warp-
//    uniformly, and only that slot is re-evaluated — with the dense kernel's FP32
//    instruction sequence, so (cost, label) stay bit-identical to cost_dense_kernel.
//    (v5 walked every slot of the warp whenever any lane held a candidate: as many
//    instructions as the main loop.)
// Per-correspondence argmin state is replicated in the 4 lanes of a quad and kept
// coherent by a quad min-reduce inside the update.
// ============================================================================
#include <cstdlib>

#include "k2_device.cuh"

namespace mh {

// hypotheses [K][12] -> split B fragments [Kpad][3][4] float2 (96 B per hypothesis):
//   [hyp][s|xn|yn][t] = {B[t][hyp], B[t+4][hyp]},  B[k] = [ahi, bhi, chi, ahi, alo, blo, bhi, clo]
// against A[k] = [xhi, yhi, 1, xlo, xhi, yhi, ylo, 1].  Rows >= K are "far" (residual ~1e36: never a hit, never an inlier).
//
// "Wild" hypotheses — some |h_i| > WILD_RATIO |h_8|, or non-finite entries (HAF estimates of outlier correspondences) —
// are kept off the tensor cores: the absolute error of a 3xTF32 product grows with |h_i| while the argmin filter's margin
// is relative to d2, so for them the filter could miss a record-breaking candidate.  They become "far" columns here, are
// appended to wild_list, and cost_argmin_wild_kernel evaluates them afterwards with the exact FP32 sequence.
constexpr float WILD_RATIO = 4.f;

__global__ void split_hyp_tf32_kernel(const float* __restrict__ hyp, int K, int Kpad, float4* __restrict__ out,
                                      int* __restrict__ wild_count, int* __restrict__ wild_list) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Kpad) return;
  float h[9];
  bool far = j >= K;
  if (!far) {
    const float4* p = reinterpret_cast<const float4*>(hyp + (size_t)j * 12);
    const float4 u = p[0], v = p[1], w = p[2];
    h[0] = u.x; h[1] = u.y; h[2] = u.z; h[3] = u.w; h[4] = v.x; h[5] = v.y; h[6] = v.z; h[7] = v.w; h[8] = w.x;
    float mx = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) mx = fmaxf(mx, fabsf(h[k]));
    if (!(mx <= WILD_RATIO * fabsf(h[8])) || !(fabsf(h[8]) < 3.0e38f)) {   // also catches NaN / Inf
      far = true;
      wild_list[atomicAdd(wild_count, 1)] = j;
    }
  }
  if (far) {
    h[0] = h[1] = h[3] = h[4] = h[6] = h[7] = 0.f; h[2] = h[5] = 1e18f; h[8] = 1.f;
  }
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    const int base = m == 0 ? 6 : m == 1 ? 0 : 3;  // s uses (h6,h7,h8); xn (h0,h1,h2); yn (h3,h4,h5)
    const float ha = h[base], hb = h[base + 1], hc = h[base + 2];
    const float ahi = __uint_as_float(to_tf32(ha)), bhi = __uint_as_float(to_tf32(hb)), chi = __uint_as_float(to_tf32(hc));
    const float alo = __uint_as_float(to_tf32(ha - ahi)), blo = __uint_as_float(to_tf32(hb - bhi)),
                clo = __uint_as_float(to_tf32(hc - chi));
    float4* d = out + ((size_t)j * 3 + m) * 2;
    d[0] = make_float4(ahi, alo, bhi, blo);
    d[1] = make_float4(chi, bhi, ahi, clo);
  }
}

// ---- mbarrier / bulk-copy primitives (sm_90+; the TMA engine's 1-D mode) ----------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  asm volatile("{\n\t.reg .pred p;\n\t"
               "W_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@!p bra W_%=;\n\t}"
               :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// non-volatile MMA: a pure function of its operands, so ptxas may hoist it above the previous block's epilogue
__device__ __forceinline__ void mma_tf32_nv(float (&d)[4], const unsigned (&a)[4], float b0, float b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)), "f"(0.f));
}

template <int PB>
struct Acc {
  float S[PB][4], XN[PB][4], YN[PB][4];
};

// WARPS warps per CTA, each owning 16 * PB correspondences; CH hypotheses per staged chunk.
// DEFER (v7): candidates that pass the interval test are not evaluated on the spot; they are appended to a per-warp
// queue in shared memory ((row, hypothesis pair), one ballot + one STS per row slot) and the queue is drained by the
// whole warp — one candidate per lane: exact FP32 re-evaluation, shared-memory atomicMin on the row's packed
// (cost, label) — whenever QTHR entries are waiting.  The rows' filters are refreshed from shared memory after a drain;
// until then they are stale, i.e. looser, so no candidate is missed and the result is unchanged.  v6 spent a third of
// its instructions in the warp-serial per-slot update; the queue makes that work lane-parallel.
// The queue is checked once per two 8-hypothesis blocks, so it holds QTHR waiting entries plus whatever two epilogues can
// add (2 blocks x 2 PB row slots x 32 lanes): no capacity test on the push path, and one inlined copy of the drain.
__host__ __device__ constexpr int queue_capacity(int PB) { return 32 + 2 * 2 * PB * 32; }

// LIST: the same kernel as the producer of sparse per-site cost lists — what a labeller consumes (mh_data_cost_fused with d_list).
// The filter is the constant interval 0 <= d2 < T (never tightened), every candidate pair goes through the queue, and the drain
// appends (label << 8 | cost) to the site's list (slot from a shared-memory counter per row; from the global counter when several
// CTAs share a row) next to the argmin update.  Entries are in drain order, not label order.
template <bool COUNT_INLIERS, int WARPS, int MINB, int PB, int CH, bool PIPE, int QTHR = 0, bool LIST = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
cost_argmin_tc_kernel(const float4* __restrict__ pts, long long N, const float* __restrict__ hyp,
                      const float2* __restrict__ hsplit, int K, int k_per_block, CostParams cp, FastOut o,
                      int use_atomic_best, int tiles_full, int tail_split, int kpb_tail) {
  constexpr int THREADS = WARPS * 32;
  constexpr unsigned CHUNK_BYTES = CH * 96;
  constexpr bool DEFER = QTHR > 0;
  constexpr int QCAP = queue_capacity(PB);
  static_assert(QTHR <= 32, "queue_capacity() assumes at most 32 waiting entries");
  static_assert(!LIST || DEFER, "the list member appends from the deferred queue");
  constexpr int WSTRIDE = QCAP + 16 * PB * (LIST ? 2 : 1);   // per warp: queue | packed best per row | LIST: entries per row
  // dynamic shared memory: [2][CH][3][4] float2 split fragments | [2][WARPS][CH] u8 per-warp inlier counts | 2 mbarriers
  //                        | DEFER: [WARPS][QCAP] candidate queues | [WARPS][16 PB] packed (cost << 16 | label) per row
  extern __shared__ __align__(128) unsigned char tc_smem[];
  float2* sB = reinterpret_cast<float2*>(tc_smem);
  unsigned char* sCnt = tc_smem + 2 * CHUNK_BYTES;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(tc_smem + 2 * CHUNK_BYTES + 2 * WARPS * CH);
  unsigned* sQ = reinterpret_cast<unsigned*>(tc_smem + 2 * CHUNK_BYTES + 2 * WARPS * CH + 16) + (threadIdx.x >> 5) * WSTRIDE;
  unsigned* sBest = sQ + QCAP;
  unsigned* sRowCnt = sBest + 16 * PB;   // LIST only
  int qcnt = 0;   // warp-uniform

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  constexpr int PTS_PER_WARP = 16 * PB, TILE = WARPS * PTS_PER_WARP;
  // Tail wave: the CTAs beyond the last full wave (blockIdx.x >= tiles_full) each take 1 / tail_split of the hypothesis range
  // of their tile, so the last, partly filled wave ends after ceil(tail tiles * split / slots) / split of a wave instead of a
  // whole one; their results merge through the same atomics as a K-split launch.
  int tile_idx = blockIdx.x, kbeg = blockIdx.y * k_per_block;          // kbeg: multiple of CH
  int kend = min(K, kbeg + k_per_block);
  if (tail_split > 1 && (int)blockIdx.x >= tiles_full) {
    const int part = blockIdx.x - tiles_full;
    tile_idx = tiles_full + part / tail_split;
    kbeg = (part % tail_split) * kpb_tail;
    kend = min(K, kbeg + kpb_tail);
    use_atomic_best = 1;
  }
  const long long tile0 = (long long)tile_idx * TILE + (long long)warp * PTS_PER_WARP;
  const int nchunks = (kend - kbeg + CH - 1) / CH;    // hsplit is padded to whole chunks

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], CHUNK_BYTES);
    bulk_g2s(sB, hsplit + (size_t)kbeg * 12, CHUNK_BYTES, &bars[0]);
  }

  // ---- per-thread correspondence state: rows g and g+8 of each of the warp's PB 16-row blocks ----------------------
  unsigned A[PB][4];
  // CM: the constant that turns the folded sum into d2 - mid for the argmin interval test.  With inlier counting the fold
  // is d2 - thr2 (its sign IS the inlier flag, independent of the row's filter state, hence of how rows are grouped into
  // warps) and CM = thr2 - mid is added afterwards; without it -mid is folded directly.
  float NX2[PB][2], NY2[PB][2], CM[PB][2], HALF[PB][2];
  auto set_filter = [&](int best_cost, float& cm, float& half) {
    float negmid;
    if (LIST) {   // every residual below the truncation threshold is a list entry: 0 <= d2 < 1.001 T, whatever the best so far
      const float hi = cp.T * 1.001f;
      negmid = -0.5f * hi;
      half = 0.5f * hi * 1.000001f + 1e-30f;
    } else {
      fast_thresholds(best_cost, cp, negmid, half);
    }
    cm = COUNT_INLIERS ? cp.thr2 + negmid : negmid;
  };
  unsigned BEST[PB][2];
  const unsigned best_init = ((unsigned)min(cp.cost_outlier, 0xffff) << 16);
#pragma unroll
  for (int pb = 0; pb < PB; ++pb)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const long long idx = tile0 + pb * 16 + g + 8 * r;
      const float4 q = pts[idx < N ? idx : N - 1];
      NX2[pb][r] = -q.z; NY2[pb][r] = -q.w;
      BEST[pb][r] = best_init;
      set_filter(cp.cost_outlier, CM[pb][r], HALF[pb][r]);
      if (idx >= N) { HALF[pb][r] = -1.f; NX2[pb][r] = NY2[pb][r] = 1e18f; }   // padding rows: d2 ~ 1e36, never a hit
      const unsigned xhi = to_tf32(q.x), yhi = to_tf32(q.y);
      const unsigned xlo = to_tf32(q.x - __uint_as_float(xhi)), ylo = to_tf32(q.y - __uint_as_float(yhi));
      const unsigned one = 0x3f800000u;
      A[pb][r] = t == 0 ? xhi : t == 1 ? yhi : t == 2 ? one : xlo;      // a0 (row g) / a1 (row g+8): k = t
      A[pb][2 + r] = t == 0 ? xhi : t == 1 ? yhi : t == 2 ? ylo : one;  // a2 / a3: k = t + 4
    }
  const u64 ONE2 = pk(1.f, 1.f), NEGTHR2 = pk(-cp.thr2, -cp.thr2);
  if (DEFER) {
    for (int i = lane; i < 16 * PB; i += 32) {
      sBest[i] = best_init;
      if (LIST) sRowCnt[i] = 0u;
    }
    __syncwarp();
  }

  auto mma_block = [&](const float2* bp, Acc<PB>& a) {   // bp -> this thread's fragment of an 8-hypothesis block
    const float2 bs = bp[0], bx = bp[4], by = bp[8];
#pragma unroll
    for (int pb = 0; pb < PB; ++pb) {
      mma_tf32_nv(a.S[pb], A[pb], bs.x, bs.y);
      mma_tf32_nv(a.XN[pb], A[pb], bx.x, bx.y);
      mma_tf32_nv(a.YN[pb], A[pb], by.x, by.y);
    }
  };

  // exact re-evaluation of hypotheses (ih0, ih0 + 1) for row slot (pb, r): entered warp-uniformly; lanes without a
  // candidate only take part in the quad reduction
  auto update_slot = [&](int pb, int r, bool mine_flag, int ih0, float& cm, float& half, unsigned& best) {
    unsigned mine = 0xffffffffu;
    if (mine_flag) {
      const long long idx = tile0 + pb * 16 + g + 8 * r;   // < N: padding rows never raise a flag
      const float4 q = __ldg(pts + idx);
      float ha[9], hb[9];
      {
        const float4* hp = reinterpret_cast<const float4*>(hyp + (size_t)min(ih0, K - 1) * 12);
        const float4 u = __ldg(hp), v = __ldg(hp + 1), w = __ldg(hp + 2);
        ha[0] = u.x; ha[1] = u.y; ha[2] = u.z; ha[3] = u.w; ha[4] = v.x; ha[5] = v.y; ha[6] = v.z; ha[7] = v.w; ha[8] = w.x;
      }
      {
        const float4* hp = reinterpret_cast<const float4*>(hyp + (size_t)min(ih0 + 1, K - 1) * 12);
        const float4 u = __ldg(hp), v = __ldg(hp + 1), w = __ldg(hp + 2);
        hb[0] = u.x; hb[1] = u.y; hb[2] = u.z; hb[3] = u.w; hb[4] = v.x; hb[5] = v.y; hb[6] = v.z; hb[7] = v.w; hb[8] = w.x;
      }
      const float da = residual(ha, q.x, q.y, q.z, q.w);
      const float db = residual(hb, q.x, q.y, q.z, q.w);
      if (ih0 < kend && da < cp.T) mine = ((unsigned)cost_in_range(da, cp) << 16) | (unsigned)(ih0 + 1);
      if (ih0 + 1 < kend && db < cp.T) mine = min(mine, ((unsigned)cost_in_range(db, cp) << 16) | (unsigned)(ih0 + 2));
    }
    mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, 1));  // the 4 lanes of a quad share rows g, g+8
    mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, 2));
    if (mine < best) {   // labels rise along the loop for a given correspondence: only strictly cheaper candidates matter
      best = mine;
      set_filter((int)(mine >> 16), cm, half);
    }
  };

  // DEFER: the whole warp evaluates the queued candidates, one per lane, then every thread refreshes its rows' filters
  auto drain = [&]() {
    __syncwarp();
    for (int i = lane; i < qcnt; i += 32) {
      const unsigned e = sQ[i];
      const int row = (int)(e >> 24), ih0 = (int)(e & 0xffffffu);
      const float4 q = __ldg(pts + tile0 + row);   // < N: padding rows never raise a flag
      float ha[9], hb[9];
      {
        const float4* hp = reinterpret_cast<const float4*>(hyp + (size_t)min(ih0, K - 1) * 12);
        const float4 u = __ldg(hp), v = __ldg(hp + 1), w = __ldg(hp + 2);
        ha[0] = u.x; ha[1] = u.y; ha[2] = u.z; ha[3] = u.w; ha[4] = v.x; ha[5] = v.y; ha[6] = v.z; ha[7] = v.w; ha[8] = w.x;
      }
      {
        const float4* hp = reinterpret_cast<const float4*>(hyp + (size_t)min(ih0 + 1, K - 1) * 12);
        const float4 u = __ldg(hp), v = __ldg(hp + 1), w = __ldg(hp + 2);
        hb[0] = u.x; hb[1] = u.y; hb[2] = u.z; hb[3] = u.w; hb[4] = v.x; hb[5] = v.y; hb[6] = v.z; hb[7] = v.w; hb[8] = w.x;
      }
      const float da = residual(ha, q.x, q.y, q.z, q.w);
      const float db = residual(hb, q.x, q.y, q.z, q.w);
      unsigned mine = 0xffffffffu;
      if (ih0 < kend && da < cp.T) mine = ((unsigned)cost_in_range(da, cp) << 16) | (unsigned)(ih0 + 1);
      if (ih0 + 1 < kend && db < cp.T) mine = min(mine, ((unsigned)cost_in_range(db, cp) << 16) | (unsigned)(ih0 + 2));
      if (LIST) {
        const long long idx = tile0 + row;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          const float d = which ? db : da;
          // a wild hypothesis (off the tensor cores, see split_hyp_tf32_kernel) is listed by cost_argmin_wild_kernel; it shows up
          // here when its pair partner raised the flag
          const float* hh = which ? hb : ha;
          float mx = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) mx = fmaxf(mx, fabsf(hh[k]));
          const bool wild = !(mx <= WILD_RATIO * fabsf(hh[8])) || !(fabsf(hh[8]) < 3.0e38f);
          if (!wild && ih0 + which < kend && d < cp.T) {
            const int slot = use_atomic_best ? atomicAdd(o.list_count + idx, 1) : (int)atomicAdd(sRowCnt + row, 1u);
            if (slot < o.kmax) o.list[idx * o.kmax + slot] = ((unsigned)(ih0 + which + 1) << 8) | (unsigned)cost_in_range(d, cp);
          }
        }
      }
      if (mine != 0xffffffffu) atomicMin(sBest + row, mine);   // (cost, label) packed: ties keep the lowest label
    }
    __syncwarp();
    if (!LIST) {
#pragma unroll
      for (int pb = 0; pb < PB; ++pb)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const unsigned b = sBest[pb * 16 + g + 8 * r];
          if (b < BEST[pb][r]) {
            BEST[pb][r] = b;
            set_filter((int)(b >> 16), CM[pb][r], HALF[pb][r]);
          }
        }
    }
    qcnt = 0;
  };

  // FP32 epilogue of one 8-hypothesis block; returns the thread's inlier counts {popc(col 2t), popc(col 2t+1)} in two bytes
  auto epilogue = [&](const Acc<PB>& a, int ih_block) -> unsigned {
    unsigned mask0 = 0u, mask1 = 0u;   // v6: shifted-in sign bits; DEFER: running counts (LEA.HI, no POPC)
    bool flag[PB][2];
    bool any = false;
#pragma unroll
    for (int pb = 0; pb < PB; ++pb)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const u64 rr = pk(rcp_approx(a.S[pb][2 * r]), rcp_approx(a.S[pb][2 * r + 1]));
        const u64 dx = fma2(pk(a.XN[pb][2 * r], a.XN[pb][2 * r + 1]), rr, pk(NX2[pb][r], NX2[pb][r]));
        const u64 dy = fma2(pk(a.YN[pb][2 * r], a.YN[pb][2 * r + 1]), rr, pk(NY2[pb][r], NY2[pb][r]));
        float ta, tb;
        if (COUNT_INLIERS) {
          const u64 vv = fma2(dx, dx, fma2(dy, dy, NEGTHR2));               // d2 - thr2: sign bit = inlier
          upk(fma2(vv, ONE2, pk(CM[pb][r], CM[pb][r])), ta, tb);             // d2 - mid
          float va, vb;
          upk(vv, va, vb);
          if (DEFER) {
            mask0 += __float_as_uint(va) >> 31;
            mask1 += __float_as_uint(vb) >> 31;
          } else {
            mask0 = __funnelshift_l(__float_as_uint(va), mask0, 1);
            mask1 = __funnelshift_l(__float_as_uint(vb), mask1, 1);
          }
        } else {
          upk(fma2(dx, dx, fma2(dy, dy, pk(CM[pb][r], CM[pb][r]))), ta, tb);   // d2 - mid
        }
        flag[pb][r] = fminf(fabsf(ta), fabsf(tb)) < HALF[pb][r];
        any = any || flag[pb][r];
      }
    if (__any_sync(0xffffffffu, any)) {   // rare after warm-up
#pragma unroll
      for (int pb = 0; pb < PB; ++pb)
#pragma unroll
        for (int r = 0; r < 2; ++r)
        {
          if (DEFER) {
            const unsigned m = __ballot_sync(0xffffffffu, flag[pb][r]);
            if (m) {
              const int n = __popc(m);
              if (flag[pb][r])
                sQ[qcnt + __popc(m & ((1u << lane) - 1u))] = ((unsigned)(pb * 16 + g + 8 * r) << 24) | (unsigned)(ih_block + 2 * t);
              qcnt += n;
            }
          } else if (__any_sync(0xffffffffu, flag[pb][r])) {
            update_slot(pb, r, flag[pb][r], ih_block + 2 * t, CM[pb][r], HALF[pb][r], BEST[pb][r]);
          }
        }
    }
    if (!COUNT_INLIERS) return 0u;
    return DEFER ? (mask0 | (mask1 << 8)) : (__popc(mask0) | (__popc(mask1) << 8));
  };

  for (int ci = 0; ci < nchunks; ++ci) {
    const int c0 = kbeg + ci * CH;
    const int buf = ci & 1;
    // buffer buf^1 (fragments and counters) was last used by chunk ci-1; every thread passed the barrier that ended it
    if (threadIdx.x == 0 && ci + 1 < nchunks) {
      mbar_expect_tx(&bars[buf ^ 1], CHUNK_BYTES);
      bulk_g2s(sB + (size_t)(buf ^ 1) * CH * 12, hsplit + (size_t)(c0 + CH) * 12, CHUNK_BYTES, &bars[buf ^ 1]);
    }
    mbar_wait(&bars[buf], (ci >> 1) & 1);

    const float2* bp = sB + (size_t)buf * CH * 12 + g * 12 + t;   // + 96 float2 per 8-hypothesis block
    unsigned char* cw = sCnt + ((size_t)buf * WARPS + warp) * CH + 2 * t;
    constexpr int NBLK = CH / 8;

    Acc<PB> acc0, acc1;
    if (PIPE) mma_block(bp, acc0);
    if (DEFER) {
      // Two epilogues per loop trip.  ptxas folds a broadcast constant (-x2, -y2, -mid, mid - thr2) into FFMA2's scalar-
      // operand form only when its {c, c} pack has a single use; with two uses it rebuilds register pairs in the loop
      // (1.5 MOVs per residual, measured).  The empty asm between the two epilogues makes the second block's constants
      // formally new values, so each pack keeps a single use — no instruction is emitted for it.
#pragma unroll 1
      for (int hb = 0; hb < NBLK; hb += 2) {
        mma_block(bp + hb * 96, acc0);
        unsigned packed = epilogue(acc0, c0 + hb * 8);
#pragma unroll
        for (int pb = 0; pb < PB; ++pb)
#pragma unroll
          for (int r = 0; r < 2; ++r)
            asm volatile("" : "+f"(NX2[pb][r]), "+f"(NY2[pb][r]), "+f"(CM[pb][r]));
        mma_block(bp + (hb + 1) * 96, acc0);
        packed |= epilogue(acc0, c0 + hb * 8 + 8) << 16;
        if (qcnt >= QTHR) drain();
        if (COUNT_INLIERS) {
          packed += __shfl_xor_sync(0xffffffffu, packed, 4);
          packed += __shfl_xor_sync(0xffffffffu, packed, 8);
          packed += __shfl_xor_sync(0xffffffffu, packed, 16);
          *reinterpret_cast<unsigned short*>(cw + hb * 8) = (unsigned short)packed;
          *reinterpret_cast<unsigned short*>(cw + hb * 8 + 8) = (unsigned short)(packed >> 16);
        }
      }
    } else {
#pragma unroll 1
    for (int hb = 0; hb < NBLK; hb += 2) {
      unsigned packed;
      if (PIPE) {   // the MMAs of the next block are in flight while this block's FP32 epilogue runs
        mma_block(bp + (hb + 1) * 96, acc1);
        packed = epilogue(acc0, c0 + hb * 8);
        if (hb + 2 < NBLK) mma_block(bp + (hb + 2) * 96, acc0);
        packed |= epilogue(acc1, c0 + hb * 8 + 8) << 16;
      } else {      // one accumulator set: latency is hidden by the other resident warps
        mma_block(bp + hb * 96, acc0);
        packed = epilogue(acc0, c0 + hb * 8);
        mma_block(bp + (hb + 1) * 96, acc0);
        packed |= epilogue(acc0, c0 + hb * 8 + 8) << 16;
      }
      if (COUNT_INLIERS) {
        // sum the byte fields over the 8 lanes that share t (<= 2 PB per lane, <= 16 PB <= 64 per field): 3 shuffles; every
        // lane then holds the warp's counts of hypotheses (2t, 2t+1) of both blocks and stores them as bytes (the 8 lanes of
        // a t-group write the same value to the same address; each (warp, hypothesis) is visited once per chunk)
        packed += __shfl_xor_sync(0xffffffffu, packed, 4);
        packed += __shfl_xor_sync(0xffffffffu, packed, 8);
        packed += __shfl_xor_sync(0xffffffffu, packed, 16);
        *reinterpret_cast<unsigned short*>(cw + hb * 8) = (unsigned short)packed;
        *reinterpret_cast<unsigned short*>(cw + hb * 8 + 8) = (unsigned short)(packed >> 16);
      }
    }
    }
    __syncthreads();   // every warp is done with fragment buffer buf; its counters are complete
    if (COUNT_INLIERS) {
      const int nh = min(CH, kend - c0);
      const unsigned char* cb = sCnt + (size_t)buf * WARPS * CH;
      for (int j = threadIdx.x; j < nh; j += THREADS) {
        int v = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) v += cb[w * CH + j];
        if (v) atomicAdd(o.inlier_count + c0 + j, v);
      }
    }
  }
  if (DEFER && qcnt > 0) drain();
  if (LIST && t == 0 && !use_atomic_best) {
#pragma unroll
    for (int pb = 0; pb < PB; ++pb)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const long long idx = tile0 + pb * 16 + g + 8 * r;
        if (idx < N) o.list_count[idx] = (int)sRowCnt[pb * 16 + g + 8 * r];
      }
  }
  if (o.best && t == 0) {  // the quad holds identical state: lane t == 0 writes
#pragma unroll
    for (int pb = 0; pb < PB; ++pb)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const long long idx = tile0 + pb * 16 + g + 8 * r;
        const unsigned b = LIST ? sBest[pb * 16 + g + 8 * r] : BEST[pb][r];
        if (idx < N && (b & 0xffffu) != 0u) {
          const u64 v = ((u64)(b >> 16) << 32) | (u64)(b & 0xffffu);
          if (use_atomic_best) atomicMin(o.best + idx, v);
          else o.best[idx] = v;
        }
      }
  }
}

// Exact FP32 pass over the wild hypotheses (a handful per thousand): one correspondence per thread, the dense kernel's
// instruction sequence, merged into the tensor-core kernel's result with a 64-bit atomicMin on (cost << 32 | label) —
// a total order, so the merged argmin is the same as a single pass over all hypotheses.
__global__ void __launch_bounds__(256) cost_argmin_wild_kernel(const float4* __restrict__ pts, long long N,
                                                                const float* __restrict__ hyp,
                                                                const int* __restrict__ wild_count,
                                                                const int* __restrict__ wild_list, CostParams cp, FastOut o) {
  const int nw = *wild_count;
  if (nw == 0) return;
  const int lane = threadIdx.x & 31;
  for (long long base = (long long)blockIdx.x * blockDim.x; base < N; base += (long long)gridDim.x * blockDim.x) {
    const long long idx = base + threadIdx.x;
    const bool live = idx < N;
    const float4 q = pts[live ? idx : N - 1];
    unsigned bc = (unsigned)cp.cost_outlier, bl = 0u;
    for (int j = 0; j < nw; ++j) {
      const int k = wild_list[j];
      const float4* hp = reinterpret_cast<const float4*>(hyp + (size_t)k * 12);
      const float4 u = __ldg(hp), v = __ldg(hp + 1), w = __ldg(hp + 2);
      const float h[9] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w, w.x};
      const float d2 = residual(h, q.x, q.y, q.z, q.w);
      if (live && d2 < cp.T) {
        const unsigned c = (unsigned)cost_in_range(d2, cp);
        if (c < bc || (c == bc && (unsigned)(k + 1) < bl)) { bc = c; bl = (unsigned)(k + 1); }   // list order is arbitrary
        if (o.list) {
          const int slot = atomicAdd(o.list_count + idx, 1);
          if (slot < o.kmax) o.list[idx * o.kmax + slot] = ((unsigned)(k + 1) << 8) | c;
        }
      }
      if (o.inlier_count) {
        const unsigned m = __ballot_sync(0xffffffffu, live && d2 < cp.thr2);
        if (lane == 0 && m) atomicAdd(o.inlier_count + k, __popc(m));
      }
    }
    if (o.best && live && bl != 0u) atomicMin(o.best + idx, ((u64)bc << 32) | (u64)bl);
  }
}

// the side pass as a launcher (also used by k2_tmem.cu)
mh_status launch_cost_argmin_wild(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, const int* d_wild_count,
                                  const int* d_wild_list, const CostParams& cp, const FastOut& fo) {
  const unsigned wild_grid = (unsigned)std::min<long long>((N + 255) / 256, 8LL * ctx->sm_count);
  cost_argmin_wild_kernel<<<wild_grid, 256, 0, ctx->stream>>>(d_pts, N, d_hyp, d_wild_count, d_wild_list, cp, fo);
  MH_LAUNCHED(ctx, "cost_argmin_wild_kernel");
  return MH_OK;
}

template <bool COUNT_INLIERS, int WARPS, int MINB, int PB, int CH, bool PIPE, int QTHR = 0, bool LIST = false>
static mh_status launch_tc(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, const CostParams& cp,
                           const FastOut& fo) {
  const int want = 2 * ctx->sm_count;
  const int tile = WARPS * 16 * PB;
  const unsigned tiles = (unsigned)((N + tile - 1) / tile);
  int ks = 1;
  if ((int)tiles < want) ks = std::min((K + CH - 1) / CH, (want + (int)tiles - 1) / (int)tiles);
  if (const char* e = std::getenv("MH_K2_KSPLIT")) ks = std::min((K + CH - 1) / CH, std::max(1, std::atoi(e)));   // measurement aid
  ks = std::max(1, ks);
  int kpb = (K + ks - 1) / ks;
  kpb = ((kpb + CH - 1) / CH) * CH;   // whole chunks per CTA
  ks = (K + kpb - 1) / kpb;
  int Kpad = ks * kpb;
  // multi-wave launches: split the hypothesis range of the tiles of the last, partial wave (see the kernel)
  int tiles_full = (int)tiles, tail_split = 1, kpb_tail = kpb;
  unsigned grid_x = tiles;
  const int slots = MINB * ctx->sm_count;
  static const bool tail_on = !(std::getenv("MH_TAIL_SPLIT") && std::atoi(std::getenv("MH_TAIL_SPLIT")) == 0);   // measurement aid
  if (tail_on && !LIST && ks == 1 && (int)tiles > slots && tiles % slots != 0) {
    const int tail = (int)(tiles % slots);
    double best_t = 1.0;
    for (int sp = 2; sp <= 16 && sp * CH <= K; sp *= 2) {
      const int kp = (((K + sp - 1) / sp + CH - 1) / CH) * CH;
      if ((sp - 1) * kp >= K) continue;   // every part must own at least one hypothesis
      const double t = (double)((tail * sp + slots - 1) / slots) / sp + 0.01 * sp;   // + a small per-part overhead
      if (t < best_t - 1e-9) { best_t = t; tail_split = sp; }
    }
    if (tail_split > 1) {
      tiles_full = (int)tiles - tail;
      kpb_tail = (((K + tail_split - 1) / tail_split + CH - 1) / CH) * CH;
      Kpad = std::max(Kpad, tail_split * kpb_tail);
      grid_x = (unsigned)tiles_full + (unsigned)tail * (unsigned)tail_split;
    }
  }
  // scratch: [Kpad x 96 B split fragments][wild count (16 B)][K wild indices]
  MH_TRY(ensure_scratch(ctx, (uint64_t)Kpad * 96 + 16 + (uint64_t)K * 4));
  float2* d_split = (float2*)ctx->scratch;
  int* d_wild_count = (int*)((char*)ctx->scratch + (size_t)Kpad * 96);
  int* d_wild_list = d_wild_count + 4;
  MH_CUDA(ctx, cudaMemsetAsync(d_wild_count, 0, 16, ctx->stream));
  split_hyp_tf32_kernel<<<(unsigned)((Kpad + 127) / 128), 128, 0, ctx->stream>>>(d_hyp, K, Kpad, (float4*)d_split, d_wild_count,
                                                                                d_wild_list);
  MH_LAUNCHED(ctx, "split_hyp_tf32_kernel");
  const size_t smem = 2 * (size_t)CH * 96 + 2 * (size_t)WARPS * CH + 16 + (QTHR > 0 ? (size_t)WARPS * (queue_capacity(PB) + 16 * PB * (LIST ? 2 : 1)) * 4 : 0);
  auto kern = cost_argmin_tc_kernel<COUNT_INLIERS, WARPS, MINB, PB, CH, PIPE, QTHR, LIST>;
  MH_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(grid_x, (unsigned)ks), WARPS * 32, smem, ctx->stream>>>(d_pts, N, d_hyp, d_split, K, kpb, cp, fo, ks > 1, tiles_full, tail_split,
                                                                      kpb_tail);
  MH_LAUNCHED(ctx, "cost_argmin_tc_kernel");
  const unsigned wild_grid = (unsigned)std::min<long long>((N + 255) / 256, 8LL * ctx->sm_count);
  cost_argmin_wild_kernel<<<wild_grid, 256, 0, ctx->stream>>>(d_pts, N, d_hyp, d_wild_count, d_wild_list, cp, fo);
  MH_LAUNCHED(ctx, "cost_argmin_wild_kernel");
  return MH_OK;
}

// the list-producing member (mh_data_cost_fused with d_list): config 55's shape with the append-drain
mh_status launch_cost_list_tc(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, const CostParams& cp,
                              const FastOut& fo) {
  return fo.inlier_count ? launch_tc<true, 4, 5, 3, 128, false, 32, true>(ctx, d_pts, N, d_hyp, K, cp, fo)
                         : launch_tc<false, 4, 5, 3, 128, false, 32, true>(ctx, d_pts, N, d_hyp, K, cp, fo);
}

// config = 30 + i, see the table; returns MH_EINVAL for an unknown config
mh_status launch_cost_argmin_tc(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, const CostParams& cp,
                                const FastOut& fo, int config) {
  const bool cnt = fo.inlier_count != nullptr;
#define TC_CASE(id, WARPS, MINB, PB, CH, PIPE, ...)                                                              \
  case id:                                                                                                       \
    return cnt ? launch_tc<true, WARPS, MINB, PB, CH, PIPE, ##__VA_ARGS__>(ctx, d_pts, N, d_hyp, K, cp, fo)      \
               : launch_tc<false, WARPS, MINB, PB, CH, PIPE, ##__VA_ARGS__>(ctx, d_pts, N, d_hyp, K, cp, fo);
  switch (config) {   // (warps per CTA, CTAs per SM, 16-row blocks per warp, hypotheses per chunk, explicit pipeline)
#ifdef MH_TUNING
    TC_CASE(30, 8, 3, 2, 256, false)
    TC_CASE(31, 8, 2, 2, 256, true)
    TC_CASE(32, 4, 6, 2, 128, false)
    TC_CASE(33, 4, 4, 3, 256, false)
    TC_CASE(34, 8, 2, 3, 256, false)
    TC_CASE(35, 4, 3, 4, 256, true)
    TC_CASE(36, 8, 1, 4, 256, true)
    TC_CASE(37, 4, 5, 2, 128, true)
    TC_CASE(38, 8, 4, 1, 128, true)
    TC_CASE(39, 4, 4, 2, 256, true)
    TC_CASE(40, 4, 7, 2, 128, false)
    TC_CASE(41, 4, 8, 1, 128, false)
    TC_CASE(42, 8, 3, 2, 128, false)
    TC_CASE(43, 2, 12, 2, 64, false)
    TC_CASE(44, 4, 6, 2, 64, false)
    TC_CASE(45, 4, 4, 3, 128, false)
    TC_CASE(46, 2, 8, 3, 64, false)
    // v7: deferred candidate queue (last argument = drain threshold)
    TC_CASE(50, 4, 7, 2, 128, false, 24)
    TC_CASE(51, 4, 8, 2, 128, false, 24)
    TC_CASE(53, 8, 3, 2, 256, false, 24)
    TC_CASE(54, 8, 4, 2, 128, false, 24)
#endif
    TC_CASE(52, 4, 6, 2, 128, false, 24)
    TC_CASE(55, 4, 5, 3, 128, false, 24)
#ifdef MH_TUNING
    TC_CASE(56, 4, 4, 4, 128, false, 24)
    TC_CASE(57, 4, 4, 4, 256, false, 24)
    TC_CASE(58, 8, 2, 4, 512, false, 24)
    TC_CASE(59, 4, 5, 3, 256, false, 24)
    TC_CASE(60, 8, 4, 2, 256, false, 24)
    TC_CASE(61, 4, 6, 3, 128, false, 24)
    TC_CASE(62, 8, 3, 3, 256, false, 24)
    TC_CASE(63, 4, 8, 2, 128, false, 32)
    TC_CASE(64, 4, 8, 2, 64, false, 24)
#endif
    default: return fail(ctx, MH_EINVAL, "unknown tensor-core fast-path config");
  }
#undef TC_CASE
}

}  // namespace mh
