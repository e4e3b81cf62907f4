// ============================================================================
// K2 v8 — data-term argmin + per-hypothesis inlier counts with the projective product on the
// 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in tensor memory).
// dataEnergy, MultiH/MultiH/MultiH.cpp:473-504; inlier scan :430-443.  The A/B partner of v7
// (cost_argmin_tc_kernel, mma.sync, k2_mma.cu): same 3xTF32-in-K product, same FP32 epilogue
// arithmetic, same filter + deferred exact update, hence the same bit-exact (cost, label).
//
// What changes is who holds the accumulators and who issues the MMAs:
//   * one CTA per SM; G epilogue warpgroups (4 warps = the 128 TMEM lanes) + one producer warp;
//   * the CTA's correspondences live in shared memory for the whole launch as UMMA A operands
//     (G x R tiles of 128 rows x K = 8: [x_hi, y_hi, 1, x_lo | x_hi, y_hi, y_lo, 1], K-major, no swizzle);
//   * hypotheses arrive by the TMA engine's bulk copy as ready-made UMMA B operands (128-hypothesis
//     super-chunks, 3-slot ring), N = 3 CHH columns per instruction: [s | x_n | y_n] of CHH hypotheses;
//   * ONE thread issues tcgen05.mma (M = 128, N = 3 CHH, K = 8) per (row tile, hypothesis chunk) into that
//     warpgroup's TMEM buffer and commits to an mbarrier; no warp spends issue slots or registers on
//     MMA fragments (v7: 0.75 mma.sync + 0.25 LDS per residual and 48 accumulator registers);
//   * an epilogue thread owns ONE correspondence per row tile (its TMEM lane) and reads CHH hypotheses'
//     (s, x_n, y_n) with three tcgen05.ld — the row's filter constants are thread-uniform, so the
//     interval test is one 3-input min per hypothesis pair and one compare per chunk;
//   * loop order: hypothesis chunk outer, row tile inner — the per-hypothesis inlier counters of a chunk
//     stay in the thread's registers across its R rows and are reduced over the warp once per chunk
//     (REDUX), summed per CTA in shared memory, flushed to global once per launch.
// ============================================================================
#include "k2_device.cuh"

namespace mh {

constexpr float WILD_RATIO_T = 4.f;   // as k2_mma.cu: |h_i| > 4 |h_8| or non-finite -> exact FP32 side pass
constexpr int TM_SCH = 128;           // hypotheses per TMA super-chunk

// all shared-memory operands below are 32-bit shared-window addresses, computed once per thread
__device__ __forceinline__ unsigned tm_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tm_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tm_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tm_mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tm_mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra W_%=;\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// the producers' form: the thread may stay suspended up to ~4 us per try (it is woken when the phase completes), so a waiting
// producer warp does not spend the epilogue warps' issue slots on polling
__device__ __forceinline__ void tm_mbar_wait_suspended(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@!p bra W_%=;\n\t}" ::"r"(bar),
      "r"(parity), "r"(4000u)
      : "memory");
}
__device__ __forceinline__ void tm_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool tm_elect_one() {   // one lane of a converged warp
  unsigned pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float4 tm_lds128(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned tm_selp(unsigned a, unsigned b, bool p) {   // p ? a : b, kept a register select
  unsigned r;
  asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\tselp.u32 %0, %1, %2, q;\n\t}" : "=r"(r) : "r"(a), "r"(b), "r"((unsigned)p));
  return r;
}
// shared-memory matrix descriptor: start address, leading / stride byte offsets (16-byte units), version 1, no swizzle
__device__ __forceinline__ unsigned long long tm_smem_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
  const unsigned long long a = (saddr & 0x3ffffu) >> 4;
  return a | ((unsigned long long)(lbo_bytes >> 4) << 16) | ((unsigned long long)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: FP32 accumulate, TF32 x TF32, both operands K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr unsigned tm_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void tm_mma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void tm_mma_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (no wait: the caller issues tcgen05.wait::ld once)
__device__ __forceinline__ void tm_ld16(unsigned taddr, float* v) {
  unsigned r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 8 consecutive 32-bit columns -> 8 registers per thread, asynchronous until tm_wait_ld
__device__ __forceinline__ void tm_ld8(unsigned taddr, unsigned* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
// tcgen05.wait::ld with the 24 loaded registers tied through it, so that no use of them can be scheduled above the wait
__device__ __forceinline__ void tm_wait_ld(unsigned (&r)[24]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23])
               :
               : "memory");
}

// hypotheses [K][12] -> UMMA B operands.  Per super-chunk of 128 hypotheses: [128 / CHH sub-chunks][2 K-chunks][3 CHH rows][4 tf32],
// row n = m CHH + j (m: 0 = s (h6,h7,h8), 1 = x_n (h0,h1,h2), 2 = y_n (h3,h4,h5); j = hypothesis within the sub-chunk),
// K-chunk 0 = (a_hi, b_hi, c_hi, a_hi), K-chunk 1 = (a_lo, b_lo, b_hi, c_lo) against A = (x_hi, y_hi, 1, x_lo | x_hi, y_hi, y_lo, 1).
// Padding and wild hypotheses become "far" columns (residual ~1e36); the wild ones are listed for the exact side pass.
template <int CHH>
__global__ void split_hyp_umma_kernel(const float* __restrict__ hyp, int K, int Kpad, float4* __restrict__ out,
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
    if (!(mx <= WILD_RATIO_T * fabsf(h[8])) || !(fabsf(h[8]) < 3.0e38f)) {
      far = true;
      wild_list[atomicAdd(wild_count, 1)] = j;
    }
  }
  if (far) {
    h[0] = h[1] = h[3] = h[4] = h[6] = h[7] = 0.f; h[2] = h[5] = 1e18f; h[8] = 1.f;
  }
  const int sc = j / TM_SCH, jj = j % TM_SCH, sub = jj / CHH, jl = jj % CHH;
  float4* base = out + ((size_t)sc * (TM_SCH / CHH) + sub) * 2 * (3 * CHH);
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    const int b0 = m == 0 ? 6 : m == 1 ? 0 : 3;
    const float ha = h[b0], hb = h[b0 + 1], hc = h[b0 + 2];
    const float ahi = __uint_as_float(to_tf32(ha)), bhi = __uint_as_float(to_tf32(hb)), chi = __uint_as_float(to_tf32(hc));
    const float alo = __uint_as_float(to_tf32(ha - ahi)), blo = __uint_as_float(to_tf32(hb - bhi)),
                clo = __uint_as_float(to_tf32(hc - chi));
    base[m * CHH + jl] = make_float4(ahi, bhi, chi, ahi);
    base[3 * CHH + m * CHH + jl] = make_float4(alo, blo, bhi, clo);
  }
}

mh_status launch_cost_argmin_wild(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, const int* d_wild_count,
                                  const int* d_wild_list, const CostParams& cp, const FastOut& fo);   // k2_mma.cu

template <int G, int R, int CHH, int NBUF, int PWARPS>
struct TmLayout {
  static constexpr int ROWS = G * R * 128;
  static constexpr int PW = PWARPS;   // producer warps: each issues the MMAs of G / PW warpgroups
  static constexpr int THREADS = G * 128 + PW * 32;
  static constexpr int QTHR = 24;
  static constexpr int QCAP = ((QTHR + (CHH / 2) * 32 + 7) / 8) * 8;
  static constexpr size_t A_BYTES = (size_t)ROWS * 32;
  static constexpr size_t B_BYTES = 3 * (size_t)TM_SCH * 96;
  static constexpr size_t OFF_B = A_BYTES;
  static constexpr size_t OFF_STATE = OFF_B + B_BYTES;
  static constexpr size_t OFF_BEST = OFF_STATE + (size_t)ROWS * 16;
  static constexpr size_t OFF_Q = OFF_BEST + (size_t)ROWS * 4;
  static constexpr size_t OFF_BARS = OFF_Q + (size_t)G * 4 * QCAP * 4;
  static constexpr int NBARS = 2 * G * NBUF + 6;   // full, empty | super-chunk landed [3], super-chunk released [3]
  static constexpr size_t OFF_CNT = OFF_BARS + (size_t)(NBARS + 1) * 8;   // + tmem base word
  static size_t bytes(int kpb, bool count) { return OFF_CNT + (count ? (size_t)kpb * 4 : 0); }
};

template <bool COUNT_INLIERS, int G, int R, int CHH, int NBUF, int PWARPS>
__global__ void __launch_bounds__(TmLayout<G, R, CHH, NBUF, PWARPS>::THREADS, 1)
cost_argmin_tmem_kernel(const float4* __restrict__ pts, long long N, const float* __restrict__ hyp, const float4* __restrict__ hsplit,
                        int K, int k_per_block, CostParams cp, FastOut o, int use_atomic_best) {
  using L = TmLayout<G, R, CHH, NBUF, PWARPS>;
  static_assert(CHH == 16, "the epilogue is written for 16 hypotheses per MMA (three tcgen05.ld x16)");
  static_assert(G * NBUF * 3 * CHH <= 512, "TMEM has 512 columns");
  static_assert(R <= 7, "per-chunk inlier counts are reduced in 8-bit fields: 32 lanes x R <= 255");
  static_assert(G % L::PW == 0, "warpgroups split evenly over the producer warps");
  constexpr int NCOL = 3 * CHH, PW = L::PW, GP = G / PW;
  constexpr int QTHR = L::QTHR, QCAP = L::QCAP;
  extern __shared__ __align__(128) unsigned char tm_smem[];
  float4* sA = reinterpret_cast<float4*>(tm_smem);                           // [G R][2][128] float4
  float4* sState = reinterpret_cast<float4*>(tm_smem + L::OFF_STATE);        // [G R 128] (-x2, -y2, CM, HALF)
  unsigned* sBest = reinterpret_cast<unsigned*>(tm_smem + L::OFF_BEST);      // [G R 128] cost << 16 | label
  unsigned* s_tmem = reinterpret_cast<unsigned*>(tm_smem + L::OFF_BARS + (size_t)L::NBARS * 8);
  unsigned* sCnt = reinterpret_cast<unsigned*>(tm_smem + L::OFF_CNT);        // [k_per_block]
  const unsigned smem0 = tm_smem_u32(tm_smem);
  const unsigned bar_full = smem0 + (unsigned)L::OFF_BARS;        // [G][NBUF]  MMA result landed in TMEM
  const unsigned bar_empty = bar_full + 8u * G * NBUF;            // [G][NBUF]  the warpgroup has read it out (4 warp arrivals)
  const unsigned bar_b = bar_empty + 8u * G * NBUF;               // [3]        hypothesis super-chunk landed in shared memory
  const unsigned bar_bfree = bar_b + 24u;                         // [3]        every MMA reading the super-chunk has completed (PW commits)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool epi = warp < 4 * G;
  const int g = warp >> 2, wq = warp & 3;
  const long long tile0 = (long long)blockIdx.x * L::ROWS;
  const int kbeg = blockIdx.y * k_per_block;
  const int kend = min(K, kbeg + k_per_block);
  const int nchunks = (kend - kbeg + CHH - 1) / CHH;
  const int nsc = (kend - kbeg + TM_SCH - 1) / TM_SCH;

  if (tid == 0) {
    for (int i = 0; i < G * NBUF; ++i) { tm_mbar_init(bar_full + 8u * i, 1); tm_mbar_init(bar_empty + 8u * i, 4); }
    for (int i = 0; i < 3; ++i) { tm_mbar_init(bar_b + 8u * i, 1); tm_mbar_init(bar_bfree + 8u * i, PW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4 * G) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tm_smem_u32(s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  auto set_filter = [&](int best_cost, float& cm, float& half) {
    float negmid;
    fast_thresholds(best_cost, cp, negmid, half);
    cm = COUNT_INLIERS ? cp.thr2 + negmid : negmid;
  };
  const unsigned best_init = ((unsigned)min(cp.cost_outlier, 0xffff) << 16);
  if (epi) {
    float cm0, half0;
    set_filter(cp.cost_outlier, cm0, half0);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int slot = (g * R + r) * 128 + wq * 32 + lane;
      const long long idx = tile0 + slot;
      const float4 q = pts[idx < N ? idx : N - 1];
      const float xhi = __uint_as_float(to_tf32(q.x)), yhi = __uint_as_float(to_tf32(q.y));
      const float xlo = __uint_as_float(to_tf32(q.x - xhi)), ylo = __uint_as_float(to_tf32(q.y - yhi));
      float4* a = sA + (size_t)(g * R + r) * 256 + wq * 32 + lane;
      a[0] = make_float4(xhi, yhi, 1.f, xlo);
      a[128] = make_float4(xhi, yhi, ylo, 1.f);
      sState[slot] = idx < N ? make_float4(-q.z, -q.w, cm0, half0) : make_float4(1e18f, 1e18f, cm0, -1.f);   // padding: never a hit
      sBest[slot] = best_init;
    }
    if (COUNT_INLIERS)
      for (int j = tid; j < k_per_block; j += G * 128) sCnt[j] = 0u;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the A operands are read by the tensor core (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = *s_tmem;

  if (!epi) {
    // ---- producers: warp 4G (+ pw) issues every MMA of warpgroups [pw GP, (pw + 1) GP); producer 0 also runs the TMA ring.
    // The warp stays converged (all lanes wait on the barriers); one elected lane issues.
    const int pw = warp - 4 * G;
    constexpr unsigned idesc = tm_idesc(128, NCOL);
    constexpr unsigned SC_BYTES = TM_SCH * 96;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(hsplit) + (size_t)kbeg * 96;
    const unsigned a_base = smem0, b_base = smem0 + (unsigned)L::OFF_B;
    const bool leader = tm_elect_one();
    if (pw == 0 && leader) {
      tm_mbar_expect_tx(bar_b, SC_BYTES);
      tm_bulk_g2s(b_base, src, SC_BYTES, bar_b);
    }
    int use = 0;
    for (int c = 0; c < nchunks; ++c) {
      const int sc = c / (TM_SCH / CHH), sub = c % (TM_SCH / CHH);
      if (sub == 0) {
        if (pw == 0 && sc + 1 < nsc) {
          // slot (sc + 1) % 3 was read by the MMAs of super-chunk sc - 2: both producers committed them to bar_bfree
          if (sc >= 2) tm_mbar_wait_suspended(bar_bfree + 8u * ((sc + 1) % 3), ((sc - 2) / 3) & 1);
          if (leader) {
            tm_mbar_expect_tx(bar_b + 8u * ((sc + 1) % 3), SC_BYTES);
            tm_bulk_g2s(b_base + (unsigned)((sc + 1) % 3) * SC_BYTES, src + (size_t)(sc + 1) * SC_BYTES, SC_BYTES, bar_b + 8u * ((sc + 1) % 3));
          }
        }
        tm_mbar_wait_suspended(bar_b + 8u * (sc % 3), (sc / 3) & 1);
      }
      const unsigned long long bd = tm_smem_desc(b_base + (unsigned)(sc % 3) * SC_BYTES + (unsigned)sub * (CHH * 96), NCOL * 16, 128);
#pragma unroll 1
      for (int r = 0; r < R; ++r, ++use) {
        const int b = use % NBUF, n = use / NBUF;
#pragma unroll
        for (int gi = 0; gi < GP; ++gi) {
          const int gg = pw * GP + gi;
          if (n > 0) tm_mbar_wait_suspended(bar_empty + 8u * (gg * NBUF + b), (n - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (leader) {
            const unsigned long long ad = tm_smem_desc(a_base + (unsigned)(gg * R + r) * 4096u, 128 * 16, 128);
            tm_mma_tf32(tmem + (unsigned)((gg * NBUF + b) * NCOL), ad, bd, idesc);
            tm_mma_commit(bar_full + 8u * (gg * NBUF + b));
          }
        }
      }
      if (leader && (sub == TM_SCH / CHH - 1 || c == nchunks - 1)) tm_mma_commit(bar_bfree + 8u * (sc % 3));
      __syncwarp();
    }
  } else {
    // ---- epilogue: thread = one TMEM lane = one correspondence of each of its warpgroup's R row tiles ------------------------
    unsigned* sQ = reinterpret_cast<unsigned*>(tm_smem + L::OFF_Q) + warp * QCAP;
    int qcnt = 0;   // warp-uniform
    const u64 ONE2 = pk(1.f, 1.f), NEGTHR2 = pk(-cp.thr2, -cp.thr2);
    const int slot0 = g * R * 128 + wq * 32;   // + r * 128 + lane
    const unsigned tlane = tmem + ((unsigned)(wq * 32) << 16) + (unsigned)(g * NBUF * NCOL);
    const unsigned state_addr = smem0 + (unsigned)L::OFF_STATE + (unsigned)(slot0 + lane) * 16u;   // + r * 2048
    const unsigned my_full = bar_full + 8u * (g * NBUF), my_empty = bar_empty + 8u * (g * NBUF);

    // the whole warp evaluates the queued candidates exactly, one per lane (the dense kernel's FP32 sequence), takes the minimum
    // per row with a shared-memory atomicMin on the packed (cost, label), and refreshes the filters of the rows that improved
    auto drain = [&]() {
      __syncwarp();
      for (int i = lane; i < qcnt; i += 32) {
        const unsigned e = sQ[i];
        const int rl = (int)(e >> 24), ih0 = (int)(e & 0xffffffu);
        const int slot = slot0 + (rl >> 5) * 128 + (rl & 31);
        const float4 q = __ldg(pts + tile0 + slot);   // < N: padding rows never raise a flag
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
        unsigned mark = 0u;
        if (mine != 0xffffffffu) {
          const unsigned old = atomicMin(sBest + slot, mine);   // (cost, label) packed: ties keep the lowest label
          if (mine < old) mark = 0x80000000u | (unsigned)slot;  // the row improved: its filter is rebuilt from its final value below
        }
        sQ[i] = mark;
      }
      __syncwarp();
      for (int i = lane; i < qcnt; i += 32) {
        const unsigned e = sQ[i];
        if (e & 0x80000000u) {
          const int slot = (int)(e & 0x7fffffffu);
          float cm, half;
          set_filter((int)(sBest[slot] >> 16), cm, half);
          *reinterpret_cast<float2*>(reinterpret_cast<float*>(sState + slot) + 2) = make_float2(cm, half);
        }
      }
      __syncwarp();
      qcnt = 0;
    };

    // Software pipeline at half-chunk granularity: while the FP32 pipe works on 8 hypotheses the next 8 are in flight from TMEM
    // (tcgen05.ld is asynchronous until tcgen05.wait::ld), and the wait for the NEXT accumulator is taken in the middle of a pass.
    const int total = nchunks * R;
    unsigned A[24], B[24];   // (s, x_n, y_n) x 8 hypotheses: first / second half of a chunk
    auto ld_half = [&](unsigned col, unsigned (&v)[24]) { tm_ld8(col, v); tm_ld8(col + CHH, v + 8); tm_ld8(col + 2 * CHH, v + 16); };
    unsigned cnt[CHH];
#pragma unroll
    for (int j = 0; j < CHH; ++j) cnt[j] = 0u;
    u64 tt[CHH / 2];
    int r = 0, c0 = kbeg;
    unsigned b = 0, ph = 0;   // TMEM buffer of this pass, parity of its full barrier
    if (total > 0) {
      tm_mbar_wait(my_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      ld_half(tlane, A);
    }
#pragma unroll 1
    for (int use = 0; use < total; ++use) {
      const float4 st = tm_lds128(state_addr + (unsigned)r * 2048u);
      const u64 NX = pk(st.x, st.x), NY = pk(st.y, st.y), CMP = pk(st.z, st.z);
      float m = 3.0e38f;
      auto compute_half = [&](const unsigned (&v)[24], int h) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const u64 rr = pk(rcp_approx(__uint_as_float(v[2 * p])), rcp_approx(__uint_as_float(v[2 * p + 1])));
          const u64 dx = fma2(pk(__uint_as_float(v[8 + 2 * p]), __uint_as_float(v[9 + 2 * p])), rr, NX);
          const u64 dy = fma2(pk(__uint_as_float(v[16 + 2 * p]), __uint_as_float(v[17 + 2 * p])), rr, NY);
          float ta, tb;
          u64 t;
          if (COUNT_INLIERS) {
            const u64 vv = fma2(dx, dx, fma2(dy, dy, NEGTHR2));   // d2 - thr2: sign bit = inlier
            t = fma2(vv, ONE2, CMP);                              // d2 - mid
            float va, vb;
            upk(vv, va, vb);
            cnt[8 * h + 2 * p] += __float_as_uint(va) >> 31;
            cnt[8 * h + 2 * p + 1] += __float_as_uint(vb) >> 31;
          } else {
            t = fma2(dx, dx, fma2(dy, dy, CMP));
          }
          tt[4 * h + p] = t;
          upk(t, ta, tb);
          m = min3(m, fabsf(ta), fabsf(tb));
        }
      };
      const unsigned col = tlane + b * NCOL;
      tm_wait_ld(A);
      ld_half(col + 8, B);
      compute_half(A, 0);
      tm_wait_ld(B);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) tm_mbar_arrive(my_empty + 8u * b);   // the accumulator is in registers: the next MMA may overwrite it
      const unsigned nb = b ^ 1u;
      ph ^= b;                                           // the parity flips when the buffer index wraps
      if (use + 1 < total) {
        tm_mbar_wait(my_full + 8u * nb, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        ld_half(tlane + nb * NCOL, A);
      }
      compute_half(B, 1);
      if (__any_sync(0xffffffffu, m < st.w)) {   // some row of the warp has a candidate among these 16 hypotheses
#pragma unroll
        for (int p = 0; p < CHH / 2; ++p) {
          float ta, tb;
          upk(tt[p], ta, tb);
          const bool hit = fminf(fabsf(ta), fabsf(tb)) < st.w;
          const unsigned mk = __ballot_sync(0xffffffffu, hit);
          if (mk) {
            if (hit) sQ[qcnt + __popc(mk & ((1u << lane) - 1u))] = ((unsigned)(r * 32 + lane) << 24) | (unsigned)(c0 + 2 * p);
            qcnt += __popc(mk);
          }
        }
        if (qcnt >= QTHR) drain();
      }
      b = nb;
      if (++r == R) {
        r = 0;
        if (COUNT_INLIERS) {
          // per-hypothesis counts of this warp's 32 R rows: four 8-bit fields per register (<= 32 R <= 224), one REDUX per register,
          // lane j < 16 adds hypothesis j's count to the CTA's shared-memory counter
          unsigned v[CHH / 4];
#pragma unroll
          for (int i = 0; i < CHH / 4; ++i)
            v[i] = __reduce_add_sync(0xffffffffu, cnt[4 * i] + (cnt[4 * i + 1] << 8) + (cnt[4 * i + 2] << 16) + (cnt[4 * i + 3] << 24));
          const unsigned lo = tm_selp(v[1], v[0], (lane & 4) != 0), hi = tm_selp(v[3], v[2], (lane & 4) != 0);
          const unsigned mine = (tm_selp(hi, lo, (lane & 8) != 0) >> ((lane & 3) * 8)) & 0xffu;
          if (lane < CHH && mine) atomicAdd(sCnt + (c0 - kbeg) + lane, mine);
#pragma unroll
          for (int j = 0; j < CHH; ++j) cnt[j] = 0u;
        }
        c0 += CHH;
      }
    }
    if (qcnt > 0) drain();
    if (o.best) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int slot = slot0 + r * 128 + lane;
        const long long idx = tile0 + slot;
        const unsigned bb = sBest[slot];
        if (idx < N && (bb & 0xffffu) != 0u) {
          const u64 v = ((u64)(bb >> 16) << 32) | (u64)(bb & 0xffffu);
          if (use_atomic_best) atomicMin(o.best + idx, v);
          else o.best[idx] = v;
        }
      }
    }
    if (COUNT_INLIERS) {
      asm volatile("bar.sync 1, %0;" ::"n"(G * 128) : "memory");
      const int nh = kend - kbeg;
      for (int j = tid; j < nh; j += G * 128) {
        const unsigned v = sCnt[j];
        if (v) atomicAdd(o.inlier_count + kbeg + j, (int)v);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4 * G) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <bool COUNT_INLIERS, int G, int R, int CHH, int NBUF, int PWARPS>
static mh_status launch_tmem(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, const CostParams& cp,
                             const FastOut& fo) {
  using L = TmLayout<G, R, CHH, NBUF, PWARPS>;
  const int want = ctx->sm_count;
  const unsigned tiles = (unsigned)((N + L::ROWS - 1) / L::ROWS);
  constexpr int KPB_MAX = 8192;   // shared-memory inlier counters
  int ks = (K + KPB_MAX - 1) / KPB_MAX;
  if ((int)tiles * ks < want) ks = std::min((K + TM_SCH - 1) / TM_SCH, (want + (int)tiles - 1) / (int)tiles);
  ks = std::max(1, ks);
  int kpb = (K + ks - 1) / ks;
  kpb = ((kpb + TM_SCH - 1) / TM_SCH) * TM_SCH;   // whole super-chunks per CTA
  ks = (K + kpb - 1) / kpb;
  const int Kpad = ks * kpb;
  MH_TRY(ensure_scratch(ctx, (uint64_t)Kpad * 96 + 16 + (uint64_t)K * 4));
  float4* d_split = (float4*)ctx->scratch;
  int* d_wild_count = (int*)((char*)ctx->scratch + (size_t)Kpad * 96);
  int* d_wild_list = d_wild_count + 4;
  MH_CUDA(ctx, cudaMemsetAsync(d_wild_count, 0, 16, ctx->stream));
  split_hyp_umma_kernel<CHH><<<(unsigned)((Kpad + 127) / 128), 128, 0, ctx->stream>>>(d_hyp, K, Kpad, d_split, d_wild_count, d_wild_list);
  MH_LAUNCHED(ctx, "split_hyp_umma_kernel");
  const size_t smem = L::bytes(kpb, COUNT_INLIERS);
  auto kern = cost_argmin_tmem_kernel<COUNT_INLIERS, G, R, CHH, NBUF, PWARPS>;
  MH_CUDA(ctx, mh_allow_max_smem(kern));
  kern<<<dim3(tiles, (unsigned)ks), L::THREADS, smem, ctx->stream>>>(d_pts, N, d_hyp, d_split, K, kpb, cp, fo, ks > 1);
  MH_LAUNCHED(ctx, "cost_argmin_tmem_kernel");
  return launch_cost_argmin_wild(ctx, d_pts, N, d_hyp, d_wild_count, d_wild_list, cp, fo);
}

// config 70 + i: (warpgroups, row tiles per warpgroup, hypotheses per MMA, TMEM buffers per warpgroup, producer warps)
mh_status launch_cost_argmin_tmem(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, const CostParams& cp,
                                  const FastOut& fo, int config) {
  const bool cnt = fo.inlier_count != nullptr;
#define TM_CASE(id, G, R, CHH, NBUF, PW)                                                              \
  case id:                                                                                         \
    return cnt ? launch_tmem<true, G, R, CHH, NBUF, PW>(ctx, d_pts, N, d_hyp, K, cp, fo)               \
               : launch_tmem<false, G, R, CHH, NBUF, PW>(ctx, d_pts, N, d_hyp, K, cp, fo);
  switch (config) {
    TM_CASE(76, 4, 4, 16, 2, 4)   // 2048 correspondences per CTA
    TM_CASE(77, 4, 2, 16, 2, 4)   // 1024 correspondences per CTA: best measured
#ifdef MH_TUNING
    TM_CASE(70, 4, 4, 16, 2, 2)
    TM_CASE(71, 4, 2, 16, 2, 2)
    TM_CASE(72, 4, 4, 16, 2, 1)
    TM_CASE(74, 2, 4, 16, 2, 2)
#endif
    default: return fail(ctx, MH_EINVAL, "unknown tcgen05 fast-path config");
  }
#undef TM_CASE
}

}  // namespace mh
