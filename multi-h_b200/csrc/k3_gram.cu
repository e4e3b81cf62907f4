// ============================================================================
// K3, L2 member — batched mean-shift with the pairwise-distance Gram on the 5th-generation
// tensor cores (tcgen05, accumulators in tensor memory).  meanshift_metric = 2.
//
// The reference's mean-shift (MeanShiftClustering.h:22-157) is sequential: one random seed
// after another, each marking the points its windows cover.  This member is the textbook
// batch formulation of the same estimator with the L2 window ||m - x||_2^2 < bw^2: EVERY point
// is a seed, all trajectories advance together, one kernel launch per window iteration, until
// no mean moves by more than 1e-3 bw (MS.h:48, 98); the converged means are then merged in
// index order into the first centre closer than bw/2, averaging as MS.h:100-120 does, and every
// point belongs to the cluster its own trajectory ended in.  It finds the modes the sequential
// algorithm finds (parity tier T3: mode sets within bw/2, cluster count equal up to a few —
// tests/test_gpu_parity.py::test_k3_meanshift_gram_tensor_core_variant), not its exact vote-based
// assignment; the exact members are metric 0 (the reference's L1 window) and 1.
//
// One window iteration = "attention-shaped": S = Q X^T (the Gram), P = [ |q|^2 + |x|^2 - 2 S < bw^2 ],
// Q' = P X / rowsum(P).  Here:
//   * Q X^T runs as tcgen05.mma kind::tf32 with the 3xTF32 split packed into K (a row holds
//     [q_hi, q_lo, q_hi], a column [x_hi, x_hi, x_lo]: hi hi + lo hi + hi lo), M = 128 queries x
//     N = 256 points per instruction group, FP32 accumulators in TMEM (2 x 256 columns, double
//     buffered), issued by ONE thread; the point tiles (operand + FP32 coordinates + norms, packed
//     once per call) arrive by the TMA engine's bulk copy (cp.async.bulk + mbarrier);
//   * the epilogue warps drain TMEM with tcgen05.ld (32 lanes x 64 columns per instruction; a thread
//     owns one query row, so P X accumulates in its registers without atomics).  The tensor-core
//     distance is a FILTER: coordinates are centred, but |q|^2 + |x|^2 - 2 q.x still cancels ~1e6
//     against a window of bw^2 = 4.84, so a pair within a margin of the window is re-evaluated
//     exactly, SUM (q_d - x_d)^2 in FP32 from shared memory — membership is exact in FP32;
//   * points and queries stay sorted by coordinate 0 (K3's preparation), so a query tile only
//     visits the point tiles its windows can reach (|q_0 - x_0| < bw).
// ============================================================================
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace mh {

constexpr int GR_M = 128;       // queries per CTA = TMEM lanes
constexpr int GR_N = 256;       // points per tile = TMEM columns of one accumulator
constexpr int GR_DP = 16;       // padded coordinate count of the FP32 copies
constexpr int GR_EPI_WARPS = 8;  // a warp reaches the TMEM lanes 32 (warp % 4) ...: two warps per lane quarter, each half the columns
constexpr int GR_THREADS = 32 * GR_EPI_WARPS + 32;   // + 1 producer warp (TMA + MMA issue)

// one point tile in global memory, exactly as it lands in shared memory (16-byte units)
template <int KC>   // KC = number of 16-byte K chunks (4 tf32 values each): K = 4 KC
struct GramTile {
  float op[KC][GR_N][4];   // UMMA operand, K-major, no swizzle: 8 rows x 16 B core matrices, row groups 128 B apart (SBO),
                           // K chunks GR_N * 16 B apart (LBO)
  float x[GR_N][GR_DP];    // centred coordinates
  float xn[GR_N];          // (1 - 4e-6) |x|^2, the filter's form of the norm (1e30 for padding rows: never inside a window)
};

struct GramProblem {
  const double* xs;     // [D][Npad] sorted rows (K3 preparation)
  const int32_t* perm;  // sorted position -> original index
  const int32_t* n_finite;
  int N, Npad, D;
  double bw;
  double* mean;         // [16] centre of the finite rows
  void* tiles;          // GramTile<KC>[ntiles]
  float* q[2];          // [Npad][GR_DP] centred means, ping-pong
  int32_t* moving;      // [iterations] != 0: some mean moved more than the stop threshold
  int32_t* inv;         // original index -> sorted position
  int dense;            // measurement aid (MH_GRAM_DENSE=1): every query tile visits every point tile — the full N x N Gram
};

__device__ __forceinline__ unsigned gr_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float gr_tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void gr_mbar_init(void* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gr_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gr_mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gr_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gr_mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gr_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gr_mbar_wait(void* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra W_%=;\n\t}" ::"r"(gr_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void gr_bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(gr_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(gr_smem_u32(bar))
               : "memory");
}
// shared-memory matrix descriptor (tcgen05): start address, leading / stride byte offsets in 16-byte units, version 1, no swizzle
__device__ __forceinline__ unsigned long long gr_smem_desc(const void* p, unsigned lbo_bytes, unsigned sbo_bytes) {
  const unsigned long long a = (gr_smem_u32(p) & 0x3ffffu) >> 4;
  return a | ((unsigned long long)(lbo_bytes >> 4) << 16) | ((unsigned long long)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: FP32 accumulate, TF32 x TF32, both operands K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr unsigned gr_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void gr_mma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                            unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void gr_mma_commit(void* bar) {   // arrives on bar when every MMA issued so far has completed
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(gr_smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive columns of 32-bit accumulators -> 32 registers per thread
__device__ __forceinline__ void gr_tmem_ld32(unsigned taddr, float (&v)[32]) {
  unsigned r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- preparation -------------------------------------------------------------------------------------------------------------------
__global__ void gram_mean_kernel(GramProblem p) {   // one CTA: centre of the finite rows, FP64, fixed order
  __shared__ double s_part[256];
  const int Nf = *p.n_finite, tid = threadIdx.x;
  for (int j = 0; j < GR_DP; ++j) {
    double s = 0.0;
    if (j < p.D)
      for (int q = tid; q < Nf; q += 256) s += p.xs[(size_t)j * p.Npad + q];
    s_part[tid] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o) s_part[tid] += s_part[tid + o];
      __syncthreads();
    }
    if (tid == 0) p.mean[j] = j < p.D && Nf > 0 ? s_part[0] / Nf : 0.0;
    __syncthreads();
  }
}

template <int KC>
__global__ void __launch_bounds__(GR_N) gram_pack_kernel(GramProblem p) {
  GramTile<KC>* tile = reinterpret_cast<GramTile<KC>*>(p.tiles) + blockIdx.x;
  const int r = threadIdx.x, pos = blockIdx.x * GR_N + r, Nf = *p.n_finite, D = p.D;
  float x[GR_DP];
  float xn = 0.f;
#pragma unroll
  for (int j = 0; j < GR_DP; ++j) {
    x[j] = (j < D && pos < Nf) ? (float)(p.xs[(size_t)j * p.Npad + pos] - p.mean[j]) : 0.f;
    xn += x[j] * x[j];
  }
  xn *= 1.f - 4e-6f;
  if (pos >= Nf) xn = 1e30f;
  // K slots: [0, D) x_hi (meets q_hi) | [D, 2D) x_hi (meets q_lo) | [2D, 3D) x_lo (meets q_hi) | zeros
  for (int k = 0; k < 4 * KC; ++k) {
    float v = 0.f;
    if (k < 3 * D) {
      const int j = k % D;
      const float hi = gr_tf32(x[j]);
      v = k < 2 * D ? hi : gr_tf32(x[j] - hi);
    }
    tile->op[k >> 2][r][k & 3] = v;
  }
#pragma unroll
  for (int j = 0; j < GR_DP; ++j) tile->x[r][j] = x[j];
  tile->xn[r] = xn;
  if (pos < p.Npad) {
#pragma unroll
    for (int j = 0; j < GR_DP; ++j) p.q[0][(size_t)pos * GR_DP + j] = x[j];
  }
  if (pos < p.N) p.inv[p.perm[pos]] = pos;
}

// ---- one window iteration of every trajectory ---------------------------------------------------------------------------------------
template <int KC>
struct GramSmem {
  GramTile<KC> tile[2];
  float a_op[KC][GR_M][4];            // the query tile as UMMA operand
  unsigned long long full[2], mma_done[2], epi_done[2];
  unsigned tmem_base;
  int t_lo, t_hi;
  float qmin[4], qmax[4];
  int moving;
  float comb[GR_M][GR_DP + 1];       // partial sums of the warps that take the upper half of the columns
};

template <int KC>
__global__ void __launch_bounds__(GR_THREADS, 1) gram_shift_kernel(GramProblem p, int it) {
  extern __shared__ __align__(1024) unsigned char gr_raw[];
  GramSmem<KC>& sm = *reinterpret_cast<GramSmem<KC>*>(gr_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Nf = *p.n_finite, D = p.D;
  const int row0 = blockIdx.x * GR_M;
  if (row0 >= Nf) return;
  const float* qin = p.q[it & 1];
  float* qout = p.q[(it & 1) ^ 1];
  const float B = (float)(p.bw * p.bw), stop2 = (float)(1e-3 * p.bw * 1e-3 * p.bw);
  const int ntiles = (Nf + GR_N - 1) / GR_N;

  // ---- set-up: barriers, TMEM, the query tile as operand, the range of point tiles the windows can reach ----------------------
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) { gr_mbar_init(&sm.full[b], 1); gr_mbar_init(&sm.mma_done[b], 1); gr_mbar_init(&sm.epi_done[b], 32 * GR_EPI_WARPS); }
    sm.moving = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == GR_EPI_WARPS) {   // one warp allocates all 512 TMEM columns (two accumulators) and gives up the permit
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(gr_smem_u32(&sm.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  float q[GR_DP];
  float qn = 0.f;
  const int rl = (warp & 3) * 32 + lane, half = warp >> 2;   // epilogue warps: TMEM lane = query row of the tile, column half
  const int row = row0 + rl;
  const bool epi = warp < GR_EPI_WARPS, live = epi && row < Nf;
  if (epi) {
#pragma unroll
    for (int j = 0; j < GR_DP; ++j) {
      q[j] = live ? qin[(size_t)row * GR_DP + j] : 0.f;
      qn += q[j] * q[j];
    }
  }
  if (epi && half == 0) {
    // K slots: [0, D) q_hi | [D, 2D) q_lo | [2D, 3D) q_hi | zeros
    for (int k = 0; k < 4 * KC; ++k) {
      float v = 0.f;
      if (k < 3 * D) {
        const int j = k % D;
        const float hi = gr_tf32(q[j]);
        v = (k >= D && k < 2 * D) ? gr_tf32(q[j] - hi) : hi;
      }
      sm.a_op[k >> 2][rl][k & 3] = v;
    }
    float lo0 = live ? q[0] : 3e38f, hi0 = live ? q[0] : -3e38f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo0 = fminf(lo0, __shfl_xor_sync(0xffffffffu, lo0, o));
      hi0 = fmaxf(hi0, __shfl_xor_sync(0xffffffffu, hi0, o));
    }
    if (lane == 0) { sm.qmin[warp] = lo0; sm.qmax[warp] = hi0; }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the operand written above is read by the tensor core (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    // sorted coordinate 0 (uncentred, FP64): the first tile holding x_0 >= qmin - bw and the last holding x_0 <= qmax + bw
    const double lo0 = (double)fminf(fminf(sm.qmin[0], sm.qmin[1]), fminf(sm.qmin[2], sm.qmin[3])) + p.mean[0] - p.bw * 1.0001 - 1e-3;
    const double hi0 = (double)fmaxf(fmaxf(sm.qmax[0], sm.qmax[1]), fmaxf(sm.qmax[2], sm.qmax[3])) + p.mean[0] + p.bw * 1.0001 + 1e-3;
    int a = 0, b = Nf;
    while (a < b) { const int m = (a + b) >> 1; if (p.xs[m] < lo0) a = m + 1; else b = m; }
    const int first = a;
    b = Nf;
    while (a < b) { const int m = (a + b) >> 1; if (p.xs[m] <= hi0) a = m + 1; else b = m; }
    sm.t_lo = p.dense ? 0 : first / GR_N;
    sm.t_hi = p.dense ? ntiles : min(ntiles, (max(a, first + 1) + GR_N - 1) / GR_N);
  }
  __syncthreads();
  const int t_lo = sm.t_lo, t_hi = sm.t_hi;
  const unsigned tmem = sm.tmem_base;
  constexpr unsigned TILE_BYTES = sizeof(GramTile<KC>);

  if (warp == GR_EPI_WARPS) {
    // ---- producer: TMA bulk loads of the point tiles, then the MMAs of each tile, all from one thread ------------------------
    if (lane == 0) {
      constexpr unsigned idesc = gr_idesc(GR_M, GR_N);
      const GramTile<KC>* tiles = reinterpret_cast<const GramTile<KC>*>(p.tiles);
      for (int t = t_lo; t < t_hi; ++t) {
        const int k = t - t_lo, buf = k & 1;
        const unsigned ph = (k >> 1) & 1;
        if (k >= 2) gr_mbar_wait(&sm.epi_done[buf], ph ^ 1);   // the epilogue has left this buffer pair (smem tile + accumulator)
        gr_mbar_expect_tx(&sm.full[buf], TILE_BYTES);
        gr_bulk_g2s(&sm.tile[buf], tiles + t, TILE_BYTES, &sm.full[buf]);
        gr_mbar_wait(&sm.full[buf], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int s = 0; s < KC / 2; ++s) {   // K = 8 per instruction: two 16-byte chunks
          const unsigned long long ad = gr_smem_desc(&sm.a_op[2 * s][0][0], GR_M * 16, 128);
          const unsigned long long bd = gr_smem_desc(&sm.tile[buf].op[2 * s][0][0], GR_N * 16, 128);
          gr_mma_tf32(tmem + buf * GR_N, ad, bd, idesc, s > 0);
        }
        gr_mma_commit(&sm.mma_done[buf]);
      }
    }
  } else {
    // ---- epilogue: a thread owns one query row (TMEM lane); windows and means on the FP32 pipe ------------------------------
    float acc[GR_DP];
#pragma unroll
    for (int j = 0; j < GR_DP; ++j) acc[j] = 0.f;
    int cnt = 0;
    const float fthr = B + 0.05f - (1.f - 4e-6f) * qn;
    for (int t = t_lo; t < t_hi; ++t) {
      const int k = t - t_lo, buf = k & 1;
      gr_mbar_wait(&sm.mma_done[buf], (k >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const GramTile<KC>& tl = sm.tile[buf];
#pragma unroll 1
      for (int c0 = half * (GR_N / 2); c0 < (half + 1) * (GR_N / 2); c0 += 32) {
        float s[32];
        gr_tmem_ld32(tmem + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)(buf * GR_N + c0), s);
        // filter: |q|^2 + |x|^2 - 2 q.x < bw^2 + 4e-6 (|q|^2 + |x|^2) + 0.05 — the 3xTF32 Gram is good to ~1e-6 of the norms, so
        // anything that close to the window is decided exactly below.  One FFMA per pair and one min per two; the 32 columns are
        // looked at one by one only when their minimum passes.
        float f[32];
        float mn = 3.0e38f;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 xn4 = *reinterpret_cast<const float4*>(&tl.xn[c0 + 4 * c4]);
          f[4 * c4] = fmaf(-2.f, s[4 * c4], xn4.x); f[4 * c4 + 1] = fmaf(-2.f, s[4 * c4 + 1], xn4.y);
          f[4 * c4 + 2] = fmaf(-2.f, s[4 * c4 + 2], xn4.z); f[4 * c4 + 3] = fmaf(-2.f, s[4 * c4 + 3], xn4.w);
          mn = fminf(mn, fminf(fminf(f[4 * c4], f[4 * c4 + 1]), fminf(f[4 * c4 + 2], f[4 * c4 + 3])));
        }
        if (mn < fthr) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            if (f[c] < fthr) {
              const float4* xr = reinterpret_cast<const float4*>(tl.x[c0 + c]);
              float e = 0.f;
              float xv[GR_DP];
#pragma unroll
              for (int j4 = 0; j4 < GR_DP / 4; ++j4) {
                const float4 v = xr[j4];
                xv[4 * j4] = v.x; xv[4 * j4 + 1] = v.y; xv[4 * j4 + 2] = v.z; xv[4 * j4 + 3] = v.w;
              }
#pragma unroll
              for (int j = 0; j < GR_DP; ++j) { const float d = q[j] - xv[j]; e = fmaf(d, d, e); }
              if (e < B) {
                ++cnt;
#pragma unroll
                for (int j = 0; j < GR_DP; ++j) acc[j] += xv[j];
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      gr_mbar_arrive(&sm.epi_done[buf]);
    }
    // the two column halves of a row meet in shared memory
    if (half == 1) {
#pragma unroll
      for (int j = 0; j < GR_DP; ++j) sm.comb[rl][j] = acc[j];
      sm.comb[rl][GR_DP] = __int_as_float(cnt);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * GR_EPI_WARPS) : "memory");
    if (half == 0) {
#pragma unroll
      for (int j = 0; j < GR_DP; ++j) acc[j] += sm.comb[rl][j];
      cnt += __float_as_int(sm.comb[rl][GR_DP]);
    }
    if (live && half == 0) {
      float shift2 = 0.f;
      const float inv = cnt > 0 ? 1.f / (float)cnt : 0.f;
#pragma unroll
      for (int j = 0; j < GR_DP; ++j) {
        const float nm = cnt > 0 ? acc[j] * inv : q[j];   // MS.h:96 (a window always holds its own mean's nearest point)
        const float d = nm - q[j];
        shift2 = fmaf(d, d, shift2);
        qout[(size_t)row * GR_DP + j] = nm;
      }
      if (shift2 >= stop2) sm.moving = 1;   // MS.h:98
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == GR_EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  if (tid == 0 && sm.moving) atomicOr(p.moving + it, 1);
}

// ---- merge the converged means in index order (MS.h:100-120) and assign -------------------------------------------------------------
// One CTA walks the points in index order (the merge is sequential by definition: a mean joins the FIRST centre closer than
// bw/2, which it then moves); for every point all 256 threads scan the centres found so far.
__global__ void __launch_bounds__(256) gram_merge_kernel(GramProblem p, int it_final, double* centres, int max_c, int32_t* assign,
                                                         int32_t* out /*[0] C, [1] overflow*/) {
  __shared__ int s_best[2];
  __shared__ double s_m[GR_DP];
  const int tid = threadIdx.x, lane = tid & 31, D = p.D, Nf = *p.n_finite;
  const float* q = p.q[it_final & 1];
  const double half = p.bw / 2;
  int C = 0, overflow = 0;
  if (tid < 2) s_best[tid] = 0x7fffffff;
  __syncthreads();
  for (int i = 0; i < p.N; ++i) {
    const int pos = p.inv[i];
    if (pos >= Nf) { if (tid == 0) assign[i] = -1; continue; }
    const int par = i & 1;   // two result slots: the reset of one overlaps the use of the other
    if (tid < D) s_m[tid] = (double)q[(size_t)pos * GR_DP + tid] + p.mean[tid];
    if (tid == 0) s_best[par ^ 1] = 0x7fffffff;
    __syncthreads();
    int best = 0x7fffffff;
    for (int c = tid; c < C && best == 0x7fffffff; c += 256) {
      double d2 = 0.0;
      for (int j = 0; j < D; ++j) { const double d = s_m[j] - centres[(size_t)c * D + j]; d2 += d * d; }
      if (sqrt(d2) < half) best = c;
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if (lane == 0 && best != 0x7fffffff) atomicMin(&s_best[par], best);
    __syncthreads();
    best = s_best[par];
    int cid = best;
    if (best == 0x7fffffff) {
      if (C < max_c) {
        cid = C;
        if (tid < D) centres[(size_t)C * D + tid] = s_m[tid];
        ++C;
      } else { overflow = 1; cid = -1; }
    } else if (tid < D) {
      centres[(size_t)cid * D + tid] = 0.5 * (centres[(size_t)cid * D + tid] + s_m[tid]);
    }
    if (tid == 0) assign[i] = cid;
    __syncthreads();   // the centre written above is read by the next point's scan
  }
  if (tid == 0) { out[0] = C; out[1] = overflow; }
}

template <int KC>
static mh_status run_gram(mh_ctx* ctx, GramProblem& p, int ntiles, int max_iters, double* d_centres, int max_c, int32_t* d_assign,
                          int32_t* d_out, int* iters_out) {
  gram_mean_kernel<<<1, 256, 0, ctx->stream>>>(p);
  MH_LAUNCHED(ctx, "gram_mean_kernel");
  gram_pack_kernel<KC><<<ntiles, GR_N, 0, ctx->stream>>>(p);
  MH_LAUNCHED(ctx, "gram_pack_kernel");
  MH_CUDA(ctx, mh_allow_max_smem(gram_shift_kernel<KC>));
  const size_t smem = sizeof(GramSmem<KC>) + 1024;
  int it = 0;
  // the stop test is read back every 4 iterations (a launch costs less than the round trip)
  for (; it < max_iters;) {
    const int chunk = std::min(4, max_iters - it);
    for (int k = 0; k < chunk; ++k, ++it) {
      gram_shift_kernel<KC><<<(p.N + GR_M - 1) / GR_M, GR_THREADS, smem, ctx->stream>>>(p, it);
      MH_LAUNCHED(ctx, "gram_shift_kernel");
    }
    int32_t mv[4] = {0, 0, 0, 0};
    MH_CUDA(ctx, cudaMemcpyAsync(mv, p.moving + (it - chunk), sizeof(int32_t) * chunk, cudaMemcpyDeviceToHost, ctx->stream));
    MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    bool done = false;
    for (int k = 0; k < chunk; ++k)
      if (!mv[k]) done = true;   // an iteration in which nothing moved: every later one is a fixed point as well
    if (done) break;
  }
  gram_merge_kernel<<<1, 256, 0, ctx->stream>>>(p, it, d_centres, max_c, d_assign, d_out);
  MH_LAUNCHED(ctx, "gram_merge_kernel");
  *iters_out = it;
  return MH_OK;
}

// d_xs / d_perm / d_nf: K3's preparation (rows sorted by coordinate 0, finite rows first).  scratch: device memory of
// gram_scratch_bytes(N, D) bytes, 1024-byte aligned.
uint64_t gram_scratch_bytes(int N, int D) {
  const int KC = 3 * D <= 32 ? 8 : 12;
  const uint64_t ntiles = (uint64_t)(N + GR_N - 1) / GR_N;
  const uint64_t tile = KC == 8 ? sizeof(GramTile<8>) : sizeof(GramTile<12>);
  const uint64_t npad = ((uint64_t)N + 31) & ~31ull;
  return 1024 + ntiles * tile + 2 * npad * GR_DP * 4 + 4096 /*moving*/ + (uint64_t)N * 4 + 256 + 1024;
}

mh_status launch_meanshift_gram(mh_ctx* ctx, const double* d_xs, const int32_t* d_perm, const int32_t* d_nf, int N, int Npad, int D,
                                double bw, void* scratch, double* d_centres, int max_c, int32_t* d_assign, int* C_out, int64_t* stats) {
  if (D > GR_DP) return fail(ctx, MH_EINVAL, "mh_meanshift (tensor-core L2): D > 16");
  const int KC = 3 * D <= 32 ? 8 : 12;
  const int ntiles = (N + GR_N - 1) / GR_N;
  const uint64_t tile = KC == 8 ? sizeof(GramTile<8>) : sizeof(GramTile<12>);
  char* base = (char*)(((uintptr_t)scratch + 1023) & ~(uintptr_t)1023);
  GramProblem p;
  p.xs = d_xs; p.perm = d_perm; p.n_finite = d_nf;
  p.N = N; p.Npad = Npad; p.D = D; p.bw = bw;
  p.dense = std::getenv("MH_GRAM_DENSE") != nullptr;
  p.tiles = base; base += (uint64_t)ntiles * tile;
  p.q[0] = (float*)base; base += (uint64_t)Npad * GR_DP * 4;
  p.q[1] = (float*)base; base += (uint64_t)Npad * GR_DP * 4;
  p.moving = (int32_t*)base; base += 4096;
  p.inv = (int32_t*)base; base += (((uint64_t)N * 4 + 255) & ~255ull);
  p.mean = (double*)base; base += 128;
  int32_t* d_out = (int32_t*)base;
  const int max_iters = 1000;   // 4096 / 4 slots of `moving`
  MH_CUDA(ctx, cudaMemsetAsync(p.moving, 0, 4096, ctx->stream));
  MH_CUDA(ctx, cudaMemsetAsync(p.q[1], 0, (uint64_t)Npad * GR_DP * 4, ctx->stream));
  int iters = 0;
  if (KC == 8) MH_TRY(run_gram<8>(ctx, p, ntiles, max_iters, d_centres, max_c, d_assign, d_out, &iters));
  else MH_TRY(run_gram<12>(ctx, p, ntiles, max_iters, d_centres, max_c, d_assign, d_out, &iters));
  int32_t out[2] = {0, 0};
  MH_CUDA(ctx, cudaMemcpyAsync(out, d_out, sizeof(out), cudaMemcpyDeviceToHost, ctx->stream));
  MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *C_out = out[0];
  if (stats) { stats[0] = N; stats[1] = iters; }
  if (out[1]) return fail(ctx, MH_ENOMEM, "mh_meanshift: more centres than max_c");
  return MH_OK;
}

}  // namespace mh
