// ============================================================================
// K2 dense — the materialised N x (K+1) data-cost matrix (dataEnergy,
// MultiH/MultiH/MultiH.cpp:473-504; layout = GCoptimization.h:336-343: site-major,
// column 0 = outlier label).  This is the HBM-bound member of the K2 family: 4 B (int32)
// or 2 B (int16) per residual leave the chip, nothing else does.
//
// cost_dense_tiled_kernel: a CTA owns a 512-hypothesis chunk of the matrix for its whole
// life — two hypotheses per thread, held as packed f32x2 register pairs, arithmetic in
// FFMA2 — and walks 8-row tiles of correspondences through a 4-stage shared-memory ring:
//   * the store warp asks the TMA engine for the next tile's correspondences
//     (cp.async.bulk global->shared, completing on the stage's "empty" mbarrier), so the
//     compute warps read them with broadcast LDS.128 at immediate offsets and never wait
//     on global memory;
//   * compute warps stage the costs row by row and arrive, one lane per warp, on the
//     stage's "full" mbarrier — there is no CTA-wide barrier, a warp may run up to four
//     tiles ahead of its neighbours;
//   * the store warp writes every staged row with ONE bulk copy
//     (cp.async.bulk.global.shared::cta), keeps two tiles of copies in flight and
//     releases a stage as soon as the TMA has read it.
//
// Every 32-byte sector of the matrix is written exactly once, whole.  Measured on B200:
// partial-sector writes (ECC read-modify-write) cost as much as ~350 B of traffic each —
// a 16-byte-aligned copy that splits sectors runs at 60 % of HBM peak, four by-hand
// partial stores per row at 56 %, sector-aligned copies at 80-90 %.  Rows of the dense
// matrix start at every 4-byte phase (K + 1 is odd for the usual even K) and chunk
// boundaries fall mid-sector, so a row copy starts at the sector boundary BELOW the
// chunk's first column and ends at the last boundary inside the chunk: the < 32 bytes
// in front belong to the previous chunk — or, for chunk 0, to the end of the previous
// matrix row — and are computed a second time by the otherwise idle store warp (scalar
// residual()/cost_of(), hypotheses preloaded in its registers) while the compute warps
// work on the tile; the chunk's trailing partial sector is left to the next chunk's
// copy.  Only the first row's head and the last row's tail are stored by hand.
// Staged rows are shifted by their sector phase so shared and global addresses agree
// mod 16; the phase pattern is tile-invariant (8-row tiles; for int16 CTAs keep to tiles
// of one parity), so staging addresses are loop-invariant registers + immediates.
//
// Chunks are counted in hypotheses, not columns: column 0 (the constant outlier cost) is
// written once in front of chunk 0's staged rows and rides along with every row copy, so
// K = 1024 is two full chunks, not 512 + 512 + 1.  A partial last chunk gets CTAs in
// proportion to its active warps; a thin one (< 128 hypotheses) goes to
// cost_dense_tail_kernel (one thread per element).
//
// float -> int: in range the cost is floor(w), 0.5 <= w <= lam + 0.5.  Instead of F2I
// (quarter-rate XU pipe, shared with MUFU.RCP) the kernel multiplies by 2^-149 rounding
// towards -inf (FMUL2.RM, two residuals per instruction): the denormal result's bit
// pattern IS the integer.  mh_diag_set_dense_variant keeps the 2^23-magic and F2I forms
// selectable as A/B evidence.
//
// The arithmetic before the conversion is instruction-for-instruction the sequence of
// residual() / cost_in_range() in k2_device.cuh (packed f32x2 ops are two independent
// IEEE operations), so the matrix is bit-identical to cost_dense_kernel's and to what
// the fused kernels' exact update evaluates.
// ============================================================================
#include "k2_device.cuh"

namespace mh {

constexpr int DT_COMPUTE = 256;              // compute threads (8 warps, 2 matrix columns each)
constexpr int DT_WARPS = DT_COMPUTE / 32;
constexpr int DT_THREADS = DT_COMPUTE + 32;  // + one store warp
constexpr int DT_P = 8;                      // correspondences (matrix rows) per tile
constexpr int DT_KC = 512;                   // matrix columns per CTA
constexpr int DT_STAGES = 4;

__device__ __forceinline__ unsigned dt_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dt_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void dt_mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void dt_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dt_mbar_wait(unsigned bar, unsigned parity) {
  asm volatile("{\n\t.reg .pred p;\n\t"
               "W_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@!p bra W_%=;\n\t}"
               :: "r"(bar), "r"(parity) : "memory");
}
template <int OFF>
__device__ __forceinline__ void dt_sts(unsigned addr, int v, int32_t*) {
  asm volatile("st.shared.b32 [%0+%1], %2;" :: "r"(addr), "n"(OFF), "r"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ void dt_sts(unsigned addr, int v, int16_t*) {
  asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tst.shared.b16 [%0+%1], lo;\n\t}"
               :: "r"(addr), "n"(OFF), "r"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ float4 dt_lds128(unsigned addr) {
  float4 q;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(addr), "n"(OFF));
  return q;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// CONV: 1 = FMUL2.RM by 2^-149 (bit pattern of the denormal = the integer), 2 = FADD2.RM of 2^23 (low mantissa bits = the
// integer), 3 = F2I.FLOOR
// HEADS_BY_COMPUTE: the elements in front of the rows (see above) are computed by the compute warps — warp w takes tile row w, its
// lanes j < phase(w) one element each, hypotheses parked in shared memory — and the store warp only issues copies.  Why: the
// store warp is ONE warp among the ~7 resident on its scheduler; with the heads (4 passes of scalar residuals for int16, 2 for
// int32) its ~250 instructions per tile at a seventh of the issue rate were the tile period of the whole CTA (measured: 3400
// cycles per tile for int16, 2660 for int32, the compute warps waiting for a free stage in 52 % of all samples).  Measured A/B
// (1M x 1025): int16 0.999 -> 0.846 ms; int32 0.789 -> 0.849 ms — with the heads on the compute warps both element sizes take
// the same time, i.e. the kernel is then bound by its arithmetic, which for int32 is slower than its store-bound 0.789 ms.  So
// the default is: compute-warp heads for int16, store-warp heads for int32.
template <typename OutT, int CONV, bool HEADS_BY_COMPUTE = (sizeof(OutT) == 2)>
__global__ void __launch_bounds__(DT_THREADS, 3)
cost_dense_tiled_kernel(const float4* __restrict__ pts, long long N, const float* __restrict__ hyp, int K,
                        OutT* __restrict__ out, CostParams cp, int Kt, int nchunks, int g_full, int g_last) {
  constexpr int SZ = (int)sizeof(OutT);
  constexpr int SEC = 32 / SZ;                  // elements per 32-byte sector
  constexpr int PITCH = DT_KC + SEC;            // staged row: sector phase (< SEC, includes chunk 0's outlier column) + 512
  constexpr int ROWB = PITCH * SZ;
  constexpr int STAGEB = DT_P * ROWB;
  constexpr int PTSB = (DT_P + 1) * 16;         // correspondences of a stage: the row before the tile + the tile
  constexpr int NPASS = DT_P * SEC / 32;        // store-warp passes over the (row, element-in-front) slots of a tile
  static_assert(ROWB % 16 == 0 && (DT_P * SEC) % 32 == 0, "staged rows stay 16-byte aligned; slots fill whole warps");
  extern __shared__ __align__(128) unsigned char dt_smem[];
  // [DT_STAGES][DT_P][PITCH] OutT | [DT_STAGES][DT_P + 1] float4 | full[DT_STAGES], empty[DT_STAGES] mbarriers
  const unsigned s_stage = dt_smem_u32(dt_smem);
  const unsigned s_pts = s_stage + DT_STAGES * STAGEB;
  const unsigned s_full = s_pts + DT_STAGES * PTSB;
  const unsigned s_empty = s_full + DT_STAGES * 8;
  const unsigned s_head = s_empty + DT_STAGES * 8;   // HEADS_BY_COMPUTE: [DT_P rows][SEC lanes][12 floats] head hypotheses

  const int tid = threadIdx.x;
  const long long L = (long long)K + 1;
  // chunk-major block map: chunks 0 .. nchunks-2 get g_full CTAs each, the last (possibly partial) chunk g_last
  int cy, bx, gdim;
  {
    const int b = blockIdx.x, nfull = (nchunks - 1) * g_full;
    if (b < nfull) { cy = b / g_full; bx = b - cy * g_full; gdim = g_full; }
    else { cy = nchunks - 1; bx = b - nfull; gdim = g_last; }
  }
  // Chunk cy = hypotheses [512 cy, 512 cy + nh) = matrix columns 1 + 512 cy ...; chunk 0 also owns column 0, the constant
  // outlier cost, which sits in front of its staged hypotheses and is written once (nobody overwrites it).  Kt = hypotheses
  // this kernel covers (a thin remainder is cost_dense_tail_kernel's).
  const int h0 = cy * DT_KC;
  const int nh = min(DT_KC, Kt - h0);
  const int lead = (cy == 0) ? 1 : 0;
  const int c_lo = h0 + 1 - lead;          // first matrix column of the chunk
  const int ncols = nh + lead;
  const bool last_chunk = cy == nchunks - 1;
  const long long NB = (N + DT_P - 1) / DT_P;
  // sector phase of tile row r = (global element index of (row, c_lo)) mod SEC: the staged position of column c_lo.  The
  // launcher keeps gdim * DT_P a multiple of SEC (or gives the CTA a single tile), so it is the same for all tiles of the CTA
  const unsigned out_elem = (unsigned)((reinterpret_cast<unsigned long long>(out) / SZ) % SEC);
  const unsigned lstep = (unsigned)(L % SEC);
  auto phase = [&](int r) {
    return (out_elem + (((unsigned)(DT_P * bx + r)) % SEC) * lstep + (unsigned)c_lo) & (SEC - 1);
  };

  if (tid == 0) {
    for (int s = 0; s < DT_STAGES; ++s) {
      dt_mbar_init(s_full + 8 * s, DT_WARPS);
      dt_mbar_init(s_empty + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (lead && tid < DT_STAGES * DT_P) {   // the outlier column of every staged row
    const int r = tid % DT_P;
    dt_sts<0>(s_stage + (tid / DT_P) * STAGEB + r * ROWB + phase(r) * SZ, cp.cost_outlier, (OutT*)nullptr);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (tid >= DT_COMPUTE) {
    // ================= store warp ===================================================================================
    const int lane = tid - DT_COMPUTE;
    // ask the TMA engine for tile `t`'s correspondences (and the one before them); the stage's "empty" barrier completes
    // when they have landed
    auto fetch_pts = [&](int t) {
      const long long pb = bx + (long long)t * gdim;
      if (pb >= NB) return;
      const int s = t % DT_STAGES;
      const long long p0 = pb * DT_P;
      const int before = (p0 > 0) ? 1 : 0;
      const unsigned bytes = (unsigned)(min((long long)DT_P, N - p0) + before) * 16u;
      dt_mbar_expect_tx(s_empty + 8 * s, bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"(s_pts + s * PTSB + (1 - before) * 16), "l"(pts + p0 - before), "r"(bytes), "r"(s_empty + 8 * s)
                   : "memory");
    };
    if (lane == 0)
      for (int t = 0; t < DT_STAGES; ++t) fetch_pts(t);
    float hs[NPASS][9];
    int s_row[NPASS], s_act[NPASS];
    if (!HEADS_BY_COMPUTE) {
    // slot = (tile row, element j in front of the chunk's first column); slot s * 32 + lane is this lane's in pass s.  The
    // element is matrix column c_lo - g + j of the same row, or — chunk 0 — column L - g + j of the row before.
#pragma unroll
    for (int s = 0; s < NPASS; ++s) {
      const int slot = s * 32 + lane, rr = slot / SEC, j = slot % SEC;
      const int g = (int)phase(rr);
      s_row[s] = rr;
      s_act[s] = j < g;
      const long long col = (cy == 0 ? L : (long long)c_lo) - g + j;   // >= 1: the tiled kernel runs with K >= 128
#pragma unroll
      for (int k = 0; k < 9; ++k) hs[s][k] = 0.f;
      if (s_act[s]) {
        const float4* hp = reinterpret_cast<const float4*>(hyp + (size_t)(col - 1) * 12);
        const float4 u = __ldg(hp), v = __ldg(hp + 1), w = __ldg(hp + 2);
        hs[s][0] = u.x; hs[s][1] = u.y; hs[s][2] = u.z; hs[s][3] = u.w; hs[s][4] = v.x; hs[s][5] = v.y; hs[s][6] = v.z;
        hs[s][7] = v.w; hs[s][8] = w.x;
      }
    }
    }
    const int my_g = (int)phase(lane % DT_P);                  // lane r < DT_P issues row r's copy
    const int my_e = (my_g + ncols) & (SEC - 1);               // elements of the row's trailing partial sector
    const bool tail_region = last_chunk && Kt < K;             // columns after this chunk belong to the tail kernel
    int it = 0;
    for (long long pb = bx; pb < NB; pb += gdim, ++it) {
      const int b = it % DT_STAGES;
      const unsigned st = s_stage + b * STAGEB;
      if (!HEADS_BY_COMPUTE) {
      // -- elements in front of the rows, while the compute warps are busy with the tile
      dt_mbar_wait(s_empty + 8 * b, (it / DT_STAGES) & 1);     // the correspondences have landed (and the stage is free)
      // (all loads, then the independent residual chains, then the stores: the passes overlap instead of queueing)
      float4 q[NPASS];
#pragma unroll
      for (int s = 0; s < NPASS; ++s)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q[s].x), "=f"(q[s].y), "=f"(q[s].z), "=f"(q[s].w)
                     : "r"(s_pts + b * PTSB + (s_row[s] + (cy > 0 ? 1 : 0)) * 16));
      int pc[NPASS];
#pragma unroll
      for (int s = 0; s < NPASS; ++s) pc[s] = cost_of(residual(hs[s], q[s].x, q[s].y, q[s].z, q[s].w), cp);
#pragma unroll
      for (int s = 0; s < NPASS; ++s) {
        const long long p = pb * DT_P + s_row[s];
        if (s_act[s] && p < N && (cy > 0 || p > 0))
          dt_sts<0>(st + s_row[s] * ROWB + ((s * 32 + lane) % SEC) * SZ, pc[s], (OutT*)nullptr);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      }
      dt_mbar_wait(s_full + 8 * b, (it / DT_STAGES) & 1);   // every compute warp has staged (and proxy-fenced) its columns
      // -- lane r < DT_P: row r = staged elements [0, g + ncols), column c_lo at g; whole sectors go out as one bulk copy
      const long long p = pb * DT_P + lane;
      if (lane < DT_P && p < N) {
        const bool first_row = (cy == 0 && p == 0);
        const bool keep_tail = last_chunk && (p == N - 1 || tail_region);
        const int lo = (first_row && my_g) ? SEC : 0;            // first staged element of the copy
        const int hi = my_g + ncols - my_e;                      // one past its last
        OutT* grow = out + p * L + c_lo - my_g;                  // global address of staged element 0
        const unsigned srow = st + lane * ROWB;
        if (hi > lo)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       :: "l"(grow + lo), "r"(srow + lo * SZ), "r"((unsigned)((hi - lo) * SZ)) : "memory");
        auto by_hand = [&](int from, int to) {
          for (int j = from; j < to; ++j) {
            if constexpr (SZ == 4) {
              int v;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(srow + j * 4));
              grow[j] = v;
            } else {
              short v;
              asm volatile("ld.shared.b16 %0, [%1];" : "=h"(v) : "r"(srow + j * 2));
              grow[j] = v;
            }
          }
        };
        if (first_row) by_hand(my_g, min(lo, my_g + ncols));     // the matrix starts mid-sector: nothing in front to complete
        if (keep_tail) by_hand(max(hi, lo), my_g + ncols);       // nobody after this chunk completes its last sector
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (it >= 1) {
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the copies of tile it-1 have READ their rows
        __syncwarp();
        if (lane == 0) fetch_pts(it - 1 + DT_STAGES);                    // stage (it-1) % STAGES goes round again
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    return;
  }

  // ================= compute warps: hypotheses h0 + tid and h0 + tid + 256 ============================================
  u64 H[9];
  float Ta, Tb;        // in-range threshold, -1 for hypotheses past the chunk
  int fara = cp.cost_far, farb = cp.cost_far;
  {
    float ha[9], hb[9];
    auto load = [&](int i, float (&h)[9], float& T) {
      const bool is_hyp = i < nh;
      T = is_hyp ? cp.T : -1.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) h[k] = (k == 8) ? 1.f : 0.f;
      if (is_hyp) {
        const float4* hp = reinterpret_cast<const float4*>(hyp + (size_t)(h0 + i) * 12);
        const float4 u = __ldg(hp), v = __ldg(hp + 1), w = __ldg(hp + 2);
        h[0] = u.x; h[1] = u.y; h[2] = u.z; h[3] = u.w; h[4] = v.x; h[5] = v.y; h[6] = v.z; h[7] = v.w; h[8] = w.x;
      }
    };
    load(tid, ha, Ta);
    load(tid + 256, hb, Tb);
#pragma unroll
    for (int k = 0; k < 9; ++k) H[k] = pk(ha[k], hb[k]);
  }
  if (CONV == 2) { fara += 0x4B000000; farb += 0x4B000000; }   // selected before the magic is stripped
  const bool warp_active = (tid & ~31) < nh;   // warps whose columns all lie past the chunk only take part in the hand-off
  const float kslope = -cp.lam * cp.inv_T;        // the same expression as cost_in_range()
  const u64 KS2 = pk(kslope, kslope), LAM2 = pk(cp.lam, cp.lam), HALF2 = pk(0.5f, 0.5f);
  const u64 DEN2 = pk(__int_as_float(1), __int_as_float(1));   // 2^-149
  const u64 MAG2 = pk(8388608.f, 8388608.f);                   // 2^23
  // HEADS_BY_COMPUTE: warp w owns the elements in front of tile row w: lane j < phase(w) computes matrix column c_lo - g + j of
  // the row (chunk 0: column L - g + j of the row before), with that column's hypothesis read back from shared memory
  const int hw = tid >> 5, hj = tid & 31;
  const int hg = (int)phase(hw);
  const bool head_lane = HEADS_BY_COMPUTE && hj < hg;
  const unsigned head_slot = s_head + (unsigned)((hw * SEC + (hj & (SEC - 1))) * 48);
  if (head_lane) {
    const long long col = (cy == 0 ? L : (long long)c_lo) - hg + hj;   // >= 1: the tiled kernel runs with K >= 128
    const float4* hp = reinterpret_cast<const float4*>(hyp + (size_t)(col - 1) * 12);
    const float4 u = __ldg(hp), v = __ldg(hp + 1), w = __ldg(hp + 2);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(head_slot), "f"(u.x), "f"(u.y), "f"(u.z), "f"(u.w) : "memory");
    asm volatile("st.shared.v4.f32 [%0+16], {%1, %2, %3, %4};" :: "r"(head_slot), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    asm volatile("st.shared.v4.f32 [%0+32], {%1, %2, %3, %4};" :: "r"(head_slot), "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w) : "memory");
  }
  // staging address of this thread's first hypothesis in the rows of stage 0
  unsigned sbase[DT_P];
#pragma unroll
  for (int j = 0; j < DT_P; ++j) sbase[j] = s_stage + j * ROWB + (phase(j) + lead + tid) * SZ;

  int it = 0;
  for (long long pb = bx; pb < NB; pb += gdim, ++it) {
    const int b = it % DT_STAGES;
    dt_mbar_wait(s_empty + 8 * b, (it / DT_STAGES) & 1);   // correspondences have landed, the stage's previous rows are read
    if (warp_active) {
      const unsigned pa = s_pts + b * PTSB + 16;   // slot 0 is the correspondence before the tile (store warp's)
      const unsigned sb = b * STAGEB;
      // one row: 10 packed FP32 instructions + 2 MUFU.RCP for the two residuals, 3 packed for the two costs
      auto row = [&](const float4 q, int& ca, int& cb) {
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
        const u64 w2 = add2(fma2(KS2, d2, LAM2), HALF2);
        float da, db;
        upk(d2, da, db);
        int ia, ib;
        if (CONV == 1) {
          u64 m2;
          asm("mul.rm.f32x2 %0, %1, %2;" : "=l"(m2) : "l"(w2), "l"(DEN2));
          float ma, mb;
          upk(m2, ma, mb);
          ia = __float_as_int(ma); ib = __float_as_int(mb);
        } else if (CONV == 2) {
          u64 m2;
          asm("add.rm.f32x2 %0, %1, %2;" : "=l"(m2) : "l"(w2), "l"(MAG2));
          float ma, mb;
          upk(m2, ma, mb);
          ia = __float_as_int(ma); ib = __float_as_int(mb);
        } else {
          float wa, wb;
          upk(w2, wa, wb);
          ia = (int)floorf(wa); ib = (int)floorf(wb);
        }
        ca = (da < Ta) ? ia : fara;   // NaN compares false -> far, as in the reference
        cb = (db < Tb) ? ib : farb;
        if (CONV == 2 && sizeof(OutT) == 4) { ca &= 0x7fffff; cb &= 0x7fffff; }   // int16 stores the low half anyway
      };
      // four rows at a time: all loads, then the four independent dependency chains (ptxas interleaves them), then the stores
      auto quad = [&](auto GC) {
        constexpr int G = decltype(GC)::value;   // first row of the group
        const float4 q0 = dt_lds128<(G + 0) * 16>(pa), q1 = dt_lds128<(G + 1) * 16>(pa);
        const float4 q2 = dt_lds128<(G + 2) * 16>(pa), q3 = dt_lds128<(G + 3) * 16>(pa);
        int a0, b0, a1, b1, a2, b2, a3, b3;
        row(q0, a0, b0); row(q1, a1, b1); row(q2, a2, b2); row(q3, a3, b3);
#define DT_STORE(R, ca, cb)                                                \
        dt_sts<0>(sbase[R] + sb, ca, (OutT*)nullptr);                      \
        dt_sts<256 * SZ>(sbase[R] + sb, cb, (OutT*)nullptr);
        DT_STORE(G + 0, a0, b0) DT_STORE(G + 1, a1, b1) DT_STORE(G + 2, a2, b2) DT_STORE(G + 3, a3, b3)
#undef DT_STORE
      };
      quad(std::integral_constant<int, 0>{});
      quad(std::integral_constant<int, 4>{});
      static_assert(DT_P == 8, "row() calls above cover 8 rows");
    }
    if (HEADS_BY_COMPUTE) {
      const long long p = pb * DT_P + hw;
      if (head_lane && p < N && (cy > 0 || p > 0)) {
        const float4 u = dt_lds128<0>(head_slot), v = dt_lds128<16>(head_slot), w = dt_lds128<32>(head_slot);
        const float h[9] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w, w.x};
        float4 q;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                     : "r"(s_pts + b * PTSB + (hw + (cy > 0 ? 1 : 0)) * 16));
        dt_sts<0>(s_stage + b * STAGEB + hw * ROWB + hj * SZ, cost_of(residual(h, q.x, q.y, q.z, q.w), cp), (OutT*)nullptr);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged costs -> visible to the TMA engine
    __syncwarp();
    if ((tid & 31) == 0) dt_mbar_arrive(s_full + 8 * b);
  }
}

// Thin column tail (fewer than DT_TAIL columns past the last full 512-column chunk, e.g. the 1025th column of K = 1024, or
// the whole matrix when K is small): one thread per element, scalar residual() / cost_of().  A 512-column CTA would idle
// 7 of its 8 compute warps on it while holding a full share of shared memory.
constexpr int DT_TAIL = 128;
template <typename OutT>
__global__ void __launch_bounds__(256)
cost_dense_tail_kernel(const float4* __restrict__ pts, long long N, const float* __restrict__ hyp, int K,
                       OutT* __restrict__ out, CostParams cp, int c_first) {
  const long long L = (long long)K + 1;
  const int tail = (int)(L - c_first);
  const long long total = N * tail;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / tail;
    const int c = c_first + (int)(i - p * tail);
    int cost = cp.cost_outlier;
    if (c > 0) {
      const float4 q = __ldg(pts + p);
      const float4* hp = reinterpret_cast<const float4*>(hyp + (size_t)(c - 1) * 12);
      const float4 u = __ldg(hp), v = __ldg(hp + 1), w = __ldg(hp + 2);
      const float h[9] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w, w.x};
      cost = cost_of(residual(h, q.x, q.y, q.z, q.w), cp);
    }
    out[p * L + c] = (OutT)cost;
  }
}

template <typename OutT, int CONV>
static mh_status launch_tiled(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, OutT* d_cost,
                              const CostParams& cp) {
  constexpr int SEC = 32 / (int)sizeof(OutT);
  const size_t smem = (size_t)DT_STAGES * DT_P * (DT_KC + SEC) * sizeof(OutT) + (size_t)DT_STAGES * (DT_P + 1) * 16 + 2 * DT_STAGES * 8 +
                      (size_t)DT_P * SEC * 48;   // + the head hypotheses
  const long long L = (long long)K + 1;
  const long long NB = (N + DT_P - 1) / DT_P;
  const int rem = K % DT_KC;
  const int Kt = (rem < DT_TAIL) ? K - rem : K;   // hypotheses of the tiled kernel (which also writes column 0)
  if (Kt > 0) {
    const int nchunks = (Kt + DT_KC - 1) / DT_KC;
    auto kern = cost_dense_tiled_kernel<OutT, CONV>;
    MH_CUDA(ctx, mh_allow_max_smem(kern));
    int occ = 1;
    MH_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, DT_THREADS, smem));
    const long long resident = (long long)std::max(1, occ) * ctx->sm_count;
    // CTAs per chunk in proportion to the chunk's active compute warps (warp w covers hypotheses [32w, 32w+32) and
    // [256+32w, 256+32w+32) of its chunk)
    const int last_h = Kt - (nchunks - 1) * DT_KC;
    const int w_last = std::min(DT_WARPS, (last_h + 31) / 32);
    const long long weight = (long long)(nchunks - 1) * DT_WARPS + w_last;
    long long g_full = 0, g_last;
    if (nchunks > 1) {
      g_full = std::max<long long>(1, std::min<long long>(NB, resident * DT_WARPS / weight));
      g_last = std::max<long long>(1, std::min<long long>(NB, resident - g_full * (nchunks - 1)));
    } else {
      g_last = std::max<long long>(1, std::min<long long>(NB, resident));
    }
    // the kernel's staging addresses assume every tile of a CTA has the same sector phases: gdim * DT_P % SEC == 0 (or a
    // single tile per CTA)
    if (DT_P % SEC != 0 && NB >= 2) {
      constexpr int M = SEC / DT_P;
      g_last = std::max<long long>(M, g_last / M * M);
      if (nchunks > 1) g_full = std::max<long long>(M, g_full / M * M);
    }
    const long long grid = g_full * (nchunks - 1) + g_last;
    if (grid > 0x7fffffffLL) return fail(ctx, MH_EINVAL, "mh_data_cost_dense: too many column chunks");
    kern<<<(unsigned)grid, DT_THREADS, smem, ctx->stream>>>(d_pts, N, d_hyp, K, d_cost, cp, Kt, nchunks,
                                                            (int)std::max<long long>(1, g_full), (int)g_last);
    MH_LAUNCHED(ctx, "cost_dense_tiled_kernel");
  }
  if (Kt < K || Kt == 0) {
    const int c_first = (Kt == 0) ? 0 : Kt + 1;
    const long long total = N * (L - c_first);
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 8));
    cost_dense_tail_kernel<OutT><<<grid, 256, 0, ctx->stream>>>(d_pts, N, d_hyp, K, d_cost, cp, c_first);
    MH_LAUNCHED(ctx, "cost_dense_tail_kernel");
  }
  return MH_OK;
}

mh_status launch_cost_dense_tiled(mh_ctx* ctx, const float4* d_pts, int64_t N, const float* d_hyp, int K, void* d_cost,
                                  int elem_bytes, int conv) {
  const CostParams cp = cost_params(ctx);
  if ((reinterpret_cast<unsigned long long>(d_cost) % (unsigned)elem_bytes) != 0)
    return fail(ctx, MH_EINVAL, "mh_data_cost_dense: output pointer must be aligned to its element size");
  if ((reinterpret_cast<unsigned long long>(d_pts) % 16) != 0)
    return fail(ctx, MH_EINVAL, "mh_data_cost_dense: correspondences must be 16-byte aligned");
  if (elem_bytes == 4) {
    if (conv == 2) return launch_tiled<int32_t, 2>(ctx, d_pts, N, d_hyp, K, (int32_t*)d_cost, cp);
    if (conv == 3) return launch_tiled<int32_t, 3>(ctx, d_pts, N, d_hyp, K, (int32_t*)d_cost, cp);
    return launch_tiled<int32_t, 1>(ctx, d_pts, N, d_hyp, K, (int32_t*)d_cost, cp);
  }
  if (conv == 2) return launch_tiled<int16_t, 2>(ctx, d_pts, N, d_hyp, K, (int16_t*)d_cost, cp);
  if (conv == 3) return launch_tiled<int16_t, 3>(ctx, d_pts, N, d_hyp, K, (int16_t*)d_cost, cp);
  return launch_tiled<int16_t, 1>(ctx, d_pts, N, d_hyp, K, (int16_t*)d_cost, cp);
}

}  // namespace mh
