// ============================================================================
// K3 — homography-space mean-shift (sm_100a), exact L1_REF member.
//
// Replaces MeanShiftClustering<double>::Cluster
// (MultiH/MultiH/moduls/mode_seeking/MeanShiftClustering.h:22-157) as called from
// EstablishStablePointSets (MultiH.cpp:654, D = 10) and MergingStep (MultiH.cpp:397, D = 6).
//
// The reference draws one random not-yet-visited seed after another (MS.h:52-56), so its
// result depends on the order in which windows mark points as visited.  What does NOT
// depend on that order is a trajectory itself: the window test (MS.h:76-85) runs over ALL
// points, visited or not, so the sequence of means started at point s is a pure function
// of s.  The algorithm therefore splits into
//
//   A. ms_trajectories_kernel — chip-wide, one warp per point: the trajectory the point
//      WOULD produce as a seed (window iterations until the mean stops moving, MS.h:62-98),
//      stored as its final mean + per neighbouring row the number of windows it was a member of.  Only short
//      trajectories over sparse windows are speculated (<= MS_CAP_A windows of <=
//      MS_RANGE_A candidates): those are the many isolated points; the few seeds inside
//      dense clusters drift for dozens of windows, and nearly all of their neighbours are
//      visited before they could ever be drawn.
//   A2. ms_heavy_kernel (N >= MS_HEAVY_MIN_N) — the seeds step A leaves out, chip-wide, one CTA per seed, with the same block-wide
//      trajectory code the replay uses; votes go to a heavy record (one byte per sorted position of a MS_HSPAN-wide span).
//      Measured at 20 000 points: 307 of 1779 drawn seeds were such trajectories and cost the one replay CTA 88 % of its
//      time (93 us each); computing all ~10^4 candidates on 148 SMs takes 5 ms and the call drops from 33.5 to 14.0 ms.
//   B. ms_replay_kernel — one CTA replays the reference's sequential part exactly: the
//      restated MSVC rand() picks the rank-th unvisited point in index order (MS.h:54-56);
//      its trajectory comes from step A's / A2's record or is computed on the spot by the whole
//      CTA; the windows mark their members visited and give them their votes (MS.h:85-93);
//      the final mean is merged into the first centre closer than bw/2 or appended
//      (MS.h:100-120); per-point votes live in sparse (cluster, votes) lists and the last
//      pass assigns every point to its best cluster (MS.h:133-146).
//
// Both halves prune: sum_j |m_j - x_j| < bw^2 implies |m_0 - x_0| < bw^2, so the points
// are sorted by coordinate 0 once and a window only scans the contiguous candidate range
// found by a warp-wide 32-ary search with the very predicate of the window test (rounding
// included); a candidate is dropped as soon as its partial sum fails (terms are
// non-negative) — the member sets are identical to a scan over all points.  The merge test
// of step B scans all centres while they are few and uses a 3-D cell hash (cell width bw/2)
// beyond that.  FP64 throughout; sums are reduced in a fixed order: reproducible run to run.
//
// A trajectory ends after MH_MS_MAX_WINDOW_ITERS window iterations at the latest (the
// reference's `while (1)` never returns when the mean cycles; oracle and kernel cap alike).
// ============================================================================
#include <cub/device/device_radix_sort.cuh>

#include <cstdio>
#include <type_traits>

#include "common.cuh"

namespace mh {

constexpr int MS_MAXD = 16;
constexpr int MS_VCAP = 16;             // inline capacity of a point's (cluster, votes) list; more goes to the spill log
constexpr int MS_TRAJ_THREADS = 256;    // step A: 8 warps per CTA, one seed per warp at a time
constexpr int MS_REPLAY_THREADS = 256;  // step B
constexpr int MS_REPLAY_WARPS = MS_REPLAY_THREADS / 32;
constexpr int MS_CAP_A = 16;            // step A speculates trajectories of up to this many windows ...
constexpr int MS_RANGE_A = 1024;        // ... each with at most this many candidates
constexpr int MS_SMEM_MASK_WORDS = 32768;  // visited bits of up to 2^20 points live in shared memory (step B)
constexpr int MS_SORT_SMALL_MAX = 4096;    // up to here one CTA sorts in shared memory; beyond: cub radix sort
constexpr int MS_BRUTE_CENTRES = 1024;     // merge test: scan all centres up to here, cell hash beyond
constexpr int MS_HSPAN = 16384;            // widest union of candidate ranges a heavy record holds (one byte per sorted position)
constexpr int MS_HHDR = MS_MAXD + 4;       // doubles per heavy record header
constexpr int MS_HEAVY_MIN_N = 2048;       // below this the replay CTA computes the few heavy trajectories itself
constexpr int MS_HEAVY_MAX_SLOTS = 32768;  // 512 MB of vote bytes at most
constexpr int MS_HASH_DIMS = 3;
constexpr int MS_STAGE1 = 3;               // coordinates summed before the first early exit of the window test

struct MsProblem {
  // ---- data ----
  const double* data;  // [N][D] row-major, as given
  double* xs;          // [D][Npad]: rows sorted by coordinate 0, finite rows first
  int32_t* perm;       // [N] sorted position -> original index
  uint32_t* mask0;     // [mask_words] initial visited bits by original index: non-finite rows and the padding
  int N, Npad, D, mask_words;
  double bandSq, stopThresh, halfBw, cellInv;
  int metric;
  // ---- speculated trajectories by ORIGINAL index (step A -> step B) ----
  double* rec;         // [N][rec_stride], layout at ms_rec_stride()
  int rec_stride;
  // ---- heavy trajectories (step A2 -> step B): the seeds step A leaves out, one CTA each, chip-wide; nullptr below MS_HEAVY_MIN_N
  int32_t* heavy_list;        // [heavy_cap] sorted positions, in the order step A found them
  int32_t* heavy_slot;        // [N] by original index: slot in the heavy pool, -1 = none
  double* heavy_hdr;          // [heavy_cap][MS_HHDR]: n | final mean | (ulo, uhi) | vbase
  unsigned char* heavy_votes; // [heavy_cap][MS_HSPAN] windows per sorted position, relative to vbase
  int heavy_cap;
  // ---- step B state ----
  int32_t* tvotes;     // [Npad] by sorted position: votes of the running on-the-spot trajectory
  int32_t* vl_n;       // [N]
  int32_t* vl_id;      // [N][MS_VCAP]
  int32_t* vl_votes;   // [N][MS_VCAP]
  int32_t* spill;      // [spill_cap][3] (point, cluster, votes)
  int spill_cap;
  int32_t* hash_head;  // [1 << hash_bits], -1 = empty
  int hash_bits;
  int32_t* node_next;  // [node_cap]
  int32_t* node_centre;
  int node_cap;
  uint32_t* gmask;     // visited bits / per-1024 counters in global memory when N > 2^20
  int32_t* gcnt1;
  uint32_t rng;
  // ---- outputs ----
  double* centres;     // [max_c][D]
  int max_c;
  int32_t* assign;     // [N]
  // ctl: [0] C  [1] flags  [2] n_finite  [3] spill_used  [4] trajectories computed on the spot  [5] heavy seeds  [6,7] trajectories (u64)
  //      [8,9] window iterations (u64)  [10] rng state
  int32_t* ctl;
};
constexpr int MS_FLAG_CENTRES = 1, MS_FLAG_SPILL = 4, MS_FLAG_NODES = 8;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// order-preserving map double -> u64
__device__ __forceinline__ unsigned long long ms_key(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ bool ms_row_finite(const double* __restrict__ row, int D) {
  bool f = true;
  for (int j = 0; j < D; ++j) f = f && isfinite(row[j]);
  return f;
}

// ---- preparation, small N: one CTA — keys, bitonic sort by (coordinate 0, index), gather, initial visited mask -----------------
__global__ void __launch_bounds__(1024) ms_prep_small_kernel(MsProblem p, int P /* power of two >= max(N, 32) */) {
  extern __shared__ __align__(16) unsigned char ps_smem[];
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(ps_smem);
  int32_t* sidx = reinterpret_cast<int32_t*>(skey + P);
  __shared__ int s_nf;
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
  if (tid == 0) s_nf = 0;
  __syncthreads();
  int nf = 0;
  for (int i = tid; i < P; i += T) {
    unsigned long long k = ~0ull;   // non-finite rows (which can neither seed nor join a window) and padding sort last
    if (i < p.N && ms_row_finite(p.data + (size_t)i * p.D, p.D)) { k = ms_key(p.data[(size_t)i * p.D]); ++nf; }
    skey[i] = k;
    sidx[i] = i;
    if (i < p.Npad) p.tvotes[i] = 0;
  }
  nf = __reduce_add_sync(0xffffffffu, nf);
  if (lane == 0 && nf) atomicAdd(&s_nf, nf);
  // initial visited mask by original index
  for (int base = (tid >> 5) * 32; base < p.mask_words * 32; base += T) {
    const int i = base + lane;
    const bool dead = i >= p.N || !ms_row_finite(p.data + (size_t)i * p.D, p.D);
    const unsigned m = __ballot_sync(0xffffffffu, dead);
    if (lane == 0) p.mask0[base >> 5] = m;
  }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (P >> 1); t += T) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
        const bool up = (i & k) == 0;
        const unsigned long long ka = skey[i], kb = skey[l];
        const int ia = sidx[i], ib = sidx[l];
        const bool a_gt_b = ka > kb || (ka == kb && ia > ib);
        if (a_gt_b == up) { skey[i] = kb; skey[l] = ka; sidx[i] = ib; sidx[l] = ia; }
      }
      __syncthreads();
    }
  for (int q = tid; q < p.N; q += T) {
    const int i = sidx[q];
    p.perm[q] = i;
    p.vl_n[i] = 0;
    for (int j = 0; j < p.D; ++j) p.xs[(size_t)j * p.Npad + q] = p.data[(size_t)i * p.D + j];
  }
  if (tid == 0) p.ctl[2] = s_nf;
}

// ---- preparation, large N: keys -> cub radix sort -> gather ----------------------------------------------------------------------
__global__ void ms_keys_kernel(MsProblem p, unsigned long long* __restrict__ keys, int32_t* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  const bool live = i < p.N;
  const bool fin = live && ms_row_finite(p.data + (size_t)i * p.D, p.D);
  if (live) {
    keys[i] = fin ? ms_key(p.data[(size_t)i * p.D]) : ~0ull;
    idx[i] = i;
    p.vl_n[i] = 0;
  }
  if (i < p.Npad) p.tvotes[i] = 0;
  const unsigned dead = __ballot_sync(0xffffffffu, !fin);
  if (lane == 0 && (i >> 5) < p.mask_words) p.mask0[i >> 5] = dead;
  const int nf = __popc(~dead);
  if (lane == 0 && nf) atomicAdd(p.ctl + 2, nf);
}
__global__ void ms_mask_tail_kernel(MsProblem p, int first_word) {   // words past the last launched warp of ms_keys_kernel
  const int w = first_word + blockIdx.x * blockDim.x + threadIdx.x;
  if (w < p.mask_words) p.mask0[w] = 0xffffffffu;
}
__global__ void ms_gather_kernel(MsProblem p, const int32_t* __restrict__ sorted_idx) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= p.N) return;
  const int i = sorted_idx[q];
  p.perm[q] = i;
  for (int j = 0; j < p.D; ++j) p.xs[(size_t)j * p.Npad + q] = p.data[(size_t)i * p.D + j];
}

// ---- the window test ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ms_term(double d, int metric) { return metric == 0 ? fabs(d) : d * d; }   // sqrt(d*d): MS.h:80-82

// The sorted rows as a kernel sees them: global memory (step A, large step B) or a shared-memory copy (small step B).
struct MsRows {
  const double* xs;   // [D][stride]
  size_t stride;
  const int32_t* perm;
};

// Candidate range [lo, hi) of a window with mean coordinate m0 among the Nf finite rows: positions whose FIRST term alone does
// not already fail the test.  Terms are non-negative and the running sum only grows, so everything outside fails the full test;
// the predicate is evaluated exactly as the window test evaluates its first term, and it is monotone along the sorted
// coordinate.  Both partition points are found together by a 32-ary search: every lane probes one position per bound and round.
__device__ __forceinline__ void ms_range(const double* xs0, int Nf, double m0, double bandSq, int metric, int lane, int& lo, int& hi,
                                         bool hint = false) {
  auto left = [&](int q) {    // "too far left": true ... true false ... false along the sorted coordinate
    const double d = m0 - xs0[q];
    return d > 0.0 && !(ms_term(d, metric) < bandSq);
  };
  auto inside = [&](int q) {  // "not yet too far right": true ... true false ... false
    const double d = m0 - xs0[q];
    return !(d < 0.0 && !(ms_term(d, metric) < bandSq));
  };
  int a1 = 0, b1 = Nf, a2 = 0, b2 = Nf;   // partition points lie in [a, b]
  if (hint) {
    // the mean moved a little: both bounds are usually within 16 positions of the previous window's (lo, hi on entry) —
    // one round of 32 unit-step probes around each; a bound that is not bracketed falls back to the search on its side
    const int w1 = max(0, min(lo - 16, Nf - 32)), w2 = max(0, min(hi - 16, Nf - 32));
    const int q1 = w1 + lane, q2 = w2 + lane;
    const bool t1 = q1 < Nf && left(q1), t2 = q2 < Nf && inside(q2);
    const int n1 = min(32, Nf - w1), n2 = min(32, Nf - w2);
    const int c1 = __popc(__ballot_sync(0xffffffffu, t1)), c2 = __popc(__ballot_sync(0xffffffffu, t2));
    if (c1 == 0) b1 = w1; else if (c1 == n1) a1 = w1 + n1; else a1 = b1 = w1 + c1;
    if (c2 == 0) b2 = w2; else if (c2 == n2) a2 = w2 + n2; else a2 = b2 = w2 + c2;
  }
  while (b1 > a1 || b2 > a2) {
    const int len1 = b1 - a1, step1 = (len1 + 31) >> 5, q1 = a1 + lane * step1;
    const int len2 = b2 - a2, step2 = (len2 + 31) >> 5, q2 = a2 + lane * step2;
    const bool t1 = len1 > 0 && q1 < b1 && left(q1), t2 = len2 > 0 && q2 < b2 && inside(q2);
    const int c1 = __popc(__ballot_sync(0xffffffffu, t1)), c2 = __popc(__ballot_sync(0xffffffffu, t2));
    if (len1 > 0) {
      if (c1 == 0) b1 = a1;
      else { const int na = a1 + (c1 - 1) * step1 + 1; b1 = min(b1, a1 + c1 * step1); a1 = na; }
    }
    if (len2 > 0) {
      if (c2 == 0) b2 = a2;
      else { const int na = a2 + (c2 - 1) * step2 + 1; b2 = min(b2, a2 + c2 * step2); a2 = na; }
    }
  }
  lo = a1;
  hi = a2;
}

// The window test of MS.h:76-85 for the row at sorted position q: s = sum_j term(mean_j - x_j) < bw^2, summed in the reference's
// order.  The first MS_STAGE1 coordinates are fetched and summed first; a partial sum that already fails ends the test (the sum
// only grows), which spares most candidates the rest of their row.  x holds the row when the test passes.
template <int DT, bool STAGED = true, typename MeanT>
__device__ __forceinline__ bool ms_member(const MsProblem& p, const MsRows& r, int q, const MeanT& mean, double (&x)[DT]) {
  const int D = p.D;
  double s = 0.0;
  if (!STAGED) {   // short candidate lists: one round trip for the whole row beats the saved traffic
#pragma unroll
    for (int j = 0; j < DT; ++j)
      if (j < D) x[j] = r.xs[(size_t)j * r.stride + q];
#pragma unroll
    for (int j = 0; j < DT; ++j)
      if (j < D) s += ms_term(mean[j] - x[j], p.metric);
    return s < p.bandSq;
  }
#pragma unroll
  for (int j = 0; j < MS_STAGE1; ++j)
    if (j < D) {
      x[j] = r.xs[(size_t)j * r.stride + q];
      s += ms_term(mean[j] - x[j], p.metric);
    }
  if (!(s < p.bandSq)) return false;
#pragma unroll
  for (int j = MS_STAGE1; j < DT; ++j)
    if (j < D) {
      x[j] = r.xs[(size_t)j * r.stride + q];
      s += ms_term(mean[j] - x[j], p.metric);
    }
  return s < p.bandSq;
}

// ---- step A: one trajectory (MS.h:62-98) per warp ----------------------------------------------------------------------------------
// Record of a speculated trajectory, MS_REC doubles per point (by ORIGINAL index):
//   [0] number of windows, or -1: not speculated (step B computes the trajectory when the point is drawn)
//   [1, D] final mean    [D + 1] union of the windows' candidate ranges (ulo, uhi as two int32)
//   [D + 2 ...] one byte per position of [ulo, uhi): in how many windows of the trajectory the row was a member (MS.h:87)
// so step B replays a speculated trajectory without touching the rows or doing any arithmetic.
constexpr int MS_VOTE_SPAN = 2048;   // widest union of candidate ranges step A records
__host__ __device__ constexpr int ms_rec_stride(int D) { return D + 2 + MS_VOTE_SPAN / 8; }

constexpr int MS_VOTE_SHORT = 512;   // step B fetches the header and this many vote bytes at once, the rest only when needed
__host__ __device__ constexpr int ms_rec_short(int D) { return D + 2 + MS_VOTE_SHORT / 8; }

struct MsWarpScratch {   // per warp of step A
  unsigned char votes[MS_VOTE_SPAN + 16];   // + 16: the record packing reads whole words
  double red[MS_MAXD * 33];                 // [coordinate][lane], padded: conflict-free both ways
  double bc[2 * MS_MAXD];                   // new mean | squared step per coordinate
};

// `mean` holds the seed (sorted position pos) on entry and the final mean on exit (replicated in all lanes); ws.votes (zero on
// entry) counts the windows per position relative to vbase.  Returns the number of windows, or -1 when the trajectory is not
// worth speculating (see the header).
template <int DT>
__device__ __forceinline__ int ms_warp_trajectory(const MsProblem& p, const MsRows& r, int Nf, int pos, double (&mean)[DT],
                                                  MsWarpScratch& ws, int& vbase, int& ulo, int& uhi, int lane) {
  const int D = p.D;
  int it = 0, lo = pos, hi = pos;   // the seed is a member of its first window: its position is the first hint
  ulo = 0x7fffffff;
  uhi = 0;
  for (;;) {
    ms_range(r.xs, Nf, mean[0], p.bandSq, p.metric, lane, lo, hi, true);
    if (it == 0) vbase = max(0, lo - (MS_VOTE_SPAN - min(hi - lo, MS_VOTE_SPAN)) / 2);   // room to drift either way
    if (it >= MS_CAP_A || hi - lo > MS_RANGE_A || lo < vbase || hi > vbase + MS_VOTE_SPAN) return -1;
    ulo = min(ulo, lo);
    uhi = max(uhi, hi);
    double acc[DT];
#pragma unroll
    for (int j = 0; j < DT; ++j) acc[j] = 0.0;
    int cnt = 0;
    auto scan = [&](auto staged) {
      for (int q = lo + lane; q < hi; q += 32) {
        double x[DT];
        if (ms_member<DT, decltype(staged)::value>(p, r, q, mean, x)) {   // MS.h:85
#pragma unroll
          for (int j = 0; j < DT; ++j)
            if (j < D) acc[j] += x[j];
          ++cnt;
          ws.votes[q - vbase] += 1;   // one lane per position and window
        }
      }
    };
    if (hi - lo > 128) scan(std::true_type{});
    else scan(std::false_type{});
    // sum over the lanes through shared memory: lane j adds up coordinate j in lane order, divides, and publishes the new mean
#pragma unroll
    for (int j = 0; j < DT; ++j)
      if (j < D) ws.red[j * 33 + lane] = acc[j];
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    __syncwarp();
    ++it;
    if (lane < D) {
      double tot = 0.0;
#pragma unroll
      for (int l = 0; l < 32; ++l) tot += ws.red[lane * 33 + l];
      const double nm = tot / (double)cnt;   // MS.h:96
      double old = 0.0;
#pragma unroll
      for (int j = 0; j < DT; ++j)
        if (lane == j) old = mean[j];
      const double d = nm - old;
      ws.bc[lane] = nm;
      ws.bc[MS_MAXD + lane] = d * d;
    }
    __syncwarp();
    double n2 = 0.0;
#pragma unroll
    for (int j = 0; j < DT; ++j)
      if (j < D) {
        n2 += ws.bc[MS_MAXD + j];
        mean[j] = ws.bc[j];
      }
    __syncwarp();
    if (sqrt(n2) < p.stopThresh || cnt == 0 || it >= MH_MS_MAX_WINDOW_ITERS) break;   // MS.h:98
  }
  return it;
}

template <int DT>
__global__ void __launch_bounds__(MS_TRAJ_THREADS) ms_trajectories_kernel(MsProblem p) {
  extern __shared__ __align__(16) unsigned char tr_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * MS_TRAJ_THREADS + threadIdx.x) >> 5, nwarps = (gridDim.x * MS_TRAJ_THREADS) >> 5;
  const int Nf = p.ctl[2], D = p.D;
  const MsRows r{p.xs, (size_t)p.Npad, p.perm};
  MsWarpScratch& ws = reinterpret_cast<MsWarpScratch*>(tr_smem)[wib];
  for (int pos = warp; pos < Nf; pos += nwarps) {
    double mean[DT];
#pragma unroll
    for (int j = 0; j < DT; ++j) mean[j] = j < D ? p.xs[(size_t)j * p.Npad + pos] : 0.0;
#pragma unroll
    for (int k = 0; k < MS_VOTE_SPAN / (32 * 16); ++k) reinterpret_cast<uint4*>(ws.votes)[lane + 32 * k] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    double* blk = p.rec + (size_t)p.perm[pos] * p.rec_stride;
    int vbase = 0, ulo, uhi;
    const int n = ms_warp_trajectory<DT>(p, r, Nf, pos, mean, ws, vbase, ulo, uhi, lane);
    double v = (double)n;
#pragma unroll
    for (int j = 0; j < DT; ++j)
      if (lane == j + 1) v = mean[j];
    if (lane == D + 1) v = __hiloint2double(uhi, ulo);
    if (lane <= D + 1) blk[lane] = v;
    if (p.heavy_slot && lane == 0) {   // not speculated here: a CTA of step A2 takes it
      int slot = -1;
      if (n < 0) {
        const int sl = atomicAdd(p.ctl + 5, 1);
        if (sl < p.heavy_cap) { slot = sl; p.heavy_list[sl] = pos; }
      }
      p.heavy_slot[p.perm[pos]] = slot;
    }
    if (n > 0) {
      // at least the part step B fetches unconditionally is written whole, so that the fetch finds its lines in L2
      uint32_t* out = reinterpret_cast<uint32_t*>(blk + D + 2);
      const unsigned char* src = ws.votes + (ulo - vbase);
      const int nbytes = uhi - ulo;
      for (int w = lane; 4 * w < max(nbytes, MS_VOTE_SHORT); w += 32)
        out[w] = 4 * w < nbytes ? (uint32_t)src[4 * w] | ((uint32_t)src[4 * w + 1] << 8) | ((uint32_t)src[4 * w + 2] << 16) |
                                      ((uint32_t)src[4 * w + 3] << 24)
                                : 0u;
    }
    __syncwarp();
  }
}

// ---- step B ------------------------------------------------------------------------------------------------------------------------
// First index of count(0..n) at which the running sum exceeds `rank`; rank is reduced by the sum in front of it (warp-wide).
template <typename F>
__device__ __forceinline__ int ms_prefix_find(F count, int n, int& rank, int lane) {
  for (int base = 0; base < n; base += 32) {
    const int v = base + lane < n ? count(base + lane) : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (rank < total) {
      const int l = __ffs(__ballot_sync(0xffffffffu, rank < incl)) - 1;
      rank -= __shfl_sync(0xffffffffu, incl - v, l);
      return base + l;
    }
    rank -= total;
  }
  return -1;
}

__device__ __forceinline__ long long ms_cell(double x, double cellInv) {
  return (long long)floor(fmin(fmax(x * cellInv, -4.0e18), 4.0e18));
}
__device__ __forceinline__ unsigned ms_hash(long long c0, long long c1, long long c2, int bits) {
  unsigned long long h = (unsigned long long)c0 * 0x9E3779B97F4A7C15ull;
  h ^= (unsigned long long)c1 * 0xC2B2AE3D27D4EB4Full + (h >> 29);
  h ^= (unsigned long long)c2 * 0x165667B19E3779F9ull + (h << 17);
  h *= 0xD6E8FEB86659FD93ull;
  return (unsigned)(h >> (64 - bits));
}

constexpr int MS_REC_MAX = ms_rec_stride(MS_MAXD);
static_assert(ms_rec_short(MS_MAXD) + MS_MAXD <= MS_REPLAY_THREADS, "the record fetch of step B is one value per thread");
struct MsShared {   // fixed part of step B's shared memory
  double rec[MS_REC_MAX];                        // the drawn seed's record (see step A)
  double cur[MS_MAXD];                           // running / final mean of a trajectory computed on the spot
  double red[MS_REPLAY_WARPS][MS_MAXD + 1];      // per-warp partial sums of an on-the-spot window
  int redc[MS_REPLAY_WARPS];
  int cnt2[1024];                                // unvisited per 2^20 points
  int cnt1[1024];                                // unvisited per 1024 points (N <= 2^20)
  int seed, remaining, merge, stop;
};

// One trajectory computed by the whole CTA, for a seed step A did not speculate (dense neighbourhood or slow drift): every window
// scans its candidate range with all threads and leaves one vote per member in tvotes (by sorted position).  sh.cur holds
// the seed on entry and the final mean on exit.  Returns the number of windows; [ulo, uhi) = union of the candidate ranges.
// `votes`: begin(it, lo, hi) -> false abandons the trajectory (returns -1; CTA-uniform), add(q) records one window for position q.
template <int DT, class SH, class Votes>
__device__ __forceinline__ int ms_block_trajectory(const MsProblem& p, const MsRows& r, Votes& votes, SH& sh, int Nf,
                                                   int& ulo, int& uhi) {
  constexpr int T = MS_REPLAY_THREADS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, D = p.D;
  ulo = 0x7fffffff;
  uhi = 0;
  int it = 0, lo = 0, hi = 0;
  for (;;) {
    double mean[DT];
#pragma unroll
    for (int j = 0; j < DT; ++j) mean[j] = j < D ? sh.cur[j] : 0.0;
    ms_range(r.xs, Nf, mean[0], p.bandSq, p.metric, lane, lo, hi, it > 0);   // every warp finds the same range
    if (!votes.begin(it, lo, hi)) return -1;
    ulo = min(ulo, lo);
    uhi = max(uhi, hi);
    double acc[DT];
#pragma unroll
    for (int j = 0; j < DT; ++j) acc[j] = 0.0;
    int cnt = 0;
    auto scan = [&](auto staged) {
      for (int q = lo + tid; q < hi; q += T) {
        double x[DT];
        if (ms_member<DT, decltype(staged)::value>(p, r, q, mean, x)) {
#pragma unroll
          for (int j = 0; j < DT; ++j)
            if (j < D) acc[j] += x[j];
          ++cnt;
          votes.add(q);   // one thread per position and window; windows are separated by barriers
        }
      }
    };
    if (hi - lo > 4 * T) scan(std::true_type{});
    else scan(std::false_type{});
    cnt = __reduce_add_sync(0xffffffffu, cnt);
#pragma unroll
    for (int j = 0; j < DT; ++j)
      if (j < D) {
        const double v = warp_sum_d(acc[j]);
        if (lane == 0) sh.red[warp][j] = v;
      }
    if (lane == 0) sh.redc[warp] = cnt;
    __syncthreads();
    ++it;
    if (warp == 0) {
      int c = 0;
      for (int w = 0; w < MS_REPLAY_WARPS; ++w) c += sh.redc[w];
      double d2 = 0.0;
      if (lane < D) {
        double tot = 0.0;
        for (int w = 0; w < MS_REPLAY_WARPS; ++w) tot += sh.red[w][lane];
        const double nm = tot / (double)c;   // MS.h:96
        const double d = nm - sh.cur[lane];
        d2 = d * d;
        sh.cur[lane] = nm;
      }
      double n2 = 0.0;
      for (int j = 0; j < D; ++j) n2 += __shfl_sync(0xffffffffu, d2, j);
      if (lane == 0) sh.stop = (sqrt(n2) < p.stopThresh || c == 0 || it >= MH_MS_MAX_WINDOW_ITERS) ? 1 : 0;   // MS.h:98
    }
    __syncthreads();
    if (sh.stop) break;
  }
  return it;
}

struct MsVotesInPlace {   // the on-the-spot trajectory of step B: votes by sorted position in global memory, consumed right away
  int32_t* tvotes;
  __device__ __forceinline__ bool begin(int, int, int) { return true; }
  __device__ __forceinline__ void add(int q) { tvotes[q] += 1; }
};
struct MsVotesHeavy {     // step A2: one byte per position of a fixed span around the first window
  unsigned char* bytes;
  int vbase;
  __device__ __forceinline__ bool begin(int it, int lo, int hi) {
    if (it == 0) vbase = max(0, lo - (MS_HSPAN - min(hi - lo, MS_HSPAN)) / 2);   // room to drift either way
    return lo >= vbase && hi <= vbase + MS_HSPAN;
  }
  __device__ __forceinline__ void add(int q) { bytes[q - vbase] += 1; }   // <= MH_MS_MAX_WINDOW_ITERS (200) windows: fits a byte
};
struct MsHeavyShared {
  double cur[MS_MAXD];
  double red[MS_REPLAY_WARPS][MS_MAXD + 1];
  int redc[MS_REPLAY_WARPS];
  int stop;
};

// ---- step A2: the trajectories step A did not speculate (dense neighbourhoods, slow drifts), one CTA per seed, all SMs -------
// Same code as the replay's on-the-spot trajectory; the windows' votes go to the seed's heavy record.  A trajectory whose windows
// leave the record's span is abandoned (n = -1) and computed by the replay if it ever draws that seed.
template <int DT>
__global__ void __launch_bounds__(MS_REPLAY_THREADS) ms_heavy_kernel(MsProblem p) {
  __shared__ MsHeavyShared sh;
  const int tid = threadIdx.x, D = p.D, Nf = p.ctl[2];
  const int nh = min(p.ctl[5], p.heavy_cap);
  const MsRows rows{p.xs, (size_t)p.Npad, p.perm};
  for (int slot = blockIdx.x; slot < nh; slot += gridDim.x) {
    const int pos = p.heavy_list[slot];
    unsigned char* bytes = p.heavy_votes + (size_t)slot * MS_HSPAN;
    for (int i = tid; i < MS_HSPAN / 16; i += MS_REPLAY_THREADS) reinterpret_cast<uint4*>(bytes)[i] = make_uint4(0, 0, 0, 0);
    if (tid < D) sh.cur[tid] = p.xs[(size_t)tid * p.Npad + pos];
    __syncthreads();
    MsVotesHeavy votes{bytes, 0};
    int ulo, uhi;
    const int n = ms_block_trajectory<DT>(p, rows, votes, sh, Nf, ulo, uhi);
    double* hdr = p.heavy_hdr + (size_t)slot * MS_HHDR;
    if (tid == 0) {
      hdr[0] = (double)n;
      hdr[D + 1] = __hiloint2double(uhi, ulo);
      hdr[D + 2] = (double)votes.vbase;
    }
    if (n > 0 && tid < D) hdr[1 + tid] = sh.cur[tid];
    __syncthreads();   // sh.cur is rewritten by the next seed
  }
}

struct MsVoteList {   // a point's (cluster, votes) list header, fetched ahead of its update
  int nl;
  int4 a, b, c, d;
};

// MODE 0: the visited bits and the centres live in shared memory; MODE 1: the visited bits do, the centres are in global
// memory; MODE 2 (N > 2^20): the bits and their counters are in global memory too.  The modes are template instances so that
// every pointer has one address space (LDS / ATOMS instead of generic accesses).
template <int DT, int MODE>
__global__ void __launch_bounds__(MS_REPLAY_THREADS, 1) ms_replay_kernel(MsProblem p) {
  extern __shared__ __align__(16) unsigned char rp_smem[];
  MsShared& sh = *reinterpret_cast<MsShared*>(rp_smem);
  uint32_t* s_masks = reinterpret_cast<uint32_t*>(rp_smem + sizeof(MsShared));

  constexpr int T = MS_REPLAY_THREADS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N, D = p.D, Nf = p.ctl[2];
  constexpr bool in_smem = MODE != 2, centres_smem = MODE == 0;
  uint32_t* mask;
  int* cnt1;
  if constexpr (in_smem) { mask = s_masks; cnt1 = sh.cnt1; }
  else { mask = p.gmask; cnt1 = p.gcnt1; }
  const int nb1 = p.mask_words >> 5, nb2 = (nb1 + 1023) >> 10;
  double* cen;   // what the merge test reads; the output array p.centres is written alongside
  if constexpr (centres_smem) cen = reinterpret_cast<double*>(s_masks + p.mask_words);
  else cen = p.centres;
  const MsRows rows{p.xs, (size_t)p.Npad, p.perm};
  int32_t* tvotes = p.tvotes;

  for (int w = tid; w < p.mask_words; w += T) mask[w] = p.mask0[w];
  __syncthreads();
  for (int b = tid; b < nb1; b += T) {
    int c = 0;
    for (int k = 0; k < 32; ++k) c += __popc(~mask[b * 32 + k]);
    cnt1[b] = c;
  }
  __syncthreads();
  for (int b2 = tid; b2 < nb2; b2 += T) {
    int c = 0;
    for (int k = b2 << 10; k < min(nb1, (b2 + 1) << 10); ++k) c += cnt1[k];
    sh.cnt2[b2] = c;
  }
  __syncthreads();
  if (tid == 0) {
    int c = 0;
    for (int b2 = 0; b2 < nb2; ++b2) c += sh.cnt2[b2];
    sh.remaining = c;
  }
  __syncthreads();

  // a window member becomes visited (MS.h:91); called by whole warps: the total is updated once per warp ...
  auto mark_visited = [&](bool member, int i) {
    bool newly = false;
    if (member) {
      const uint32_t bit = 1u << (i & 31);
      newly = !(atomicOr(&mask[i >> 5], bit) & bit);
      if (newly) {
        atomicSub(&cnt1[i >> 10], 1);
        if (nb2 > 1) atomicSub(&sh.cnt2[i >> 20], 1);
      }
    }
    const unsigned nm = __ballot_sync(0xffffffffu, newly);
    if (nm && lane == __ffs(nm) - 1) atomicSub(&sh.remaining, __popc(nm));
  };
  int flags = 0;
  // ... and its votes go to the trajectory's cluster (MS.h:87, 114, 119).  The list header is one count + four 16-byte loads,
  // issued before the cluster id is known.
  auto fetch_list = [&](int i) {
    MsVoteList l;
    const int4* idp = reinterpret_cast<const int4*>(p.vl_id + (size_t)i * MS_VCAP);
    l.nl = p.vl_n[i];
    l.a = idp[0]; l.b = idp[1]; l.c = idp[2]; l.d = idp[3];
    return l;
  };
  auto add_votes = [&](int i, const MsVoteList& l, int cid, int v) {
    const int ids[16] = {l.a.x, l.a.y, l.a.z, l.a.w, l.b.x, l.b.y, l.b.z, l.b.w, l.c.x, l.c.y, l.c.z, l.c.w, l.d.x, l.d.y, l.d.z, l.d.w};
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) m |= (ids[k] == cid ? 1u : 0u) << k;
    m &= (1u << l.nl) - 1u;
    if (m) {
      atomicAdd(p.vl_votes + (size_t)i * MS_VCAP + (__ffs(m) - 1), v);
    } else if (l.nl < MS_VCAP) {
      p.vl_id[(size_t)i * MS_VCAP + l.nl] = cid;
      p.vl_votes[(size_t)i * MS_VCAP + l.nl] = v;
      p.vl_n[i] = l.nl + 1;
    } else {
      const int s = atomicAdd(p.ctl + 3, 1);
      if (s < p.spill_cap) { p.spill[3 * (size_t)s] = i; p.spill[3 * (size_t)s + 1] = cid; p.spill[3 * (size_t)s + 2] = v; }
      else flags |= MS_FLAG_SPILL;
    }
  };
  // squared distance of the running final mean to a centre, summed in the reference's order (norm(myMean - clustCent), MS.h:103)
  auto centre_d2 = [&](const double* fm, const double* c) {
    double cv[DT];
#pragma unroll
    for (int j = 0; j < DT; ++j)
      if (j < D) cv[j] = c[j];
    double d2 = 0.0;
#pragma unroll
    for (int j = 0; j < DT; ++j)
      if (j < D) {
        const double d = fm[j] - cv[j];
        d2 += d * d;
      }
    return d2;
  };
  // sqrt(d2) < bw/2, with the square root only where d2 is within rounding distance of (bw/2)^2
  const double hb2 = p.halfBw * p.halfBw, hb2_lo = hb2 * (1.0 - 1e-12), hb2_hi = hb2 * (1.0 + 1e-12);
  auto is_close = [&](double d2) { return d2 < hb2_lo || (d2 < hb2_hi && sqrt(d2) < p.halfBw); };

  uint32_t hold = p.rng;
  int C = 0, nnodes = 0, on_the_spot = 0;
  unsigned long long traj = 0, iters = 0;
#ifdef MS_PROFILE
  long long prof[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pt = clock64();
#define MS_TICK(k) do { const long long t_ = clock64(); prof[k] += t_ - pt; pt = t_; } while (0)
#else
#define MS_TICK(k) do { } while (0)
#endif
  for (;;) {
    // ---- the next seed: the rank-th unvisited point in index order (MS.h:54-56) ----------------------------------------
    if (warp == 0) {
      const int remaining = sh.remaining;
      int seed = -1;
      if (remaining > 0) {
        hold = hold * 214013u + 2531011u;   // MSVC rand()
        // tempInd = round(rand() / RAND_MAX * (size - 1)) (MS.h:54-55) in integers: with r = rand() and m = size - 1 the exact
        // product r m / 32767 is never half-way between integers (32767 (2k + 1) is odd, 2 r m even) and lies at least
        // 1 / 65534 away from such a point, far more than the two roundings of the double expression can move it for any
        // m < 2^31 — so round-to-nearest of the exact quotient is the reference's result.
        // Split m = 32767 q + t: r m / 32767 = r q + r t / 32767 with r t < 2^30, so 32-bit arithmetic suffices.
        const unsigned r = (hold >> 16) & 0x7fffu, m = (unsigned)(remaining - 1);
        int rank = (int)(r * (m / 32767u) + (2u * r * (m % 32767u) + 32767u) / 65534u);
        MS_TICK(8);
        const int b2 = nb2 > 1 ? ms_prefix_find([&](int k) { return sh.cnt2[k]; }, nb2, rank, lane) : 0;
        const int b1 = nb1 > 1 ? (b2 << 10) + ms_prefix_find([&](int k) { return cnt1[(b2 << 10) + k]; }, min(1024, nb1 - (b2 << 10)), rank, lane) : 0;
        MS_TICK(9);
        const int l = ms_prefix_find([&](int k) { return __popc(~mask[b1 * 32 + k]); }, 32, rank, lane);
        const uint32_t wsel = ~mask[b1 * 32 + l];
        // the rank-th (0-based) set bit of wsel: lane k looks at bit k
        const bool hit = ((wsel >> lane) & 1u) && __popc(wsel & ((2u << lane) - 1u)) == rank + 1;
        seed = (b1 * 32 + l) * 32 + (__ffs(__ballot_sync(0xffffffffu, hit)) - 1);
        MS_TICK(10);
      }
      if (lane == 0) { sh.seed = seed; sh.merge = 0x7fffffff; }
    }
    __syncthreads();
    MS_TICK(0);
    const int seed = sh.seed;
    if (seed < 0) break;
    // the seed's record (header + the first vote bytes) and its own coordinates: one round trip, whatever the record says
    {
      const int short_len = ms_rec_short(D);
      if (tid < short_len) sh.rec[tid] = p.rec[(size_t)seed * p.rec_stride + tid];
      else if (tid < short_len + D) sh.cur[tid - short_len] = p.data[(size_t)seed * D + (tid - short_len)];
    }
    __syncthreads();
    MS_TICK(1);
    int n = (int)sh.rec[0];
    const bool speculated = n > 0;
    int ulo = __double2loint(sh.rec[D + 1]), uhi = __double2hiint(sh.rec[D + 1]);
    const unsigned char* hvotes = nullptr;   // a heavy record of step A2: vote bytes in global memory, relative to hbase
    int hbase = 0;
    if (!speculated && p.heavy_slot) {
      const int slot = p.heavy_slot[seed];
      if (slot >= 0) {
        const double* hdr = p.heavy_hdr + (size_t)slot * MS_HHDR;
        const int nh = (int)hdr[0];
        if (nh > 0) {
          n = nh;
          const double ur = hdr[D + 1];
          ulo = __double2loint(ur); uhi = __double2hiint(ur);
          hbase = (int)hdr[D + 2];
          hvotes = p.heavy_votes + (size_t)slot * MS_HSPAN;
          if (tid < D) sh.cur[tid] = hdr[1 + tid];   // the final mean
          __syncthreads();
        }
      }
    }
    if (!speculated && !hvotes) {
      MsVotesInPlace inplace{tvotes};
      n = ms_block_trajectory<DT>(p, rows, inplace, sh, Nf, ulo, uhi);
      ++on_the_spot;
    } else if (speculated && uhi - ulo > MS_VOTE_SHORT) {   // a wide trajectory: the rest of its vote bytes
      for (int i = ms_rec_short(D) + tid; i < D + 2 + (uhi - ulo + 7) / 8; i += T) sh.rec[i] = p.rec[(size_t)seed * p.rec_stride + i];
      __syncthreads();
    }
    MS_TICK(2);
    const double* fm = speculated ? sh.rec + 1 : sh.cur;   // final mean
    const unsigned char* rvotes = reinterpret_cast<const unsigned char*>(sh.rec + D + 2);
    ++traj;
    iters += (unsigned long long)n;

    // votes of position q: as recorded by step A, or as left behind by the on-the-spot trajectory
    auto votes_of = [&](int q) {
      if (speculated) return (int)rvotes[q - ulo];
      if (hvotes) return (int)hvotes[q - hbase];
      const int v = tvotes[q];
      if (v) tvotes[q] = 0;
      return v;
    };
    // first candidate of this thread: its loads overlap the merge test
    const int q0 = ulo + tid;
    int v0 = 0, i0 = 0;
    MsVoteList l0;
    l0.nl = 0;
    if (q0 < uhi) {
      v0 = votes_of(q0);
      if (v0) i0 = rows.perm[q0];
    }
    MS_TICK(11);
    mark_visited(v0 != 0, i0);
    MS_TICK(12);
    if (v0) l0 = fetch_list(i0);
    MS_TICK(13);

    MS_TICK(3);
    // ---- merge test: the FIRST centre closer than bw/2 (MS.h:100-109) ----------------------------------------------------
    {
      int cand = 0x7fffffff;
      if (C < MS_BRUTE_CENTRES) {
        for (int c = tid; c < C && cand == 0x7fffffff; c += T)
          if (is_close(centre_d2(fm, cen + (size_t)c * D))) cand = c;   // ascending c per thread: the first hit is its lowest
      } else if (warp == 0 && lane < 27) {
        // a centre within bw/2 differs by less than one cell (width bw/2) in every hashed coordinate
        long long cell[3] = {0, 0, 0};
        bool valid = true;
        int o = lane;
#pragma unroll
        for (int j = 0; j < MS_HASH_DIMS; ++j) {
          if (j < D) cell[j] = ms_cell(fm[j], p.cellInv) + (o % 3 - 1);
          else if (o % 3 != 1) valid = false;   // fewer than 3 dimensions: one neighbour cell per missing one
          o /= 3;
        }
        if (valid) {
          for (int node = p.hash_head[ms_hash(cell[0], cell[1], cell[2], p.hash_bits)]; node >= 0; node = p.node_next[node]) {
            const int c = p.node_centre[node];
            if (is_close(centre_d2(fm, cen + (size_t)c * D))) cand = min(cand, c);
          }
        }
      }
      cand = __reduce_min_sync(0xffffffffu, cand);
      if (lane == 0 && cand != 0x7fffffff) atomicMin(&sh.merge, cand);
    }
    __syncthreads();
    MS_TICK(4);
    const int mergeWith = sh.merge;
    const bool merged = mergeWith != 0x7fffffff;
    const int cid = merged ? mergeWith : C;
    const bool room = merged || C < p.max_c;
    if (!room) flags |= MS_FLAG_CENTRES;

    // ---- centre update (MS.h:110-120) and its cell-hash entry: the last warp (it rarely holds candidates), one lane per
    // coordinate, while the others record the votes ----------------------------------------------------------------------
    if (warp == MS_REPLAY_WARPS - 1 && room) {
      double nv = 0.0;
      long long oc = 0, nc = 0;
      if (lane < D) {
        const double old = merged ? cen[(size_t)cid * D + lane] : 0.0;
        nv = merged ? 0.5 * (old + fm[lane]) : fm[lane];
        p.centres[(size_t)cid * D + lane] = nv;
        if constexpr (centres_smem) cen[(size_t)cid * D + lane] = nv;
        oc = ms_cell(old, p.cellInv);
        nc = ms_cell(nv, p.cellInv);
      }
      // The cell hash exists only once there are too many centres to scan (see below): stale entries stay behind in a moved
      // centre's old cell — every hit is verified against the centre's current coordinates.
      if (C >= MS_BRUTE_CENTRES) {
        const bool hashed = lane < D && lane < MS_HASH_DIMS;
        const bool fin = __all_sync(0xffffffffu, !hashed || isfinite(nv));
        const bool moved = __any_sync(0xffffffffu, hashed && oc != nc);
        const long long n0 = __shfl_sync(0xffffffffu, nc, 0), n1 = __shfl_sync(0xffffffffu, D > 1 ? nc : 0, 1),
                        n2 = __shfl_sync(0xffffffffu, D > 2 ? nc : 0, 2);
        if (lane == 0 && fin && (!merged || moved)) {
          if (nnodes < p.node_cap) {
            const unsigned h = ms_hash(n0, n1, n2, p.hash_bits);
            p.node_centre[nnodes] = cid;
            p.node_next[nnodes] = p.hash_head[h];
            p.hash_head[h] = nnodes;
            ++nnodes;
          } else {
            flags |= MS_FLAG_NODES;
          }
        }
      }
    }

    MS_TICK(5);
    // ---- votes and visited flags of every window of the trajectory (MS.h:85-93, 114, 119) -----------------------------
    if (v0 && room) add_votes(i0, l0, cid, v0);
    for (int qb = ulo + T + warp * 32; qb < uhi; qb += T) {   // whole warps: mark_visited aggregates per warp
      const int q = qb + lane;
      int v = 0, i = 0;
      if (q < uhi) {
        v = votes_of(q);
        if (v) i = rows.perm[q];
      }
      mark_visited(v != 0, i);
      if (v && room) add_votes(i, fetch_list(i), cid, v);
    }
    if (!merged && room) ++C;
    MS_TICK(6);
    __syncthreads();   // closes the step: visited counts, vote lists and the centre are in place for the next seed
    MS_TICK(7);
    if (!merged && room && C == MS_BRUTE_CENTRES) {
      // from here on the merge test uses the cell hash: enter the centres found so far (node c = centre c; later entries
      // follow from MS_BRUTE_CENTRES on)
      for (int c = tid; c < C; c += T) {
        const double* cc = cen + (size_t)c * D;
        bool fin = true;
        long long cl[3] = {0, 0, 0};
#pragma unroll
        for (int j = 0; j < MS_HASH_DIMS; ++j)
          if (j < D) { cl[j] = ms_cell(cc[j], p.cellInv); fin = fin && isfinite(cc[j]); }
        if (fin) {
          p.node_centre[c] = c;
          p.node_next[c] = atomicExch(&p.hash_head[ms_hash(cl[0], cl[1], cl[2], p.hash_bits)], c);
        }
      }
      nnodes = MS_BRUTE_CENTRES;
      __syncthreads();
    }
  }

  // ---- final assignment: most votes, the first (lowest) cluster wins ties (MS.h:133-146) ----------------------------------
  const int S = min(p.ctl[3], p.spill_cap);
  for (int i = tid; i < N; i += T) {
    const int nl = p.vl_n[i];
    int best = 0, bid = -1;
    for (int k = 0; k < nl; ++k) {
      const int v = p.vl_votes[(size_t)i * MS_VCAP + k], id = p.vl_id[(size_t)i * MS_VCAP + k];
      if (v > best || (v == best && v > 0 && id < bid)) { best = v; bid = id; }
    }
    if (nl == MS_VCAP && S > 0) {   // clusters that did not fit the inline list: sum their spill entries
      for (int e = 0; e < S; ++e) {
        if (p.spill[3 * (size_t)e] != i) continue;
        const int id = p.spill[3 * (size_t)e + 1];
        int tot = 0;
        bool first = true;
        for (int f = 0; f < S; ++f)
          if (p.spill[3 * (size_t)f] == i && p.spill[3 * (size_t)f + 1] == id) {
            if (f < e) first = false;
            tot += p.spill[3 * (size_t)f + 2];
          }
        if (first && (tot > best || (tot == best && tot > 0 && id < bid))) { best = tot; bid = id; }
      }
    }
    p.assign[i] = bid;
  }
  if (flags) atomicOr(p.ctl + 1, flags);
  if (tid == 0) {
    p.ctl[0] = C;
    p.ctl[4] = on_the_spot;
    *reinterpret_cast<unsigned long long*>(p.ctl + 6) = traj;
    *reinterpret_cast<unsigned long long*>(p.ctl + 8) = iters;
    p.ctl[10] = (int32_t)hold;
#ifdef MS_PROFILE
    for (int k = 0; k < 16; ++k) reinterpret_cast<long long*>(p.ctl + 16)[k] = prof[k];
#endif
  }
}

template <int DT>
static mh_status launch_ms_kernels(mh_ctx* ctx, const MsProblem& p, size_t replay_smem, int mode) {
  const int blocks = std::max(1, std::min((p.N + 7) / 8, ctx->sm_count * 8));
  constexpr size_t traj_smem = sizeof(MsWarpScratch) * (MS_TRAJ_THREADS / 32);
  MH_CUDA(ctx, mh_allow_max_smem(ms_trajectories_kernel<DT>));
  ms_trajectories_kernel<DT><<<blocks, MS_TRAJ_THREADS, traj_smem, ctx->stream>>>(p);
  MH_LAUNCHED(ctx, "ms_trajectories_kernel");
  if (p.heavy_slot) {
    ms_heavy_kernel<DT><<<ctx->sm_count * 4, MS_REPLAY_THREADS, 0, ctx->stream>>>(p);
    MH_LAUNCHED(ctx, "ms_heavy_kernel");
  }
  auto kern = mode == 0 ? ms_replay_kernel<DT, 0> : mode == 1 ? ms_replay_kernel<DT, 1> : ms_replay_kernel<DT, 2>;
  MH_CUDA(ctx, mh_allow_max_smem(kern));
  kern<<<1, MS_REPLAY_THREADS, replay_smem, ctx->stream>>>(p);
  MH_LAUNCHED(ctx, "ms_replay_kernel");
  return MH_OK;
}

uint64_t gram_scratch_bytes(int N, int D);   // k3_gram.cu: the batched L2 member on the tensor cores (metric 2)
mh_status launch_meanshift_gram(mh_ctx* ctx, const double* d_xs, const int32_t* d_perm, const int32_t* d_nf, int N, int Npad, int D,
                                double bw, void* scratch, double* d_centres, int max_c, int32_t* d_assign, int* C_out, int64_t* stats);

mh_status launch_meanshift(mh_ctx* ctx, const double* d_feat, int N, int D, double bw, int metric, uint32_t* rng_state,
                           double* d_centres, int max_c, int32_t* d_assign, int* C_out, int64_t* stats) {
  if (D > MS_MAXD) return fail(ctx, MH_EINVAL, "mh_meanshift: D > 16");
  const int Npad = (N + 31) & ~31;
  const int mask_words = ((N + 1023) >> 10) << 5;   // whole 1024-point blocks; the padding counts as visited
  int spill_cap = std::max(N, 4096);
  int node_cap = 2 * std::max(max_c, 1) + 64;
  int hash_bits = 10;
  while ((1 << hash_bits) < 2 * max_c && hash_bits < 26) ++hash_bits;
  const bool small = N <= MS_SORT_SMALL_MAX;
  size_t cub_bytes = 0;
  if (!small)
    MH_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                                 (const int32_t*)nullptr, (int32_t*)nullptr, N, 0, 64, ctx->stream));
  struct { int32_t C, flags, nf, spill_used, on_the_spot, pad0; unsigned long long traj, iters; uint32_t rng; int32_t pad1; } out;
  static_assert(sizeof(out) == 48, "layout of MsProblem::ctl");

  for (int attempt = 0;; ++attempt) {
    uint64_t off = 0;
    auto take = [&](uint64_t bytes) { uint64_t o = off; off = (off + bytes + 255) & ~uint64_t(255); return o; };
    const uint64_t o_ctl = take(256);
    const uint64_t o_xs = take(sizeof(double) * (uint64_t)D * Npad);
    const uint64_t o_perm = take(sizeof(int32_t) * (uint64_t)N);
    const uint64_t o_mask0 = take(sizeof(uint32_t) * (uint64_t)mask_words);
    const uint64_t o_gmask = take(mask_words > MS_SMEM_MASK_WORDS ? sizeof(uint32_t) * (uint64_t)mask_words : 0);
    const uint64_t o_gcnt1 = take(mask_words > MS_SMEM_MASK_WORDS ? sizeof(int32_t) * (uint64_t)(mask_words >> 5) : 0);
    const int rec_stride = ms_rec_stride(D);
    const uint64_t o_rec = take(sizeof(double) * (uint64_t)N * rec_stride);
    const int heavy_cap = (N >= MS_HEAVY_MIN_N && metric != 2) ? std::min(N, MS_HEAVY_MAX_SLOTS) : 0;
    const uint64_t o_hlist = take(sizeof(int32_t) * (uint64_t)heavy_cap);
    const uint64_t o_hslot = take(heavy_cap ? sizeof(int32_t) * (uint64_t)N : 0);
    const uint64_t o_hhdr = take(sizeof(double) * (uint64_t)heavy_cap * MS_HHDR);
    const uint64_t o_hvotes = take((uint64_t)heavy_cap * MS_HSPAN);
    const uint64_t o_tv = take(sizeof(int32_t) * (uint64_t)Npad);
    const uint64_t o_vn = take(sizeof(int32_t) * (uint64_t)N);
    const uint64_t o_vid = take(sizeof(int32_t) * (uint64_t)N * MS_VCAP);
    const uint64_t o_vv = take(sizeof(int32_t) * (uint64_t)N * MS_VCAP);
    const uint64_t o_spill = take(sizeof(int32_t) * 3 * (uint64_t)spill_cap);
    const uint64_t o_head = take(sizeof(int32_t) * (uint64_t)(1u << hash_bits));
    const uint64_t o_nnext = take(sizeof(int32_t) * (uint64_t)node_cap);
    const uint64_t o_ncen = take(sizeof(int32_t) * (uint64_t)node_cap);
    const uint64_t o_keys = take(small ? 0 : sizeof(unsigned long long) * 2 * (uint64_t)N);
    const uint64_t o_idx = take(small ? 0 : sizeof(int32_t) * 2 * (uint64_t)N);
    const uint64_t o_cub = take(cub_bytes);
    const uint64_t o_gram = take(metric == 2 ? gram_scratch_bytes(N, D) : 0);
    MH_TRY(ensure_scratch(ctx, off));
    char* base = (char*)ctx->scratch;

    MsProblem p;
    p.data = d_feat;
    p.xs = (double*)(base + o_xs);
    p.perm = (int32_t*)(base + o_perm);
    p.mask0 = (uint32_t*)(base + o_mask0);
    p.N = N; p.Npad = Npad; p.D = D; p.mask_words = mask_words;
    p.bandSq = bw * bw;            // MS.h:31
    p.stopThresh = 1e-3 * bw;      // MS.h:48
    p.halfBw = bw / 2;             // MS.h:104
    p.cellInv = 1.0 / (p.halfBw * (1.0 + 1e-9));
    p.metric = metric;
    p.rec = (double*)(base + o_rec);
    p.rec_stride = rec_stride;
    p.heavy_cap = heavy_cap;
    p.heavy_list = heavy_cap ? (int32_t*)(base + o_hlist) : nullptr;
    p.heavy_slot = heavy_cap ? (int32_t*)(base + o_hslot) : nullptr;
    p.heavy_hdr = heavy_cap ? (double*)(base + o_hhdr) : nullptr;
    p.heavy_votes = heavy_cap ? (unsigned char*)(base + o_hvotes) : nullptr;
    p.tvotes = (int32_t*)(base + o_tv);
    p.vl_n = (int32_t*)(base + o_vn);
    p.vl_id = (int32_t*)(base + o_vid);
    p.vl_votes = (int32_t*)(base + o_vv);
    p.spill = (int32_t*)(base + o_spill);
    p.spill_cap = spill_cap;
    p.hash_head = (int32_t*)(base + o_head);
    p.hash_bits = hash_bits;
    p.node_next = (int32_t*)(base + o_nnext);
    p.node_centre = (int32_t*)(base + o_ncen);
    p.node_cap = node_cap;
    p.gmask = (uint32_t*)(base + o_gmask);
    p.gcnt1 = (int32_t*)(base + o_gcnt1);
    p.rng = rng_state ? *rng_state : 1u;
    p.centres = d_centres;
    p.max_c = max_c;
    p.assign = d_assign;
    p.ctl = (int32_t*)(base + o_ctl);

    MH_CUDA(ctx, cudaMemsetAsync(base + o_ctl, 0, 256, ctx->stream));
    MH_CUDA(ctx, cudaMemsetAsync(base + o_head, 0xff, sizeof(int32_t) * ((size_t)1 << hash_bits), ctx->stream));
    if (small) {
      int P = 32;
      while (P < N) P <<= 1;
      const size_t smem = (size_t)P * 12;
      MH_CUDA(ctx, mh_allow_max_smem(ms_prep_small_kernel));
      ms_prep_small_kernel<<<1, std::min(1024, std::max(32, P / 2)), smem, ctx->stream>>>(p, P);
      MH_LAUNCHED(ctx, "ms_prep_small_kernel");
    } else {
      unsigned long long* keys = (unsigned long long*)(base + o_keys);
      int32_t* idx = (int32_t*)(base + o_idx);
      const int kb = (Npad + 255) / 256;
      ms_keys_kernel<<<kb, 256, 0, ctx->stream>>>(p, keys, idx);
      MH_LAUNCHED(ctx, "ms_keys_kernel");
      const int first_tail = kb * 8;   // words written by ms_keys_kernel: one per launched warp
      if (first_tail < mask_words) {
        ms_mask_tail_kernel<<<(mask_words - first_tail + 255) / 256, 256, 0, ctx->stream>>>(p, first_tail);
        MH_LAUNCHED(ctx, "ms_mask_tail_kernel");
      }
      MH_CUDA(ctx, cub::DeviceRadixSort::SortPairs(base + o_cub, cub_bytes, keys, keys + N, idx, idx + N, N, 0, 64, ctx->stream));
      ctx->launches += 1;
      ms_gather_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(p, idx + N);
      MH_LAUNCHED(ctx, "ms_gather_kernel");
    }
    if (metric == 2)   // all seeds at once, L2 window, Gram on the tensor cores: k3_gram.cu
      return launch_meanshift_gram(ctx, p.xs, p.perm, p.ctl + 2, N, Npad, D, bw, base + o_gram, d_centres, max_c, d_assign, C_out, stats);
    size_t replay_smem = sizeof(MsShared) + sizeof(uint32_t) * (size_t)std::min(mask_words, MS_SMEM_MASK_WORDS);
    const size_t centre_bytes = sizeof(double) * (size_t)D * std::max(max_c, 1);
    const bool centres_smem = mask_words <= MS_SMEM_MASK_WORDS && replay_smem + centre_bytes <= 200 * 1024;
    if (centres_smem) replay_smem += centre_bytes;
    const int mode = centres_smem ? 0 : mask_words <= MS_SMEM_MASK_WORDS ? 1 : 2;
    if (D <= 6) MH_TRY(launch_ms_kernels<6>(ctx, p, replay_smem, mode));
    else if (D <= 10) MH_TRY(launch_ms_kernels<10>(ctx, p, replay_smem, mode));
    else MH_TRY(launch_ms_kernels<16>(ctx, p, replay_smem, mode));
    MH_CUDA(ctx, cudaMemcpyAsync(&out, base + o_ctl, sizeof(out), cudaMemcpyDeviceToHost, ctx->stream));
    MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
#ifdef MS_PROFILE
    {
      long long prof[16];
      cudaMemcpy(prof, base + o_ctl + 64, sizeof(prof), cudaMemcpyDeviceToHost);
      std::fprintf(stderr, "[ms profile+] rank %.0f scan1 %.0f word %.0f | votes+perm %.0f mark %.0f fetch %.0f\n", prof[8] / (double)out.traj,
                   prof[9] / (double)out.traj, prof[10] / (double)out.traj, prof[11] / (double)out.traj, prof[12] / (double)out.traj,
                   prof[13] / (double)out.traj);
      std::fprintf(stderr, "[ms profile] N=%d steps=%llu on_the_spot=%d cycles/step: seed %.0f rec %.0f traj %.0f first %.0f merge %.0f centre %.0f votes %.0f bar %.0f\n",
                   N, out.traj, out.on_the_spot, prof[0] / (double)out.traj, prof[1] / (double)out.traj, prof[2] / (double)out.traj,
                   prof[3] / (double)out.traj, prof[4] / (double)out.traj, prof[5] / (double)out.traj, prof[6] / (double)out.traj,
                   prof[7] / (double)out.traj);
    }
#endif
    if (!(out.flags & (MS_FLAG_SPILL | MS_FLAG_NODES))) break;
    if (attempt >= 5) return fail(ctx, MH_ENOMEM, "mh_meanshift: working storage still too small after 6 attempts");
    // rare: very many clusters per point or very many centre moves — grow what ran out and repeat (same seed, same result)
    if (out.flags & MS_FLAG_SPILL) spill_cap = std::max(4 * spill_cap, out.spill_used + 1024);
    if (out.flags & MS_FLAG_NODES) node_cap *= 4;
  }
  *C_out = out.C;
  if (stats) { stats[0] = (int64_t)out.traj; stats[1] = (int64_t)out.iters; }
  if (rng_state) *rng_state = out.rng;
  if (out.flags & MS_FLAG_CENTRES) return fail(ctx, MH_ENOMEM, "mh_meanshift: more centres than max_c");
  return MH_OK;
}

}  // namespace mh
