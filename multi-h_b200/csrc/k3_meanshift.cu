// ============================================================================
// K3 — homography-space mean-shift (sm_100a).
//
// Replaces MeanShiftClustering<double>::Cluster
// (MultiH/MultiH/moduls/mode_seeking/MeanShiftClustering.h:22-157) as called from
// EstablishStablePointSets (MultiH.cpp:654, D = 10) and MergingStep (MultiH.cpp:397, D = 6).
//
// The reference algorithm is inherently sequential: trajectories start one after
// another from a random not-yet-visited point and every window iteration marks
// the points it covers as visited, which changes the pool the next seed is drawn
// from.  To return the reference's clustering (not merely a statistically similar
// one) the whole algorithm runs as ONE persistent cooperative kernel: the data
// stay resident (L2), each window iteration is a chip-wide pass — L1 window test
// (sum_j |mean_j - x_ij| < bw^2, MS.h:76-85), flat-kernel mean — with a
// deterministic two-level reduction (warp shuffle -> CTA -> fixed-order sum of
// per-CTA partials) and one grid barrier; seeds come from the restated MSVC
// rand() so the visiting order equals the reference's.  FP64 throughout.
// A window iteration costs a grid barrier (~2-3 us) instead of the reference's
// O(N*D) scalar loop + repmat allocation (MS.h:66-94).
// ============================================================================
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mh {

constexpr int MS_THREADS = 256;
constexpr int MS_MAXD = 16;
constexpr int MS_VCAP = 24;  // per-point capacity of the sparse (cluster, votes) list

struct MsState {
  // inputs
  const double* dataT;  // [D][Npad] SoA copy
  int N, Npad, D;
  double bandSq, stopThresh, halfBw;
  int metric;
  uint32_t rng;
  // work buffers
  uint8_t* visited;      // N (1 = visited or non-finite row)
  int32_t* tvotes;       // N, votes of the running trajectory
  int32_t* vl_id;        // N x VCAP
  int32_t* vl_votes;     // N x VCAP
  int32_t* vl_n;         // N
  double* partial;       // [2][blocks][MS_MAXD + 1]
  int32_t* block_unvisited;  // [blocks]
  double* seed_mean;     // MS_MAXD
  // outputs
  double* centres;       // [max_c][D]
  int max_c;
  int32_t* assign;       // N
  int32_t* out_C;        // [0] = C, [1] = overflow flag
  unsigned long long* out_stats;  // trajectories, window iterations
  uint32_t* out_rng;
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void ms_transpose_kernel(const double* __restrict__ data, int N, int Npad, int D, double* __restrict__ dataT,
                                    uint8_t* __restrict__ visited, int32_t* __restrict__ vl_n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  bool finite = true;
  for (int j = 0; j < D; ++j) {
    const double v = data[(size_t)i * D + j];
    finite = finite && isfinite(v);
    dataT[(size_t)j * Npad + i] = v;
  }
  visited[i] = finite ? 0 : 1;  // non-finite rows can neither seed nor join a window (the reference would spin on them)
  vl_n[i] = 0;
}

// number of not-yet-visited points of [lo, hi) -> block_unvisited[b]
template <int MS_THREADS>
__device__ __forceinline__ void publish_unvisited(const MsState& st, int lo, int hi, int b, int* s_int) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int c = 0;
  for (int i = lo + tid; i < hi; i += MS_THREADS) c += st.visited[i] ? 0 : 1;
  c = __reduce_add_sync(0xffffffffu, c);
  __syncthreads();
  if (lane == 0) s_int[warp] = c;
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int w = 0; w < MS_THREADS / 32; ++w) t += s_int[w];
    st.block_unvisited[b] = t;
  }
}

// SINGLE: the whole problem fits one CTA's stride loop (N <= MS_SINGLE_MAX: the bundled pairs' 1-2k correspondences, and
// every merging step's K hypotheses, cfg5's 5k-correspondence pairs) — a plain launch whose "grid barrier" is __syncthreads(): ~1 us per window iteration
// instead of ~5 us, which is what the latency-bound small cases pay for.
constexpr int MS_SINGLE_MAX = 8192;
constexpr int MS_SINGLE_THREADS = 512;
template <bool SINGLE, int MS_THREADS>
__global__ void __launch_bounds__(MS_THREADS) meanshift_kernel(MsState st) {
  auto grid_sync = [] {
    if constexpr (SINGLE) __syncthreads();
    else cg::this_grid().sync();
  };
  __shared__ double s_mean[MS_MAXD];
  __shared__ double s_red[MS_THREADS / 32][MS_MAXD + 1];
  __shared__ double s_new[MS_MAXD + 1];
  __shared__ int s_int[MS_THREADS / 32 + 2];

  const int D = st.D, N = st.N, Npad = st.Npad;
  const int nb = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // contiguous ownership: CTA b owns [lo, hi)
  const int chunk = (N + nb - 1) / nb;
  const int lo = min(N, b * chunk), hi = min(N, lo + chunk);

  uint32_t hold = st.rng;
  int C = 0;
  unsigned long long traj = 0, iters = 0;
  int overflow = 0;

  publish_unvisited<MS_THREADS>(st, lo, hi, b, s_int);
  grid_sync();
  for (;;) {
    // ---- every CTA derives the same seed rank (MS.h:54-56) -------------------------
    int remaining = 0;
    for (int k = 0; k < nb; ++k) remaining += st.block_unvisited[k];
    if (remaining == 0) break;
    hold = hold * 214013u + 2531011u;  // MSVC rand()
    const double rnd = (double)((hold >> 16) & 0x7fff) / 32767.0;
    int rank = (int)round(rnd * (double)(remaining - 1));
    int owner = 0;
    for (; owner < nb; ++owner) {
      const int c = st.block_unvisited[owner];
      if (rank < c) break;
      rank -= c;
    }
    if (b == owner) {
      // find the rank-th unvisited point of [lo, hi) in ascending index order
      __syncthreads();
      if (tid < 32) {
        int seen = 0, found = -1;
        for (int base = lo; base < hi && found < 0; base += 32) {
          const int i = base + lane;
          const bool u = i < hi && !st.visited[i];
          const unsigned m = __ballot_sync(0xffffffffu, u);
          const int c = __popc(m);
          if (rank < seen + c) {
            // the (rank - seen)-th set bit of m
            int want = rank - seen;
            unsigned mm = m;
            while (want--) mm &= mm - 1;
            found = base + __ffs(mm) - 1;
          }
          seen += c;
        }
        if (lane == 0) s_int[MS_THREADS / 32] = found;
      }
      __syncthreads();
      const int seed = s_int[MS_THREADS / 32];
      if (tid < D) st.seed_mean[tid] = st.dataT[(size_t)tid * Npad + seed];
    }
    for (int i = lo + tid; i < hi; i += MS_THREADS) st.tvotes[i] = 0;
    grid_sync();
    if (tid < D) s_mean[tid] = st.seed_mean[tid];
    __syncthreads();
    ++traj;

    // ---- window iterations (MS.h:62-98) -------------------------------------------
    // The reference loops until the mean stops moving; with its L1 window the mean can cycle and the reference never
    // returns (oracle/multih_oracle.cpp orc_meanshift).  A trajectory ends after MH_MS_MAX_WINDOW_ITERS iterations at the
    // latest, keeping its current mean — identical in oracle and product.
    for (int window_iters = 1;; ++window_iters) {
      const int par = (int)(iters & 1ull);
      ++iters;
      double acc[MS_MAXD + 1];
#pragma unroll
      for (int j = 0; j <= MS_MAXD; ++j) acc[j] = 0.0;
      for (int i = lo + tid; i < hi; i += MS_THREADS) {
        double x[MS_MAXD];
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < MS_MAXD; ++j)
          if (j < D) {
            x[j] = st.dataT[(size_t)j * Npad + i];
            const double d = s_mean[j] - x[j];
            s += st.metric == 0 ? fabs(d) : d * d;  // sqrt(d*d) summed: L1 (MS.h:80-82)
          }
        if (s < st.bandSq) {  // MS.h:85 (NaN rows never pass)
          st.tvotes[i] += 1;
          st.visited[i] = 1;
#pragma unroll
          for (int j = 0; j < MS_MAXD; ++j)
            if (j < D) acc[j] += x[j];
          acc[MS_MAXD] += 1.0;
        }
      }
      // CTA reduction, fixed order
#pragma unroll
      for (int j = 0; j <= MS_MAXD; ++j)
        if (j < D || j == MS_MAXD) {
          const double v = warp_sum_d(acc[j]);
          if (lane == 0) s_red[warp][j] = v;
        }
      __syncthreads();
      if (tid <= MS_MAXD && (tid < D || tid == MS_MAXD)) {
        double v = 0.0;
        for (int w = 0; w < MS_THREADS / 32; ++w) v += s_red[w][tid];
        st.partial[((size_t)par * nb + b) * (MS_MAXD + 1) + tid] = v;
      }
      grid_sync();
      // every CTA sums the per-CTA partials in the same order -> identical new mean everywhere
      if (warp == 0) {
        for (int j = 0; j <= MS_MAXD; ++j) {
          if (!(j < D || j == MS_MAXD)) continue;
          double v = 0.0;
          for (int k = lane; k < nb; k += 32) v += st.partial[((size_t)par * nb + k) * (MS_MAXD + 1) + j];
          v = warp_sum_d(v);
          if (lane == 0) s_new[j] = v;
        }
      }
      __syncthreads();
      const double cnt = s_new[MS_MAXD];
      double n2 = 0.0;
      double nm[MS_MAXD];
#pragma unroll
      for (int j = 0; j < MS_MAXD; ++j)
        if (j < D) {
          nm[j] = s_new[j] / cnt;  // MS.h:96
          const double d = nm[j] - s_mean[j];
          n2 += d * d;
        }
      __syncthreads();
      if (tid < D) s_mean[tid] = nm[tid];
      __syncthreads();
      if (sqrt(n2) < st.stopThresh || !(cnt > 0.0) || window_iters >= MH_MS_MAX_WINDOW_ITERS) break;  // MS.h:98 (cnt == 0 cannot happen for finite seeds)
    }

    // ---- merge into the first centre closer than bw/2, else append (MS.h:100-120) ---
    int mergeWith = 0x7fffffff;
    for (int c = tid; c < C; c += MS_THREADS) {
      double d2 = 0.0;
      for (int j = 0; j < D; ++j) {
        const double d = s_mean[j] - st.centres[(size_t)c * D + j];
        d2 += d * d;
      }
      if (sqrt(d2) < st.halfBw) mergeWith = min(mergeWith, c);
    }
    mergeWith = __reduce_min_sync(0xffffffffu, mergeWith);
    if (lane == 0) s_int[warp] = mergeWith;
    __syncthreads();
    mergeWith = 0x7fffffff;
    for (int w = 0; w < MS_THREADS / 32; ++w) mergeWith = min(mergeWith, s_int[w]);
    __syncthreads();
    const bool merged = mergeWith != 0x7fffffff;
    const int cid = merged ? mergeWith : C;
    const bool room = merged || C < st.max_c;
    // votes of this trajectory go to cluster cid (MS.h:114,119)
    if (room)
      for (int i = lo + tid; i < hi; i += MS_THREADS) {
        const int v = st.tvotes[i];
        if (v == 0) continue;
        const int n = st.vl_n[i];
        int32_t* ids = st.vl_id + (size_t)i * MS_VCAP;
        int32_t* vs = st.vl_votes + (size_t)i * MS_VCAP;
        int k = 0;
        for (; k < n; ++k)
          if (ids[k] == cid) break;
        if (k < n) vs[k] += v;
        else if (n < MS_VCAP) { ids[n] = cid; vs[n] = v; st.vl_n[i] = n + 1; }
        else overflow = 1;
      }
    else overflow = 1;
    // One barrier closes the trajectory: it publishes the new unvisited counts for the next seed draw AND orders the
    // centre update below after every CTA's read of `centres` above.
    publish_unvisited<MS_THREADS>(st, lo, hi, b, s_int);
    grid_sync();
    if (b == 0 && tid < D && room) {
      double* c = st.centres + (size_t)cid * D;
      c[tid] = merged ? 0.5 * (c[tid] + s_mean[tid]) : s_mean[tid];  // MS.h:113 / :118
    }
    if (!merged && room) ++C;
    // centres are next read after at least one more grid barrier (the seed barrier of the next trajectory)
  }

  // ---- final assignment: most votes, first (lowest id) wins ties (MS.h:133-146) -------
  for (int i = lo + tid; i < hi; i += MS_THREADS) {
    const int n = st.vl_n[i];
    int best = 0, bid = -1;
    for (int k = 0; k < n; ++k) {
      const int v = st.vl_votes[(size_t)i * MS_VCAP + k], id = st.vl_id[(size_t)i * MS_VCAP + k];
      if (v > best || (v == best && v > 0 && id < bid)) { best = v; bid = id; }
    }
    st.assign[i] = bid;
  }
  if (overflow) atomicExch(st.out_C + 1, 1);
  if (b == 0 && tid == 0) {
    st.out_C[0] = C;
    st.out_stats[0] = traj;
    st.out_stats[1] = iters;
    st.out_rng[0] = hold;
  }
}

mh_status launch_meanshift(mh_ctx* ctx, const double* d_feat, int N, int D, double bw, int metric, uint32_t* rng_state,
                           double* d_centres, int max_c, int32_t* d_assign, int* C_out, int64_t* stats) {
  if (D > MS_MAXD) return fail(ctx, MH_EINVAL, "mh_meanshift: D > 16");
  int dev_coop = 0;
  MH_CUDA(ctx, cudaDeviceGetAttribute(&dev_coop, cudaDevAttrCooperativeLaunch, ctx->device));
  if (!dev_coop) return fail(ctx, MH_ECUDA, "device lacks cooperative launch");
  int per_sm = 0;
  MH_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, meanshift_kernel<false, MS_THREADS>, MS_THREADS, 0));
  if (per_sm < 1) return fail(ctx, MH_ECUDA, "meanshift kernel does not fit an SM");
  const bool single = N <= MS_SINGLE_MAX;
  int blocks = single ? 1 : std::min(ctx->sm_count, (N + MS_THREADS - 1) / MS_THREADS);
  blocks = std::max(1, blocks);

  const int Npad = (N + 31) & ~31;
  // scratch layout
  uint64_t off = 0;
  auto take = [&](uint64_t bytes) { uint64_t o = off; off = (off + bytes + 255) & ~uint64_t(255); return o; };
  const uint64_t o_dataT = take(sizeof(double) * (uint64_t)D * Npad);
  const uint64_t o_vis = take((uint64_t)N);
  const uint64_t o_tv = take(sizeof(int32_t) * (uint64_t)N);
  const uint64_t o_vid = take(sizeof(int32_t) * (uint64_t)N * MS_VCAP);
  const uint64_t o_vv = take(sizeof(int32_t) * (uint64_t)N * MS_VCAP);
  const uint64_t o_vn = take(sizeof(int32_t) * (uint64_t)N);
  const uint64_t o_part = take(sizeof(double) * 2 * (uint64_t)blocks * (MS_MAXD + 1));
  const uint64_t o_bu = take(sizeof(int32_t) * (uint64_t)blocks);
  const uint64_t o_seed = take(sizeof(double) * MS_MAXD);
  const uint64_t o_out = take(64);
  MH_TRY(ensure_scratch(ctx, off));
  char* base = (char*)ctx->scratch;

  MsState st;
  st.dataT = (const double*)(base + o_dataT);
  st.N = N; st.Npad = Npad; st.D = D;
  st.bandSq = bw * bw;            // MS.h:31
  st.stopThresh = 1e-3 * bw;      // MS.h:48
  st.halfBw = bw / 2;             // MS.h:104
  st.metric = metric;
  st.rng = rng_state ? *rng_state : 1u;
  st.visited = (uint8_t*)(base + o_vis);
  st.tvotes = (int32_t*)(base + o_tv);
  st.vl_id = (int32_t*)(base + o_vid);
  st.vl_votes = (int32_t*)(base + o_vv);
  st.vl_n = (int32_t*)(base + o_vn);
  st.partial = (double*)(base + o_part);
  st.block_unvisited = (int32_t*)(base + o_bu);
  st.seed_mean = (double*)(base + o_seed);
  st.centres = d_centres;
  st.max_c = max_c;
  st.assign = d_assign;
  st.out_C = (int32_t*)(base + o_out);
  st.out_stats = (unsigned long long*)(base + o_out + 16);
  st.out_rng = (uint32_t*)(base + o_out + 32);

  MH_CUDA(ctx, cudaMemsetAsync(base + o_out, 0, 64, ctx->stream));
  ms_transpose_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(d_feat, N, Npad, D, (double*)(base + o_dataT), st.visited,
                                                                st.vl_n);
  MH_LAUNCHED(ctx, "ms_transpose_kernel");
  if (single) {
    if (N <= 1024) meanshift_kernel<true, MS_THREADS><<<1, MS_THREADS, 0, ctx->stream>>>(st);
    else meanshift_kernel<true, MS_SINGLE_THREADS><<<1, MS_SINGLE_THREADS, 0, ctx->stream>>>(st);
    MH_LAUNCHED(ctx, "meanshift_kernel<single>");
  } else {
    void* args[] = {&st};
    MH_CUDA(ctx, cudaLaunchCooperativeKernel((void*)meanshift_kernel<false, MS_THREADS>, dim3(blocks), dim3(MS_THREADS), args, 0, ctx->stream));
    ++ctx->launches;
  }
  struct { int32_t C, overflow, pad0, pad1; unsigned long long traj, iters; uint32_t rng; } out;
  MH_CUDA(ctx, cudaMemcpyAsync(&out, base + o_out, 36, cudaMemcpyDeviceToHost, ctx->stream));
  MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *C_out = out.C;
  if (stats) { stats[0] = (int64_t)out.traj; stats[1] = (int64_t)out.iters; }
  if (rng_state) *rng_state = out.rng;
  if (out.overflow) return fail(ctx, MH_ENOMEM, "mh_meanshift: centre or vote-list capacity exceeded");
  return MH_OK;
}

}  // namespace mh
