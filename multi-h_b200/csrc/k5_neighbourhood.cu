// ============================================================================
// K5 — neighbourhood graph on the GPU.
//
// Replaces the FlannBasedMatcher::radiusMatch call of ClusterMergingAndLabeling
// (MultiH/MultiH/MultiH.cpp:231-253) with the exactly defined set the host version
// (host_graphcut.cpp radius_neighbourhood) and the oracle (orc_radius_neighbours) return:
// for every correspondence, on float (x1, y1, x2, y2), the `max_neighbours` nearest other
// correspondences (ties by index) among those with d^2 <= radius^2, listed in ascending
// index order.
//
// Uniform grid over (x1, y1) (counting sort: histogram -> exclusive scan -> scatter), then one
// thread per query walks the rings of cells around its own cell, nearest ring first, keeping
// the best `max_neighbours` candidates ordered by (d^2, index) in a small sorted array, and
// stops as soon as no farther ring can beat the worst kept candidate — the host version's
// search, one query per thread instead of one after another.  d^2 is accumulated with
// explicit __fmul_rn / __fadd_rn in the host's left-to-right order (no FMA contraction), so
// the comparisons — hence the neighbour sets — are bit-identical to the host's.
// Output is a fixed-stride table [N][max_neighbours] + counts; the caller compacts it to CSR.
// ============================================================================
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace mh {

constexpr int NB_MAXK = 64;   // largest supported max_neighbours (the reference's effective value is 31)

struct NbGrid {
  float minx, miny, cw, ch, cmin, r2;
  int gx, gy, maxn;
};

__device__ __forceinline__ int nb_cx(const NbGrid& g, float x) { return min(g.gx - 1, max(0, (int)((x - g.minx) / g.cw))); }
__device__ __forceinline__ int nb_cy(const NbGrid& g, float y) { return min(g.gy - 1, max(0, (int)((y - g.miny) / g.ch))); }

__global__ void nb_count_kernel(const float4* __restrict__ p, int N, NbGrid g, int* __restrict__ cell_of, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float4 q = p[i];
  const int c = nb_cy(g, q.y) * g.gx + nb_cx(g, q.x);
  cell_of[i] = c;
  atomicAdd(counts + c, 1);
}

__global__ void nb_scatter_kernel(const int* __restrict__ cell_of, int N, const int* __restrict__ cstart, int* __restrict__ cursor,
                                  int* __restrict__ order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int c = cell_of[i];
  order[cstart[c] + atomicAdd(cursor + c, 1)] = i;   // order inside a cell is arbitrary: the result does not depend on it
}

__global__ void __launch_bounds__(128)
nb_search_kernel(const float4* __restrict__ p, int N, NbGrid g, const int* __restrict__ cstart, const int* __restrict__ order,
                 int32_t* __restrict__ out /*[N][maxn]*/, int32_t* __restrict__ out_n /*[N]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float4 a = p[i];
  const int ix = nb_cx(g, a.x), iy = nb_cy(g, a.y);
  const int K = g.maxn;
  float kd[NB_MAXK];   // kept candidates, ascending by (d2, index)
  int ki[NB_MAXK];
  int n = 0;
  const int max_ring = max(g.gx, g.gy);
  for (int ring = 0; ring <= max_ring; ++ring) {
    if (ring > 0) {
      const float reach = (float)(ring - 1) * g.cmin;   // every site of this ring is at least this far away in (x1, y1)
      const float reach2 = __fmul_rn(reach, reach);
      if (reach2 > g.r2) break;
      if (n == K && reach2 > kd[K - 1]) break;
    }
    if (ring > max(max(ix, g.gx - 1 - ix), max(iy, g.gy - 1 - iy))) break;   // past the grid
    const int y0 = iy - ring, y1 = iy + ring, x0 = ix - ring, x1 = ix + ring;
    // cells of the ring: rows y0 and y1 in full, columns x0 and x1 between them
    const int side = 2 * ring + 1;
    const int ncells = ring == 0 ? 1 : 4 * side - 4;
    for (int t = 0; t < ncells; ++t) {
      int xx, yy;
      if (ring == 0) { xx = ix; yy = iy; }
      else if (t < side) { xx = x0 + t; yy = y0; }
      else if (t < 2 * side) { xx = x0 + (t - side); yy = y1; }
      else { const int u = t - 2 * side; yy = y0 + 1 + (u >> 1); xx = (u & 1) ? x1 : x0; }
      if (xx < 0 || yy < 0 || xx >= g.gx || yy >= g.gy) continue;
      const int c = yy * g.gx + xx;
      const int kb = cstart[c], ke = cstart[c + 1];
      for (int k = kb; k < ke; ++k) {
        const int j = order[k];
        if (j == i) continue;
        const float4 b = p[j];
        const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
        // ((d0^2 + d1^2) + d2^2) + d3^2, each product and sum rounded on its own, as the host evaluates it
        const float d = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), __fmul_rn(d3, d3));
        if (d > g.r2) continue;
        if (n == K && !(d < kd[K - 1] || (d == kd[K - 1] && j < ki[K - 1]))) continue;   // not better than the worst kept
        int pos = (n < K) ? n : K - 1;   // slot that is free / dropped
        while (pos > 0 && (d < kd[pos - 1] || (d == kd[pos - 1] && j < ki[pos - 1]))) {
          kd[pos] = kd[pos - 1]; ki[pos] = ki[pos - 1];
          --pos;
        }
        kd[pos] = d; ki[pos] = j;
        if (n < K) ++n;
      }
    }
  }
  // ascending index order (MultiH.cpp:532-540 walks the match lists; the host version sorts too)
  for (int s = 1; s < n; ++s) {
    const int v = ki[s];
    int t = s;
    while (t > 0 && ki[t - 1] > v) { ki[t] = ki[t - 1]; --t; }
    ki[t] = v;
  }
  for (int s = 0; s < n; ++s) out[(size_t)i * K + s] = ki[s];
  out_n[i] = n;
}

// host points (N x 4 FP64, pixels) -> CSR neighbourhood; same contract as radius_neighbourhood() for max_neighbours in
// [1, NB_MAXK].  `adj` may be NULL (count only).
mh_status neighbourhood_device(mh_ctx* ctx, const double* pts, int N, double radius, int max_neighbours, int64_t* offsets,
                               int32_t* adj, int64_t* total_out) {
  if (max_neighbours < 1 || max_neighbours > NB_MAXK) return fail(ctx, MH_EINVAL, "neighbourhood_device: max_neighbours out of range");
  if (N <= 0) { if (offsets) offsets[0] = 0; if (total_out) *total_out = 0; return MH_OK; }
  // float copy + bounds, exactly as the host version derives its grid
  const float r = (float)radius;
  NbGrid g;
  g.r2 = r * r;
  g.maxn = max_neighbours;
  std::vector<float> p(4 * (size_t)N);
  float minx = 1e30f, miny = 1e30f, maxx = -1e30f, maxy = -1e30f;
  for (int i = 0; i < N; ++i) {
    for (int k = 0; k < 4; ++k) p[4 * (size_t)i + k] = (float)pts[4 * (size_t)i + k];
    minx = std::min(minx, p[4 * (size_t)i]); maxx = std::max(maxx, p[4 * (size_t)i]);
    miny = std::min(miny, p[4 * (size_t)i + 1]); maxy = std::max(maxy, p[4 * (size_t)i + 1]);
  }
  float cell = std::max(r, 1e-6f);
  const double area = std::max(1e-12, (double)(maxx - minx) * (double)(maxy - miny));
  cell = std::min(cell, (float)std::sqrt(area * 4.0 / N));   // ~4 sites per cell, never above the radius
  cell = std::max(cell, 1e-6f);
  g.gx = std::max(1, std::min(4096, (int)((maxx - minx) / cell) + 1));
  g.gy = std::max(1, std::min(4096, (int)((maxy - miny) / cell) + 1));
  g.cw = std::max(cell, (maxx - minx) / g.gx + 1e-6f);
  g.ch = std::max(cell, (maxy - miny) / g.gy + 1e-6f);
  g.cmin = std::min(g.cw, g.ch);
  g.minx = minx; g.miny = miny;
  const size_t ncell = (size_t)g.gx * g.gy;

  size_t cub_bytes = 0;
  MH_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const int*)nullptr, (int*)nullptr, (int)(ncell + 1), ctx->stream));
  uint64_t off = 0;
  auto take = [&](uint64_t bytes) { uint64_t o = off; off = (off + bytes + 255) & ~uint64_t(255); return o; };
  const uint64_t o_p = take(sizeof(float4) * (uint64_t)N);
  const uint64_t o_cell = take(sizeof(int) * (uint64_t)N);
  const uint64_t o_order = take(sizeof(int) * (uint64_t)N);
  const uint64_t o_counts = take(sizeof(int) * (ncell + 1));
  const uint64_t o_cstart = take(sizeof(int) * (ncell + 1));
  const uint64_t o_cursor = take(sizeof(int) * (ncell + 1));
  const uint64_t o_cub = take(cub_bytes);
  const uint64_t o_out = take(sizeof(int32_t) * (uint64_t)N * max_neighbours);
  const uint64_t o_n = take(sizeof(int32_t) * (uint64_t)N);
  MH_TRY(ensure_staging(ctx, off));
  char* base = (char*)ctx->staging;
  float4* d_p = (float4*)(base + o_p);
  int *d_cell = (int*)(base + o_cell), *d_order = (int*)(base + o_order), *d_counts = (int*)(base + o_counts),
      *d_cstart = (int*)(base + o_cstart), *d_cursor = (int*)(base + o_cursor);
  int32_t *d_out = (int32_t*)(base + o_out), *d_n = (int32_t*)(base + o_n);

  MH_CUDA(ctx, cudaMemcpyAsync(d_p, p.data(), sizeof(float4) * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
  MH_CUDA(ctx, cudaMemsetAsync(d_counts, 0, sizeof(int) * (ncell + 1), ctx->stream));
  MH_CUDA(ctx, cudaMemsetAsync(d_cursor, 0, sizeof(int) * (ncell + 1), ctx->stream));
  const unsigned blocks = (unsigned)((N + 255) / 256);
  nb_count_kernel<<<blocks, 256, 0, ctx->stream>>>(d_p, N, g, d_cell, d_counts);
  MH_LAUNCHED(ctx, "nb_count_kernel");
  MH_CUDA(ctx, cub::DeviceScan::ExclusiveSum(base + o_cub, cub_bytes, d_counts, d_cstart, (int)(ncell + 1), ctx->stream));
  nb_scatter_kernel<<<blocks, 256, 0, ctx->stream>>>(d_cell, N, d_cstart, d_cursor, d_order);
  MH_LAUNCHED(ctx, "nb_scatter_kernel");
  nb_search_kernel<<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>(d_p, N, g, d_cstart, d_order, d_out, d_n);
  MH_LAUNCHED(ctx, "nb_search_kernel");

  std::vector<int32_t> cnt(N);
  MH_CUDA(ctx, cudaMemcpyAsync(cnt.data(), d_n, sizeof(int32_t) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<int32_t> table;
  if (adj) {
    table.resize((size_t)N * max_neighbours);
    MH_CUDA(ctx, cudaMemcpyAsync(table.data(), d_out, sizeof(int32_t) * table.size(), cudaMemcpyDeviceToHost, ctx->stream));
  }
  MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int64_t total = 0;
  for (int i = 0; i < N; ++i) {
    if (offsets) offsets[i] = total;
    if (adj) std::memcpy(adj + total, table.data() + (size_t)i * max_neighbours, sizeof(int32_t) * (size_t)cnt[i]);
    total += cnt[i];
  }
  if (offsets) offsets[N] = total;
  if (total_out) *total_out = total;
  return MH_OK;
}

}  // namespace mh
