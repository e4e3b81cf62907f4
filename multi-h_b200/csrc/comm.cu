// ============================================================================
// comm.cu — the multi-GPU layer of the library: one context = one device + one stream + one NCCL
// communicator.  Correspondences are sharded over the ranks, hypotheses are
// replicated; a sharded hot pass is
//     ncclBroadcast  hypotheses K x 12 FP32 from rank 0         (overlaps K1 on a second stream)
//     K2 fused cost / argmin / inlier counts on the local shard; K1 per-correspondence HAF behind it on a second stream (fills K2's tail)
//     K4 refit statistics of the local shard, inlier counts packed into their pad column
//     ncclAllReduce(sum) of the K x 12 FP64 statistics           (overlaps the NEXT pass: double-buffered)
//     K4 batched eigen-solves, redundantly on every rank.
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already has — PyTorch ships
// one — or of MH_NCCL_LIB), so the library has no link-time dependency and single-GPU users never
// touch it.  The communicator's unique id is produced on rank 0 (mh_comm_unique_id) and carried
// to the other ranks by whatever the caller has (MPI, a file, torch.distributed): 128 bytes.
// ============================================================================
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
enum { kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string error;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {std::getenv("MH_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);            // the copy the process already loaded (e.g. torch's)
      if (!api.handle) api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) { api.error = "libnccl.so.2 not found (set MH_NCCL_LIB)"; return; }
    auto sym = [&](const char* s) { void* p = dlsym(api.handle, s); if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + s; return p; };
    api.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
    api.Broadcast = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclBroadcast");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclAllReduce");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  });
  return api;
}

mh_status nccl_fail(mh_ctx* ctx, int rc, const char* what) {
  NcclApi& a = nccl();
  return mh::fail(ctx, MH_ENCCL, std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "NCCL error"));
}
#define MH_NCCL(ctx, call)                                   \
  do {                                                       \
    const int _rc = (call);                                  \
    if (_rc != 0) return nccl_fail((ctx), _rc, #call);       \
  } while (0)

}  // namespace

using namespace mh;

extern "C" {

mh_status mh_comm_unique_id(void* id128) {
  if (!id128) return MH_EINVAL;
  NcclApi& a = nccl();
  if (!a.error.empty()) return MH_ENCCL;
  NcclUniqueId id;
  if (a.GetUniqueId(&id) != 0) return MH_ENCCL;
  std::memcpy(id128, &id, sizeof(id));
  return MH_OK;
}

mh_status mh_comm_init(mh_ctx* ctx, const void* id128, int32_t rank, int32_t world) {
  if (!ctx) return MH_EINVAL;
  if (!id128 || world < 1 || rank < 0 || rank >= world) return fail(ctx, MH_EINVAL, "mh_comm_init: bad arguments");
  NcclApi& a = nccl();
  if (!a.error.empty()) return fail(ctx, MH_ENCCL, a.error);
  if (ctx->comm) return fail(ctx, MH_EINVAL, "mh_comm_init: the context already has a communicator");
  MH_CUDA(ctx, cudaSetDevice(ctx->device));
  NcclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  NcclComm comm = nullptr;
  MH_NCCL(ctx, a.CommInitRank(&comm, world, id, rank));
  ctx->comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  MH_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  for (cudaEvent_t& e : ctx->comm_ev) MH_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return MH_OK;
}

mh_status mh_comm_destroy(mh_ctx* ctx) {
  if (!ctx) return MH_EINVAL;
  if (!ctx->comm) return MH_OK;
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->comm_stream);
  nccl().CommDestroy((NcclComm)ctx->comm);
  ctx->comm = nullptr;
  for (cudaEvent_t& e : ctx->comm_ev)
    if (e) { cudaEventDestroy(e); e = nullptr; }
  if (ctx->comm_stream) { cudaStreamDestroy(ctx->comm_stream); ctx->comm_stream = nullptr; }
  if (ctx->comm_acc) { cudaFree(ctx->comm_acc); ctx->comm_acc = nullptr; ctx->comm_acc_k = 0; }
  ctx->comm_world = 1; ctx->comm_rank = 0; ctx->comm_pending = -1;
  return MH_OK;
}

void mh_step_release(mh_ctx* ctx) {   // called by mh_destroy: the second stream of mh_step_sharded
  if (!ctx) return;
  if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); ctx->aux_stream = nullptr; }
  for (cudaEvent_t& e : ctx->aux_ev)
    if (e) { cudaEventDestroy(e); e = nullptr; }
}

int32_t mh_comm_rank(const mh_ctx* ctx) { return ctx ? ctx->comm_rank : 0; }
int32_t mh_comm_world(const mh_ctx* ctx) { return ctx ? ctx->comm_world : 1; }

mh_status mh_comm_broadcast(mh_ctx* ctx, void* d_buf, uint64_t bytes, int32_t root) {
  if (!ctx) return MH_EINVAL;
  if (!ctx->comm) return ctx->comm_world == 1 ? MH_OK : fail(ctx, MH_ENCCL, "no communicator");
  MH_NCCL(ctx, nccl().Broadcast(d_buf, d_buf, bytes / 4, kNcclFloat32, root, (NcclComm)ctx->comm, ctx->stream));
  return MH_OK;
}

mh_status mh_comm_allreduce_sum_f64(mh_ctx* ctx, void* d_buf, uint64_t count) {
  if (!ctx) return MH_EINVAL;
  if (!ctx->comm) return MH_OK;
  MH_NCCL(ctx, nccl().AllReduce(d_buf, d_buf, count, kNcclFloat64, kNcclSum, (NcclComm)ctx->comm, ctx->stream));
  return MH_OK;
}

// completes the pass whose all-reduce is still in flight: unpack the whole-scene inlier counts, solve the K eigen-problems
mh_status mh_step_sharded_finish(mh_ctx* ctx) {
  if (!ctx) return MH_EINVAL;
  if (ctx->comm_pending < 0) return MH_OK;
  const int b = ctx->comm_pending;
  ctx->comm_pending = -1;
  double* acc = ctx->comm_acc + (size_t)b * 12 * ctx->comm_acc_k;
  MH_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->comm_ev[2 + b], 0));        // the reduction has landed
  if (ctx->pend_inliers) MH_TRY(launch_pack_inlier_counts(ctx, (int32_t*)ctx->pend_inliers, ctx->pend_K, acc, 1));
  MH_TRY(launch_refit_haf_solve(ctx, acc, ctx->pend_K, (float*)ctx->pend_ref, nullptr));
  return MH_OK;
}

// One sharded hot pass (see the header of this file).  d_hyp [K][12] is read on rank 0 and overwritten elsewhere (broadcast);
// d_hyp_pt [n_local][12], d_best u64 [n_local], d_labels i32 [n_local] are the local shard's K1 / K2 outputs; d_inliers i32 [K]
// and d_ref [K][12] receive the WHOLE-SCENE inlier counts and refined homographies — valid after the next mh_step_sharded or
// mh_step_sharded_finish (the reduction of pass i overlaps the kernels of pass i + 1).  Without a communicator (world 1) the
// same call runs the single-GPU pass.  ev_k2_begin / ev_k2_end: optional cudaEvent_t recorded around the K2 launch (measurement).
mh_status mh_step_sharded(mh_ctx* ctx, const void* d_pts, const void* d_aff, int64_t n_local, void* d_hyp, int32_t K, void* d_hyp_pt,
                          void* d_best, void* d_labels, void* d_inliers, void* d_ref, void* ev_k2_begin, void* ev_k2_end) {
  if (!ctx) return MH_EINVAL;
  if (!ctx->have_geom) return fail(ctx, MH_EINVAL, "call mh_set_geometry first");
  if (n_local < 0 || K <= 0 || !d_pts || !d_aff || !d_hyp || !d_hyp_pt || !d_best || !d_labels || !d_ref)
    return fail(ctx, MH_EINVAL, "mh_step_sharded: bad arguments");
  NcclApi& a = nccl();
  const bool multi = ctx->comm != nullptr;
  if (ctx->comm_acc_k < K) {   // statistics buffers: two, so that a reduction can be in flight while the next pass accumulates
    MH_TRY(mh_step_sharded_finish(ctx));
    MH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->comm_acc) cudaFree(ctx->comm_acc);
    MH_CUDA(ctx, cudaMalloc(&ctx->comm_acc, sizeof(double) * 2 * 12 * (size_t)K));
    ctx->comm_acc_k = K;
  }
  const int b = ctx->comm_cur;
  double* acc = ctx->comm_acc + (size_t)b * 12 * ctx->comm_acc_k;
  if (!ctx->aux_stream) {
    MH_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    for (cudaEvent_t& e : ctx->aux_ev) MH_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  MH_CUDA(ctx, cudaEventRecord(ctx->aux_ev[0], ctx->stream));   // everything before this pass (inputs uploaded, previous pass done)
  if (multi) {   // hypotheses from rank 0, on the communication stream
    MH_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->aux_ev[0], 0));
    MH_NCCL(ctx, a.Broadcast(d_hyp, d_hyp, (size_t)K * 12, kNcclFloat32, 0, (NcclComm)ctx->comm, ctx->comm_stream));
    MH_CUDA(ctx, cudaEventRecord(ctx->comm_ev[1], ctx->comm_stream));
    MH_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->comm_ev[1], 0));   // K2 needs the hypotheses
  }
  // K1 depends on nothing in this pass and nothing in it depends on K1, so it runs on a second stream.  One GPU: it starts when
  // K2 has finished and runs beside the K4 chain (label CSR, segmented statistics: small, latency-bound grids that leave most SMs
  // idle) — K2 itself keeps the chip to itself, which also keeps its event-timed duration (the roofline figure) clean.  Several
  // GPUs: it is enqueued first and runs under the hypothesis broadcast that K2 has to wait for.  The pass's stream joins it
  // before the solves.
  auto k1_on_second_stream = [&]() -> mh_status {
    cudaStream_t main_stream = ctx->stream;
    if (!multi) MH_CUDA(ctx, cudaEventRecord(ctx->aux_ev[0], ctx->stream));   // K2 done
    MH_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_ev[0], 0));
    ctx->stream = ctx->aux_stream;
    const mh_status st = launch_haf(ctx, (const float4*)d_pts, (const float4*)d_aff, n_local, (float*)d_hyp_pt, 0);       // K1
    ctx->stream = main_stream;
    if (st != MH_OK) return st;
    MH_CUDA(ctx, cudaEventRecord(ctx->aux_ev[1], ctx->aux_stream));
    return MH_OK;
  };
  if (multi) MH_TRY(k1_on_second_stream());
  if (ev_k2_begin) MH_CUDA(ctx, cudaEventRecord((cudaEvent_t)ev_k2_begin, ctx->stream));
  MH_TRY(launch_cost_fused(ctx, (const float4*)d_pts, n_local, (const float*)d_hyp, K, 0, nullptr, nullptr,
                           (unsigned long long*)d_best, (int32_t*)d_inliers));                                             // K2
  if (ev_k2_end) MH_CUDA(ctx, cudaEventRecord((cudaEvent_t)ev_k2_end, ctx->stream));
  if (!multi) MH_TRY(k1_on_second_stream());
  MH_TRY(launch_labels_from_best(ctx, (const unsigned long long*)d_best, n_local, (int32_t*)d_labels));
  MH_CUDA(ctx, cudaMemcpyAsync(d_ref, d_hyp, sizeof(float) * 12 * (size_t)K, cudaMemcpyDeviceToDevice, ctx->stream));     // labels without members keep theirs
  MH_TRY(launch_refit_haf_accumulate(ctx, (const float4*)d_pts, (const float4*)d_aff, (const int32_t*)d_labels, n_local, K, acc));   // K4 statistics
  MH_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->aux_ev[1], 0));                                                       // K1 joined
  if (!multi) return launch_refit_haf_solve(ctx, acc, K, (float*)d_ref, nullptr);
  if (d_inliers) MH_TRY(launch_pack_inlier_counts(ctx, (int32_t*)d_inliers, K, acc, 0));
  MH_CUDA(ctx, cudaEventRecord(ctx->comm_ev[0], ctx->stream));
  MH_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->comm_ev[0], 0));
  MH_NCCL(ctx, a.AllReduce(acc, acc, (size_t)K * 12, kNcclFloat64, kNcclSum, (NcclComm)ctx->comm, ctx->comm_stream));
  MH_CUDA(ctx, cudaEventRecord(ctx->comm_ev[2 + b], ctx->comm_stream));
  // the previous pass's reduction has had a whole pass to land: finish it now, then remember this one
  const int prev_b = ctx->comm_pending;
  if (prev_b >= 0) MH_TRY(mh_step_sharded_finish(ctx));
  ctx->comm_pending = b;
  ctx->pend_K = K; ctx->pend_inliers = d_inliers; ctx->pend_ref = d_ref;
  ctx->comm_cur = b ^ 1;
  return MH_OK;
}

}  // extern "C"
