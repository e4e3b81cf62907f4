"""world_size-2 gloo tests (CPU) of the multi-GPU layer: sharding, the all-reduce of per-shard statistics and the label
all-gather reproduce the single-process result.  Per-shard statistics come from the oracle (no GPU here); on the GPU box
the same collectives carry the kernels' outputs (bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions(mh):
    for n in (0, 1, 7, 4194304, 4194305):
        for world in (1, 2, 3, 8):
            r = [mh.dist.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import multih_b200 as m
    from oracle import oracle as orc

    sc = m.scenes.make_scene(4001, 5, seed=31)  # odd size -> uneven shards
    lo, hi = m.dist.shard_range(len(sc.pts), rank, world)
    K = 5
    hyp = torch.from_numpy(sc.planes.copy()) if rank == 0 else torch.zeros((K, 9), dtype=torch.float64)
    m.dist.broadcast_hypotheses(hyp, 0)
    H = hyp.numpy()
    # per-shard K2 (argmin labels + inlier counts) and K4 statistics
    _, arg, cnt = orc.data_cost_sweep(sc.pts[lo:hi], H)
    labels = arg - 1
    _, M10, n = orc.refit_haf(sc.pts[lo:hi], sc.aff[lo:hi], labels, K, sc.F)
    acc = torch.zeros((K, 12), dtype=torch.float64)
    acc[:, :10] = torch.from_numpy(M10); acc[:, 10] = torch.from_numpy(n.astype(np.float64))
    m.dist.allreduce_sum(acc)
    counts = m.dist.allreduce_sum(torch.from_numpy(cnt.copy()))
    all_labels = m.dist.allgather_labels(torch.from_numpy(labels.copy()), len(sc.pts))
    if rank == 0:
        q.put((acc.numpy(), counts.numpy(), all_labels.numpy(), H))
    dist.destroy_process_group()


def test_two_rank_statistics_match_single_process(mh, orc):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    acc, counts, labels, H = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc = mh.scenes.make_scene(4001, 5, seed=31)
    assert np.array_equal(H, sc.planes)                      # broadcast
    _, arg, cnt = orc.data_cost_sweep(sc.pts, sc.planes)
    assert np.array_equal(labels, arg - 1)                   # all-gather, global order, uneven shards
    assert np.array_equal(counts, cnt)                       # all-reduce of inlier counts
    _, M10, n = orc.refit_haf(sc.pts, sc.aff, arg - 1, 5, sc.F)
    assert np.array_equal(acc[:, 10], n.astype(np.float64))
    assert np.allclose(acc[:, :10], M10, rtol=1e-12)         # all-reduce of refit statistics


def test_comm_layer_argument_checks_and_nccl_binding(mh):
    """csrc/comm.cu without a GPU: null arguments are refused, a null context reads as a single-rank world, and rank 0's
    128-byte NCCL id is produced by the NCCL the process already has (torch's) — run-time binding, no link dependency."""
    import ctypes as C

    L = mh.capi.lib()
    assert L.mh_comm_unique_id(None) == mh.capi.MH_EINVAL
    assert L.mh_comm_world(None) == 1 and L.mh_comm_rank(None) == 0
    assert L.mh_comm_init(None, None, 0, 1) == mh.capi.MH_EINVAL
    assert L.mh_step_sharded_finish(None) == mh.capi.MH_EINVAL
    buf = (C.c_char * 128)()
    st = L.mh_comm_unique_id(buf)
    assert st in (mh.capi.MH_OK, mh.capi.MH_ENCCL)   # ENCCL only where ncclGetUniqueId itself cannot run (no NCCL / no driver)
    if st == mh.capi.MH_OK:
        assert any(b != b"\x00" for b in buf)
