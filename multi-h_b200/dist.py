"""Multi-GPU layer: one process per GPU, correspondences block-sharded, hypotheses replicated.

The path shards per correspondence (K1, K2, K4-accumulate are independent per site), so the only exchanges are tiny and
latency-bound: broadcast of the hypothesis block (K x 12 f32), all-reduce of per-hypothesis inlier counts (K i32) and of
the per-label refit statistics (K x 12 f64), all-gather of labels when the host graph-cut needs the full view.  They run
over torch.distributed — NCCL on NVLink/NVSwitch on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; the first n % world ranks hold one extra correspondence."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_hypotheses(hyp: torch.Tensor, src: int = 0) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(hyp, src=src)
    return hyp


def allreduce_sum(t: torch.Tensor) -> torch.Tensor:
    """per-hypothesis inlier counts / per-label refit statistics: plain sums over the correspondence shards"""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t


def allgather_labels(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """labels of all shards in global correspondence order (shards may differ in size by one)"""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return local
    world = dist.get_world_size()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    pad = torch.full((cap,), -2, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([out[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)])
