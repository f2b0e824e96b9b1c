"""Multi-GPU plumbing of the KLT path: static block partition of independent image pairs + final gather.

The path shards embarrassingly: features and pairs are independent
(/root/reference src/algorithm.cpp:43 touches index i only), so there is NO data-path collective;
the only exchange is the gather of (x, y, flag) per feature at the end.  One process per GPU,
torch.distributed for the plumbing (nccl on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [start, stop) of `n_units` owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world) or n_units < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_units, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(n_units: int, world: int) -> list[int]:
    return [shard_range(n_units, r, world)[1] - shard_range(n_units, r, world)[0] for r in range(world)]


def gather_results(kp2_local: torch.Tensor, succ_local: torch.Tensor, n_units: int, group=None):
    """Final gather: every rank passes its [units_local, n, 2] float32 positions and [units_local, n]
    uint8 flags (block partition of n_units); returns the full arrays in unit order on every rank.
    Uneven shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()):
        return kp2_local, succ_local
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_units, world)
    biggest = max(sizes)
    n = kp2_local.shape[1] if kp2_local.dim() > 1 else 0

    def padded(t, tail_shape):
        if t.shape[0] == biggest:
            return t.contiguous()
        pad = torch.zeros((biggest - t.shape[0],) + tail_shape, dtype=t.dtype, device=t.device)
        return torch.cat([t, pad], 0).contiguous()

    kp = padded(kp2_local, (n, 2))
    sc = padded(succ_local, (n,))
    kp_all = [torch.empty_like(kp) for _ in range(world)]
    sc_all = [torch.empty_like(sc) for _ in range(world)]
    dist.all_gather(kp_all, kp, group=group)
    dist.all_gather(sc_all, sc, group=group)
    kp_full = torch.cat([kp_all[r][:sizes[r]] for r in range(world)], 0)
    sc_full = torch.cat([sc_all[r][:sizes[r]] for r in range(world)], 0)
    return kp_full, sc_full
