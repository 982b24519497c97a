"""Multi-GPU plumbing of the render path: rays shard embarrassingly (one process per GPU, replicated
weights, per-view feature maps), the rendered per-ray buffers are returned with ONE collective.

The reference has no collective on the render path (inference is single-GPU);
the only cross-ray coupling is render_depth's call-global clip range (models/model.py:277), which a
ray-sharded call reproduces by all-reducing (min, max) of the sample steps between the field pass
and the finish pass.  Works with the NCCL backend on GPUs and with gloo on CPU tensors (tests).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

PACK_ORDER = ("rgb", "depth", "flow", "jbar", "p", "pw")


def ray_shard(n_rays: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of the flattened (view, ray) index owned by ``rank``; the
    remainder goes to the low ranks, so shard sizes differ by at most one ray."""
    base, rem = divmod(n_rays, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_widths(action_dim: int) -> Dict[str, int]:
    return {"rgb": 3, "depth": 1, "flow": 2, "jbar": 3 * action_dim, "p": 3, "pw": 3}


def pack_outputs(out: Dict[str, torch.Tensor]) -> torch.Tensor:
    """(n, 12 + 3A) row-major struct-of-rays buffer: one gather instead of six."""
    return torch.cat([out[k].reshape(-1, out[k].shape[-1]) for k in PACK_ORDER], dim=1).contiguous()


def unpack_outputs(buf: torch.Tensor, action_dim: int) -> Dict[str, torch.Tensor]:
    res, c = {}, 0
    for k, w in pack_widths(action_dim).items():
        res[k] = buf[:, c:c + w]
        c += w
    return res


def allreduce_minmax(minmax: torch.Tensor, group=None) -> torch.Tensor:
    """In-place all-reduce of the (min, max) pair produced by njf_field_pass."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        lo, hi = minmax[0:1].clone(), minmax[1:2].clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        minmax[0:1].copy_(lo)
        minmax[1:2].copy_(hi)
    return minmax


def gather_rendered(packed: torch.Tensor, n_rays_total: int, group=None) -> torch.Tensor:
    """All-gather the per-rank packed ray buffers into the full (n_rays_total, C) frame buffer.
    Shards may differ by one ray, so ranks pad to the largest shard for the single collective."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return packed
    C = packed.shape[1]
    sizes = [ray_shard(n_rays_total, r, world) for r in range(world)]
    mx = max(b - a for a, b in sizes)
    padded = packed if packed.shape[0] == mx else torch.cat([packed, packed.new_zeros(mx - packed.shape[0], C)])
    out = packed.new_empty(world * mx, C)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    return torch.cat([out[r * mx: r * mx + (b - a)] for r, (a, b) in enumerate(sizes)], dim=0)
