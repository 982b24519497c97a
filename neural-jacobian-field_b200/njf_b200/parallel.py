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


def shard_views(start: int, stop: int, rays_per_view: int) -> Tuple[int, int]:
    """Views [v0, v1) touched by the flattened ray range [start, stop)."""
    return start // rays_per_view, (stop - 1) // rays_per_view + 1


def render_sharded(fld, maps, Hf: int, Wf: int, cams, origins: torch.Tensor, dirs: torch.Tensor, z_near, z_far, action,
                   s_prop, s_nerf: int, n_views: int, rays_per_view: int, rank: int, world: int, group=None,
                   gather: bool = True, **kw):
    """One rank's part of a ray-sharded B-view call (BASELINE config 4): rays [start, stop) of the
    flattened (view, ray) space, the reference's call-global depth clip reproduced by all-reducing (min, max) of
    the sample steps between njf_field_pass and njf_finish_pass, and the packed per-ray struct that
    njf_finish_pass writes gathered with ONE collective.  ``origins`` / ``dirs``: this rank's (n, 3) rays;
    cameras / z_near / z_far / action / ``maps``: all ``n_views`` views (only the views the range touches are read).
    Returns (RenderResult of the shard, gathered (n_views * rays_per_view, 12 + 3A) frame or None)."""
    from .render import render

    start, stop = ray_shard(n_views * rays_per_view, rank, world)
    assert origins.shape[0] == stop - start, "origins must hold exactly this rank's rays"
    res = render(fld, maps, Hf, Wf, cams, origins[None], dirs[None], z_near, z_far, action, s_prop, s_nerf, packed=True,
                 minmax_hook=lambda mm: allreduce_minmax(mm, group), ray_range=(n_views, rays_per_view, start), **kw)
    frame = gather_rendered(res.packed[0], n_views * rays_per_view, group) if gather else None
    return res, frame


def gather_rendered(packed: torch.Tensor, n_rays_total: int, group=None, dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """The per-rank packed ray buffers -> the full (n_rays_total, C) frame buffer with ONE collective: a gather to
    rank ``dst`` (the other ranks return None; moves world x less data than an all-gather) or, with dst=None, an
    all-gather.  Shards may differ by one ray, so ranks pad to the largest shard."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return packed
    rank = dist.get_rank(group)
    C = packed.shape[1]
    sizes = [ray_shard(n_rays_total, r, world) for r in range(world)]
    mx = max(b - a for a, b in sizes)
    padded = packed if packed.shape[0] == mx else torch.cat([packed, packed.new_zeros(mx - packed.shape[0], C)])
    padded = padded.contiguous()
    if dst is None:
        out = packed.new_empty(world * mx, C)
        dist.all_gather_into_tensor(out, padded, group=group)
    else:
        out = packed.new_empty(world * mx, C) if rank == dst else None
        dist.gather(padded, list(out.view(world, mx, C).unbind(0)) if rank == dst else None, dst=dst, group=group)
        if rank != dst:
            return None
    if all(b - a == mx for a, b in sizes):
        return out
    return torch.cat([out[r * mx: r * mx + (b - a)] for r, (a, b) in enumerate(sizes)], dim=0)


def gather_rows(rows: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather of row-sharded (n_local, C) buffers (ray_shard split of n_total rows) on every rank."""
    return gather_rendered(rows, n_total, group, dst=None)
