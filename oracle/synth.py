"""TEST INFRASTRUCTURE: the seeded synthetic-data scheme shared with the product's benchmarks
(``njf_b200/synth.py``: weights keyed by (seed, crc32(state-dict key)), camera rig, layer shapes) plus CPU
restatements of the reference's ray generation that the oracle-side scripts and tests use to make inputs."""
from __future__ import annotations

import os
import sys

import torch

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neural-jacobian-field_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
from njf_b200.synth import *  # noqa: F401,F403,E402
from njf_b200.synth import ALLEGRO_INTRINSICS_PX, synth_tensor  # noqa: F401,E402  (explicit: not covered by *)


def pixel_grid(height: int, width: int) -> torch.Tensor:
    """Normalised pixel centres, row-major, x fastest (geometry.py:117-134)."""
    x = (torch.arange(width, dtype=torch.float32) + 0.5) / width
    y = (torch.arange(height, dtype=torch.float32) + 0.5) / height
    return torch.stack(torch.meshgrid(x, y, indexing="xy"), dim=-1).reshape(-1, 2)


def world_rays(coords_xy: torch.Tensor, k_norm: torch.Tensor, c2w: torch.Tensor):
    """origins, unit directions for normalised pixel coords (geometry.py:170-203)."""
    ones = torch.ones_like(coords_xy[..., :1])
    d = torch.cat([coords_xy, ones], -1) @ torch.inverse(k_norm).T
    d = d / d.norm(dim=-1, keepdim=True)
    d = d @ c2w[:3, :3].T
    o = c2w[:3, 3].expand_as(d)
    return o.contiguous(), d.contiguous()
