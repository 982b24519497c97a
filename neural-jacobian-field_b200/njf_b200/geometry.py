"""Ray-bundle geometry of the render path, mirroring the reference's function names
(neural_jacobian_field/rendering/geometry.py:117-134, 76-114, 170-203).  The ray arithmetic runs in
``njf_make_rays`` (csrc/rays.cu); CPU tensors are rejected (there is no CPU fallback)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib, api


def get_pixel_coordinates(height: int, width: int, device: torch.device = torch.device("cpu")) -> Tuple[Tensor, Tensor]:
    """Normalised (0..1) xy pixel centres (H, W, 2) and (row, col) selectors (H, W, 2), geometry.py:117-134.
    Index bookkeeping only (no kernel): ray order is row-major, x fastest."""
    # built on the host and moved: torch's CUDA division by a scalar multiplies by the reciprocal, which differs
    # from the reference's CPU result (and from the kernel's correctly rounded division) in the last bit
    row = torch.arange(height)
    col = torch.arange(width)
    selector = torch.stack(torch.meshgrid(row, col, indexing="ij"), dim=-1).to(device)
    x = (col + 0.5) / width
    y = (row + 0.5) / height
    coordinates = torch.stack(torch.meshgrid(x, y, indexing="xy"), dim=-1).to(device)
    return coordinates, selector


def _rays(coordinates_xy: Optional[Tensor], intrinsics: Tensor, cam2world: Tensor, hw: Optional[Tuple[int, int]],
          want_z: bool):
    if not intrinsics.is_cuda:
        raise _lib.NjfError("njf_b200.geometry: ray generation runs on the GPU only (no CPU fallback)")
    L = api._declare()
    dev = intrinsics.device
    B = intrinsics.shape[0]
    k = intrinsics.detach().contiguous().float()
    c2w = cam2world.detach().to(dev).contiguous().float()
    if coordinates_xy is not None:
        xy = coordinates_xy.detach().to(dev).contiguous().float()
        if xy.dim() != 3 or xy.shape[0] != B or xy.shape[-1] != 2:
            raise _lib.NjfError(f"coordinates_xy must be (camera, ray, 2), got {tuple(xy.shape)}")
        R, H, W = xy.shape[1], 0, 0
    else:
        H, W = hw
        xy, R = None, H * W
    o = torch.empty(B, R, 3, device=dev, dtype=torch.float32)
    d = torch.empty(B, R, 3, device=dev, dtype=torch.float32)
    z = torch.empty(B, R, 1, device=dev, dtype=torch.float32) if want_z else None
    with torch.cuda.device(dev):
        _lib.check(L.njf_make_rays(api.dptr(k), api.dptr(c2w), api.dptr(xy) if xy is not None else None, B, R, H, W,
                                   api.dptr(o), api.dptr(d), api.dptr(z) if z is not None else None, api.stream_ptr()))
    return o, d, z


def get_world_rays(coordinates_xy: Tensor, intrinsics: Tensor, cam2world: Tensor) -> Tuple[Tensor, Tensor]:
    """origins, unit directions (camera, ray, 3) -- geometry.py:76-114."""
    o, d, _ = _rays(coordinates_xy, intrinsics, cam2world, None, False)
    return o, d


def get_world_rays_with_z(coordinates_xy: Tensor, intrinsics: Tensor, cam2world: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """origins, unit directions, camera-space z of the unit direction -- geometry.py:170-203."""
    return _rays(coordinates_xy, intrinsics, cam2world, None, True)


def get_world_rays_grid(height: int, width: int, intrinsics: Tensor, cam2world: Tensor, with_z: bool = False):
    """get_pixel_coordinates + get_world_rays(_with_z) fused: the H x W grid of pixel centres is generated inside
    the kernel (no coordinate tensor is materialised)."""
    o, d, z = _rays(None, intrinsics, cam2world, (height, width), with_z)
    return (o, d, z) if with_z else (o, d)
