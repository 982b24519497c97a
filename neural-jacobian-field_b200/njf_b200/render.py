"""Typed Python entry to the fused render path (njf_render_forward and its stage entry points)."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field as dc_field
from typing import List, Optional, Sequence

import torch

from . import _lib, api


@dataclass
class RenderResult:
    rgb: Optional[torch.Tensor] = None            # (B,R,3)
    depth: Optional[torch.Tensor] = None          # (B,R,1)
    flow: Optional[torch.Tensor] = None           # (B,R,2)
    jbar: Optional[torch.Tensor] = None           # (B,R,3A)
    p: Optional[torch.Tensor] = None              # (B,R,3)
    pw: Optional[torch.Tensor] = None             # (B,R,3)
    steps: Optional[torch.Tensor] = None          # (B,R,S)
    weights: Optional[torch.Tensor] = None        # (B,R,S)
    sigma: Optional[torch.Tensor] = None          # (B,R,S,1)
    jac: Optional[torch.Tensor] = None            # (B,R,S,3A)
    positions: Optional[torch.Tensor] = None      # (B,R,S,3)
    rgb_samples: Optional[torch.Tensor] = None    # (B,R,S,3)
    prop_weights: List[torch.Tensor] = dc_field(default_factory=list)   # per level (B,R,S_l)
    level_bins: List[torch.Tensor] = dc_field(default_factory=list)     # per level (B,R,n_l+1) spacing bins
    level_inds: List[torch.Tensor] = dc_field(default_factory=list)     # per level (B,R,n_l+1) int32
    bins0: Optional[torch.Tensor] = None
    minmax: Optional[torch.Tensor] = None         # (2,) call-global (min, max) of steps (the depth clip range)
    packed: Optional[torch.Tensor] = None         # (B,R,12+3A) rgb|depth|flow|jbar|p|pw (multi-GPU gather buffer)


def render(fld: api.Field, maps: torch.Tensor, Hf: int, Wf: int, cams: api.NjfCameras, origins: torch.Tensor,
           dirs: torch.Tensor, z_near: torch.Tensor, z_far: torch.Tensor, action: torch.Tensor,
           s_prop: Sequence[int], s_nerf: int, *, vis: bool = True, per_sample: bool = False,
           sampler_outputs: bool = False, bins0: Optional[torch.Tensor] = None,
           us: Optional[Sequence[torch.Tensor]] = None, anneal: float = 1.0,
           sum_vec_width: Optional[int] = None, final_bins: Optional[torch.Tensor] = None,
           host_near_far: Optional[Sequence[torch.Tensor]] = None, packed: bool = False,
           workspace: Optional[torch.Tensor] = None, minmax_hook=None,
           ray_range: Optional[Sequence[int]] = None) -> RenderResult:
    """One fused render of B x R rays.  All tensors live on the field's CUDA device.

    ``final_bins`` (B,R,s_nerf+1): skip the proposal levels and render the field at the given
    spacing-domain bins (stage-wise parity tests / externally supplied samples).
    ``workspace``: caller-owned uint8 scratch (any size >= fld.workspace_min_bytes); default = the field's
    per-stream buffer of fld.workspace_bytes(...).
    ``minmax_hook(minmax)``: called between the field pass and the finish pass with the (2,) device tensor of the
    call's (min, max) sample distance -- a ray-sharded render all-reduces it there so that every shard applies
    the reference's call-global depth clip (models/model.py:277); forces the staged entry points.
    ``ray_range = (n_views, rays_per_view, ray_offset)``: a ray-sharded call -- ``origins`` / ``dirs`` are
    (1, n, 3) and hold rays [ray_offset, ray_offset + n) of the flattened (view, ray) space of an
    n_views x rays_per_view call; cameras, z_near, z_far, action and ``maps`` cover all n_views views."""
    L = api._declare()
    dev = origins.device
    fld._check_device(origins, "ray origins")
    fld._check_device(maps, "hoisted maps")
    B, R = origins.shape[:2]
    Bv, Rv = B, R            # views / rays per view as the kernels see them
    if ray_range is not None:
        Bv, Rv, ray_offset = (int(v) for v in ray_range)
        if B != 1:
            raise _lib.NjfError("a ray-sharded call takes origins / dirs of shape (1, n, 3)")
    A = fld.action_dim
    f32 = dict(device=dev, dtype=torch.float32)
    origins, dirs = origins.contiguous().float(), dirs.contiguous().float()
    z_near, z_far = z_near.contiguous().float(), z_far.contiguous().float()
    action = action.contiguous().float()
    if len(s_prop) != fld.n_proposal:
        raise _lib.NjfError(f"{len(s_prop)} proposal levels requested, field has {fld.n_proposal}")
    if bins0 is None or us is None:   # eval mode: the reference's own linspace tables (host -> device copy)
        tab_bins0, tab_us = api.eval_tables(s_prop, s_nerf, dev)
    bins0 = tab_bins0 if bins0 is None else bins0.contiguous().float()
    us = tab_us if us is None else [u.contiguous().float() for u in us]

    res = RenderResult(bins0=bins0)
    a = api.NjfRenderArgs()
    a.B, a.R, a.n_levels, a.s_nerf = Bv, Rv, len(s_prop), int(s_nerf)
    if ray_range is not None:
        a.ray_offset, a.n_rays = ray_offset, R
    for i, s in enumerate(s_prop):
        a.s_prop[i] = int(s)
    a.origins, a.dirs = api.dptr(origins), api.dptr(dirs)
    a.z_near, a.z_far, a.action = api.dptr(z_near), api.dptr(z_far), api.dptr(action)
    if host_near_far is not None:  # host copies let the kernels read per-view constants from the constant bank
        hn, hf = (t.detach().to("cpu", torch.float32).contiguous() for t in host_near_far)
        a.h_z_near, a.h_z_far = hn.data_ptr(), hf.data_ptr()
    a.bins0 = api.dptr(bins0)
    a.bins0_stride = 0 if bins0.dim() == 1 else bins0.shape[-1]
    a.anneal = float(anneal)
    a.sum_vec_width = api.default_sum_vec_width() if sum_vec_width is None else int(sum_vec_width)
    a.maps, a.Hf, a.Wf = api.dptr(maps), int(Hf), int(Wf)
    keep = [origins, dirs, z_near, z_far, action, bins0, maps] + list(us)
    if host_near_far is not None:
        keep += [hn, hf]
    for lvl in range(len(s_prop)):
        n = s_prop[lvl + 1] if lvl + 1 < len(s_prop) else s_nerf
        u = us[lvl]
        a.u[lvl] = api.dptr(u)
        a.u_stride[lvl] = 0 if u.dim() == 1 else n + 1
        lb = torch.empty(B, R, n + 1, **f32)
        res.level_bins.append(lb)
        a.level_bins[lvl] = api.dptr(lb)
        if sampler_outputs:
            pw = torch.empty(B, R, s_prop[lvl], **f32)
            li = torch.empty(B, R, n + 1, device=dev, dtype=torch.int32)
            res.prop_weights.append(pw)
            res.level_inds.append(li)
            a.prop_weights[lvl] = api.dptr(pw)
            a.level_inds[lvl] = api.dptr(li)
    minmax = torch.empty(2, **f32)
    a.minmax = api.dptr(minmax)
    res.minmax = minmax
    if workspace is None:
        workspace = fld.workspace(fld.workspace_bytes(B, R, s_prop, s_nerf))
    assert workspace.is_cuda and workspace.dtype == torch.uint8 and workspace.is_contiguous()
    a.workspace, a.workspace_bytes = workspace.data_ptr(), workspace.numel()
    keep.append(workspace)
    res.rgb = torch.empty(B, R, 3, **f32)
    res.depth = torch.empty(B, R, 1, **f32)
    res.flow = torch.empty(B, R, 2, **f32)
    res.jbar = torch.empty(B, R, 3 * A, **f32)
    res.p = torch.empty(B, R, 3, **f32)
    res.pw = torch.empty(B, R, 3, **f32)
    a.rgb, a.depth, a.flow = api.dptr(res.rgb), api.dptr(res.depth), api.dptr(res.flow)
    a.jbar, a.p, a.pw = api.dptr(res.jbar), api.dptr(res.p), api.dptr(res.pw)
    if packed:
        res.packed = torch.empty(B, R, 12 + 3 * A, **f32)
        a.packed = api.dptr(res.packed)
    if vis or per_sample:
        res.steps = torch.empty(B, R, s_nerf, **f32)
        res.weights = torch.empty(B, R, s_nerf, **f32)
        a.steps, a.weights = api.dptr(res.steps), api.dptr(res.weights)
    if per_sample:
        res.sigma = torch.empty(B, R, s_nerf, 1, **f32)
        res.jac = torch.empty(B, R, s_nerf, 3 * A, **f32)
        res.positions = torch.empty(B, R, s_nerf, 3, **f32)
        res.rgb_samples = torch.empty(B, R, s_nerf, 3, **f32)
        a.sigma, a.jac = api.dptr(res.sigma), api.dptr(res.jac)
        a.positions, a.rgb_samples = api.dptr(res.positions), api.dptr(res.rgb_samples)
    st = api.stream_ptr()
    h = fld.handle
    if final_bins is None and minmax_hook is None:
        _lib.check(L.njf_render_forward(h, ctypes.byref(cams), ctypes.byref(a), st))
    elif final_bins is None:
        bins, stride = bins0, a.bins0_stride
        for lvl in range(len(s_prop)):
            _lib.check(L.njf_proposal_pass(h, ctypes.byref(cams), ctypes.byref(a), lvl, api.dptr(bins), stride, st))
            bins = res.level_bins[lvl]
            stride = bins.shape[-1]
        _lib.check(L.njf_field_pass(h, ctypes.byref(cams), ctypes.byref(a), api.dptr(bins), stride, st))
        minmax_hook(minmax)
        _lib.check(L.njf_finish_pass(h, ctypes.byref(cams), ctypes.byref(a), st))
    else:
        fb = final_bins.contiguous().float()
        keep.append(fb)
        stride = 0 if fb.dim() == 1 else fb.shape[-1]
        _lib.check(L.njf_field_pass(h, ctypes.byref(cams), ctypes.byref(a), api.dptr(fb), stride, st))
        if minmax_hook is not None:
            minmax_hook(minmax)
        _lib.check(L.njf_finish_pass(h, ctypes.byref(cams), ctypes.byref(a), st))
    res._keep = keep  # inputs stay alive until the caller is done with the (async) result
    return res
