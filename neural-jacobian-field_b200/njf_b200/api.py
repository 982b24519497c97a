"""ctypes mirrors of the structs in ``include/njf_b200.h`` and thin typed wrappers.

Everything here is plumbing: pointers of torch CUDA tensors are handed to ``libnjf_b200.so``;
no rendering arithmetic happens in Python.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p
from typing import Dict, Optional, Sequence

import torch

from . import _lib

NJF_MAX_LEVELS = 4
HEAD_TRANSFORMER, HEAD_MLP = 0, 1
HEADS = {"jacobian_transformer": HEAD_TRANSFORMER, "jacobian_mlp": HEAD_MLP}
SH_CONVENTIONS = {"tcnn": 0, "nerfstudio_torch": 1}


class NjfFieldDesc(Structure):
    _fields_ = [("head", c_int), ("action_dim", c_int), ("n_proposal", c_int), ("encoder_dim", c_int),
                ("sh_fp16_round", c_int), ("sh_convention", c_int)]


class NjfTensor(Structure):
    _fields_ = [("name", c_char_p), ("data", c_void_p), ("numel", c_int64)]


class NjfCameras(Structure):
    _fields_ = [("ctxt_w2c", c_void_p), ("ctxt_k", c_void_p), ("trgt_w2c", c_void_p), ("trgt_k_px", c_void_p),
                ("h_ctxt_w2c", c_void_p), ("h_ctxt_k", c_void_p)]


class NjfRenderArgs(Structure):
    _fields_ = [
        ("B", c_int), ("R", c_int), ("n_levels", c_int), ("s_prop", c_int * NJF_MAX_LEVELS), ("s_nerf", c_int),
        ("origins", c_void_p), ("dirs", c_void_p), ("z_near", c_void_p), ("z_far", c_void_p),
        ("h_z_near", c_void_p), ("h_z_far", c_void_p), ("action", c_void_p),
        ("bins0", c_void_p), ("bins0_stride", c_int),
        ("u", c_void_p * NJF_MAX_LEVELS), ("u_stride", c_int * NJF_MAX_LEVELS),
        ("anneal", c_float), ("sum_vec_width", c_int),
        ("maps", c_void_p), ("Hf", c_int), ("Wf", c_int),
        ("rgb", c_void_p), ("depth", c_void_p), ("flow", c_void_p), ("jbar", c_void_p), ("p", c_void_p),
        ("pw", c_void_p),
        ("steps", c_void_p), ("weights", c_void_p), ("sigma", c_void_p), ("jac", c_void_p),
        ("positions", c_void_p), ("rgb_samples", c_void_p),
        ("prop_weights", c_void_p * NJF_MAX_LEVELS), ("level_bins", c_void_p * NJF_MAX_LEVELS),
        ("level_inds", c_void_p * NJF_MAX_LEVELS),
        ("minmax", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_size_t), ("packed", c_void_p),
        ("ray_offset", c_int), ("n_rays", c_int),
    ]


def _declare():
    L = _lib.lib()
    if getattr(L, "_njf_declared", False):
        return L
    L.njf_field_create.restype = c_int
    L.njf_field_create.argtypes = [POINTER(NjfFieldDesc), POINTER(NjfTensor), c_int, POINTER(c_void_p)]
    L.njf_field_destroy.restype = None
    L.njf_field_destroy.argtypes = [c_void_p]
    L.njf_field_update_head.restype = c_int
    L.njf_field_update_head.argtypes = [c_void_p, POINTER(NjfTensor), c_int, c_void_p]
    L.njf_hoisted_bytes.restype = c_size_t
    L.njf_hoisted_bytes.argtypes = [c_void_p, c_int, c_int, c_int]
    L.njf_hoist_features.restype = c_int
    L.njf_hoist_features.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
    L.njf_hoist_features_nhwc16.restype = c_int
    L.njf_hoist_features_nhwc16.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    L.njf_hoist_features_views.restype = c_int
    L.njf_hoist_features_views.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    for name in ("njf_render_forward", "njf_finish_pass"):
        fn = getattr(L, name)
        fn.restype = c_int
        fn.argtypes = [c_void_p, POINTER(NjfCameras), POINTER(NjfRenderArgs), c_void_p]
    L.njf_proposal_pass.restype = c_int
    L.njf_proposal_pass.argtypes = [c_void_p, POINTER(NjfCameras), POINTER(NjfRenderArgs), c_int, c_void_p, c_int,
                                    c_void_p]
    L.njf_field_pass.restype = c_int
    L.njf_field_pass.argtypes = [c_void_p, POINTER(NjfCameras), POINTER(NjfRenderArgs), c_void_p, c_int, c_void_p]
    L.njf_query_points.restype = c_int
    L.njf_query_points.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    L.njf_query_proposal_density.restype = c_int
    L.njf_query_proposal_density.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                                             c_int, c_void_p, c_void_p]
    L.njf_query_workspace_bytes.restype = c_size_t
    L.njf_query_workspace_bytes.argtypes = [c_void_p, c_int, c_int]
    L.njf_workspace_bytes.restype = c_size_t
    L.njf_workspace_bytes.argtypes = [c_void_p, c_int, c_int, c_int, POINTER(c_int), c_int]
    L.njf_workspace_min_bytes.restype = c_size_t
    L.njf_workspace_min_bytes.argtypes = [c_void_p, c_int, POINTER(c_int), c_int]
    L.njf_invert_poses.restype = c_int
    L.njf_invert_poses.argtypes = [c_void_p, c_void_p, c_int, c_void_p]
    L.njf_point_features.restype = c_int
    L.njf_point_features.argtypes = [c_void_p] * 4 + [c_int] * 5 + [c_void_p, c_void_p, c_void_p]
    L.njf_pdf_sample.restype = c_int
    L.njf_pdf_sample.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int,
                                 c_void_p, c_void_p, c_void_p]
    L.njf_transmittance_weights.restype = c_int
    L.njf_transmittance_weights.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]
    L.njf_flow_from_encoding.restype = c_int
    L.njf_flow_from_encoding.argtypes = [c_void_p] * 5 + [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
    L.njf_make_rays.restype = c_int
    L.njf_make_rays.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    L.njf_flow_gn_terms.restype = c_int
    L.njf_flow_gn_terms.argtypes = [c_void_p] * 7 + [c_int, c_int, c_int] + [c_void_p] * 5
    L.njf_flow_gn_workspace_doubles.restype = c_int
    L.njf_flow_gn_workspace_doubles.argtypes = [c_int]
    L.njf_debug_field_timing.restype = c_int
    L.njf_debug_field_timing.argtypes = [c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    L._njf_declared = True
    return L


def stream_ptr() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def dptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device pointer of a contiguous CUDA tensor expected"
    return t.data_ptr()


def default_sum_vec_width() -> int:
    """ATen's CPU fp32 ``sum`` uses 8-lane vectors on AVX2 and AVX512 builds; reproducing that order
    makes the PDF sampler bit-identical to the reference's CPU path given identical weights."""
    return 8 if torch.backends.cpu.get_cpu_capability() in ("AVX2", "AVX512") else 0


class Field:
    """Owner of an ``NjfField*`` (packed decoder + proposal-network weights on the current device)."""

    def __init__(self, head: str, action_dim: int, n_proposal: int, weights: Dict[str, torch.Tensor],
                 sh_fp16_round: bool = True, encoder_dim: int = 512, sh_convention: str = "tcnn"):
        if head not in HEADS:
            raise _lib.NjfError(f"decoder '{head}' has no B200 kernel (supported: {sorted(HEADS)})")
        if sh_convention not in SH_CONVENTIONS:
            raise _lib.NjfError(f"sh_convention '{sh_convention}' unknown (supported: {sorted(SH_CONVENTIONS)})")
        L = _declare()
        self.head, self.action_dim, self.n_proposal = head, int(action_dim), int(n_proposal)
        desc = NjfFieldDesc(HEADS[head], int(action_dim), int(n_proposal), int(encoder_dim), int(bool(sh_fp16_round)),
                            SH_CONVENTIONS[sh_convention])
        keep = []
        arr = (NjfTensor * len(weights))()
        for i, (k, v) in enumerate(weights.items()):
            t = v.detach().to(device="cpu", dtype=torch.float32).contiguous()
            keep.append(t)
            arr[i] = NjfTensor(k.encode(), t.data_ptr(), t.numel())
        h = c_void_p()
        _lib.check(L.njf_field_create(ctypes.byref(desc), arr, len(weights), ctypes.byref(h)))
        self._h = h
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._ws: Dict[int, torch.Tensor] = {}   # caller-owned workspaces, one per stream (grow-only)

    @property
    def handle(self) -> c_void_p:
        return self._h

    def update_head(self, weights: Dict[str, torch.Tensor]) -> None:
        """Re-pack only the cross-attention Jacobian head in place (njf_field_update_head): what an optimiser step of the
        action phase changes.  ``weights``: the ``decoder.jacobian_*`` tensors."""
        L = _declare()
        keep = []
        arr = (NjfTensor * len(weights))()
        for i, (k, v) in enumerate(weights.items()):
            t = v.detach().to(device="cpu", dtype=torch.float32).contiguous()
            keep.append(t)
            arr[i] = NjfTensor(k.encode(), t.data_ptr(), t.numel())
        with torch.cuda.device(self.device):
            _lib.check(L.njf_field_update_head(self._h, arr, len(weights), stream_ptr()))

    def _check_device(self, t: torch.Tensor, what: str) -> None:
        if t.device != self.device:
            raise _lib.NjfError(f"{what} lives on {t.device}, the field was packed on {self.device}")

    def workspace_bytes(self, B: int, R: int, s_prop: Sequence[int], s_nerf: int) -> int:
        L = _declare()
        arr = (c_int * NJF_MAX_LEVELS)(*[int(s) for s in s_prop])
        return int(L.njf_workspace_bytes(self._h, int(B), int(R), len(s_prop), arr, int(s_nerf)))

    def workspace_min_bytes(self, s_prop: Sequence[int], s_nerf: int) -> int:
        L = _declare()
        arr = (c_int * NJF_MAX_LEVELS)(*[int(s) for s in s_prop])
        return int(L.njf_workspace_min_bytes(self._h, len(s_prop), arr, int(s_nerf)))

    def workspace(self, nbytes: int) -> torch.Tensor:
        """The library allocates nothing inside a pass; this is the host mirror's scratch for the CURRENT stream
        (one buffer per stream, so two streams can render from one field concurrently), grown on demand."""
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def hoist(self, feat_nchw: torch.Tensor) -> torch.Tensor:
        """(B,512,Hf,Wf) fp32 encoder output -> opaque fp16 hoisted maps (uint8 buffer)."""
        L = _declare()
        assert feat_nchw.is_cuda and feat_nchw.dtype == torch.float32
        self._check_device(feat_nchw, "feature map")
        feat_nchw = feat_nchw.contiguous()
        B, C, Hf, Wf = feat_nchw.shape
        if C != 512:
            raise _lib.NjfError(f"feature map has {C} channels, kernels are built for 512")
        nbytes = L.njf_hoisted_bytes(self._h, B, Hf, Wf)
        maps = torch.empty(nbytes, dtype=torch.uint8, device=feat_nchw.device)
        _lib.check(L.njf_hoist_features(self._h, feat_nchw.data_ptr(), B, Hf, Wf, maps.data_ptr(), stream_ptr()))
        return maps

    def hoist_nhwc16(self, feat_nhwc: torch.Tensor, view0: int = 0, n_views_total: Optional[int] = None,
                     maps: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B,Hf,Wf,512) fp16 NHWC encoder output (EncoderResnet.forward_nhwc_half) -> hoisted maps."""
        L = _declare()
        assert feat_nhwc.is_cuda and feat_nhwc.dtype == torch.float16
        self._check_device(feat_nhwc, "feature map")
        feat_nhwc = feat_nhwc.contiguous()
        Bl, Hf, Wf, C = feat_nhwc.shape
        if C != 512:
            raise _lib.NjfError(f"feature map has {C} channels, kernels are built for 512")
        n_views_total = Bl if n_views_total is None else int(n_views_total)
        if maps is None:
            maps = torch.empty(L.njf_hoisted_bytes(self._h, n_views_total, Hf, Wf), dtype=torch.uint8, device=feat_nhwc.device)
        _lib.check(L.njf_hoist_features_nhwc16(self._h, feat_nhwc.data_ptr(), Bl, int(view0), n_views_total, Hf, Wf,
                                               maps.data_ptr(), stream_ptr()))
        return maps

    def hoist_views(self, feat_nchw: torch.Tensor, view0: int, n_views_total: int,
                    maps: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Hoist the given views into slots [view0, view0 + B_local) of a map buffer laid out for ``n_views_total``
        views (allocated here if ``maps`` is None): what a rank of a ray-sharded multi-view call does for the views
        its ray range touches."""
        L = _declare()
        assert feat_nchw.is_cuda and feat_nchw.dtype == torch.float32
        self._check_device(feat_nchw, "feature map")
        feat_nchw = feat_nchw.contiguous()
        Bl, C, Hf, Wf = feat_nchw.shape
        if C != 512:
            raise _lib.NjfError(f"feature map has {C} channels, kernels are built for 512")
        if maps is None:
            maps = torch.empty(L.njf_hoisted_bytes(self._h, n_views_total, Hf, Wf), dtype=torch.uint8, device=feat_nchw.device)
        _lib.check(L.njf_hoist_features_views(self._h, feat_nchw.data_ptr(), Bl, int(view0), int(n_views_total), Hf, Wf,
                                              maps.data_ptr(), stream_ptr()))
        return maps

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().njf_field_destroy(self._h)
                self._h = None
        except Exception:
            pass


def make_cameras(ctxt_c2w, ctxt_k, trgt_c2w, trgt_k_px, device):
    """The reference inverts the 4x4 poses with torch.inverse inside the path (geometry.py:59-65).
    Host inputs: inverted on the host (fp32, CPU LAPACK like the CPU reference); the host copies also let the
    kernels carry the per-view constants in their parameter block.  CUDA inputs: inverted on the device by
    njf_invert_poses -- no host round trip, no synchronisation (what a CUDA-graphed frame uses)."""
    if ctxt_c2w.is_cuda:
        L = _declare()
        f = lambda t: t.detach().to(device, torch.float32).contiguous()
        cc, ck = f(ctxt_c2w), f(ctxt_k)
        cw = torch.empty_like(cc)
        _lib.check(L.njf_invert_poses(dptr(cc), dptr(cw), cc.shape[0], stream_ptr()))
        tw = tk = None
        tc = None
        if trgt_c2w is not None:
            tc = f(trgt_c2w)
            tw = torch.empty_like(tc)
            _lib.check(L.njf_invert_poses(dptr(tc), dptr(tw), tc.shape[0], stream_ptr()))
            tk = f(trgt_k_px)
        cams = NjfCameras(dptr(cw), dptr(ck), dptr(tw), dptr(tk), None, None)
        return cams, (cw, ck, tw, tk, cc, tc)
    f = lambda t: t.detach().to("cpu", torch.float32)
    cw_h = torch.inverse(f(ctxt_c2w)).contiguous()
    ck_h = f(ctxt_k).contiguous()
    cw = cw_h.to(device, non_blocking=True)
    tw = torch.inverse(f(trgt_c2w)).contiguous().to(device, non_blocking=True) if trgt_c2w is not None else None
    ck = ck_h.to(device, non_blocking=True)
    tk = f(trgt_k_px).contiguous().to(device, non_blocking=True) if trgt_k_px is not None else None
    cams = NjfCameras(dptr(cw), dptr(ck), dptr(tw), dptr(tk), cw_h.data_ptr(), ck_h.data_ptr())
    return cams, (cw, ck, tw, tk, cw_h, ck_h)


def query_points(fld: "Field", w2c, k_norm, maps, Hf, Wf, points, want_jac=True, dirs=None):
    """njf_query_points: density head (+ Jacobian head, + colour head when view directions are given) at explicit
    world points (B,N,3).  Returns (sigma, geo, jac) or, with ``dirs``, (sigma, geo, jac, rgb)."""
    L = _declare()
    B, N = points.shape[:2]
    A = fld.action_dim
    o = dict(device=points.device, dtype=torch.float32)
    sigma, geo = torch.empty(B, N, 1, **o), torch.empty(B, N, 15, **o)
    jac = torch.empty(B, N, 3 * A, **o) if want_jac else None
    rgb = torch.empty(B, N, 3, **o) if dirs is not None else None
    ws = fld.workspace(int(L.njf_query_workspace_bytes(fld.handle, B, N)))
    _lib.check(L.njf_query_points(fld.handle, dptr(w2c), dptr(k_norm), dptr(maps), int(Hf), int(Wf), dptr(points),
                                  dptr(dirs), B, N, dptr(sigma), dptr(geo), dptr(jac), dptr(rgb), ws.data_ptr(), ws.numel(),
                                  stream_ptr()))
    return (sigma, geo, jac) if dirs is None else (sigma, geo, jac, rgb)


def query_proposal_density(fld: "Field", level, w2c, k_norm, maps, Hf, Wf, points):
    """njf_query_proposal_density: proposal network ``level`` at explicit world points (B,N,3) -> (B,N,1)."""
    L = _declare()
    B, N = points.shape[:2]
    sigma = torch.empty(B, N, 1, device=points.device, dtype=torch.float32)
    _lib.check(L.njf_query_proposal_density(fld.handle, int(level), dptr(w2c), dptr(k_norm), dptr(maps), int(Hf), int(Wf),
                                            dptr(points), B, N, dptr(sigma), stream_ptr()))
    return sigma


def eval_tables(s_prop: Sequence[int], s_nerf: int, device):
    """Eval-mode sampling tables, computed with the same torch calls as the reference so that
    they are bit-identical: level-0 bins (ray_samplers.py:214) and the PDF positions u (:404-407)."""
    bins0 = torch.linspace(0.0, 1.0, s_prop[0] + 1).to(device)
    us = []
    for lvl in range(len(s_prop)):
        n = s_prop[lvl + 1] if lvl + 1 < len(s_prop) else s_nerf
        nb = n + 1
        u = torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb) + 1.0 / (2 * nb)
        us.append(u.contiguous().to(device))
    return bins0, us


def pdf_sample(weights, bins_in, u, n_out, anneal=1.0, sum_vec_width=None, want_inds=True):
    """PDFSampler on the GPU. weights (N,S); bins_in (S+1,) or (N,S+1); u (n_out+1,) or (N,n_out+1)."""
    L = _declare()
    N, S = weights.shape
    weights = weights.contiguous()
    bins_in, u = bins_in.contiguous(), u.contiguous()
    bs = 0 if bins_in.dim() == 1 else S + 1
    us = 0 if u.dim() == 1 else n_out + 1
    out = torch.empty(N, n_out + 1, device=weights.device, dtype=torch.float32)
    inds = torch.empty(N, n_out + 1, device=weights.device, dtype=torch.int32) if want_inds else None
    sv = default_sum_vec_width() if sum_vec_width is None else sum_vec_width
    _lib.check(L.njf_pdf_sample(dptr(weights), dptr(bins_in), bs, dptr(u), us, N, S, n_out, float(anneal), sv,
                                dptr(out), dptr(inds), stream_ptr()))
    return out, inds


def transmittance_weights(deltas, sigma):
    L = _declare()
    N, S = deltas.shape
    out = torch.empty_like(deltas)
    _lib.check(L.njf_transmittance_weights(dptr(deltas.contiguous()), dptr(sigma.contiguous()), N, S, dptr(out),
                                           stream_ptr()))
    return out
