"""Action-phase training through the fused render path.

The reference's second training phase (``dataset.mode == "action"``) freezes every parameter except the Jacobian head
(models/model_wrapper.py:75-85, models/decoder/action_decoder_jacobian.py:251-258) and minimises a masked MSE on the
rendered optical flow (model_wrapper.py:148-163).  Sample placement and the transmittance weights then carry no
gradient, and the whole backward pass is

    g_flow -> g_pw -> g_Jbar = u (x) g_pw -> g_J_s = w_s g_Jbar -> cross-attention head -> q0 -> W_q, b_q .

Forward = the same fused kernels as inference (stratified jitter comes in as tensors); backward = the kernels in
``csrc/xf_backward.cu``.  Python only chains the gradients of the FOLDED head matrices (what the kernels multiply with)
back to the state-dict parameters through ``fold_head`` -- a few 64x64 products, differentiable torch ops that mirror
``csrc/field.cu`` -- and forms d W_q[:, 63:] = (gathered-gradient map)^T . features with one library GEMM.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import Tensor

from . import _lib, api
from .render import render

TILE_ROWS = 128
XF_TILE_BYTES = 8 * TILE_ROWS * 16 + TILE_ROWS * 4


def _declare():
    L = api._declare()
    if getattr(L, "_njf_train_declared", False):
        return L
    c_int, c_void_p, c_size_t = ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t
    L.njf_xf_folded_floats.restype = c_int
    L.njf_xf_folded_floats.argtypes = []
    L.njf_xf_backward_workspace_bytes.restype = c_size_t
    L.njf_xf_backward_workspace_bytes.argtypes = [c_int]
    L.njf_xf_backward.restype = c_int
    L.njf_xf_backward.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_void_p]
    L.njf_query_backward.restype = c_int
    L.njf_query_backward.argtypes = [ctypes.POINTER(api.NjfCameras), ctypes.POINTER(api.NjfRenderArgs), c_void_p, c_int,
                                     c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    L.njf_flow_backward.restype = c_int
    L.njf_flow_backward.argtypes = [c_void_p] * 7 + [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
    L._njf_train_declared = True
    return L


def tile_count(n_rays: int, s_nerf: int) -> int:
    """128-row tiles of a field pass (csrc/render.cu make_geom)."""
    g = TILE_ROWS // s_nerf if s_nerf <= TILE_ROWS else 1
    t = 1 if s_nerf <= TILE_ROWS else -(-s_nerf // TILE_ROWS)
    return -(-n_rays // g) * t


def fold_head(decoder, action_dim: int) -> Tensor:
    """The matrices the cross-attention kernels multiply with, from the reference's parameters (differentiable):
    keys / values of the index embedding folded into M1 = scale W_q^T K and M2 = W_out V, PreNorm's LayerNorm affine
    folded into the following linear map (transformer.py:14-21, 63-82; csrc/field.cu does the same in fp64 for the
    forward).  Packed as njf_xf_backward expects; natural-base softmax (no log2 e factor)."""
    A = action_dim
    emb = decoder.jacobian_index_embedding[0]                      # (A, 64)
    parts: List[Tensor] = []
    for layer in decoder.jacobian_attn_decoder.layers:
        attn, ff = layer[0], layer[1]
        g1, be1 = attn.norm.weight, attn.norm.bias
        wq, wkv = attn.fn.to_q.weight, attn.fn.to_kv.weight        # (512,64), (1024,64)
        wo, bo = attn.fn.to_out[0].weight, attn.fn.to_out[0].bias   # (64,512), (64,)
        K = (emb @ wkv[:512].t()).view(A, 8, 64)                    # (a, h, d)
        V = (emb @ wkv[512:].t()).view(A, 8, 64)
        T = torch.einsum("hdk,ahd->hak", wq.view(8, 64, 64), K) * (64 ** -0.5)   # logits = T . LN(x)
        m1 = F.pad(T * g1, (0, 0, 0, 8 - A)).reshape(64, 64)        # rows h*8 + a
        m1b = F.pad(T @ be1, (0, 8 - A)).reshape(64)
        m2 = F.pad(torch.einsum("ohd,ahd->oha", wo.view(64, 8, 64), V), (0, 8 - A)).reshape(64, 64)
        g2, be2 = ff.norm.weight, ff.norm.bias
        w1, b1 = ff.fn.net[0].weight, ff.fn.net[0].bias
        w2, b2 = ff.fn.net[3].weight, ff.fn.net[3].bias
        parts += [m1.reshape(-1), m1b, m2.reshape(-1), bo, (w1 * g2).reshape(-1), b1 + w1 @ be2, w2.reshape(-1), b2]
    wh, bh = decoder.jacobian_head.weight, decoder.jacobian_head.bias   # (3A,64), (3A,)
    parts += [F.pad(wh, (0, 0, 0, 64 - 3 * A)).reshape(-1), F.pad(bh, (0, 64 - 3 * A))]
    return torch.cat([p.float() for p in parts])


class _RenderJacobianHead(torch.autograd.Function):
    """Fused render forward; backward w.r.t. the folded head matrices, jacobian_query_mlp and the action."""

    @staticmethod
    def forward(ctx, folded, wq_weight, wq_bias, action, fld, maps, feats, cams, keep, origins, dirs, z_near, z_far,
                s_prop, s_nerf, bins0, us, anneal, holder):
        B, R = origins.shape[:2]
        n_tiles = tile_count(B * R, s_nerf)
        # a workspace of its own: the hand-over (q0 of every sample + sample weights) must survive until backward,
        # and njf_xf_backward needs it in one launch group
        need = max(fld.workspace_bytes(B, R, s_prop, s_nerf), n_tiles * XF_TILE_BYTES,
                   B * R * max(s_prop) * 4)
        ws = torch.empty(need + 256, dtype=torch.uint8, device=origins.device)
        res = render(fld, maps, feats.shape[-2], feats.shape[-1], cams, origins, dirs, z_near, z_far, action.detach(),
                     s_prop, s_nerf, vis=True, sampler_outputs=True, bins0=bins0, us=us, anneal=anneal, workspace=ws)
        if holder is not None:
            holder.append(res)   # the caller reads the non-differentiable outputs (weights, bins, ...) from here
        ctx.fld, ctx.cams, ctx.keep, ctx.res, ctx.ws = fld, cams, keep, res, ws
        ctx.geom = (B, R, tuple(s_prop), int(s_nerf), n_tiles, feats.shape[-2], feats.shape[-1])
        ctx.save_for_backward(folded.detach(), action.detach(), feats, origins, dirs, z_near, z_far)
        ctx.mark_non_differentiable(res.rgb, res.depth, res.p)
        ctx.set_materialize_grads(False)
        return res.flow, res.pw, res.jbar, res.rgb, res.depth, res.p

    @staticmethod
    def backward(ctx, g_flow, g_pw, g_jbar, *_):
        L = _declare()
        folded, action, feats, origins, dirs, z_near, z_far = ctx.saved_tensors
        fld, cams, res = ctx.fld, ctx.cams, ctx.res
        B, R, s_prop, s_nerf, n_tiles, Hf, Wf = ctx.geom
        A = fld.action_dim
        dev = origins.device
        f32 = dict(device=dev, dtype=torch.float32)
        st = api.stream_ptr()
        c = lambda t: None if t is None else t.contiguous().float()
        g_flow, g_pw, g_jbar = c(g_flow), c(g_pw), c(g_jbar)
        if g_flow is None:
            g_flow = torch.zeros(B, R, 2, **f32)
        gj = torch.empty(B, R, 3 * A, **f32)
        ga = torch.empty(B, A, **f32)
        _lib.check(L.njf_flow_backward(api.dptr(g_flow), api.dptr(g_pw), api.dptr(res.jbar), api.dptr(res.p),
                                       api.dptr(action), cams.trgt_w2c, cams.trgt_k_px, B * R, R, A, api.dptr(gj),
                                       api.dptr(ga), st))
        if g_jbar is not None:
            gj = gj + g_jbar
        g_folded = torch.zeros_like(folded)
        g_q0 = torch.empty(n_tiles * TILE_ROWS, 64, **f32)
        bws = torch.empty(int(L.njf_xf_backward_workspace_bytes(n_tiles)), dtype=torch.uint8, device=dev)
        _lib.check(L.njf_xf_backward(api.dptr(folded), A, ctx.ws.data_ptr(), n_tiles, B * R, s_nerf, api.dptr(gj),
                                     api.dptr(g_folded), api.dptr(g_q0), bws.data_ptr(), bws.numel(), st))
        a = api.NjfRenderArgs()
        a.B, a.R, a.s_nerf, a.Hf, a.Wf = B, R, s_nerf, Hf, Wf
        a.origins, a.dirs, a.z_near, a.z_far = api.dptr(origins), api.dptr(dirs), api.dptr(z_near), api.dptr(z_far)
        fb = res.level_bins[-1]
        g_wq_enc = torch.zeros(64, 64, **f32)
        g_bq = torch.zeros(64, **f32)
        g_map = torch.zeros(B, Hf * Wf, 64, **f32)
        _lib.check(L.njf_query_backward(ctypes.byref(cams), ctypes.byref(a), api.dptr(fb), fb.shape[-1], api.dptr(g_q0),
                                        n_tiles, api.dptr(g_wq_enc), api.dptr(g_bq), api.dptr(g_map), st))
        # d W_q[:, 63:] = sum over views and pixels of g_map^T . features -- one plain library GEMM per step
        g_wq_feat = torch.einsum("bpn,bcp->nc", g_map, feats.reshape(B, feats.shape[1], Hf * Wf))
        g_wq = torch.cat([g_wq_enc[:, :63], g_wq_feat], dim=1)
        ctx.ws = ctx.res = None
        return (g_folded, g_wq, g_bq, ga) + (None,) * 15


def stratified_tables(s_prop: Sequence[int], s_nerf: int, B: int, R: int, single_jitter: bool, device,
                      generator: Optional[torch.Generator] = None):
    """Train-mode sampling tables drawn with the reference's own torch calls (rendering/ray_samplers.py:214-233
    level-0 stratified bins, :389-401 jittered PDF positions)."""
    rand = lambda n: torch.rand((B, R, 1 if single_jitter else n), dtype=torch.float32, device=device, generator=generator)
    s0 = s_prop[0]
    bins = torch.linspace(0.0, 1.0, s0 + 1).to(device)[None, ...]
    t_rand = rand(s0 + 1)
    centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
    upper = torch.cat([centers, bins[..., -1:]], -1)
    lower = torch.cat([bins[..., :1], centers], -1)
    bins0 = (lower + (upper - lower) * t_rand).contiguous()
    us = []
    for lvl in range(len(s_prop)):
        n = s_prop[lvl + 1] if lvl + 1 < len(s_prop) else s_nerf
        nb = n + 1
        u = torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb, device=device)
        u = u.expand((B, R, nb))
        us.append((u + rand(n + 1) / nb).contiguous())
    return bins0, us


def train_tables(model, s_prop: Sequence[int], s_nerf: int, B: int, R: int, device):
    """The step's sampling tables: ``model.jitter_tables`` when set (moved to the device), else freshly drawn."""
    jt = getattr(model, "jitter_tables", None)
    if jt is not None:
        bins0, us = jt
        f = lambda t: t.detach().to(device, torch.float32).contiguous()
        return f(bins0), [f(u) for u in us]
    return stratified_tables(s_prop, s_nerf, B, R, model.cfg.rendering.single_jitter, device,
                             generator=getattr(model, "jitter_generator", None))
