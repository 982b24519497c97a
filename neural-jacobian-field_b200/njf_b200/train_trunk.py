"""Training of the ResnetFC trunks: the perception phase of ``ModelWrapper.training_step``
(models/model_wrapper.py:116-146 -- rgb / depth / interlevel / distortion losses; gradients into the density head, the
colour head, the proposal networks and, through the feature map, the encoder) and the action phase of the MLP Jacobian
head (model_wrapper.py:75-85, 148-163 with ``jacobian_mlp``).

The fused tcgen05 render kernels keep no activations, so a training step evaluates the trunks layer by layer with the
kernels of ``csrc/trunk_train.cu`` (fp32 SIMT, or tcgen05 kind::tf32 when torch's float32 matmul precision is not
"highest" -- see ``tensor_cores``; activations stay in HBM tensors) and ``torch.autograd`` does the book-keeping
between them, the way it chains ``nn.Linear`` calls in the reference (model_components/resnet_fc.py:70-79, 130-154):

* ``_Linear``      y = relu?(x) W^T + b (+ residual)      njf_train_linear / njf_train_linear_wgrad
* ``_GatherMaps``  z_b = bilinear taps of lin_z map b     njf_train_gather / njf_train_scatter (adjoint, atomics)
* ``sample_setup`` world point -> NeRFEncoding + taps     njf_train_sample_setup (no gradient: cameras are data)
* ``_TruncExp``    trunc_exp (model_components/activations.py:13-38, clamped backward)

``lin_z`` is hoisted onto the feature map exactly as at inference -- lin_z(bilinear(f)) = bilinear(lin_z(f)) -- the
per-pixel maps being ONE plain library GEMM over the NHWC encoder output per trunk (differentiable: its backward
produces d lin_z and the gradient that flows on into the encoder).  Sample placement (stratified jitter, PDF
resampling on the detached proposal weights: ray_samplers.py:219-233, 351-451) runs on the sampler kernels; transmittance
weights, compositing and the flow projection are the reference's own few elementwise torch ops on (B,R,S) tensors.
When only ``decoder.jacobian_head`` (MLP head) is trainable, the frozen trunks render on the fused kernels and only the
Jacobian trunk runs here (``_forward_train_mlp_head``).

There is no CPU path: every entry point needs the CUDA library.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from . import _lib, api


def _declare():
    L = api._declare()
    if getattr(L, "_njf_trunk_train_declared", False):
        return L
    c_int, c_void_p = ctypes.c_int, ctypes.c_void_p
    L.njf_train_sample_setup.restype = c_int
    L.njf_train_sample_setup.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                         c_void_p, c_void_p]
    for fn in (L.njf_train_gather, L.njf_train_scatter):
        fn.restype = c_int
        fn.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    L.njf_train_linear.restype = c_int
    L.njf_train_linear.argtypes = [c_void_p] * 6 + [c_int] * 6 + [c_void_p]
    L.njf_train_linear_wgrad.restype = c_int
    L.njf_train_linear_wgrad.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                         c_void_p]
    L.njf_train_sh16.restype = c_int
    L.njf_train_sh16.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
    L._njf_trunk_train_declared = True
    return L


def _f32c(t: Tensor) -> Tensor:
    return t.contiguous().float()


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def tensor_cores() -> bool:
    """The trunk GEMMs follow torch's own float32 matmul switch: "highest" (torch's default) -> fp32 SIMT kernels,
    "high" / "medium" (what the reference's train.py:64-65 selects) -> tcgen05 kind::tf32."""
    return torch.get_float32_matmul_precision() != "highest"


# ----------------------------------------------------------------------------- kernels as autograd functions
class _Linear(torch.autograd.Function):
    """y[M,N] = act(x[M,K]) . w[N,K]^T + b (+ residual), act = ReLU when relu_in (the reference applies the activation
    to a layer's INPUT: resnet_fc.py:70-79).  N, K multiples of 4 in [4,128]."""

    @staticmethod
    def forward(ctx, x, w, b, relu_in, residual):
        L = _declare()
        x, w = _f32c(x), _f32c(w)
        b = None if b is None else _f32c(b)
        residual = None if residual is None else _f32c(residual)
        M, K = x.shape
        N = w.shape[0]
        y = torch.empty(M, N, device=x.device, dtype=torch.float32)
        _lib.check(L.njf_train_linear(api.dptr(x), api.dptr(w), api.dptr(b), api.dptr(residual), None, api.dptr(y),
                                      M, N, K, 1, int(relu_in), int(tensor_cores()), api.stream_ptr()))
        ctx.save_for_backward(x, w)
        ctx.relu_in, ctx.has_b = bool(relu_in), b is not None
        return y

    @staticmethod
    def backward(ctx, g):
        L = _declare()
        x, w = ctx.saved_tensors
        g = _f32c(g)
        M, K = x.shape
        N = w.shape[0]
        st = api.stream_ptr()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty(M, K, device=x.device, dtype=torch.float32)
            _lib.check(L.njf_train_linear(api.dptr(g), api.dptr(w), None, None, api.dptr(x) if ctx.relu_in else None,
                                          api.dptr(gx), M, K, N, 0, 0, int(tensor_cores()), st))
        if ctx.needs_input_grad[1] or (ctx.has_b and ctx.needs_input_grad[2]):
            gw = torch.zeros(N, K, device=x.device, dtype=torch.float32)
            gb = torch.zeros(N, device=x.device, dtype=torch.float32) if ctx.has_b else None
            _lib.check(L.njf_train_linear_wgrad(api.dptr(g), api.dptr(x), M, N, K, int(ctx.relu_in), api.dptr(gw),
                                                api.dptr(gb), int(tensor_cores()), st))
        return gx, gw, gb, None, (g if ctx.needs_input_grad[4] else None)


class _GatherMaps(torch.autograd.Function):
    """z_b[m, c] = sum_t tap_w[m, t] maps[tap_pix[m, t], 128 b + c] for every 128-channel group b of the maps (one
    contiguous (M,128) tensor per lin_z layer); backward scatters the groups' gradients into one zeroed map-shaped
    gradient."""

    @staticmethod
    def forward(ctx, maps, tap_pix, tap_w):
        L = _declare()
        maps = _f32c(maps)
        CH = maps.shape[-1]
        M = tap_pix.shape[0]
        outs = []
        for b in range(CH // 128):
            out = torch.empty(M, 128, device=maps.device, dtype=torch.float32)
            _lib.check(L.njf_train_gather(api.dptr(maps), api.dptr(tap_pix), api.dptr(tap_w), M, CH, 128 * b, 128,
                                          api.dptr(out), api.stream_ptr()))
            outs.append(out)
        ctx.save_for_backward(tap_pix, tap_w)
        ctx.map_shape = tuple(maps.shape)
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        L = _declare()
        tap_pix, tap_w = ctx.saved_tensors
        CH = ctx.map_shape[-1]
        dmap = torch.zeros(ctx.map_shape, device=tap_w.device, dtype=torch.float32)
        for b, g in enumerate(gs):
            if g is None:
                continue
            g = _f32c(g)
            _lib.check(L.njf_train_scatter(api.dptr(g), api.dptr(tap_pix), api.dptr(tap_w), g.shape[0], CH, 128 * b, 128,
                                           api.dptr(dmap), api.stream_ptr()))
        return dmap, None, None


class _TruncExp(torch.autograd.Function):
    """model_components/activations.py:13-38."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        return g * torch.exp(torch.clamp(ctx.saved_tensors[0], min=-15, max=15))


def sample_setup(w2c: Tensor, k_norm: Tensor, points: Tensor, Hf: int, Wf: int) -> Tuple[Tensor, Tensor, Tensor]:
    """points (B,N,3) world space -> enc (B*N,64), tap_pix (B*N,4) int32, tap_w (B*N,4)."""
    L = _declare()
    B, N = points.shape[:2]
    pts = _f32c(points.detach())
    dev = pts.device
    enc = torch.empty(B * N, 64, device=dev, dtype=torch.float32)
    pix = torch.empty(B * N, 4, device=dev, dtype=torch.int32)
    tw = torch.empty(B * N, 4, device=dev, dtype=torch.float32)
    _lib.check(L.njf_train_sample_setup(api.dptr(w2c), api.dptr(k_norm), api.dptr(pts), B, N, int(Hf), int(Wf),
                                        api.dptr(enc), api.dptr(pix), api.dptr(tw), api.stream_ptr()))
    return enc, pix, tw


def sh16(dirs: Tensor, convention: str, fp16_round: bool) -> Tensor:
    """SH degree 4 of unit view directions (M,3) -> (M,16) (action_decoder_jacobian.py:24-30, 284)."""
    L = _declare()
    d = _f32c(dirs.detach())
    out = torch.empty(d.shape[0], 16, device=d.device, dtype=torch.float32)
    conv = api.SH_CONVENTIONS[convention]
    _lib.check(L.njf_train_sh16(api.dptr(d), d.shape[0], conv, int(bool(fp16_round)), api.dptr(out), api.stream_ptr()))
    return out


# ----------------------------------------------------------------------------- trunks
def lin_z_maps(trunk, feat_nhwc: Tensor) -> Tensor:
    """The three lin_z layers of a ResnetFC applied per PIXEL of the encoder output (B,Hf,Wf,512) -> (B,Hf,Wf,384):
    one plain GEMM; bilinear interpolation then commutes with the affine map (tap weights sum to one)."""
    w = torch.cat([l.weight for l in trunk.lin_z], 0)
    b = torch.cat([l.bias for l in trunk.lin_z], 0)
    return F.linear(feat_nhwc, w, b)


def linear(x: Tensor, w: Tensor, b: Optional[Tensor], relu_in: bool, residual: Optional[Tensor] = None) -> Tensor:
    """A Linear layer of arbitrary (<= 128) width on the training kernels: zero-pads both sizes to multiples of 4."""
    N, K = w.shape
    Np, Kp = _pad4(N), _pad4(K)
    if x.shape[1] != Kp:
        x = F.pad(x, (0, Kp - x.shape[1]))
    if (Np, Kp) != (N, K):
        w = F.pad(w, (0, Kp - K, 0, Np - N))
        b = None if b is None else F.pad(b, (0, Np - N))
    y = _Linear.apply(x, w, b, relu_in, residual)
    return y if Np == N else y[:, :N]


def resnet_fc(trunk, enc: Tensor, z: Tensor) -> Tensor:
    """ResnetFC.forward (model_components/resnet_fc.py:130-154; block :70-79; beta = 0 -> ReLU) on M sample rows:
    enc (M,64) positional encoding (63 + zero column), z = one (M,128) gathered lin_z map per combine layer."""
    x = linear(enc, trunk.lin_in.weight, trunk.lin_in.bias, False)
    for b, blk in enumerate(trunk.blocks):
        if b < len(trunk.lin_z):
            x = x + z[b]
        net = linear(x, blk.fc_0.weight, blk.fc_0.bias, True)
        x = linear(net, blk.fc_1.weight, blk.fc_1.bias, True, residual=x)
    return linear(x, trunk.lin_out.weight, trunk.lin_out.bias, True)


def transmittance_weights(deltas: Tensor, densities: Tensor) -> Tensor:
    """RaySamples.get_weights (rendering/ray_samplers.py:77-101); (..., S, 1) tensors."""
    delta_density = torch.where(deltas > 0, deltas * densities, torch.zeros_like(densities))
    alphas = 1 - torch.exp(-delta_density)
    t = torch.cumsum(delta_density[..., :-1, :], dim=-2)
    t = torch.cat([torch.zeros((*t.shape[:-2], 1, 1), device=densities.device), t], dim=-2)
    return alphas * torch.exp(-t)


def project(points: Tensor, w2c: Tensor, k: Tensor) -> Tensor:
    """project_world_coords_to_camera (rendering/geometry.py:206-215, deprecated_project :137-154): (B,N,3) -> (B,N,2)."""
    cam = torch.einsum("bij,bnj->bni", w2c[:, :3, :3], points) + w2c[:, None, :3, 3]
    xyw = torch.einsum("bij,bnj->bni", k, cam)
    return (xyw / (xyw[..., -1:] + 1e-9))[..., :2]


# ----------------------------------------------------------------------------- Model.forward in train mode
def forward_train(model, camera_input, rendering_input, robot_input, compute_vis_features: bool):
    """Model.forward (models/model.py:316-396) with autograd through every ResnetFC trunk, the colour head and the
    encoder.  Returns what njf_b200.Model._forward_train assembles into a ModelOutput:
    dict(rgb, depth, flow, jbar, steps, weights, p, pw, weights_list, bins_list)."""
    from .train import train_tables

    dev = model._device()
    r = model.cfg.rendering
    s_prop, s_nerf = tuple(r.num_proposal_samples), int(r.num_nerf_samples)
    mv = lambda t: t.detach().to(dev, torch.float32).contiguous()
    o, d = mv(rendering_input.origins), mv(rendering_input.directions)
    zn, zf = mv(rendering_input.z_near), mv(rendering_input.z_far)
    B, R = o.shape[:2]
    action = robot_input.robot_action.to(dev, torch.float32)
    head, A = model._head_and_dim()
    arm = model._mode() == "arm"

    if (head == "jacobian_mlp" and not arm and torch.is_grad_enabled()
            and not any(p.requires_grad for n, p in model.named_parameters() if not n.startswith("decoder.jacobian_head."))):
        return _forward_train_mlp_head(model, camera_input, robot_input, o, d, zn, zf, action, A, s_prop, s_nerf,
                                       compute_vis_features)

    feats = model.encoder.forward(camera_input.input_image.to(dev)).float()       # (B,512,Hf,Wf), autograd as configured
    Hf, Wf = feats.shape[-2:]
    fmap = feats.permute(0, 2, 3, 1).contiguous()                                  # NHWC view of channels-last storage
    cams, keep = api.make_cameras(camera_input.ctxt_extrinsics, camera_input.ctxt_intrinsics,
                                  camera_input.trgt_extrinsics, camera_input.trgt_intrinsics, dev)
    cw, ck, tw, tk = keep[:4]
    bins, us = train_tables(model, s_prop, s_nerf, B, R, dev)
    near, far = zn[:, None, None], zf[:, None, None]
    euclid = lambda b: b * far + (1 - b) * near                                    # ray_samplers.py:242-245

    def samples(b):
        e = euclid(b)
        starts, ends = e[..., :-1, None], e[..., 1:, None]
        pos = o[..., None, :] + d[..., None, :] * (starts + ends) / 2            # RaySamples.get_positions :48-55
        S = pos.shape[2]
        enc, pix, tapw = sample_setup(cw, ck, pos.reshape(B, R * S, 3), Hf, Wf)
        return starts, ends, pos, enc, pix, tapw

    # ProposalNetworkSampler.generate_ray_samples (ray_samplers.py:497-552)
    updated = model._steps_since_update > model._update_schedule(model._step) or model._step < 10
    weights_list: List[Tensor] = []
    bins_list: List[Tensor] = []
    for lvl, net in enumerate(model.proposal_networks):
        starts, ends, pos, enc, pix, tapw = samples(bins)
        with torch.set_grad_enabled(torch.is_grad_enabled() and updated):
            trunk = net.density_head
            z = _GatherMaps.apply(lin_z_maps(trunk, fmap).reshape(-1, 384), pix, tapw)
            sigma = _TruncExp.apply(resnet_fc(trunk, enc, z) - 1.0).reshape(B, R, -1, 1)
        w = transmittance_weights(ends - starts, sigma)
        weights_list.append(w)
        bins_list.append(bins)
        n = s_prop[lvl + 1] if lvl + 1 < len(s_prop) else s_nerf
        S = w.shape[2]
        nb, _ = api.pdf_sample(w.detach().reshape(B * R, S), bins.reshape(B * R, S + 1), us[lvl].reshape(B * R, n + 1), n,
                               anneal=model._anneal, want_inds=False)
        bins = nb.reshape(B, R, n + 1)
    if updated:
        model._steps_since_update = 0

    # decoder.forward (action_decoder_jacobian.py:147-215)
    starts, ends, pos, enc, pix, tapw = samples(bins)
    dec = model.decoder
    z = _GatherMaps.apply(lin_z_maps(dec.density_head, fmap).reshape(-1, 384), pix, tapw)
    out = resnet_fc(dec.density_head, enc, z)                                     # (M,16): 15 geometry features + density
    geo, pre = out[:, :15], out[:, 15:16]
    sigma = _TruncExp.apply(pre - 1.0).reshape(B, R, s_nerf, 1)
    dirs_s = d[..., None, :].expand(B, R, s_nerf, 3).reshape(-1, 3)
    sh = sh16(dirs_s, model.sh_convention, model.sh_fp16_round)
    ch = dec.color_head
    h = linear(torch.cat([geo, sh], -1), ch[0].weight, ch[0].bias, False)
    h = linear(h, ch[2].weight, ch[2].bias, True)
    rgb_s = torch.sigmoid(linear(h, ch[4].weight, ch[4].bias, True)).reshape(B, R, s_nerf, 3)
    if head == "jacobian_mlp":       # compute_jacobian (:324-337): a second ResnetFC on the same (features, encoding)
        jt = dec.jacobian_head_arm if arm else dec.jacobian_head
        zj = _GatherMaps.apply(lin_z_maps(jt, fmap).reshape(-1, 384), pix, tapw)
        jac = resnet_fc(jt, enc, zj).reshape(B, R, s_nerf, 3 * A)
    else:
        # the cross-attention head carries no gradient on this path (the perception losses do not see it; its action
        # phase trains through train._RenderJacobianHead): evaluated by the fused query kernel
        with torch.no_grad():
            fld = model._field_for_head_queries()   # trunks may be stale: the Jacobian head does not read them
            maps16 = fld.hoist(feats.detach().contiguous())
            _, _, jac = api.query_points(fld, cw, ck, maps16, Hf, Wf, pos.reshape(B, R * s_nerf, 3).contiguous())
        jac = jac.reshape(B, R, s_nerf, 3 * A)
    weights = transmittance_weights(ends - starts, sigma)                         # model.py:351
    weights_list.append(weights)
    bins_list.append(bins)
    # render_rgb / render_depth / render_action_features / render_optical_flow (model.py:257-314)
    steps = (starts + ends) / 2
    depth = torch.sum(weights * steps, dim=-2) / (torch.sum(weights, -2) + 1e-10)
    depth = torch.clip(depth, steps.min(), steps.max())
    flow_s = torch.einsum("brsad,ba->brsd", jac.reshape(B, R, s_nerf, A, 3), action)   # compute_flow :128-145
    p = torch.sum(weights * pos, dim=-2)
    pw = torch.sum(weights * (pos + flow_s), dim=-2)
    flow = project(pw, tw, tk) - project(p, tw, tk)
    return dict(rgb=torch.sum(weights * rgb_s, dim=-2), depth=depth, flow=flow,
                jbar=torch.sum(weights * jac, dim=-2) if compute_vis_features else None,
                steps=steps.squeeze(-1), weights=weights.squeeze(-1), p=p, pw=pw,
                weights_list=weights_list, bins_list=bins_list, near=near, far=far)


def _forward_train_mlp_head(model, camera_input, robot_input, o, d, zn, zf, action, A, s_prop, s_nerf, compute_vis_features):
    """Action phase of the MLP Jacobian head (models/model_wrapper.py:75-85, 148-163): everything but
    ``decoder.jacobian_head`` is frozen, so sample placement, densities, weights and colours carry no gradient and come
    from the FUSED render (same kernels as inference, jittered tables); only the Jacobian trunk runs layer by layer under
    autograd, at the fused render's final sample positions."""
    from .render import render
    from .train import train_tables

    dev = o.device
    r = model.cfg.rendering
    B, R = o.shape[:2]
    with torch.no_grad():
        feats = model.encoder.forward(camera_input.input_image.to(dev)).float().contiguous()
        Hf, Wf = feats.shape[-2:]
        fld = model._field_with_current_trunks()
        maps16 = fld.hoist(feats)
        cams, keep = api.make_cameras(camera_input.ctxt_extrinsics, camera_input.ctxt_intrinsics,
                                      camera_input.trgt_extrinsics, camera_input.trgt_intrinsics, dev)
        cw, ck, tw, tk = keep[:4]
        bins0, us = train_tables(model, s_prop, s_nerf, B, R, dev)
        res = render(fld, maps16, Hf, Wf, cams, o, d, zn, zf, action.detach().contiguous(), s_prop, s_nerf, vis=True,
                     sampler_outputs=True, bins0=bins0, us=us, anneal=model._anneal)
        res._cams = keep
        near, far = zn[:, None, None], zf[:, None, None]
        fb = res.level_bins[-1]
        e = fb * far + (1 - fb) * near
        starts, ends = e[..., :-1, None], e[..., 1:, None]
        pos = o[..., None, :] + d[..., None, :] * (starts + ends) / 2
        enc, pix, tapw = sample_setup(cw, ck, pos.reshape(B, R * s_nerf, 3), Hf, Wf)
        fmap = feats.permute(0, 2, 3, 1).contiguous()
        weights = res.weights[..., None]
    jt = model.decoder.jacobian_head
    zj = _GatherMaps.apply(lin_z_maps(jt, fmap).reshape(-1, 384), pix, tapw)     # d lin_z through the map GEMM's backward
    jac = resnet_fc(jt, enc, zj).reshape(B, R, s_nerf, 3 * A)
    flow_s = torch.einsum("brsad,ba->brsd", jac.reshape(B, R, s_nerf, A, 3), action)
    p = torch.sum(weights * pos, dim=-2)
    pw = torch.sum(weights * (pos + flow_s), dim=-2)
    flow = project(pw, tw, tk) - project(p, tw, tk)
    if model._steps_since_update > model._update_schedule(model._step) or model._step < 10:
        model._steps_since_update = 0
    return dict(rgb=res.rgb, depth=res.depth, flow=flow,
                jbar=torch.sum(weights * jac, dim=-2) if compute_vis_features else None,
                steps=res.steps, weights=res.weights, p=p, pw=pw,
                weights_list=[w[..., None] for w in res.prop_weights] + [weights],
                bins_list=[bins0] + list(res.level_bins), near=near, far=far)
