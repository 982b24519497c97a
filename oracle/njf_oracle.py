"""TEST INFRASTRUCTURE -- CPU oracle for the neural-jacobian-field volumetric-rendering hot path.

A restatement (torch CPU, fp32) of what the reference computes on the path
ray bundle -> proposal sampling -> pixel-aligned feature gather -> encodings -> sigma / rgb /
Jacobian heads -> alpha compositing, written as plain functions over a flat ``{state-dict key:
tensor}`` weight dict.  Every function cites the reference file:line it follows (paths relative to
``/root/reference/project/neural_jacobian_field``).

PARITY PINNING: the reference has no tests and no checkpoints, so this
oracle is pinned against the REFERENCE ITSELF, imported unmodified in the build container via
``oracle/ref_shim.py``; ``oracle/make_golden.py`` stores its outputs in ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` checks this file against them.  The two third-party encodings
(nerfstudio NeRFEncoding, tiny-cuda-nn SH) are restated from their published algorithms and are
"parity unpinned" (see ref_shim.py header).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference``
legs may import this module -- never the product path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

W = Dict[str, torch.Tensor]


@dataclass
class FieldSpec:
    head: str                 # "jacobian_transformer" | "jacobian_mlp"
    action_dim: int
    n_blocks: int = 5
    combine_layer: int = 3
    geo_dim: int = 15
    heads: int = 8
    dim_head: int = 64
    depth: int = 3
    sh_fp16_round: bool = True
    sh_convention: str = "tcnn"   # or "nerfstudio_torch" (nerfstudio's torch fallback)


# ----------------------------------------------------------------------------- encodings
def posenc(x: torch.Tensor, num_freq: int = 10) -> torch.Tensor:
    """nerfstudio NeRFEncoding(in_dim=3, num_frequencies=10, min 0, max 9, include_input=True)
    (call sites models/decoder/action_decoder_jacobian.py:275-282, density_decoder.py:31-38):
    [sin(2 pi x 2^k) (dim-major, freq-minor) ; sin(. + pi/2) ; x] -> 63 columns."""
    freqs = (2.0 ** torch.linspace(0.0, num_freq - 1.0, num_freq)).to(x.device)
    t = ((2.0 * torch.pi * x)[..., None] * freqs).reshape(*x.shape[:-1], -1)
    return torch.cat([torch.sin(torch.cat([t, t + torch.pi / 2.0], -1)), x], -1)


_SH = (0.28209479177387814, 0.48860251190291987, 1.0925484305920792, 0.94617469575755997,
       0.31539156525251999, 0.54627421529603959, 0.59004358992664352, 2.8906114426405538,
       0.45704579946446572, 0.3731763325901154, 1.4453057213202769)


def sh4(d01: torch.Tensor, fp16_round: bool = True, convention: str = "tcnn") -> torch.Tensor:
    """tiny-cuda-nn SphericalHarmonics degree 4 of (2*d01 - 1), output through fp16
    (SHEncoding(levels=4, implementation="tcnn"), action_decoder_jacobian.py:194-199, :284).
    convention="nerfstudio_torch": what nerfstudio's SHEncoding computes when it falls back to its torch
    implementation (components_from_spherical_harmonics, restated from the published source): the direction tensor
    is used AS PASSED (here the [0,1]-normalised one) and every component carries a positive leading sign."""
    if convention == "nerfstudio_torch":
        x, y, z = d01.unbind(-1)
        xx, yy, zz = x * x, y * y, z * z
        c = _SH
        o = torch.stack([
            torch.full_like(x, c[0]), c[1] * y, c[1] * z, c[1] * x,
            c[2] * x * y, c[2] * y * z, c[3] * zz - c[4], c[2] * x * z, c[5] * (xx - yy),
            c[6] * y * (3 * xx - yy), c[7] * x * y * z, c[8] * y * (5 * zz - 1),
            c[9] * z * (5 * zz - 3), c[8] * x * (5 * zz - 1), c[10] * z * (xx - yy),
            c[6] * x * (xx - 3 * yy)], -1)
        return o.half().float() if fp16_round else o
    d = d01 * 2.0 - 1.0
    x, y, z = d.unbind(-1)
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    c = _SH
    o = torch.stack([
        torch.full_like(x, c[0]), -c[1] * y, c[1] * z, -c[1] * x,
        c[2] * xy, -c[2] * yz, c[3] * z2 - c[4], -c[2] * xz, c[5] * x2 - c[5] * y2,
        c[6] * y * (-3.0 * x2 + y2), c[7] * xy * z, c[8] * y * (1.0 - 5.0 * z2),
        c[9] * z * (5.0 * z2 - 3.0), c[8] * x * (1.0 - 5.0 * z2), c[10] * z * (x2 - y2),
        c[6] * x * (-x2 + 3.0 * y2)], -1)
    return o.half().float() if fp16_round else o


# ----------------------------------------------------------------------------- gather
def project_to_context(xyz: torch.Tensor, c2w: torch.Tensor, k_norm: torch.Tensor):
    """World -> context camera -> normalised image coords.
    model_components/pixel_aligned_features.py:18-21, rendering/geometry.py:59-65, :137-154.
    xyz (B,N,3), c2w (B,4,4), k_norm (B,3,3) -> x_cam (B,N,3), uv (B,N,2)."""
    w2c = torch.inverse(c2w)
    hom = torch.cat([xyz, torch.ones_like(xyz[..., :1])], -1)
    cam = torch.einsum("bij,bnj->bni", w2c, hom)
    xyw = torch.einsum("bij,bnj->bni", k_norm, cam[..., :3])
    uv = xyw / (xyw[..., -1:] + 1e-9)
    return cam[..., :3], uv[..., :2]


def bilinear_border(feat: torch.Tensor, uv: torch.Tensor) -> torch.Tensor:
    """F.grid_sample(feat, (uv-0.5)*2, align_corners=True, padding_mode="border") written out
    (pixel_aligned_features.py:24-34).  feat (B,C,H,W), uv (B,N,2) in [0,1] -> (B,N,C)."""
    B, C, H, Wd = feat.shape
    g = (uv - 0.5) * 2.0
    ix = ((g[..., 0] + 1.0) / 2.0) * (Wd - 1)
    iy = ((g[..., 1] + 1.0) / 2.0) * (H - 1)
    ix = ix.clamp(0.0, float(Wd - 1))
    iy = iy.clamp(0.0, float(H - 1))
    x0, y0 = torch.floor(ix), torch.floor(iy)
    x1, y1 = x0 + 1.0, y0 + 1.0
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    flat = feat.permute(0, 2, 3, 1).reshape(B, H * Wd, C)
    out = torch.zeros(B, uv.shape[1], C, dtype=feat.dtype, device=feat.device)
    for xx, yy, ww in ((x0, y0, w_nw), (x1, y0, w_ne), (x0, y1, w_sw), (x1, y1, w_se)):
        inb = (xx <= Wd - 1) & (yy <= H - 1)  # xx,yy >= 0 after the border clamp
        idx = (yy.clamp(max=H - 1) * Wd + xx.clamp(max=Wd - 1)).long()
        tap = torch.gather(flat, 1, idx[..., None].expand(-1, -1, C))
        out = out + tap * (ww * inb)[..., None]
    return out


def pixel_aligned(xyz, c2w, k_norm, feat):
    """get_pixel_aligned_features (pixel_aligned_features.py:11-35): features + camera-space xyz."""
    cam, uv = project_to_context(xyz, c2w, k_norm)
    return bilinear_border(feat, uv), cam


# ----------------------------------------------------------------------------- MLPs
def linear(w: W, name: str, x: torch.Tensor) -> torch.Tensor:
    b = w.get(name + ".bias")
    return F.linear(x, w[name + ".weight"], b)


def resnet_fc(w: W, p: str, z: torch.Tensor, x: torch.Tensor, spec: FieldSpec) -> torch.Tensor:
    """ResnetFC.forward (model_components/resnet_fc.py:130-154; block :70-79), beta=0 -> ReLU."""
    h = linear(w, f"{p}.lin_in", x)
    for b in range(spec.n_blocks):
        if b < spec.combine_layer:
            h = h + linear(w, f"{p}.lin_z.{b}", z)
        net = linear(w, f"{p}.blocks.{b}.fc_0", torch.relu(h))
        h = h + linear(w, f"{p}.blocks.{b}.fc_1", torch.relu(net))
    return linear(w, f"{p}.lin_out", torch.relu(h))


def density_activation(x: torch.Tensor) -> torch.Tensor:
    """trunc_exp(x - 1) forward (model_components/activations.py:13-38)."""
    return torch.exp(x - 1.0)


def attention_block(w: W, p: str, x: torch.Tensor, z: torch.Tensor, spec: FieldSpec) -> torch.Tensor:
    """PreNorm(Attention) with external keys/values (model_components/transformer.py:14-21, :63-82)."""
    h, d = spec.heads, spec.dim_head
    xn = F.layer_norm(x, (x.shape[-1],), w[f"{p}.norm.weight"], w[f"{p}.norm.bias"])
    q = F.linear(xn, w[f"{p}.fn.to_q.weight"])
    kv = F.linear(z, w[f"{p}.fn.to_kv.weight"])
    k, v = kv[..., : h * d], kv[..., h * d:]
    B, N = q.shape[:2]
    q = q.reshape(B, N, h, d).transpose(1, 2)
    k = k.reshape(z.shape[0], -1, h, d).transpose(1, 2)
    v = v.reshape(z.shape[0], -1, h, d).transpose(1, 2)
    att = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * d ** -0.5, dim=-1)
    out = torch.matmul(att, v).transpose(1, 2).reshape(B, N, h * d)
    return linear(w, f"{p}.fn.to_out.0", out)


def feed_forward(w: W, p: str, x: torch.Tensor) -> torch.Tensor:
    """PreNorm(FeedForward): Linear -> GELU -> Linear (transformer.py:24-36)."""
    xn = F.layer_norm(x, (x.shape[-1],), w[f"{p}.norm.weight"], w[f"{p}.norm.bias"])
    return linear(w, f"{p}.fn.net.3", F.gelu(linear(w, f"{p}.fn.net.0", xn)))


def jacobian_head(w: W, feats: torch.Tensor, enc: torch.Tensor, spec: FieldSpec) -> torch.Tensor:
    """compute_jacobian: MLP variant (action_decoder_jacobian.py:324-337) or cross-attention
    variant (:418-446; Transformer.forward transformer.py:119-135)."""
    if spec.head == "jacobian_mlp":
        return resnet_fc(w, "decoder.jacobian_head", feats, enc, spec)
    x = linear(w, "decoder.jacobian_query_mlp", torch.cat([enc, feats], -1))
    z = w["decoder.jacobian_index_embedding"]
    for l in range(spec.depth):
        p = f"decoder.jacobian_attn_decoder.layers.{l}"
        x = attention_block(w, f"{p}.0", x, z, spec) + x
        x = feed_forward(w, f"{p}.1", x) + x
    return linear(w, "decoder.jacobian_head", x)


def color_head(w: W, geo: torch.Tensor, sh: torch.Tensor) -> torch.Tensor:
    """color_head Sequential (action_decoder_jacobian.py:315-322, use :208)."""
    h = torch.relu(linear(w, "decoder.color_head.0", torch.cat([geo, sh], -1)))
    h = torch.relu(linear(w, "decoder.color_head.2", h))
    return torch.sigmoid(linear(w, "decoder.color_head.4", h))


# ----------------------------------------------------------------------------- sampling
def spacing_to_euclid(bins: torch.Tensor, near: torch.Tensor, far: torch.Tensor) -> torch.Tensor:
    """UniformSampler: x*s_far + (1-x)*s_near (rendering/ray_samplers.py:238-245)."""
    return bins * far + (1 - bins) * near


def uniform_bins(n_rays_shape: Sequence[int], s: int) -> torch.Tensor:
    """Eval-mode SpacedSampler bins (ray_samplers.py:214-235)."""
    return torch.linspace(0.0, 1.0, s + 1)[None, ...].repeat(*n_rays_shape, 1)


def stratified_bins(t_rand: torch.Tensor, s: int) -> torch.Tensor:
    """Train-mode SpacedSampler bins (ray_samplers.py:214-233): t_rand (...,s+1) or (...,1) uniform [0,1)."""
    bins = torch.linspace(0.0, 1.0, s + 1).to(t_rand.device)[None, ...]
    centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
    upper = torch.cat([centers, bins[..., -1:]], -1)
    lower = torch.cat([bins[..., :1], centers], -1)
    return lower + (upper - lower) * t_rand


def stratified_u(rand: torch.Tensor, n_samples: int) -> torch.Tensor:
    """Train-mode PDFSampler positions (ray_samplers.py:389-401): rand (...,n_samples+1) or (...,1)."""
    nb = n_samples + 1
    u = torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb, device=rand.device)
    u = u.expand((*rand.shape[:-1], nb))
    return (u + rand / nb).contiguous()


def transmittance_weights(deltas: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
    """RaySamples.get_weights (ray_samplers.py:77-101). deltas, sigma (...,S,1)."""
    dd = torch.where(deltas > 0, deltas * sigma, torch.zeros_like(sigma))
    alphas = 1 - torch.exp(-dd)
    t = torch.cumsum(dd[..., :-1, :], dim=-2)
    t = torch.cat([torch.zeros_like(t[..., :1, :]), t], dim=-2)
    return alphas * torch.exp(-t)


def pdf_resample(weights: torch.Tensor, bins: torch.Tensor, n_samples: int,
                 u: Optional[torch.Tensor] = None, padding: float = 0.01, eps: float = 1e-5):
    """Eval-mode PDFSampler (ray_samplers.py:351-451, include_original=False).
    weights (...,S) ; bins (...,S+1) spacing-domain bin edges.  Returns new bins (...,n_samples+1)
    and the integer indices ``inds`` (the searchsorted result) for bit-exact checks."""
    nb = n_samples + 1
    wgt = weights + padding
    wsum = torch.sum(wgt, dim=-1, keepdim=True)
    pad = torch.relu(eps - wsum)
    wgt = wgt + pad / wgt.shape[-1]
    wsum = wsum + pad
    pdf = wgt / wsum
    cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, dim=-1))
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
    if u is None:
        u = (torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb) + 1.0 / (2 * nb)).to(cdf.device)
        u = u.expand(size=(*cdf.shape[:-1], nb))
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, 0, bins.shape[-1] - 1)
    above = torch.clamp(inds, 0, bins.shape[-1] - 1)
    c0, b0 = torch.gather(cdf, -1, below), torch.gather(bins, -1, below)
    c1, b1 = torch.gather(cdf, -1, above), torch.gather(bins, -1, above)
    t = torch.clip(torch.nan_to_num((u - c0) / (c1 - c0), 0), 0, 1)
    return b0 + t * (b1 - b0), inds


# ----------------------------------------------------------------------------- the path
def proposal_density(w: W, i: int, xyz: torch.Tensor, feat, c2w, k_norm, spec: FieldSpec) -> torch.Tensor:
    """DensityDecoderMlp.get_density (models/decoder/density_decoder.py:45-71). xyz (B,N,3) -> (B,N,1)."""
    z, cam = pixel_aligned(xyz, c2w, k_norm, feat)
    out = resnet_fc(w, f"proposal_networks.{i}.density_head", z, posenc(cam), spec)
    return density_activation(out)


def field_heads(w: W, xyz: torch.Tensor, feat, c2w, k_norm, spec: FieldSpec):
    """compute_density + compute_jacobian (action_decoder_jacobian.py:92-119, :128-145).
    Returns sigma (B,N,1), geo (B,N,15), jac (B,N,3A), plus the gathered feats / encoding."""
    z, cam = pixel_aligned(xyz, c2w, k_norm, feat)
    enc = posenc(cam)
    out = resnet_fc(w, "decoder.density_head", z, enc, spec)
    geo, pre = out[..., : spec.geo_dim], out[..., spec.geo_dim:]
    return density_activation(pre), geo, jacobian_head(w, z, enc, spec), z, enc


def pixel_coordinates(height: int, width: int):
    """get_pixel_coordinates (rendering/geometry.py:117-134): normalised xy centres (H,W,2), (row,col) selectors."""
    row, col = torch.arange(height), torch.arange(width)
    selector = torch.stack(torch.meshgrid(row, col, indexing="ij"), dim=-1)
    xy = torch.stack(torch.meshgrid((col + 0.5) / width, (row + 0.5) / height, indexing="xy"), dim=-1)
    return xy, selector


def world_rays_with_z(coords_xy: torch.Tensor, k_norm: torch.Tensor, c2w: torch.Tensor):
    """get_world_rays_with_z (rendering/geometry.py:170-203; unproject :42-56, transform_cam2world :68-73):
    coords (B,R,2), k_norm (B,3,3), c2w (B,4,4) -> origins (B,R,3), unit dirs (B,R,3), z (B,R,1)."""
    hom = torch.cat([coords_xy, torch.ones_like(coords_xy[..., :1])], -1)
    d = torch.einsum("cij,crj->cri", torch.inverse(k_norm), hom)
    d = d / d.norm(dim=-1, keepdim=True)
    z = d[..., -1:]
    dw = torch.einsum("cij,crj->cri", c2w[:, :3, :3], d)
    o = c2w[:, None, :3, 3].expand_as(dw)
    return o.contiguous(), dw.contiguous(), z


def positions_of(origins, dirs, starts, ends):
    """RaySamples.get_positions (ray_samplers.py:48-55)."""
    return origins[..., None, :] + dirs[..., None, :] * (starts + ends) / 2


def project_px(p: torch.Tensor, c2w: torch.Tensor, k_px: torch.Tensor) -> torch.Tensor:
    """project_world_coords_to_camera (rendering/geometry.py:206-215) with pixel-unit K."""
    _, uv = project_to_context(p, c2w, k_px)
    return uv


def render_forward(w: W, spec: FieldSpec, feat: torch.Tensor, ctxt_c2w, ctxt_k, trgt_c2w, trgt_k_px,
                   origins, dirs, z_near, z_far, action, s_prop: Sequence[int], s_nerf: int,
                   anneal: float = 1.0, bins0: Optional[torch.Tensor] = None,
                   us: Optional[Sequence[torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """Model.forward with compute_vis_features=True (models/model.py:316-396), encoder output ``feat``
    (B,512,Hf,Wf) given.  origins/dirs (B,R,3); z_near/z_far (B,); action (B,A).  Eval mode by default; train mode =
    the caller passes the stratified level-0 bins (``stratified_bins``) and per-level PDF positions
    (``stratified_u``) drawn from its own torch.rand tensors, plus the proposal-weight ``anneal``.
    Returns composites and the per-sample intermediates used by the tests."""
    B, R = origins.shape[:2]
    near = torch.ones_like(origins[..., :1]) * z_near[:, None, None]   # model.py:215-226
    far = torch.ones_like(origins[..., :1]) * z_far[:, None, None]
    bins = uniform_bins((B, R), s_prop[0]).to(origins.device) if bins0 is None else bins0
    out: Dict[str, torch.Tensor] = {}
    prop_w: List[torch.Tensor] = []
    prop_bins: List[torch.Tensor] = [bins]
    # ProposalNetworkSampler.generate_ray_samples (ray_samplers.py:497-552)
    for lvl in range(len(s_prop) + 1):
        is_prop = lvl < len(s_prop)
        if lvl > 0:
            n = s_prop[lvl] if is_prop else s_nerf
            bins, inds = pdf_resample(torch.pow(prop_w[-1], anneal)[..., 0], bins, n,
                                      u=None if us is None else us[lvl - 1])
            bins = bins.detach()
            prop_bins.append(bins)
            out[f"inds_{lvl}"] = inds
        e = spacing_to_euclid(bins, near, far)
        starts, ends = e[..., :-1, None], e[..., 1:, None]
        if is_prop:
            pos = positions_of(origins, dirs, starts, ends)
            S = pos.shape[2]
            sig = proposal_density(w, lvl, pos.reshape(B, R * S, 3), feat, ctxt_c2w, ctxt_k, spec)
            prop_w.append(transmittance_weights(ends - starts, sig.reshape(B, R, S, 1)))
    pos = positions_of(origins, dirs, starts, ends)                     # (B,R,S,3)
    S = s_nerf
    sigma, geo, jac, _, _ = field_heads(w, pos.reshape(B, R * S, 3), feat, ctxt_c2w, ctxt_k, spec)
    sigma, geo, jac = (t.reshape(B, R, S, -1) for t in (sigma, geo, jac))
    # flow = J u (action_decoder_jacobian.py:135-143): J laid out (action_dim, spatial_dim)
    flow = torch.einsum("brsad,ba->brsd", jac.reshape(B, R, S, spec.action_dim, 3), action)
    d01 = (dirs[..., None, :].expand(pos.shape) + 1.0) / 2.0           # :24-30
    rgb = color_head(w, geo, sh4(d01, spec.sh_fp16_round, spec.sh_convention))
    weights = transmittance_weights(ends - starts, sigma)               # model.py:351
    steps = (starts + ends) / 2
    depth = torch.sum(weights * steps, dim=-2) / (torch.sum(weights, -2) + 1e-10)   # :271-279
    depth = torch.clip(depth, steps.min(), steps.max())
    p = torch.sum(weights * pos, dim=-2)                                # :288-314
    pw = torch.sum(weights * (pos + flow), dim=-2)
    out.update(
        rgb=torch.sum(weights * rgb, dim=-2), depth=depth,
        optical_flow=project_px(pw, trgt_c2w, trgt_k_px) - project_px(p, trgt_c2w, trgt_k_px),
        action_features=torch.sum(weights * jac, dim=-2), steps=steps.squeeze(-1),
        weights=weights.squeeze(-1), ray_positions=p, ray_positions_warped=pw,
        sigma=sigma, jacobian=jac, rgb_samples=rgb, positions=pos, final_bins=bins,
        proposal_weights=prop_w[-1].squeeze(-1) if prop_w else torch.zeros(0, device=origins.device),
    )
    for i, (pw_, pb_) in enumerate(zip(prop_w, prop_bins)):   # ModelTrainingOutput (model.py:377-382)
        out[f"prop_weights_{i}"] = pw_.squeeze(-1)
        out[f"prop_bins_{i}"] = pb_
    return out


def infer_optical_flow(jac, weights, positions, action, trgt_c2w, trgt_k_px):
    """Model.infer_optical_flow (models/model.py:497-525) on an encode_image result."""
    B, R, S, _ = jac.shape
    flow = torch.einsum("brsad,ba->brsd", jac.reshape(B, R, S, action.shape[-1], 3), action)
    p = torch.sum(weights * positions, dim=-2)
    pw = torch.sum(weights * (positions + flow), dim=-2)
    return project_px(pw, trgt_c2w, trgt_k_px) - project_px(p, trgt_c2w, trgt_k_px)


def flow_gn_terms(jac, weights, positions, action, trgt_c2w, trgt_k_px, target_flow):
    """Gauss-Newton normal equations of the notebooks' inverse-dynamics objective
    sum_i |infer_optical_flow(u)_i - target_i|^2 (2_inverse_dynamics.ipynb cell 26 minimises it with Adam):
    G = d flow / d u by autograd through the reference formulation; returns H (B,A,A), g (B,A), loss (B) in fp64."""
    B = jac.shape[0]
    A = action.shape[-1]
    Hs, gs, ls = [], [], []
    for b in range(B):
        sl = slice(b, b + 1)
        f = lambda u: infer_optical_flow(jac[sl].double(), weights[sl].double(), positions[sl].double(), u[None],
                                         trgt_c2w[sl].double(), trgt_k_px[sl].double())[0]
        u = action[b].double()
        G = torch.autograd.functional.jacobian(f, u)          # (R, 2, A)
        r = f(u) - target_flow[b].double()
        Hs.append(torch.einsum("ria,rib->ab", G, G))
        gs.append(torch.einsum("ria,ri->a", G, r))
        ls.append((r ** 2).sum())
    return torch.stack(Hs), torch.stack(gs), torch.stack(ls)


def encoder_resnet34(w: W, image: torch.Tensor, prefix: str = "encoder.model.") -> torch.Tensor:
    """EncoderResnet.forward (models/encoder/encoder_resnet.py:53-86): resnet34 conv1..layer3 (eval BN),
    bilinear upsampling to the conv1 resolution, channel concat -> 512 channels at H/2 x W/2."""
    import torchvision

    net = torchvision.models.resnet34(weights=None)
    sd = {k[len(prefix):]: v for k, v in w.items() if k.startswith(prefix)}
    net.load_state_dict(sd, strict=False)
    net.eval()
    with torch.no_grad():
        x = net.relu(net.bn1(net.conv1(image)))
        lat = [x]
        x = net.layer1(net.maxpool(x)); lat.append(x)
        x = net.layer2(x); lat.append(x)
        x = net.layer3(x); lat.append(x)
        sz = lat[0].shape[-2:]
        return torch.cat([F.interpolate(t, sz, mode="bilinear", align_corners=False) for t in lat], 1)


_ = (math, Tuple)
