"""Higher-precision proposal mode ("bit-exact sample indexing").

The fused proposal kernel multiplies fp16 operands (fp32 accumulation); its densities differ from the fp32 reference by
~1e-3 relative, which moves a CDF value across a sample position ``u`` for ~0.2 % of the PDF-sampler's ``searchsorted``
indices (profiles/parity_r02.json).  TF32 has the same 10-bit mantissa as fp16, so it cannot help; what decides the
indices has to be computed in fp32.  This module does exactly that for the PROPOSAL levels only: the fp32 kernels of
the training path (csrc/trunk_train.cu, forced to their SIMT variants) evaluate every proposal network's ResnetFC
layer by layer on fp32 lin_z maps, ``njf_transmittance_weights`` (fp64 scan) and ``njf_pdf_sample`` (bit-exact given
identical weights) place the samples, and the final level -- densities, colours, Jacobians, compositing -- stays on
the fused tcgen05 field pass, which receives the bins through ``render(..., final_bins=...)``.

Cost: the proposal levels run ~25x slower than the fused kernel (a 400x400 frame goes from ~33 ms to ~0.3 s), so
this is an opt-in verification / reproducibility mode (``Model.precise_proposal = True``), not the default.
"""
from __future__ import annotations

from contextlib import contextmanager
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import api, train_trunk as TT


def trunk_from_state_dict(sd: Dict[str, Tensor], prefix: str, device) -> SimpleNamespace:
    """A ResnetFC parameter view (lin_in / lin_z / blocks / lin_out with .weight / .bias) over state-dict tensors."""
    lin = lambda name: SimpleNamespace(weight=sd[f"{prefix}.{name}.weight"].to(device, torch.float32),
                                       bias=sd[f"{prefix}.{name}.bias"].to(device, torch.float32))
    n_blocks = 1 + max(int(k[len(prefix) + 8:].split(".")[0]) for k in sd if k.startswith(prefix + ".blocks."))
    n_z = 1 + max(int(k[len(prefix) + 7:].split(".")[0]) for k in sd if k.startswith(prefix + ".lin_z."))
    return SimpleNamespace(lin_in=lin("lin_in"), lin_out=lin("lin_out"), lin_z=[lin(f"lin_z.{i}") for i in range(n_z)],
                           blocks=[SimpleNamespace(fc_0=lin(f"blocks.{b}.fc_0"), fc_1=lin(f"blocks.{b}.fc_1"))
                                   for b in range(n_blocks)])


@contextmanager
def _fp32_everywhere():
    """fp32 SIMT trunk kernels and fp32 library GEMMs, whatever the process-wide matmul precision is."""
    prev = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("highest")
    try:
        yield
    finally:
        torch.set_float32_matmul_precision(prev)


@torch.no_grad()
def proposal_bins_fp32(trunks: Sequence, feat_nchw: Tensor, w2c: Tensor, k_norm: Tensor, origins: Tensor, dirs: Tensor,
                       z_near: Tensor, z_far: Tensor, s_prop: Sequence[int], s_nerf: int,
                       bins0: Optional[Tensor] = None, us: Optional[Sequence[Tensor]] = None, anneal: float = 1.0,
                       samples_per_chunk: int = 1 << 21) -> Tuple[Tensor, List[Tensor], List[Tensor], List[Tensor]]:
    """ProposalNetworkSampler.generate_ray_samples (rendering/ray_samplers.py:497-552) with fp32 proposal densities.

    trunks: one ResnetFC parameter container per proposal level (``net.density_head`` or ``trunk_from_state_dict``);
    feat_nchw (B,512,Hf,Wf) encoder output; w2c (B,4,4) inverse context poses, k_norm (B,3,3); origins / dirs (B,R,3).
    Returns (final_bins (B,R,s_nerf+1), per-level output bins, per-level searchsorted indices (int32),
    per-level proposal weights (B,R,S))."""
    dev = origins.device
    B, R = origins.shape[:2]
    Hf, Wf = feat_nchw.shape[-2:]
    if bins0 is None or us is None:
        tb, tu = api.eval_tables(s_prop, s_nerf, dev)
        bins0 = tb if bins0 is None else bins0
        us = tu if us is None else us
    near, far = z_near[:, None, None], z_far[:, None, None]
    level_bins = [torch.empty(B, R, (s_prop[l + 1] if l + 1 < len(s_prop) else s_nerf) + 1, device=dev) for l in range(len(s_prop))]
    level_inds = [torch.empty(lb.shape, device=dev, dtype=torch.int32) for lb in level_bins]
    prop_w = [torch.empty(B, R, s, device=dev) for s in s_prop]
    with _fp32_everywhere():
        fmap = feat_nchw.permute(0, 2, 3, 1).contiguous().float()
        maps = [TT.lin_z_maps(t, fmap).reshape(-1, 128 * len(t.lin_z)) for t in trunks]
        rc = max(1, samples_per_chunk // (B * max(list(s_prop))))
        for r0 in range(0, R, rc):
            r1 = min(R, r0 + rc)
            o, d = origins[:, r0:r1], dirs[:, r0:r1]
            n = r1 - r0
            bins = bins0 if bins0.dim() == 1 else bins0[:, r0:r1]
            for lvl, trunk in enumerate(trunks):
                S = s_prop[lvl]
                b = bins.expand(B, n, S + 1) if bins.dim() == 1 else bins
                e = b * far + (1 - b) * near                                        # ray_samplers.py:242-245
                starts, ends = e[..., :-1, None], e[..., 1:, None]
                pos = o[..., None, :] + d[..., None, :] * (starts + ends) / 2      # :48-55
                enc, pix, tapw = TT.sample_setup(w2c, k_norm, pos.reshape(B, n * S, 3), Hf, Wf)
                z = TT._GatherMaps.apply(maps[lvl], pix, tapw)
                sigma = torch.exp(TT.resnet_fc(trunk, enc, z) - 1.0).reshape(B * n, S)
                w = api.transmittance_weights((ends - starts).reshape(B * n, S).contiguous(), sigma.contiguous())
                prop_w[lvl][:, r0:r1] = w.reshape(B, n, S)
                n_out = s_prop[lvl + 1] if lvl + 1 < len(s_prop) else s_nerf
                u = us[lvl] if us[lvl].dim() == 1 else us[lvl][:, r0:r1].reshape(B * n, n_out + 1)
                bi = bins if bins.dim() == 1 else bins.reshape(B * n, S + 1)
                nb, inds = api.pdf_sample(w, bi, u, n_out, anneal=anneal)
                bins = nb.reshape(B, n, n_out + 1)
                level_bins[lvl][:, r0:r1] = bins
                level_inds[lvl][:, r0:r1] = inds.reshape(B, n, n_out + 1)
    return level_bins[-1], level_bins, level_inds, prop_w
