"""TEST INFRASTRUCTURE (not product code): import the UNMODIFIED reference from /root/reference.

The reference (sizhe-li/neural-jacobian-field) is pure Python but depends on packages that
are not installed here (nerfstudio, tiny-cuda-nn, omegaconf).  This module installs the
smallest possible ``sys.modules`` stand-ins for exactly the symbols the hot path imports and
then imports ``neural_jacobian_field.models.model`` from ``/root/reference/project``.

It exists ONLY to (a) generate the golden vectors under ``tests/golden`` (``make_golden.py``)
and (b) validate ``oracle/njf_oracle.py`` inside this container.  ``/root/reference`` does not
exist on the GPU box, so nothing on the GPU test / bench path may import this file.

Third-party arithmetic restated here (source NOT under /root/reference; upstream unpinned,
``install.sh:23-24``) -- parity for these two encodings is therefore "unpinned":
  * nerfstudio ``NeRFEncoding`` (torch implementation): s = 2*pi*x; freqs = 2**linspace(min,max,n);
    t = (s[...,None]*freqs).flatten(-2); enc = sin(cat[t, t+pi/2]); include_input -> cat[enc, x].
  * tiny-cuda-nn ``SphericalHarmonics`` degree 4 evaluated on (2*x-1) with fp16 output
    (nerfstudio ``SHEncoding(levels=4, implementation="tcnn")``).
"""
from __future__ import annotations

import math
import sys
import types

import torch
from torch import nn

REFERENCE_PROJECT = "/root/reference/project"


def sh4_tcnn(d: torch.Tensor) -> torch.Tensor:
    """Degree-4 real spherical harmonics, tiny-cuda-nn sign/ordering convention. d: (...,3) in [-1,1]."""
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    out = [
        torch.full_like(x, 0.28209479177387814),
        -0.48860251190291987 * y,
        0.48860251190291987 * z,
        -0.48860251190291987 * x,
        1.0925484305920792 * xy,
        -1.0925484305920792 * yz,
        0.94617469575755997 * z2 - 0.31539156525251999,
        -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2,
        0.59004358992664352 * y * (-3.0 * x2 + y2),
        2.8906114426405538 * xy * z,
        0.45704579946446572 * y * (1.0 - 5.0 * z2),
        0.3731763325901154 * z * (5.0 * z2 - 3.0),
        0.45704579946446572 * x * (1.0 - 5.0 * z2),
        1.4453057213202769 * z * (x2 - y2),
        0.59004358992664352 * x * (-x2 + 3.0 * y2),
    ]
    return torch.stack(out, dim=-1)


class NeRFEncoding(nn.Module):
    def __init__(self, in_dim, num_frequencies, min_freq_exp, max_freq_exp, include_input=False,
                 implementation="torch"):
        super().__init__()
        self.in_dim, self.num_frequencies = in_dim, num_frequencies
        self.min_freq, self.max_freq, self.include_input = min_freq_exp, max_freq_exp, include_input

    def get_out_dim(self):
        return self.in_dim * self.num_frequencies * 2 + (self.in_dim if self.include_input else 0)

    def forward(self, in_tensor):
        scaled = 2 * torch.pi * in_tensor
        freqs = 2 ** torch.linspace(self.min_freq, self.max_freq, self.num_frequencies, device=in_tensor.device)
        scaled = (scaled[..., None] * freqs).view(*scaled.shape[:-1], -1)
        enc = torch.sin(torch.cat([scaled, scaled + torch.pi / 2.0], dim=-1))
        if self.include_input:
            enc = torch.cat([enc, in_tensor], dim=-1)
        return enc


class SHEncoding(nn.Module):
    """tcnn convention; output rounded through fp16 like tcnn's half-precision encodings."""

    fp16_round = True

    def __init__(self, levels=4, implementation="tcnn"):
        super().__init__()
        assert levels == 4
        self.levels = levels

    def get_out_dim(self):
        return self.levels**2

    def forward(self, in_tensor):
        out = sh4_tcnn(in_tensor * 2.0 - 1.0)
        if self.fp16_round:
            out = out.to(torch.float16)
        return out


def _apply_depth_colormap(depth, *args, **kwargs):
    d = depth.float()
    lo, hi = d.min(), d.max()
    g = (d - lo) / (hi - lo + 1e-10)
    return torch.cat([g, g, g], dim=-1)


def install() -> None:
    """Insert the stand-in modules and put the reference project on sys.path (idempotent)."""
    if "nerfstudio.field_components.encodings" in sys.modules:
        return

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    mod("nerfstudio")
    mod("nerfstudio.field_components")
    enc = mod("nerfstudio.field_components.encodings")
    enc.NeRFEncoding, enc.SHEncoding = NeRFEncoding, SHEncoding
    mod("nerfstudio.utils")
    cm = mod("nerfstudio.utils.colormaps")
    cm.apply_depth_colormap = _apply_depth_colormap
    mod("nerfstudio.cameras")
    cu = mod("nerfstudio.cameras.camera_utils")
    cu.normalize_with_norm = lambda x, dim: (x / x.norm(dim=dim, keepdim=True), x.norm(dim=dim, keepdim=True))
    if "omegaconf" not in sys.modules:
        try:
            import omegaconf  # noqa: F401
        except Exception:
            oc = mod("omegaconf")
            oc.DictConfig = dict
    if REFERENCE_PROJECT not in sys.path:
        sys.path.insert(0, REFERENCE_PROJECT)


def reference_modules():
    """Returns the reference's model module (imports it on first use)."""
    install()
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import neural_jacobian_field.models.model as ref_model  # type: ignore
    return ref_model


def build_reference_cfg(action_dim: int, head: str, s_prop, s_nerf: int, single_jitter: bool = False):
    """ModelCfg with the shipped values (configurations/model/model_allegro.yaml / model_toy_arm.yaml)."""
    install()
    ref_model = reference_modules()
    from neural_jacobian_field.model_components.resnet_fc import MlpCfg  # type: ignore
    from neural_jacobian_field.models.decoder.action_decoder_jacobian import (  # type: ignore
        ActionDecoderJacobianMlpCfg,
        ActionDecoderJacobianTransformerCfg,
        TransformerCfg,
    )
    from neural_jacobian_field.models.decoder.density_decoder import DensityDecoderMlpCfg  # type: ignore
    from neural_jacobian_field.models.encoder.encoder_resnet import EncoderResnetCfg  # type: ignore

    mlp = MlpCfg(5, 128, 3, "mean", 0.0)
    if head == "jacobian_transformer":
        dec = ActionDecoderJacobianTransformerCfg(
            name="jacobian_transformer", mlp=mlp, transformer=TransformerCfg(64, 64, 8, 3, 64))
    elif head == "jacobian_mlp":
        dec = ActionDecoderJacobianMlpCfg(name="jacobian_mlp", mlp=mlp)
    else:
        raise ValueError(head)
    return ref_model.ModelCfg(
        action_dim=action_dim,
        rendering=ref_model.RenderingCfg(tuple(s_prop), s_nerf, bool(single_jitter), 5000, 5, True, 1000, 10.0),
        encoder=EncoderResnetCfg("resnet", "bilinear", 4, True, "batch"),
        density_decoder=DensityDecoderMlpCfg("density_mlp", mlp),
        action_decoder=dec,
    )


_ = math  # keep import (documentation of constants)
