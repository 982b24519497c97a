"""Parameter containers with the reference's state-dict names, plus the plugin registries.

The reference's plugin surface for this path is Python: three name->class registries
(``ENCODERS`` models/encoder/__init__.py:7-16; ``DENSITY_DECODERS`` / ``ACTION_DECODERS``
models/decoder/__init__.py:11-19) and the config dataclasses that select them.  The classes below
keep those names, constructor signatures and parameter names (the checkpoint contract,
SURVEY.md section 8b), but hold NO arithmetic: rendering happens in libnjf_b200.so, which
``njf_b200.model.Model`` drives.  Initial values follow the reference's init distributions.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal, Optional

import torch
import torch.nn.functional as F
from torch import nn


# ----------------------------------------------------------------------------- config dataclasses
@dataclass
class MlpCfg:  # model_components/resnet_fc.py:11-17
    n_blocks: int = 5
    d_hidden: int = 128
    combine_layer: int = 3
    combine_type: Literal["mean"] = "mean"
    beta: float = 0.0


@dataclass
class TransformerCfg:  # action_decoder_jacobian.py:33-39
    attn_feat_dim: int = 64
    attn_head_dim: int = 64
    num_attn_heads: int = 8
    attn_depth: int = 3
    attn_mlp_dim: int = 64


@dataclass
class DensityDecoderMlpCfg:  # density_decoder.py:16-20
    name: Literal["density_mlp"]
    mlp: MlpCfg
    num_frequencies: int = 10


@dataclass
class ActionDecoderJacobianMlpCfg:  # action_decoder_jacobian.py:42-49
    name: Literal["jacobian_mlp"]
    mlp: MlpCfg
    num_frequencies: int = 10
    geometry_feature_dim: int = 15
    use_arm_model: bool = False
    arm_action_dim: Optional[int] = None


@dataclass
class ActionDecoderJacobianTransformerCfg:  # action_decoder_jacobian.py:52-60
    name: Literal["jacobian_transformer"]
    mlp: MlpCfg
    transformer: TransformerCfg
    num_frequencies: int = 10
    geometry_feature_dim: int = 15
    use_arm_model: bool = False
    arm_action_dim: Optional[int] = None


@dataclass
class EncoderResnetCfg:  # models/encoder/encoder_resnet.py:15-21
    name: Literal["resnet"] = "resnet"
    upsample_interp: Literal["bilinear"] = "bilinear"
    num_layers: int = 4
    use_first_pool: bool = True
    norm_type: Literal["batch", "instance", "group", "none"] = "batch"


def _check_mlp(cfg: MlpCfg, who: str) -> None:
    if (cfg.n_blocks, cfg.d_hidden, cfg.combine_layer) != (5, 128, 3) or cfg.beta > 0:
        raise NotImplementedError(
            f"{who}: the sm_100a kernels are specialised for the shipped MLP shape "
            f"(n_blocks=5, d_hidden=128, combine_layer=3, beta=0); got {cfg}")


# ----------------------------------------------------------------------------- decoder-level plugin surface
@dataclass
class DecoderOutput:  # models/decoder/action_decoder.py:19-24
    density: torch.Tensor
    color: torch.Tensor
    flow: torch.Tensor
    action_features: torch.Tensor


@dataclass
class DecoderFeatureOnlyOutput:  # action_decoder.py:27-30
    density: torch.Tensor
    action_features: torch.Tensor


@dataclass
class DensityHeadOutput:  # action_decoder_jacobian.py:63-68
    density: torch.Tensor
    density_features: torch.Tensor
    xyz_features: torch.Tensor
    pixel_aligned_features: torch.Tensor


class _StandaloneKernels:
    """What lets a B200 decoder object stand in for the reference's OWN decoder inside the reference's OWN ``Model``
    (registered under the reference's DENSITY_DECODERS / ACTION_DECODERS by ``njf_b200.plugin``): the per-point
    methods the reference calls -- ``get_density`` (density_decoder.py:45-71), ``forward`` / ``encode_image`` /
    ``compute_density`` (action_decoder_jacobian.py:92-249) -- evaluated by the fused point-query kernels
    (njf_query_points / njf_query_proposal_density).  Inference only: the reference's sampler, PDF resampling and
    compositing stay in its torch code, so this path is for drop-in compatibility, not speed; ``njf_b200.Model``
    is the fused path."""

    _sk_field = None
    _sk_key = None
    _sk_maps = None
    _sk_maps_key = None

    def _sk_parts(self):  # -> (head, action_dim, {field key: tensor})  implemented by the two decoder kinds
        raise NotImplementedError

    def _sk_get_field(self):
        from . import api, synth

        params = list(self.parameters())
        if not params or params[0].device.type != "cuda":
            from ._lib import NjfError
            raise NjfError("the B200 decoders run on a CUDA device only (there is no CPU fallback)")
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._sk_field is None or self._sk_key != key:
            head, A, mine = self._sk_parts()
            weights = {}
            for k, shp in synth.field_shapes(head, A, n_proposal=1).items():   # parts this module does not own: zeros
                weights[k] = mine[k].detach() if k in mine else torch.zeros(shp)
            with torch.cuda.device(params[0].device):
                self._sk_field = api.Field(head, A, 1, weights)
            self._sk_key, self._sk_maps_key = key, None
        return self._sk_field

    def _sk_context(self, pixel_encoding):
        """(field, hoisted maps, w2c, K, Hf, Wf) for a reference-style PixelEncoding (features, extrinsics, intrinsics)."""
        from . import api

        fld = self._sk_get_field()
        feats = pixel_encoding.features
        mkey = (feats.data_ptr(), feats._version, tuple(feats.shape))
        if self._sk_maps_key != mkey:
            self._sk_maps = fld.hoist(feats.detach().to(fld.device, torch.float32).contiguous())
            self._sk_maps_key = mkey
        cams, keep = api.make_cameras(pixel_encoding.extrinsics.to(fld.device), pixel_encoding.intrinsics.to(fld.device),
                                      None, None, fld.device)
        return fld, self._sk_maps, keep[0], keep[1], feats.shape[-2], feats.shape[-1]

    def _sk_check_inference(self):
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("the decoder-level B200 plugin is inference-only; train through njf_b200.Model "
                                      "(action phase) -- see INTEGRATION.md")


# ----------------------------------------------------------------------------- containers
class ResnetBlockParams(nn.Module):
    def __init__(self, d: int):
        super().__init__()
        self.fc_0 = nn.Linear(d, d)
        self.fc_1 = nn.Linear(d, d)
        nn.init.constant_(self.fc_0.bias, 0.0)
        nn.init.kaiming_normal_(self.fc_0.weight, a=0, mode="fan_in")
        nn.init.constant_(self.fc_1.bias, 0.0)
        nn.init.zeros_(self.fc_1.weight)


class ResnetFCParams(nn.Module):
    """Parameters of the reference's ResnetFC (model_components/resnet_fc.py:82-128)."""

    def __init__(self, cfg: MlpCfg, d_in: int, d_latent: int, d_out: int):
        super().__init__()
        self.lin_in = nn.Linear(d_in, cfg.d_hidden)
        self.lin_out = nn.Linear(cfg.d_hidden, d_out)
        self.blocks = nn.ModuleList([ResnetBlockParams(cfg.d_hidden) for _ in range(cfg.n_blocks)])
        self.lin_z = nn.ModuleList([nn.Linear(d_latent, cfg.d_hidden) for _ in range(min(cfg.combine_layer, cfg.n_blocks))])
        for lin in [self.lin_in, self.lin_out, *self.lin_z]:
            nn.init.constant_(lin.bias, 0.0)
            nn.init.kaiming_normal_(lin.weight, a=0, mode="fan_in")


class _AttnParams(nn.Module):
    def __init__(self, dim, heads, dim_head, kv_dim):
        super().__init__()
        inner = heads * dim_head
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(kv_dim, inner * 2, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(0.0))


class _FFParams(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Dropout(0.0), nn.Linear(hidden, dim),
                                 nn.Dropout(0.0))


class _PreNormParams(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn


class TransformerParams(nn.Module):
    """model_components/transformer.py:85-117 (cross-attention: selfatt=False)."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim, kv_dim):
        super().__init__()
        self.layers = nn.ModuleList([
            nn.ModuleList([_PreNormParams(dim, _AttnParams(dim, heads, dim_head, kv_dim)),
                           _PreNormParams(dim, _FFParams(dim, mlp_dim))])
            for _ in range(depth)])


def _init_jacobian(m):  # action_decoder_jacobian.py:78-83
    if type(m) == nn.Linear:
        nn.init.normal_(m.weight, mean=0.0, std=1e-4)
        if m.bias is not None:
            nn.init.normal_(m.bias, mean=0.0, std=1e-4)


def _color_head(geo: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(geo + 16, 64), nn.ReLU(), nn.Linear(64, 64), nn.ReLU(), nn.Linear(64, 3),
                         nn.Sigmoid())


class DensityDecoderMlp(nn.Module, _StandaloneKernels):
    """Proposal-network density field (models/decoder/density_decoder.py:23-43)."""

    def __init__(self, cfg: DensityDecoderMlpCfg, encoder_dim: int):
        super().__init__()
        _check_mlp(cfg.mlp, "density_mlp")
        if cfg.num_frequencies != 10:
            raise NotImplementedError("kernels are specialised for num_frequencies=10")
        self.cfg = cfg
        self.density_head = ResnetFCParams(cfg.mlp, 63, encoder_dim, 1)

    def _sk_parts(self):
        return "jacobian_mlp", 1, {"proposal_networks.0." + k: v for k, v in self.state_dict().items()}

    def get_density(self, world_space_xyz: torch.Tensor, pixel_encoding) -> torch.Tensor:
        """density_decoder.py:45-71: (batch, ray, sample, 3) world points -> (batch, ray, sample, 1) densities."""
        from . import api

        self._sk_check_inference()
        fld, maps, w2c, kn, Hf, Wf = self._sk_context(pixel_encoding)
        B, R, S = world_space_xyz.shape[:3]
        pts = world_space_xyz.detach().to(fld.device, torch.float32).reshape(B, R * S, 3).contiguous()
        with torch.cuda.device(fld.device):
            sigma = api.query_proposal_density(fld, 0, w2c, kn, maps, Hf, Wf, pts)
        return sigma.reshape(B, R, S, 1).to(world_space_xyz.device)


class ActionDecoderJacobian(nn.Module, _StandaloneKernels):
    """Common surface of the two Jacobian decoders (action_decoder_jacobian.py:86-258)."""

    spatial_dim: int = 3
    action_param_glob_pattern = "jacobian"

    def _sk_parts(self):
        arm = getattr(self, "mode", "regular") == "arm"
        sd = {}
        for k, v in self.state_dict().items():
            if arm:
                if k.startswith("jacobian_head_arm."):
                    sd["decoder.jacobian_head." + k[len("jacobian_head_arm."):]] = v
                elif "jacobian" not in k:
                    sd["decoder." + k] = v
            elif "jacobian_head_arm" not in k:
                sd["decoder." + k] = v
        if arm:
            return "jacobian_mlp", int(self.cfg.arm_action_dim), sd
        return self.cfg.name, int(self.action_dim), sd

    def _sk_get_field(self):   # the decoder's mode is part of what is packed
        if getattr(self, "_sk_mode", None) != getattr(self, "mode", "regular"):
            self._sk_field, self._sk_mode = None, getattr(self, "mode", "regular")
        return super()._sk_get_field()

    def _sk_query(self, xyz_flat: torch.Tensor, dirs_flat, pixel_encoding):
        from . import api

        self._sk_check_inference()
        fld, maps, w2c, kn, Hf, Wf = self._sk_context(pixel_encoding)
        pts = xyz_flat.detach().to(fld.device, torch.float32).contiguous()
        dirs = None if dirs_flat is None else dirs_flat.detach().to(fld.device, torch.float32).contiguous()
        with torch.cuda.device(fld.device):
            return api.query_points(fld, w2c, kn, maps, Hf, Wf, pts, dirs=dirs), (fld, w2c, kn, Hf, Wf, pts)

    def compute_density(self, world_space_xyz: torch.Tensor, pixel_encoding) -> DensityHeadOutput:
        """action_decoder_jacobian.py:92-119 at (batch, n, 3) world points."""
        from . import _lib, api

        out, (fld, w2c, kn, Hf, Wf, pts) = self._sk_query(world_space_xyz, None, pixel_encoding)
        sigma, geo = out[0], out[1]
        B, N = pts.shape[:2]
        feats = pixel_encoding.features.detach().to(fld.device, torch.float32).contiguous()
        C = feats.shape[1]
        xyzf = torch.empty(B, N, 63, device=fld.device)
        pixf = torch.empty(B, N, C, device=fld.device)
        L = api._declare()
        with torch.cuda.device(fld.device):
            _lib.check(L.njf_point_features(api.dptr(feats), api.dptr(w2c), api.dptr(kn), api.dptr(pts), B, N, C, Hf, Wf,
                                            api.dptr(xyzf), api.dptr(pixf), api.stream_ptr()))
        dev = world_space_xyz.device
        return DensityHeadOutput(density=sigma.to(dev), density_features=geo.to(dev), xyz_features=xyzf.to(dev),
                                 pixel_aligned_features=pixf.to(dev))

    def forward(self, world_space_xyz: torch.Tensor, world_space_dir: torch.Tensor, pixel_encoding) -> DecoderOutput:
        """action_decoder_jacobian.py:147-215: density, colour, J and flow = J u at (batch, ray, sample, 3) points."""
        B, R, S = world_space_xyz.shape[:3]
        (sigma, geo, jac, rgb), (fld, *_rest) = self._sk_query(world_space_xyz.reshape(B, R * S, 3),
                                                               world_space_dir.reshape(B, R * S, 3), pixel_encoding)
        A = fld.action_dim
        action = pixel_encoding.action.detach().to(fld.device, torch.float32)
        flow = torch.einsum("bnad,ba->bnd", jac.reshape(B, R * S, A, 3), action)   # :135-143 (tiny, per call)
        dev = world_space_xyz.device
        r = lambda t: t.reshape(B, R, S, -1).to(dev)
        return DecoderOutput(density=r(sigma), color=r(rgb), flow=r(flow), action_features=r(jac))

    def encode_image(self, world_space_xyz: torch.Tensor, pixel_encoding) -> DecoderFeatureOnlyOutput:
        """action_decoder_jacobian.py:217-249."""
        B, R, S = world_space_xyz.shape[:3]
        (sigma, geo, jac), _ = self._sk_query(world_space_xyz.reshape(B, R * S, 3), None, pixel_encoding)
        dev = world_space_xyz.device
        return DecoderFeatureOnlyOutput(density=sigma.reshape(B, R, S, 1).to(dev),
                                        action_features=jac.reshape(B, R, S, -1).to(dev))

    def switch_mode(self, mode: Literal["regular", "arm"]):
        """action_decoder_jacobian.py:89-90.  "arm" renders with ``jacobian_head_arm`` (a ResnetFC Jacobian head of
        3 * arm_action_dim outputs): njf_b200.Model packs that head for the MLP-head kernel."""
        if mode not in ("regular", "arm"):
            raise ValueError(f"unknown decoder mode '{mode}'")
        if mode == "arm" and not getattr(self.cfg, "use_arm_model", False):
            raise ValueError("mode 'arm' needs cfg.use_arm_model=True")
        self.mode = mode

    def freeze_non_action_parameters(self) -> int:
        counts = 0
        for name, param in self.named_parameters():
            if self.action_param_glob_pattern not in name:
                param.requires_grad = False
                counts += 1
        return counts

    def _common(self, cfg, action_dim, encoder_dim):
        _check_mlp(cfg.mlp, cfg.name)
        if cfg.num_frequencies != 10 or cfg.geometry_feature_dim != 15:
            raise NotImplementedError("kernels are specialised for num_frequencies=10, geometry_feature_dim=15")
        self.cfg = cfg
        self.action_dim = action_dim
        self.mode = "regular"
        self.density_head = ResnetFCParams(cfg.mlp, 63, encoder_dim, cfg.geometry_feature_dim + 1)


class ActionDecoderJacobianMLP(ActionDecoderJacobian):  # action_decoder_jacobian.py:261-337
    action_param_glob_pattern = "jacobian_head"

    def __init__(self, cfg: ActionDecoderJacobianMlpCfg, action_dim: int, encoder_dim: int):
        super().__init__()
        self._common(cfg, action_dim, encoder_dim)
        self.jacobian_head = ResnetFCParams(cfg.mlp, 63, encoder_dim, 3 * action_dim)
        self.jacobian_head.apply(_init_jacobian)
        if cfg.use_arm_model:
            self.jacobian_head_arm = ResnetFCParams(cfg.mlp, 63, encoder_dim, 3 * cfg.arm_action_dim)
            self.jacobian_head_arm.apply(_init_jacobian)
        self.color_head = _color_head(cfg.geometry_feature_dim)


class ActionDecoderJacobianTransformer(ActionDecoderJacobian):  # action_decoder_jacobian.py:340-446
    action_param_glob_pattern = "jacobian"

    def __init__(self, cfg: ActionDecoderJacobianTransformerCfg, action_dim: int, encoder_dim: int):
        super().__init__()
        self._common(cfg, action_dim, encoder_dim)
        t = cfg.transformer
        if (t.attn_feat_dim, t.attn_head_dim, t.num_attn_heads, t.attn_depth, t.attn_mlp_dim) != (64, 64, 8, 3, 64):
            raise NotImplementedError(f"kernels are specialised for the shipped transformer (64/64/8/3/64); got {t}")
        self.jacobian_index_embedding = nn.Parameter(torch.randn(1, action_dim, t.attn_feat_dim), requires_grad=True)
        self.jacobian_query_mlp = nn.Linear(encoder_dim + 63, t.attn_feat_dim)
        self.jacobian_attn_decoder = TransformerParams(t.attn_feat_dim, t.attn_depth, t.num_attn_heads,
                                                       t.attn_head_dim, t.attn_mlp_dim, t.attn_feat_dim)
        self.jacobian_head = nn.Linear(t.attn_feat_dim, 3 * action_dim)
        self.jacobian_head.apply(_init_jacobian)
        if cfg.use_arm_model:
            self.jacobian_head_arm = ResnetFCParams(cfg.mlp, 63, encoder_dim, 3 * cfg.arm_action_dim)
            self.jacobian_head_arm.apply(_init_jacobian)
        self.color_head = _color_head(cfg.geometry_feature_dim)


class EncoderResnet(nn.Module):
    """PixelNeRF image encoder (models/encoder/encoder_resnet.py:24-89): torchvision resnet34 trunk
    (conv1..layer3), bilinear upsampling to the conv1 resolution, channel concat (512 ch at H/2 x W/2).
    Runs once per image on cuDNN -- outside the hot path; its NCHW fp32
    output is the input contract of njf_hoist_features."""

    def __init__(self, cfg: EncoderResnetCfg):
        super().__init__()
        import torchvision

        if cfg.norm_type != "batch" or cfg.num_layers != 4 or not cfg.use_first_pool:
            raise NotImplementedError(f"encoder config {cfg} not supported (shipped: batch norm, 4 layers, first pool)")
        self.cfg = cfg
        self.model = torchvision.models.resnet34(weights=None)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward_nhwc_half(self, rgb: torch.Tensor) -> torch.Tensor:
        """The same trunk under fp16 autocast, returned as the (B,Hf,Wf,512) fp16 NHWC tensor that
        njf_hoist_features_nhwc16 consumes.  The convolutions then multiply fp16 operands with fp32
        accumulation -- the operand precision cuDNN's default TF32 convolutions already have -- and the hoist GEMM,
        which rounds its input to fp16 anyway, reads the map without a conversion pass."""
        with torch.autocast("cuda", dtype=torch.float16):
            f = self.forward(rgb)
        return f.to(torch.float16).permute(0, 2, 3, 1).contiguous()   # a view when the storage is channels-last

    def forward(self, rgb: torch.Tensor) -> torch.Tensor:
        m = self.model
        if rgb.is_cuda:   # NHWC storage: the layout cuDNN's tensor-core convolutions run in natively (values unchanged)
            if not getattr(self, "_channels_last", False):
                m.to(memory_format=torch.channels_last)
                self._channels_last = True
            rgb = rgb.contiguous(memory_format=torch.channels_last)
        x = m.relu(m.bn1(m.conv1(rgb)))
        lat = [x]
        x = m.layer1(m.maxpool(x)); lat.append(x)
        x = m.layer2(x); lat.append(x)
        x = m.layer3(x); lat.append(x)
        sz = lat[0].shape[-2:]
        return torch.cat([F.interpolate(t, sz, mode=self.cfg.upsample_interp, align_corners=False) for t in lat], 1)

    def get_output_dim(self) -> int:
        return 512


# ----------------------------------------------------------------------------- registries
ENCODERS = {"resnet": EncoderResnet}
DENSITY_DECODERS = {"density_mlp": DensityDecoderMlp}
class ActionDecoderFlowMlp(nn.Module):
    """Registry slot of the reference's ablation decoder (models/decoder/action_decoder_flow.py:64-286: a direct
    scene-flow MLP conditioned on the action, no Jacobian).  No shipped config selects it and it has no B200 kernel
    ("keep registry slot only"): constructing it fails loudly."""

    def __init__(self, cfg=None, action_dim: int = 0, encoder_dim: int = 0):
        raise NotImplementedError("action decoder 'flow_mlp' (ablation baseline) has no B200 kernel; "
                                  "use 'jacobian_mlp' or 'jacobian_transformer'")


ACTION_DECODERS = {"jacobian_mlp": ActionDecoderJacobianMLP, "jacobian_transformer": ActionDecoderJacobianTransformer,
                   "flow_mlp": ActionDecoderFlowMlp}

EncoderCfg = EncoderResnetCfg
DensityDecoderCfg = DensityDecoderMlpCfg
ActionDecoderCfg = ActionDecoderJacobianMlpCfg | ActionDecoderJacobianTransformerCfg


def get_encoder(cfg):
    return ENCODERS[cfg.name](cfg)


def get_density_decoder(cfg, encoder_dim: int):
    return DENSITY_DECODERS[cfg.name](cfg=cfg, encoder_dim=encoder_dim)


def get_action_decoder(cfg, action_dim: int, encoder_dim: int):
    if cfg.name not in ACTION_DECODERS:
        raise NotImplementedError(f"action decoder '{cfg.name}' has no B200 kernel (supported: {sorted(ACTION_DECODERS)})")
    return ACTION_DECODERS[cfg.name](cfg=cfg, action_dim=action_dim, encoder_dim=encoder_dim)
