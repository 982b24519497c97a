"""``Model`` -- API-identical mirror of ``neural_jacobian_field.models.model.Model``
(project/neural_jacobian_field/models/model.py:147-628) whose rendering arithmetic runs in
``libnjf_b200.so``.

Same constructor (``Model(cfg: ModelCfg)``), same input/output dataclasses, same state-dict keys
(``encoder.model.*``, ``proposal_networks.N.density_head.*``, ``decoder.*``), same methods:
``forward``, ``patch_render``, ``encode_image``, ``infer_optical_flow``, ``compute_pixel_encoding``,
``step_before_iter`` / ``step_after_iter``.  Inputs may live on the host: they are copied to the
model's CUDA device (that copy is what bench.py's ``e2e`` number includes).

There is no CPU / PyTorch fallback: without a CUDA device or without the built extension every
rendering call raises.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from . import _lib, api
from .modules import (ActionDecoderCfg, DensityDecoderCfg, EncoderCfg, get_action_decoder, get_density_decoder,
                      get_encoder)
from .render import RenderResult, render


# ----------------------------------------------------------------------------- dataclasses (model.py:35-144)
@dataclass
class RenderingCfg:
    num_proposal_samples: Tuple[int, ...]
    num_nerf_samples: int
    single_jitter: bool = False
    proposal_warmup: int = 5000
    proposal_update_every: int = 5
    use_proposal_weight_anneal: bool = True
    proposal_weights_anneal_max_num_iters: int = 1000
    proposal_weights_anneal_slope: float = 10.0


@dataclass
class ModelCfg:
    action_dim: int
    rendering: RenderingCfg
    encoder: EncoderCfg
    density_decoder: DensityDecoderCfg
    action_decoder: ActionDecoderCfg


@dataclass
class CameraInput:
    input_image: Tensor       # (B,3,H,W)
    ctxt_extrinsics: Tensor   # (B,4,4) camera-to-world, relative to the context camera
    ctxt_intrinsics: Tensor   # (B,3,3) normalised
    trgt_extrinsics: Tensor   # (B,4,4)
    trgt_intrinsics: Tensor   # (B,3,3) pixel units


@dataclass
class RenderingInput:
    origins: Tensor      # (B,R,3)
    directions: Tensor   # (B,R,3)
    z_near: Tensor       # (B,)
    z_far: Tensor        # (B,)


@dataclass
class RobotInput:
    robot_action: Tensor  # (B,A)


@dataclass
class ModelInput:
    camera_input: CameraInput
    rendering_input: RenderingInput
    robot_input: RobotInput


@dataclass
class ModelStandardOutput:
    rgb: Tensor
    depth: Tensor
    optical_flow: Tensor


@dataclass
class SampleBins:
    """What the training losses read from the reference's RaySamples (rendering/ray_samplers.py:28-46):
    spacing-domain and euclidean bin edges of one sampling level."""
    spacing_starts: Tensor  # (B,R,S,1)
    spacing_ends: Tensor    # (B,R,S,1)
    starts: Optional[Tensor] = None  # (B,R,S,1) euclidean
    ends: Optional[Tensor] = None    # (B,R,S,1)


@dataclass
class ModelTrainingOutput:
    weights_list: List[Tensor]
    ray_samples_list: List[SampleBins]


@dataclass
class ModelVisOutput:
    action_features: Tensor
    ray_positions: Tensor
    ray_positions_warped: Tensor
    weights: Tensor
    steps: Tensor


@dataclass
class ModelOutput:
    standard_output: ModelStandardOutput
    training_output: Optional[ModelTrainingOutput]
    vis_output: Optional[ModelVisOutput]


@dataclass
class ModelInferenceEncoding:
    density: Tensor                # (B,R,S,1)
    action_features: Tensor        # (B,R,S,3A)
    weights: Tensor                # (B,R,S,1)
    ray_samples_positions: Tensor  # (B,R,S,3)
    # collapsed form used by infer_optical_flow (flow is linear in the action, so
    # sum_s w (x + J u) = p + Jbar^T u): computed by the field kernel in the same pass
    jbar: Optional[Tensor] = None  # (B,R,3A)
    p: Optional[Tensor] = None     # (B,R,3)


@dataclass
class DensityHeadOutput:  # action_decoder_jacobian.py:63-68
    density: Tensor
    density_features: Tensor
    xyz_features: Tensor
    pixel_aligned_features: Tensor


@dataclass
class PixelEncoding:
    features: Tensor
    extrinsics: Tensor
    intrinsics: Tensor
    action: Tensor
    hoisted: Optional[Tensor] = None  # njf_hoist_features output for ``features``


@dataclass
class RenderingOutput:
    rgb: Tensor
    depth_raw: Tensor
    depth_rgb: Tensor
    flow_raw: Tensor
    flow_rgb: Tensor
    ray_positions: Tensor
    ray_positions_warped: Tensor
    action_features: Tensor
    steps: Tensor
    weights: Tensor


def apply_depth_colormap(depth: Tensor) -> Tensor:
    """Visual-only stand-in for nerfstudio.utils.colormaps.apply_depth_colormap (model.py:607):
    min-max normalised depth through a 5-knot turbo-like ramp."""
    d = depth.float()
    lo, hi = d.min(), d.max()
    t = ((d - lo) / (hi - lo + 1e-10)).clamp(0, 1)
    knots = torch.tensor([[0.19, 0.07, 0.23], [0.16, 0.57, 0.96], [0.48, 0.99, 0.35], [0.98, 0.70, 0.17],
                          [0.48, 0.01, 0.01]], device=d.device)
    x = t * 4.0
    i = x.floor().clamp(max=3).long()
    f = x - i
    return knots[i[..., 0]] * (1 - f) + knots[i[..., 0] + 1] * f


class _FlowFromEncoding(torch.autograd.Function):
    """optical flow of the collapsed encoding; forward in libnjf_b200.so, analytic backward wrt the action."""

    @staticmethod
    def forward(ctx, action, jbar, p, w2c, kpx):
        L = api._declare()
        B, R = p.shape[:2]
        A = action.shape[-1]
        flow = torch.empty(B, R, 2, device=p.device, dtype=torch.float32)
        pw = torch.empty(B, R, 3, device=p.device, dtype=torch.float32)
        act = action.detach().contiguous().float()
        _lib.check(L.njf_flow_from_encoding(api.dptr(jbar), api.dptr(p), api.dptr(act), api.dptr(w2c), api.dptr(kpx),
                                            B * R, R, A, api.dptr(flow), api.dptr(pw), api.stream_ptr()))
        ctx.save_for_backward(jbar, p, act, w2c, kpx)
        ctx.A = A
        return flow

    @staticmethod
    def backward(ctx, g):
        # d flow / d action through proj(p + Jbar^T u): njf_flow_backward (csrc/xf_backward.cu), one thread per ray
        from . import train as T

        L = T._declare()
        jbar, p, act, w2c, kpx = ctx.saved_tensors
        B, R = p.shape[:2]
        ga = torch.empty(B, ctx.A, device=p.device, dtype=torch.float32)
        _lib.check(L.njf_flow_backward(api.dptr(g.contiguous().float()), None, api.dptr(jbar), api.dptr(p), api.dptr(act),
                                       api.dptr(w2c), api.dptr(kpx), B * R, R, ctx.A, None, api.dptr(ga),
                                       api.stream_ptr()))
        return ga, None, None, None, None


class _FrameGraph:
    """Encoder + hoist + pose inversion + proposal / field / finish passes of ONE frame shape captured into a CUDA
    graph (SURVEY.md 7.2 "CUDA-graph the whole frame", 8f-3): a frame is the host->device copies of its inputs into
    static buffers plus one graph launch.  Possible because the library allocates nothing inside a pass and reads the
    cameras through device pointers here (no per-view constants baked into kernel parameters)."""

    NAMES = ("image", "ctxt_c2w", "ctxt_k", "trgt_c2w", "trgt_k", "origins", "dirs", "z_near", "z_far", "action")

    def __init__(self, model: "Model", fld: api.Field, cam: CameraInput, rin: RenderingInput, rob: RobotInput, vis: bool):
        dev = model._device()
        self.dev = dev
        r = model.cfg.rendering
        self.s_prop, self.s_nerf = tuple(r.num_proposal_samples), int(r.num_nerf_samples)
        src = self._sources(cam, rin, rob)
        self.inp = {k: torch.empty(tuple(t.shape), dtype=torch.float32, device=dev) for k, t in zip(self.NAMES, src)}
        B, R = rin.origins.shape[:2]
        with torch.cuda.device(dev):
            self.bins0, self.us = api.eval_tables(self.s_prop, self.s_nerf, dev)
            self.ws = torch.empty(fld.workspace_bytes(B, R, self.s_prop, self.s_nerf), dtype=torch.uint8, device=dev)
            self._copy_in(src)
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(2):   # warm-up: cuDNN algorithm choice, function attributes, allocator
                    self._body(model, fld, vis)
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph), torch.no_grad():
                self.res = self._body(model, fld, vis)

    @staticmethod
    def _sources(cam, rin, rob):
        return (cam.input_image, cam.ctxt_extrinsics, cam.ctxt_intrinsics, cam.trgt_extrinsics, cam.trgt_intrinsics,
                rin.origins, rin.directions, rin.z_near, rin.z_far, rob.robot_action)

    def _copy_in(self, src):
        for k, t in zip(self.NAMES, src):
            self.inp[k].copy_(t.detach(), non_blocking=True)

    def _body(self, model, fld, vis) -> RenderResult:
        i = self.inp
        if getattr(model, "encoder_half", False):
            fh = model.encoder.forward_nhwc_half(i["image"])
            maps = fld.hoist_nhwc16(fh)
            feats = fh.permute(0, 3, 1, 2)
        else:
            feats = model.encoder.forward(i["image"]).float().contiguous()
            maps = fld.hoist(feats)
        cams, keep = api.make_cameras(i["ctxt_c2w"], i["ctxt_k"], i["trgt_c2w"], i["trgt_k"], self.dev)
        Hf, Wf = feats.shape[-2:]
        res = render(fld, maps, Hf, Wf, cams, i["origins"], i["dirs"], i["z_near"], i["z_far"], i["action"], self.s_prop,
                     self.s_nerf, vis=vis, bins0=self.bins0, us=self.us, anneal=model._anneal, workspace=self.ws)
        res._cams = (keep, feats, maps)
        return res

    def run(self, cam, rin, rob) -> RenderResult:
        with torch.cuda.device(self.dev):
            self._copy_in(self._sources(cam, rin, rob))
            self.graph.replay()
        return self.res   # static buffers: overwritten by the next replay of this shape


class Model(nn.Module):
    def __init__(self, cfg: ModelCfg):
        super().__init__()
        self.cfg = cfg
        self.encoder = get_encoder(cfg.encoder)
        self.decoder = get_action_decoder(cfg.action_decoder, action_dim=cfg.action_dim,
                                          encoder_dim=self.encoder.get_output_dim())
        n_prop = len(cfg.rendering.num_proposal_samples)
        if not 1 <= n_prop <= api.NJF_MAX_LEVELS:
            raise NotImplementedError(f"{n_prop} proposal levels (supported: 1..{api.NJF_MAX_LEVELS})")
        self.proposal_networks = nn.ModuleList(
            [get_density_decoder(cfg.density_decoder, encoder_dim=self.encoder.get_output_dim()) for _ in range(n_prop)])
        self._anneal = 1.0
        self._step = 0
        self._steps_since_update = 0
        self._field: Optional[api.Field] = None
        self._field_key = None
        self._field_key_head = None
        self.sh_fp16_round = True  # tiny-cuda-nn's SH encoding returns fp16
        self.sh_convention = "tcnn"  # or "nerfstudio_torch" (what nerfstudio computes without tiny-cuda-nn)
        self.cuda_graph = False    # eval-mode forward of a fixed shape as ONE CUDA-graph launch (encoder + hoist + render)
        # eval-mode proposal levels in fp32 (njf_b200/precise.py): sample indices follow the fp32 reference to ~1e-5
        # instead of ~2e-3, at ~10x the frame time; the final level stays on the fused tcgen05 field pass
        self.precise_proposal = False
        # eval-mode encoder under fp16 autocast, emitting the NHWC fp16 map the hoist kernel copies straight into its
        # operand tiles; features differ from the fp32 / TF32 encoder at the 1e-3 level
        self.encoder_half = False
        self._graphs: Dict[tuple, "_FrameGraph"] = {}
        self.output_device: Optional[torch.device] = None   # None: results go back to where the rays came from (reference)
        self.jitter_generator: Optional[torch.Generator] = None   # train-mode stratified jitter (None = torch's global CUDA RNG)
        # (bins0, [u per level]) to use INSTEAD of drawing jitter (train.stratified_tables' layout): replays a step of
        # another implementation, e.g. tables drawn on the CPU with the reference's seed (tests/golden/train_*.npz)
        self.jitter_tables = None

    # ------------------------------------------------------------------ training-schedule hooks (model.py:201-213)
    def step_before_iter(self, step):
        r = self.cfg.rendering
        if r.use_proposal_weight_anneal:
            n = r.proposal_weights_anneal_max_num_iters
            train_frac = np.clip(step / n, 0, 1)
            b = r.proposal_weights_anneal_slope
            self._anneal = float((b * train_frac) / ((b - 1) * train_frac + 1))

    def step_after_iter(self, step):
        if self.cfg.rendering.use_proposal_weight_anneal:
            self._step = step
            self._steps_since_update += 1

    # ------------------------------------------------------------------ packed-weight cache
    def _mode(self) -> str:
        return getattr(self.decoder, "mode", "regular")

    def _head_and_dim(self) -> Tuple[str, int]:
        """Kernel head and action dimension for the decoder's current mode.  mode "arm"
        (action_decoder_jacobian.py:306-313, 331, 400-407, 439) swaps the Jacobian head for ``jacobian_head_arm``,
        a ResnetFC with 3 * arm_action_dim outputs: that is the MLP-head kernel on the arm head's weights."""
        if self._mode() == "arm":
            if not getattr(self.decoder.cfg, "use_arm_model", False):
                raise _lib.NjfError("decoder mode 'arm' needs cfg.use_arm_model=True (jacobian_head_arm)")
            return "jacobian_mlp", int(self.decoder.cfg.arm_action_dim)
        return self.cfg.action_decoder.name, int(self.cfg.action_dim)

    def _hot_state(self) -> Dict[str, Tensor]:
        arm = self._mode() == "arm"
        sd = {}
        for k, v in self.state_dict().items():
            if not (k.startswith("decoder.") or k.startswith("proposal_networks.")):
                continue
            if arm:
                if k.startswith("decoder.jacobian_head_arm."):
                    sd["decoder.jacobian_head." + k[len("decoder.jacobian_head_arm."):]] = v
                elif "jacobian" not in k:
                    sd[k] = v
            elif "jacobian_head_arm" not in k:
                sd[k] = v
        return sd

    def field(self) -> api.Field:
        """Packed weights for the kernels; re-packed whenever a hot-path parameter or a packing option changed.  When only
        the cross-attention Jacobian head changed (an optimiser step of the action phase) the head is re-packed in place
        (njf_field_update_head) instead of rebuilding the whole field."""
        dev = self._device()
        is_head = lambda n: n.startswith("decoder.") and "jacobian" in n and "jacobian_head_arm" not in n
        named = [(n, p) for n, p in self.named_parameters() if not n.startswith("encoder.")]
        ver = lambda sel: tuple((p.data_ptr(), p._version) for n, p in named if sel(n))
        key = (dev, self._mode(), bool(self.sh_fp16_round), self.sh_convention, ver(lambda n: not is_head(n)))
        key_head = ver(is_head)
        if self._field is None or self._field_key != key:
            head, A = self._head_and_dim()
            with torch.cuda.device(dev):
                self._field = api.Field(head, A, len(self.proposal_networks), self._hot_state(),
                                        sh_fp16_round=self.sh_fp16_round, sh_convention=self.sh_convention)
            self._field_key, self._field_key_head = key, key_head
            self._graphs.clear()
        elif self._field_key_head != key_head:
            if self.cfg.action_decoder.name == "jacobian_transformer" and self._mode() == "regular":
                self._field.update_head({n: p for n, p in self.state_dict().items() if is_head(n)})
                self._field_key_head = key_head
            else:
                self._field = None
                return self.field()
        return self._field

    def _field_with_current_trunks(self) -> api.Field:
        """A packed field whose TRUNKS (proposal networks, density head, colour head) are current while its Jacobian head
        may be stale: what the MLP head's action phase renders densities / weights / colours with (everything but the
        head is frozen there, model_wrapper.py:75-85) without re-packing the field after every optimiser step.  The
        Jacobian outputs of such a render are not used."""
        if self._field is not None and self._mode() == "regular":
            is_head = lambda n: n.startswith("decoder.") and "jacobian" in n and "jacobian_head_arm" not in n
            key = (self._device(), self._mode(), bool(self.sh_fp16_round), self.sh_convention,
                   tuple((p.data_ptr(), p._version) for n, p in self.named_parameters()
                         if not n.startswith("encoder.") and not is_head(n)))
            if self._field_key == key:
                return self._field
        return self.field()

    def _field_for_head_queries(self) -> api.Field:
        """A packed field whose CROSS-ATTENTION HEAD (query MLP, attention / feed-forward layers, index embedding,
        jacobian_head) is current, whatever the state of the trunks: what the trunk-training forward needs to evaluate
        the per-sample Jacobians for the (gradient-free) flow output without re-packing every trunk after every
        optimiser step -- the Jacobian head's output does not depend on the density trunk."""
        if (self._field is None or self.cfg.action_decoder.name != "jacobian_transformer" or self._mode() != "regular"
                or self._field_key[:4] != (self._device(), self._mode(), bool(self.sh_fp16_round), self.sh_convention)):
            return self.field()
        is_head = lambda n: n.startswith("decoder.") and "jacobian" in n and "jacobian_head_arm" not in n
        key_head = tuple((p.data_ptr(), p._version) for n, p in self.named_parameters()
                         if not n.startswith("encoder.") and is_head(n))
        if self._field_key_head != key_head:
            self._field.update_head({n: p for n, p in self.state_dict().items() if is_head(n)})
            self._field_key_head = key_head
        return self._field

    def _device(self) -> torch.device:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.NjfError("njf_b200.Model renders on a CUDA device only (move the module with .cuda()); "
                                "there is no CPU fallback")
        return dev

    def _check_mode(self):
        if self._mode() not in ("regular", "arm"):
            raise NotImplementedError(f"decoder mode '{self._mode()}' unknown")

    def _trainable_outside_head(self) -> List[str]:
        pat = self.decoder.action_param_glob_pattern
        return [n for n, p in self.named_parameters()
                if p.requires_grad and not (n.startswith("decoder.") and pat in n[len("decoder."):])]

    # ------------------------------------------------------------------ encoding
    def _encode(self, camera_input: CameraInput, robot_input: RobotInput) -> PixelEncoding:
        dev = self._device()
        img = camera_input.input_image.to(dev, non_blocking=True)
        with torch.no_grad():
            if getattr(self, "encoder_half", False) and not self.training:
                fh = self.encoder.forward_nhwc_half(img)
                hoisted = self.field().hoist_nhwc16(fh)
                feats = fh.permute(0, 3, 1, 2)   # (B,512,Hf,Wf)-shaped fp16 view; by-product consumers convert on demand
            else:
                feats = self.encoder.forward(img).float().contiguous()
                hoisted = self.field().hoist(feats)
        return PixelEncoding(features=feats, extrinsics=camera_input.ctxt_extrinsics,
                             intrinsics=camera_input.ctxt_intrinsics, action=robot_input.robot_action, hoisted=hoisted)

    def compute_pixel_encoding(self, camera_input, rendering_input, robot_input) -> PixelEncoding:
        return self._encode(camera_input, robot_input)

    def _render(self, pe: PixelEncoding, camera_input: CameraInput, rendering_input: RenderingInput,
                robot_input: RobotInput, **kw) -> RenderResult:
        dev = self._device()
        r = self.cfg.rendering
        cams, keep = api.make_cameras(pe.extrinsics, pe.intrinsics, camera_input.trgt_extrinsics,
                                      camera_input.trgt_intrinsics, dev)
        mv = lambda t: t.detach().to(dev, torch.float32, non_blocking=True)
        Hf, Wf = pe.features.shape[-2:]
        host_nf = ((rendering_input.z_near, rendering_input.z_far)
                   if rendering_input.z_near.device.type == "cpu" and not pe.extrinsics.is_cuda else None)
        with torch.cuda.device(dev):
            if getattr(self, "precise_proposal", False) and "final_bins" not in kw:
                from . import precise

                fb, lb, li, pw = precise.proposal_bins_fp32(
                    [n.density_head for n in self.proposal_networks], pe.features, keep[0], keep[1],
                    mv(rendering_input.origins), mv(rendering_input.directions), mv(rendering_input.z_near),
                    mv(rendering_input.z_far), tuple(r.num_proposal_samples), r.num_nerf_samples, anneal=self._anneal)
                kw["final_bins"] = fb
            res = render(self.field(), pe.hoisted, Hf, Wf, cams, mv(rendering_input.origins),
                         mv(rendering_input.directions), mv(rendering_input.z_near), mv(rendering_input.z_far),
                         mv(robot_input.robot_action), tuple(r.num_proposal_samples), r.num_nerf_samples,
                         anneal=self._anneal, host_near_far=host_nf, **kw)
        res._cams = keep
        return res

    # ------------------------------------------------------------------ forward (model.py:316-396)
    def forward(self, camera_input: CameraInput, rendering_input: RenderingInput, robot_input: RobotInput,
                compute_vis_features: bool = False) -> ModelOutput:
        self._check_mode()
        if self.training:
            return self._forward_train(camera_input, rendering_input, robot_input, compute_vis_features)
        out_dev = self.output_device or rendering_input.origins.device
        if self.cuda_graph and getattr(self, "precise_proposal", False):
            raise _lib.NjfError("precise_proposal runs outside the CUDA-graph frame: set cuda_graph = False")
        if self.cuda_graph:
            res = self._graph_frame(camera_input, rendering_input, robot_input, compute_vis_features)
            if out_dev.type == "cuda":   # the graph's outputs are static buffers: hand CUDA callers their own copy
                res = RenderResult(**{k: (v.clone() if isinstance(v, Tensor) else v) for k, v in vars(res).items()
                                      if k in RenderResult.__dataclass_fields__})
        else:
            pe = self._encode(camera_input, robot_input)
            res = self._render(pe, camera_input, rendering_input, robot_input, vis=compute_vis_features)
        return self._model_output(res, out_dev, compute_vis_features, None)

    @staticmethod
    def _model_output(res, out_dev, compute_vis_features, training_output) -> ModelOutput:
        back = (lambda t: t) if out_dev.type == "cuda" else (lambda t: t.to(out_dev))
        out = ModelOutput(
            standard_output=ModelStandardOutput(rgb=back(res.rgb), depth=back(res.depth), optical_flow=back(res.flow)),
            training_output=training_output, vis_output=None)
        if compute_vis_features:
            out.vis_output = ModelVisOutput(action_features=back(res.jbar), steps=back(res.steps),
                                            weights=back(res.weights), ray_positions=back(res.p),
                                            ray_positions_warped=back(res.pw))
        return out

    # ------------------------------------------------------------------ train-mode forward (model.py:316-396 with
    # self.training: stratified jitter, ray_samplers.py:219-233 / 389-401; ModelTrainingOutput, model.py:377-382)
    def _forward_train(self, camera_input, rendering_input, robot_input, compute_vis_features) -> ModelOutput:
        from . import train as T

        dev = self._device()
        r = self.cfg.rendering
        s_prop, s_nerf = tuple(r.num_proposal_samples), int(r.num_nerf_samples)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if need_grad and (self._trainable_outside_head() or self.cfg.action_decoder.name != "jacobian_transformer"
                          or self._mode() != "regular"):
            # perception phase (density / colour / proposal networks / encoder trainable) or an MLP Jacobian head:
            # the trunks run layer by layer on the fp32 training kernels (csrc/trunk_train.cu) under autograd
            return self._forward_train_trunks(camera_input, rendering_input, robot_input, compute_vis_features)
        out_dev = self.output_device or rendering_input.origins.device
        pe = self._encode(camera_input, robot_input)
        mv = lambda t: t.detach().to(dev, torch.float32, non_blocking=True).contiguous()
        o, d = mv(rendering_input.origins), mv(rendering_input.directions)
        zn, zf = mv(rendering_input.z_near), mv(rendering_input.z_far)
        B, R = o.shape[:2]
        # RNG stays in torch: seed torch / set `jitter_generator` to reproduce a step
        bins0, us = T.train_tables(self, s_prop, s_nerf, B, R, dev)
        action = robot_input.robot_action.to(dev, torch.float32)
        cams, keep = api.make_cameras(pe.extrinsics, pe.intrinsics, camera_input.trgt_extrinsics,
                                      camera_input.trgt_intrinsics, dev)
        Hf, Wf = pe.features.shape[-2:]
        with torch.cuda.device(dev):
            if need_grad:
                holder: list = []
                folded = T.fold_head(self.decoder, self.cfg.action_dim)
                wq = self.decoder.jacobian_query_mlp
                flow, pw, jbar, _, _, _ = T._RenderJacobianHead.apply(
                    folded, wq.weight, wq.bias, action, self.field(), pe.hoisted, pe.features, cams, keep, o, d, zn, zf,
                    s_prop, s_nerf, bins0, us, self._anneal, holder)
                res = holder[0]
                res_flow, res_pw, res_jbar = flow, pw, jbar
            else:
                res = render(self.field(), pe.hoisted, Hf, Wf, cams, o, d, zn, zf, action.detach().contiguous(), s_prop,
                             s_nerf, vis=True, sampler_outputs=True, bins0=bins0, us=us, anneal=self._anneal)
                res_flow, res_pw, res_jbar = res.flow, res.pw, res.jbar
        res._cams = keep
        # ModelTrainingOutput: weights and sample bins of every proposal level and of the final level
        near, far = zn[:, None, None], zf[:, None, None]
        euclid = lambda b: b * far + (1 - b) * near                       # ray_samplers.py:242-245
        level_bins = [bins0] + list(res.level_bins)                        # spacing-domain edges per level
        weights_list, samples_list = [], []
        for lvl, b in enumerate(level_bins):
            w = res.prop_weights[lvl] if lvl < len(s_prop) else res.weights
            e = euclid(b)
            weights_list.append(w[..., None])
            samples_list.append((b, e))
        back = (lambda t: t) if out_dev.type == "cuda" else (lambda t: t.to(out_dev))
        samples_list = [SampleBins(spacing_starts=back(b[..., :-1, None]), spacing_ends=back(b[..., 1:, None]),
                                   starts=back(e[..., :-1, None]), ends=back(e[..., 1:, None])) for b, e in samples_list]
        out = ModelOutput(
            standard_output=ModelStandardOutput(rgb=back(res.rgb), depth=back(res.depth), optical_flow=back(res_flow)),
            training_output=ModelTrainingOutput(weights_list=[back(w) for w in weights_list], ray_samples_list=samples_list),
            vis_output=None)
        if compute_vis_features:
            out.vis_output = ModelVisOutput(action_features=back(res_jbar), steps=back(res.steps),
                                            weights=back(res.weights), ray_positions=back(res.p),
                                            ray_positions_warped=back(res_pw))
        return out

    def _update_schedule(self, step) -> float:
        """models/model.py:181-189: how many steps may pass between proposal-network updates."""
        r = self.cfg.rendering
        return float(np.clip(np.interp(step, [0, r.proposal_warmup], [0, r.proposal_update_every]), 1,
                             r.proposal_update_every))

    def _forward_train_trunks(self, camera_input, rendering_input, robot_input, compute_vis_features) -> ModelOutput:
        from . import train_trunk as TT

        out_dev = self.output_device or rendering_input.origins.device
        with torch.cuda.device(self._device()):
            t = TT.forward_train(self, camera_input, rendering_input, robot_input, compute_vis_features)
        back = (lambda v: v) if out_dev.type == "cuda" else (lambda v: v.to(out_dev))
        euclid = lambda b: b * t["far"] + (1 - b) * t["near"]
        samples_list = []
        for b in t["bins_list"]:
            e = euclid(b)
            samples_list.append(SampleBins(spacing_starts=back(b[..., :-1, None]), spacing_ends=back(b[..., 1:, None]),
                                           starts=back(e[..., :-1, None]), ends=back(e[..., 1:, None])))
        out = ModelOutput(
            standard_output=ModelStandardOutput(rgb=back(t["rgb"]), depth=back(t["depth"]), optical_flow=back(t["flow"])),
            training_output=ModelTrainingOutput(weights_list=[back(w) for w in t["weights_list"]],
                                                ray_samples_list=samples_list),
            vis_output=None)
        if compute_vis_features:
            out.vis_output = ModelVisOutput(action_features=back(t["jbar"]), steps=back(t["steps"]),
                                            weights=back(t["weights"]), ray_positions=back(t["p"]),
                                            ray_positions_warped=back(t["pw"]))
        return out

    # ------------------------------------------------------------------ one CUDA-graph launch per frame
    def _graph_frame(self, camera_input, rendering_input, robot_input, vis: bool) -> RenderResult:
        dev = self._device()
        fld = self.field()
        B, R = rendering_input.origins.shape[:2]
        key = (tuple(camera_input.input_image.shape), B, R, bool(vis), float(self._anneal), bool(getattr(self, "encoder_half", False)),
               tuple((p.data_ptr(), p._version) for p in self.encoder.parameters()))
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= 4:
                self._graphs.clear()
            g = self._graphs[key] = _FrameGraph(self, fld, camera_input, rendering_input, robot_input, vis)
        return g.run(camera_input, rendering_input, robot_input)

    # ------------------------------------------------------------------ inverse-dynamics helpers (model.py:458-525)
    def encode_image(self, camera_input, rendering_input, robot_input) -> ModelInferenceEncoding:
        self._check_mode()
        pe = self._encode(camera_input, robot_input)
        res = self._render(pe, camera_input, rendering_input, robot_input, vis=True, per_sample=True)
        return ModelInferenceEncoding(density=res.sigma, action_features=res.jac, weights=res.weights[..., None],
                                      ray_samples_positions=res.positions, jbar=res.jbar, p=res.p)

    def infer_optical_flow(self, model_inference_encoding: ModelInferenceEncoding, camera_input: CameraInput,
                           robot_input: RobotInput) -> Tensor:
        assert "jacobian" in self.cfg.action_decoder.name
        enc = model_inference_encoding
        dev = enc.weights.device
        if enc.jbar is None or enc.p is None:
            raise _lib.NjfError("encoding lacks the collapsed (jbar, p) fields; produce it with Model.encode_image")
        f = lambda t: t.detach().to("cpu", torch.float32)
        w2c = torch.inverse(f(camera_input.trgt_extrinsics)).contiguous().to(dev)
        kpx = f(camera_input.trgt_intrinsics).contiguous().to(dev)
        action = robot_input.robot_action.to(dev)
        return _FlowFromEncoding.apply(action, enc.jbar.contiguous(), enc.p.contiguous(), w2c, kpx)

    # ------------------------------------------------------------------ patch_render (model.py:527-628)
    @torch.no_grad()
    def patch_render(self, camera_input: CameraInput, rendering_input: RenderingInput, robot_input: RobotInput,
                     patch_size: int = 2048, render_height: int = 480, render_width: int = 640,
                     verbose: bool = False) -> RenderingOutput:
        """Same contract as the reference, including the per-patch depth clip range (model.py:277 couples
        the rays of one forward call), but the image encoder and the hoisted maps are computed ONCE per
        frame instead of once per patch.  ``patch_size=None`` renders the frame in one launch."""
        from torchvision.utils import flow_to_image

        self._check_mode()
        pe = self._encode(camera_input, robot_input)
        num_rays = rendering_input.origins.shape[1]
        step = num_rays if patch_size is None else patch_size
        keys = ["rgb", "depth_raw", "flow_raw", "action_features", "steps", "weights", "ray_positions",
                "ray_positions_warped"]
        acc = {k: [] for k in keys}
        for s in range(0, num_rays, step):
            ri = RenderingInput(origins=rendering_input.origins[:, s:s + step], directions=rendering_input.directions[:, s:s + step],
                                z_near=rendering_input.z_near, z_far=rendering_input.z_far)
            res = self._render(pe, camera_input, ri, robot_input, vis=True)
            for k, v in zip(keys, (res.rgb, res.depth, res.flow, res.jbar, res.steps, res.weights, res.p, res.pw)):
                acc[k].append(v)
        out = {k: torch.cat(v, dim=1).reshape(v[0].shape[0], render_height, render_width, -1) for k, v in acc.items()}
        out["depth_rgb"] = apply_depth_colormap(out["depth_raw"])
        out["flow_rgb"] = flow_to_image(out["flow_raw"].permute(0, 3, 1, 2).contiguous()).permute(0, 2, 3, 1)
        return RenderingOutput(**out)

    # ------------------------------------------------------------------ point queries (model.py:416-456)
    def compute_density(self, world_space_xyz: Tensor, pixel_encoding: PixelEncoding):
        """Density head (+ Jacobian head) at explicit world-space points (B,N,3).  Returns
        ``(DensityHeadOutput, extras)`` like the reference; ``extras["jacobian_head_output"]`` is (B,N,3A)."""
        self._check_mode()
        L = api._declare()
        dev = self._device()
        if pixel_encoding.hoisted is None:
            pixel_encoding.hoisted = self.field().hoist(pixel_encoding.features.to(dev).float().contiguous())
        pts = world_space_xyz.detach().to(dev, torch.float32).contiguous()
        B, N = pts.shape[:2]
        A = self.cfg.action_dim
        f = lambda t: t.detach().to("cpu", torch.float32)
        w2c = torch.inverse(f(pixel_encoding.extrinsics)).contiguous().to(dev)
        kn = f(pixel_encoding.intrinsics).contiguous().to(dev)
        feats = pixel_encoding.features.to(dev).float().contiguous()
        C, Hf, Wf = feats.shape[1:]
        o = dict(device=dev, dtype=torch.float32)
        xyzf, pixf = torch.empty(B, N, 63, **o), torch.empty(B, N, C, **o)
        with torch.cuda.device(dev):
            sigma, geo, jac = api.query_points(self.field(), w2c, kn, pixel_encoding.hoisted, Hf, Wf, pts)
            _lib.check(L.njf_point_features(api.dptr(feats), api.dptr(w2c), api.dptr(kn), api.dptr(pts), B, N, C, Hf, Wf,
                                            api.dptr(xyzf), api.dptr(pixf), api.stream_ptr()))
        out = DensityHeadOutput(density=sigma, density_features=geo, xyz_features=xyzf, pixel_aligned_features=pixf)
        return out, {"jacobian_head_output": jac}
