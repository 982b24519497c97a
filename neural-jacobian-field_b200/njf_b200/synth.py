"""Deterministic synthetic weights, cameras and shapes, keyed by state-dict name (benchmarks, the smoke test and
the parity fixtures all draw from here; no rendering arithmetic lives in this module).

No checkpoint of the reference exists offline (``notebooks/real_world/inference_demo_data/
real_world_pretrained_ckpts/placeholder.txt`` is empty) and its random init depends on
construction order, so golden vectors and GPU parity tests share THIS scheme instead: every
tensor is drawn from a numpy Generator seeded by (seed, crc32(key)), with a per-kind scale
that gives "trained-like" dynamic range (non-zero fc_1, Jacobian head well above the
reference's N(0,1e-4) init, moderate densities so that weights spread along the ray).
"""
from __future__ import annotations

import zlib
from typing import Dict, Tuple

import numpy as np
import torch


def _rng(seed: int, key: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(key.encode())])


def synth_tensor(key: str, shape: Tuple[int, ...], seed: int, regime: str = "trained") -> torch.Tensor:
    g = _rng(seed, key)
    shape = tuple(int(s) for s in shape)
    leaf = key.split(".")[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    n = lambda: g.standard_normal(shape).astype(np.float32)
    is_norm = (".bn" in key or "bn1" in key or "downsample.1" in key or ".norm." in key)
    if leaf == "running_mean":
        v = 0.1 * n()
    elif leaf == "running_var":
        v = (1.0 + 0.2 * g.random(shape)).astype(np.float32)
    elif is_norm and leaf == "weight":
        v = 1.0 + 0.1 * n()
    elif is_norm and leaf == "bias":
        v = 0.05 * n()
    elif key.endswith("jacobian_index_embedding"):
        v = n()
    elif len(shape) == 4:  # conv: kaiming fan_out
        v = n() * np.float32(np.sqrt(2.0 / (shape[0] * shape[2] * shape[3])))
    elif len(shape) == 2:
        fan_in = shape[1]
        scale = np.sqrt(2.0 / fan_in)
        if "fc_1" in key:
            scale *= 0.5
        if "lin_z" in key:
            scale *= 0.5
        if "density_head.lin_out" in key:
            scale *= 0.35
        if "jacobian_head" in key or "jacobian_query_mlp" in key:
            if regime == "init_like" and "jacobian_head" in key:
                scale = 1e-4          # the reference's own init (action_decoder_jacobian.py:78-83)
            else:
                scale *= 0.5
            if key in ("decoder.jacobian_head.weight", "decoder.jacobian_head.lin_out.weight") and regime != "init_like":
                scale *= 0.1          # keeps composited flows at a few pixels
        if ".to_q." in key or ".to_kv." in key or ".to_out." in key or ".net." in key:
            scale = np.sqrt(1.0 / fan_in)
        v = n() * np.float32(scale)
        # Spectral decay over the 10 positional-encoding octaves (columns are dim-major,
        # freq-minor, sin block then cos block then raw xyz): a trained field is smooth, a white
        # spectrum makes every per-sample value hang on fp32 noise x 2*pi*512 (chaotic parity).
        if regime == "trained" and (key.endswith("lin_in.weight") or key.endswith("jacobian_query_mlp.weight")):
            decay = np.ones(shape[1], dtype=np.float32)
            decay[:60] = 2.0 ** (-(np.arange(60) % 10)).astype(np.float32)
            v = v * decay[None, :] * np.float32(2.0)
    elif len(shape) == 1:
        v = 0.05 * n()
        if "jacobian_head" in key and regime == "init_like":
            v = 1e-4 * n()
    else:
        v = n()
    return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))


def synth_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int, regime: str = "trained") -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(k, tuple(s), seed, regime) for k, s in shapes.items()}


# --------------------------------------------------------------------------- parameter shapes
def resnet_fc_shapes(prefix: str, d_in: int, d_latent: int, d_out: int, d_hidden=128, n_blocks=5, combine=3):
    s = {f"{prefix}.lin_in.weight": (d_hidden, d_in), f"{prefix}.lin_in.bias": (d_hidden,),
         f"{prefix}.lin_out.weight": (d_out, d_hidden), f"{prefix}.lin_out.bias": (d_out,)}
    for b in range(n_blocks):
        for fc in ("fc_0", "fc_1"):
            s[f"{prefix}.blocks.{b}.{fc}.weight"] = (d_hidden, d_hidden)
            s[f"{prefix}.blocks.{b}.{fc}.bias"] = (d_hidden,)
    for k in range(min(combine, n_blocks)):
        s[f"{prefix}.lin_z.{k}.weight"] = (d_hidden, d_latent)
        s[f"{prefix}.lin_z.{k}.bias"] = (d_hidden,)
    return s


def field_shapes(head: str, action_dim: int, n_proposal: int = 1, encoder_dim: int = 512) -> Dict[str, Tuple[int, ...]]:
    """Shapes of every hot-path parameter (decoder + proposal networks), reference state-dict names
    (``models/decoder/action_decoder_jacobian.py:261-416``, ``density_decoder.py:23-43``)."""
    s: Dict[str, Tuple[int, ...]] = {}
    for i in range(n_proposal):
        s.update(resnet_fc_shapes(f"proposal_networks.{i}.density_head", 63, encoder_dim, 1))
    s.update(resnet_fc_shapes("decoder.density_head", 63, encoder_dim, 16))
    if head == "jacobian_mlp":
        s.update(resnet_fc_shapes("decoder.jacobian_head", 63, encoder_dim, 3 * action_dim))
    elif head == "jacobian_transformer":
        s["decoder.jacobian_index_embedding"] = (1, action_dim, 64)
        s["decoder.jacobian_query_mlp.weight"] = (64, encoder_dim + 63)
        s["decoder.jacobian_query_mlp.bias"] = (64,)
        for l in range(3):
            p = f"decoder.jacobian_attn_decoder.layers.{l}"
            s[f"{p}.0.norm.weight"] = (64,)
            s[f"{p}.0.norm.bias"] = (64,)
            s[f"{p}.0.fn.to_q.weight"] = (512, 64)
            s[f"{p}.0.fn.to_kv.weight"] = (1024, 64)
            s[f"{p}.0.fn.to_out.0.weight"] = (64, 512)
            s[f"{p}.0.fn.to_out.0.bias"] = (64,)
            s[f"{p}.1.norm.weight"] = (64,)
            s[f"{p}.1.norm.bias"] = (64,)
            s[f"{p}.1.fn.net.0.weight"] = (64, 64)
            s[f"{p}.1.fn.net.0.bias"] = (64,)
            s[f"{p}.1.fn.net.3.weight"] = (64, 64)
            s[f"{p}.1.fn.net.3.bias"] = (64,)
        s["decoder.jacobian_head.weight"] = (3 * action_dim, 64)
        s["decoder.jacobian_head.bias"] = (3 * action_dim,)
    else:
        raise ValueError(head)
    s["decoder.color_head.0.weight"] = (64, 31)
    s["decoder.color_head.0.bias"] = (64,)
    s["decoder.color_head.2.weight"] = (64, 64)
    s["decoder.color_head.2.bias"] = (64,)
    s["decoder.color_head.4.weight"] = (3, 64)
    s["decoder.color_head.4.bias"] = (3,)
    return s


# --------------------------------------------------------------------------- synthetic scene
# Two of the twelve OPENCV cameras of the Allegro rig, rounded (the real rig lives in
# notebooks/real_world/dataset_configs/allegro_config.json: fl ~608, c ~(317.5, 239.1), 480x640).
ALLEGRO_INTRINSICS_PX = dict(fl_x=608.6, fl_y=608.1, cx=317.5, cy=239.1, w=640, h=480)


def normalized_intrinsics(fl_x, fl_y, cx, cy, w, h) -> torch.Tensor:
    """Intrinsics with rows divided by (w, h) (dataset.py:283-294 convention)."""
    return torch.tensor([[fl_x / w, 0.0, cx / w], [0.0, fl_y / h, cy / h], [0.0, 0.0, 1.0]], dtype=torch.float32)


def look_at_c2w(eye, target, up=(0.0, -1.0, 0.0)) -> torch.Tensor:
    """OpenCV camera-to-world (x right, y down, z forward)."""
    eye = np.asarray(eye, dtype=np.float64)
    f = np.asarray(target, dtype=np.float64) - eye
    f /= np.linalg.norm(f)
    r = np.cross(-np.asarray(up, dtype=np.float64), f)   # x = y(down) x z(forward)
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = r, d, f, eye
    return torch.from_numpy(m.astype(np.float32))


def relative_target_pose(view: int = 1) -> torch.Tensor:
    """Target camera-to-world RELATIVE to the context camera (ctxt c2w == I, dataset.py:321-327):
    a camera orbiting a point ~1.2 m in front of the context camera."""
    ang = 0.35 * view
    eye = (1.2 * np.sin(ang), -0.15 * view, 1.2 - 1.2 * np.cos(ang))
    return look_at_c2w(eye, (0.0, 0.0, 1.2))
