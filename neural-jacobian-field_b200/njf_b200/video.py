"""Host-side glue either side of the render path: the dataset's camera conventions as a device-side
"camera rig" object, and the validation-video renderer.

* ``CameraRig``  -- what ``DatasetCommon`` does with the rig file per sample (data/dataset/dataset.py:277-327,
  utils/convention.py:9-13, 110-125): OpenCV camera-to-world poses flipped by ``post_process_camera_to_world``,
  intrinsics normalised by the image size, poses made relative to the context camera (``get_relative_transform``), the
  target intrinsics de-normalised to pixels for the flow projection.  All cameras of the rig live on the device once;
  a (context, target) pair is two index operations.
* ``render_interpolated_view`` -- ``ModelWrapper.render_interpolated_view`` (models/model_wrapper.py:213-327): target pose and
  intrinsics interpolated from the target to the context camera with the cosine ease, one fused render per frame
  (rays generated on the device by ``njf_make_rays``; with ``model.cuda_graph`` every frame is one graph launch), depth
  colour-mapped, flow turned into an image.  ``interpolate_pose`` is the reference's axis-angle interpolation
  (visualization/view_interpolation.py:9-36) written with torch ops instead of scipy.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor

from . import geometry
from .model import CameraInput, Model, RenderingInput, RobotInput, apply_depth_colormap


# ----------------------------------------------------------------------------- pose interpolation
def _rotvec_from_matrix(r: Tensor) -> Tensor:
    """Axis-angle vector of a rotation matrix (scipy Rotation.from_matrix(...).as_rotvec())."""
    r = r.double()
    cos = ((r[0, 0] + r[1, 1] + r[2, 2]) - 1.0) / 2.0
    angle = torch.arccos(cos.clamp(-1.0, 1.0))
    axis = torch.stack([r[2, 1] - r[1, 2], r[0, 2] - r[2, 0], r[1, 0] - r[0, 1]])
    s = axis.norm()
    if float(s) < 1e-12:
        if float(angle) < 1e-6:
            return torch.zeros(3, dtype=torch.float64, device=r.device)
        # angle ~ pi: the axis is the eigenvector of eigenvalue 1
        w, v = torch.linalg.eigh((r + r.T) / 2.0)
        return v[:, -1] * angle
    return axis / s * angle


def _matrix_from_rotvec(v: Tensor) -> Tensor:
    """Rodrigues' formula (scipy Rotation.from_rotvec(...).as_matrix())."""
    angle = v.norm()
    if float(angle) < 1e-12:
        return torch.eye(3, dtype=torch.float64, device=v.device)
    k = v / angle
    K = torch.zeros(3, 3, dtype=torch.float64, device=v.device)
    K[0, 1], K[0, 2], K[1, 0], K[1, 2], K[2, 0], K[2, 1] = -k[2], k[1], k[2], -k[0], -k[1], k[0]
    return torch.eye(3, dtype=torch.float64, device=v.device) + torch.sin(angle) * K + (1.0 - torch.cos(angle)) * (K @ K)


@torch.no_grad()
def interpolate_pose(initial: Tensor, final: Tensor, t: float) -> Tensor:
    """visualization/view_interpolation.py:9-36: rotate by the fraction t of the relative rotation, lerp the position."""
    r_initial, r_final = initial[:3, :3], final[:3, :3]
    r_relative = _matrix_from_rotvec(_rotvec_from_matrix(r_final @ r_initial.T) * t).to(final.dtype)
    result = torch.zeros_like(initial)
    result[3, 3] = 1
    result[:3, :3] = r_relative @ r_initial
    result[:3, 3] = initial[:3, 3] + (final[:3, 3] - initial[:3, 3]) * t
    return result


@torch.no_grad()
def interpolate_intrinsics(initial: Tensor, final: Tensor, t: float) -> Tensor:
    """view_interpolation.py:39-45."""
    return initial + (final - initial) * t


# ----------------------------------------------------------------------------- camera rig
def post_process_camera_to_world(c2w: Tensor) -> Tensor:
    """utils/convention.py:9-13: flip the y and z axes of the camera frame (OpenGL <-> OpenCV)."""
    conversion = torch.eye(4, dtype=torch.float32, device=c2w.device)
    conversion[1:3, 1:3] *= -1
    return c2w @ conversion


def denormalize_intrinsics(intrinsics: Tensor, width: int, height: int) -> Tensor:
    """utils/convention.py:110-125."""
    k = intrinsics.clone()
    k[..., 0, :] *= width
    k[..., 1, :] *= height
    return k


@dataclass
class CameraPair:
    ctxt_extrinsics: Tensor   # (1,4,4) == identity: poses are relative to the context camera
    ctxt_intrinsics: Tensor   # (1,3,3) normalised
    trgt_extrinsics: Tensor   # (1,4,4)
    trgt_intrinsics: Tensor   # (1,3,3) normalised (ray generation)
    trgt_intrinsics_px: Tensor  # (1,3,3) pixel units (CameraInput.trgt_intrinsics, model_wrapper.py:527-531)
    height: int
    width: int


class CameraRig:
    """All cameras of a rig file (``{"cameras": [{fl_x, fl_y, cx, cy, w, h, transform_matrix}, ...]}``, the layout of
    notebooks/real_world/dataset_configs/*_config.json) on one device, in the dataset's conventions."""

    def __init__(self, cameras: Sequence[dict], device="cpu"):
        c2w, k, hw = [], [], []
        for cam in cameras:
            m = torch.tensor(cam["transform_matrix"], dtype=torch.float32)
            if m.shape == (3, 4):
                m = torch.cat([m, torch.tensor([[0.0, 0.0, 0.0, 1.0]])], 0)
            c2w.append(m)
            kk = torch.eye(3)
            kk[0, 0], kk[1, 1], kk[0, 2], kk[1, 2] = cam["fl_x"], cam["fl_y"], cam["cx"], cam["cy"]
            k.append(kk)
            hw.append((int(cam["h"]), int(cam["w"])))
        self.hw: List[tuple] = hw
        self.device = torch.device(device)
        raw = torch.stack(c2w).to(self.device)
        self.c2w = post_process_camera_to_world(raw)                       # dataset.py:277-281 load_extrinsics
        kn = torch.stack(k).to(self.device)
        wh = torch.tensor([[w, h] for h, w in hw], dtype=torch.float32, device=self.device)
        kn[:, :2] = kn[:, :2] / wh[:, :, None]                             # dataset.py:283-294 load_intrinsics
        self.k_norm = kn

    def __len__(self) -> int:
        return len(self.hw)

    def pair(self, ctxt: int, trgt: int) -> CameraPair:
        """dataset.py:321-327 get_relative_transform: both poses expressed in the context camera's frame."""
        inv = torch.inverse(self.c2w[ctxt])
        h, w = self.hw[trgt]
        rel = lambda m: torch.einsum("ij, jk -> ik", inv, m)   # the reference's own contraction (bit-faithful)
        return CameraPair(ctxt_extrinsics=rel(self.c2w[ctxt])[None], ctxt_intrinsics=self.k_norm[ctxt][None],
                          trgt_extrinsics=rel(self.c2w[trgt])[None], trgt_intrinsics=self.k_norm[trgt][None],
                          trgt_intrinsics_px=denormalize_intrinsics(self.k_norm[trgt], w, h)[None], height=h, width=w)


# ----------------------------------------------------------------------------- validation video
@torch.no_grad()
def render_interpolated_view(model: Model, context_image: Tensor, cams: CameraPair, robot_action: Tensor, z_near: Tensor,
                             z_far: Tensor, image_height: int, image_width: int, num_frames: int = 30,
                             coordinates: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """models/model_wrapper.py:213-327 with batch size 1.  Returns videos shaped (1, T, C, H, W) under "rgb", "depth"
    (colour-mapped) and "optical_flow" (flow image), plus the raw first-frame (target-view) "pred_depth_t0" (1,1,H,W) and
    "pred_flow_t0" (1,2,H,W)."""
    from torchvision.utils import flow_to_image

    dev = model._device()
    was_training = model.training
    model.eval()
    f = lambda t: t.to(dev, torch.float32)
    if coordinates is None:
        xy, _ = geometry.get_pixel_coordinates(image_height, image_width, device=dev)
        coordinates = xy.reshape(1, -1, 2)
    frames: Dict[str, List[Tensor]] = {"rgb": [], "depth": [], "optical_flow": []}
    out: Dict[str, Tensor] = {}
    chw = lambda t: t.reshape(1, image_height, image_width, -1).permute(0, 3, 1, 2)
    for i, tl in enumerate(torch.linspace(0, 1, num_frames)):
        t = (math.cos(math.pi * (tl.item() + 1)) + 1) / 2                 # smoothing, model_wrapper.py:232
        c2w = interpolate_pose(cams.trgt_extrinsics[0], cams.ctxt_extrinsics[0], t)[None]
        k = interpolate_intrinsics(cams.trgt_intrinsics[0], cams.ctxt_intrinsics[0], t)[None]
        origins, directions, _ = geometry.get_world_rays_with_z(f(coordinates), f(k), f(c2w))
        mo = model.forward(
            CameraInput(input_image=f(context_image), ctxt_extrinsics=f(cams.ctxt_extrinsics),
                        ctxt_intrinsics=f(cams.ctxt_intrinsics), trgt_extrinsics=f(c2w),
                        trgt_intrinsics=f(denormalize_intrinsics(k, width=image_width, height=image_height))),
            RenderingInput(origins=origins, directions=directions, z_near=f(z_near), z_far=f(z_far)),
            RobotInput(robot_action=f(robot_action)), compute_vis_features=False)
        so = mo.standard_output
        if i == 0:
            out["pred_depth_t0"] = chw(so.depth).clone()
            out["pred_flow_t0"] = chw(so.optical_flow).clone()
        frames["rgb"].append(chw(so.rgb).clone())
        frames["depth"].append(apply_depth_colormap(so.depth.reshape(1, image_height, image_width, 1)).permute(0, 3, 1, 2))
        frames["optical_flow"].append(flow_to_image(chw(so.optical_flow).contiguous()))
    for k_, v in frames.items():
        out[k_] = torch.stack(v, dim=1)                                    # (1, T, C, H, W)
    if was_training:
        model.train()
    return out
