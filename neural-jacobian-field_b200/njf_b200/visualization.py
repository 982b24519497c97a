"""Joint-sensitivity maps of a rendered Jacobian image, as used by the reference's visualisation notebooks
(neural_jacobian_field/inference/jacobian_color_map.py:53-109).  Plain tensor ops on whatever device the
Jacobian image lives on (post-processing of the render output, not part of the per-sample path)."""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
from torch import Tensor

# per-action display colours of the released models (jacobian_color_map.py:13-50; configuration data)
JACOBIAN_COLORMAP: Dict[str, List[List[float]]] = {
    "model_allegro": [[0.0, 0.5, 0.5], [0, 1, 0], [0.8, 0.1, 0.1], [0.8, 0.0, 0.8], [0.0, 0.8, 0], [1.0, 0.8, 0],
                      [1, 1, 0], [1, 0.0, 0.0]],
    "model_toy_arm": [[0.5, 0.8, 0.2], [0.9, 0.2, 0.0], [0, 0.8, 0], [1.0, 0.0, 1.0], [0, 0, 1], [0.1, 0.9, 0.7]],
    "model_pneumatic_hand_only": [[0, 0, 1], [0.9, 0.2, 0.0], [0, 0.9, 0], [1.0, 0.0, 1.0], [0.1, 0.9, 0.7],
                                  [0.5, 0.8, 0.2]],
}
JACOBIAN_COLORMAP["model_allegro_transformer"] = JACOBIAN_COLORMAP["model_allegro"]


def _minmax01(x: Tensor, dims) -> Tensor:
    lo = x.amin(dim=dims, keepdim=True)
    hi = x.amax(dim=dims, keepdim=True)
    return ((x - lo) / (hi - lo + 1e-10)).clip(0, 1)


def compute_joint_sensitivity(jacobians: Tensor, extrinsics: Optional[Tensor] = None, mode: int = 0) -> Tensor:
    """(..., H, W, 3A) rendered Jacobians -> (..., A, H, W) per-action sensitivity in [0, 1]: the norm of each
    action's 3-vector (optionally rotated by ``extrinsics`` (..., 4, 4) first), min-max normalised per action
    over the image; ``mode == 1`` returns ``clip(1.1 - s)``."""
    J = jacobians.reshape(*jacobians.shape[:-1], -1, 3)                       # (..., H, W, A, 3)
    if extrinsics is not None:
        # direction vectors (homogeneous w = 0): only the rotation block acts.  Like the reference's einsum the
        # leading dims of `extrinsics` broadcast right-aligned against (..., H, W, A): pass (4, 4) or
        # (batch, 1, 1, 1, 4, 4)
        J = torch.einsum("...ij,...j->...i", extrinsics[..., :3, :3], J)
    s = torch.linalg.vector_norm(J, dim=-1)                                   # (..., H, W, A)
    s = s.movedim(-1, -3)                                                     # (..., A, H, W)
    s = _minmax01(s, (-2, -1))
    if mode == 1:
        s = (1.1 - s).clip(0, 1)
    return s


def visualize_joint_sensitivity(sensitivity: Tensor, color_map: Tensor) -> np.ndarray:
    """(..., A, H, W) sensitivities x (3, A) colours -> uint8 image (..., H, W, 3), white where nothing moves."""
    img = torch.einsum("...ahw,ca->...chw", sensitivity, color_map.to(sensitivity))
    img = _minmax01(img, (-2, -1))
    img = img.movedim(-3, -1).cpu().numpy()
    return ((1 - img) * 255).astype(np.uint8)
