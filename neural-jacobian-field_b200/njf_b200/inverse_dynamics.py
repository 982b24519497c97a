"""Inverse dynamics on a ``Model.encode_image`` result: which action reproduces an observed optical flow?

The reference solves it with an Adam loop over ``Model.infer_optical_flow`` (notebooks/real_world/
2_inverse_dynamics.ipynb, models/model.py:497-525) and notes that a least-squares solver would make it real-time.
Because the composited flow is ``proj(p + Jbar^T u) - proj(p)`` the problem is a small
non-linear least squares in the A action components: ``njf_flow_gn_terms`` accumulates the A x A normal equations
over all query rays in one kernel and the damped Gauss-Newton step is an A x A solve."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib, api


@dataclass
class GaussNewtonTerms:
    H: Tensor     # (B, A, A) float64
    g: Tensor     # (B, A)    float64
    loss: Tensor  # (B,)      float64  sum_i w_i |flow_i - target_i|^2


def _target_cameras(camera_input, dev):
    f = lambda t: t.detach().to("cpu", torch.float32)
    w2c = torch.inverse(f(camera_input.trgt_extrinsics)).contiguous().to(dev)
    kpx = f(camera_input.trgt_intrinsics).contiguous().to(dev)
    return w2c, kpx


def gauss_newton_terms(encoding, camera_input, action: Tensor, target_flow: Tensor,
                       ray_weight: Optional[Tensor] = None) -> GaussNewtonTerms:
    """Normal equations at ``action`` (B, A) for ``target_flow`` (B, R, 2) in pixels."""
    if encoding.jbar is None or encoding.p is None:
        raise _lib.NjfError("encoding lacks the collapsed (jbar, p) fields; produce it with Model.encode_image")
    L = api._declare()
    dev = encoding.p.device
    B, R = encoding.p.shape[:2]
    A = action.shape[-1]
    w2c, kpx = _target_cameras(camera_input, dev)
    act = action.detach().to(dev, torch.float32).contiguous()
    tgt = target_flow.detach().to(dev, torch.float32).contiguous()
    wgt = None if ray_weight is None else ray_weight.detach().to(dev, torch.float32).contiguous()
    f64 = dict(device=dev, dtype=torch.float64)
    ws = torch.empty(L.njf_flow_gn_workspace_doubles(B), **f64)
    H, g, loss = torch.empty(B, A, A, **f64), torch.empty(B, A, **f64), torch.empty(B, **f64)
    with torch.cuda.device(dev):
        _lib.check(L.njf_flow_gn_terms(api.dptr(encoding.jbar.contiguous()), api.dptr(encoding.p.contiguous()), api.dptr(act),
                                       api.dptr(w2c), api.dptr(kpx), api.dptr(tgt), api.dptr(wgt) if wgt is not None else None,
                                       B * R, R, A, api.dptr(ws), api.dptr(H), api.dptr(g), api.dptr(loss), api.stream_ptr()))
    return GaussNewtonTerms(H, g, loss)


def solve_action(encoding, camera_input, target_flow: Tensor, action0: Tensor, iters: int = 8, damping: float = 1e-6,
                 ray_weight: Optional[Tensor] = None, action_prior_weight: float = 0.0) -> Tuple[Tensor, List[float]]:
    """Levenberg-Marquardt on the action: returns (action (B, A) float32, loss history).  ``damping`` scales the
    diagonal of H (relative LM damping, adapted x10 / x0.3 on rejected / accepted steps);
    ``action_prior_weight`` adds lambda |u - action0|^2 (the notebooks' regulariser)."""
    u = action0.detach().to(encoding.p.device, torch.float64).clone()
    u0 = u.clone()
    lam = float(damping)
    hist: List[float] = []
    t = gauss_newton_terms(encoding, camera_input, u.float(), target_flow, ray_weight)
    cost = t.loss + action_prior_weight * ((u - u0) ** 2).sum(-1)
    hist.append(float(cost.sum()))
    eye = torch.eye(u.shape[-1], device=u.device, dtype=torch.float64)
    for _ in range(iters):
        Hd = t.H + action_prior_weight * eye
        gd = t.g + action_prior_weight * (u - u0)
        Hl = Hd + lam * (torch.diag_embed(torch.diagonal(Hd, dim1=-2, dim2=-1)) + 1e-12 * eye)
        step = torch.linalg.solve(Hl, -gd[..., None])[..., 0]
        cand = u + step
        tc = gauss_newton_terms(encoding, camera_input, cand.float(), target_flow, ray_weight)
        cc = tc.loss + action_prior_weight * ((cand - u0) ** 2).sum(-1)
        if float(cc.sum()) <= float(cost.sum()):
            u, t, cost, lam = cand, tc, cc, max(lam * 0.3, 1e-12)
        else:
            lam *= 10.0
        hist.append(float(cost.sum()))
    return u.float(), hist
