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
                       ray_weight: Optional[Tensor] = None, _cams=None) -> GaussNewtonTerms:
    """Normal equations at ``action`` (B, A) for ``target_flow`` (B, R, 2) in pixels.
    (``_cams``: target (w2c, K_px) already on the device -- the solver computes them once.)"""
    if encoding.jbar is None or encoding.p is None:
        raise _lib.NjfError("encoding lacks the collapsed (jbar, p) fields; produce it with Model.encode_image")
    L = api._declare()
    dev = encoding.p.device
    B, R = encoding.p.shape[:2]
    A = action.shape[-1]
    w2c, kpx = _cams if _cams is not None else _target_cameras(camera_input, dev)
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


def levenberg_marquardt(terms_fn, u0: Tensor, iters: int = 8, damping: float = 1e-6,
                        prior_weight: float = 0.0) -> Tuple[Tensor, Tensor]:
    """Batched Levenberg-Marquardt driver on normal equations: ``terms_fn(u (B,A) float64) -> (H (B,A,A), g (B,A),
    loss (B,))``.  Every decision (accept / reject, damping x0.3 / x10) is taken per batch row with tensor ops on
    the device of ``u0`` -- no host synchronisation inside the loop.  Returns (u (B,A) float64, history
    (iters+1, B) of the accepted cost)."""
    u = u0.detach().to(torch.float64).clone()
    ref = u.clone()
    A = u.shape[-1]
    eye = torch.eye(A, device=u.device, dtype=torch.float64)
    lam = torch.full((u.shape[0],), float(damping), device=u.device, dtype=torch.float64)
    H, g, loss = terms_fn(u)
    cost = loss + prior_weight * ((u - ref) ** 2).sum(-1)
    hist = [cost]
    for _ in range(iters):
        Hd = H + prior_weight * eye
        gd = g + prior_weight * (u - ref)
        diag = torch.diag_embed(torch.diagonal(Hd, dim1=-2, dim2=-1)) + 1e-12 * eye
        step = torch.linalg.solve(Hd + lam[:, None, None] * diag, -gd[..., None])[..., 0]
        cand = u + step
        Hc, gc, lc = terms_fn(cand)
        cc = lc + prior_weight * ((cand - ref) ** 2).sum(-1)
        ok = cc <= cost   # NaN candidates are rejected
        u = torch.where(ok[:, None], cand, u)
        H = torch.where(ok[:, None, None], Hc, H)
        g = torch.where(ok[:, None], gc, g)
        cost = torch.where(ok, cc, cost)
        lam = torch.where(ok, (lam * 0.3).clamp_min(1e-12), lam * 10.0)
        hist.append(cost)
    return u, torch.stack(hist)


def solve_action(encoding, camera_input, target_flow: Tensor, action0: Tensor, iters: int = 8, damping: float = 1e-6,
                 ray_weight: Optional[Tensor] = None, action_prior_weight: float = 0.0) -> Tuple[Tensor, List[float]]:
    """Levenberg-Marquardt on the action: returns (action (B, A) float32, loss history summed over the batch).
    ``damping`` scales the diagonal of H (relative LM damping, adapted x10 / x0.3 on rejected / accepted steps);
    ``action_prior_weight`` adds lambda |u - action0|^2 (the notebooks' regulariser)."""

    cams = _target_cameras(camera_input, encoding.p.device)
    tgt = target_flow.detach().to(encoding.p.device, torch.float32).contiguous()

    def terms(u):
        t = gauss_newton_terms(encoding, camera_input, u.float(), tgt, ray_weight, _cams=cams)
        return t.H, t.g, t.loss

    u, hist = levenberg_marquardt(terms, action0.to(encoding.p.device), iters, damping, action_prior_weight)
    return u.float(), [float(v) for v in hist.sum(-1).cpu()]
