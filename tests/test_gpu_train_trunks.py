"""Training of the ResnetFC trunks: csrc/trunk_train.cu through njf_b200/train_trunk.py.

* kernel level: linear forward / input gradient / weight gradient, bilinear gather / scatter, per-sample set-up and
  SH-16 against plain torch fp32 restatements of the same op (tolerances: fp32 round-off of a K <= 128 dot product,
  1e-5 relative);
* perception phase (models/model_wrapper.py:116-146): d loss / d theta for EVERY parameter of the proposal networks,
  the density head, the colour head and the encoder against torch autograd through the CPU oracle on the same jitter
  tables -- relative L2 error per tensor <= 2e-3 (both sides are fp32; what remains is summation order and the
  occasional PDF-sample tie that falls the other way);
* action phase with the MLP Jacobian head (model_wrapper.py:75-85, 148-163): the same check for jacobian_head.* (frozen
  trunks render on the fused kernels, only the Jacobian trunk runs on the layer kernels; tolerance 1e-2, measured 6e-3).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import O, synth
from test_gpu_model import DEV, _model
from test_gpu_train import _flow_loss, _freeze_like_model_wrapper, _inputs

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------- kernels
@pytest.mark.parametrize("M,N,K,relu,res", [(1, 4, 4, False, False), (300, 128, 64, False, False), (257, 128, 128, True, True),
                                            (1000, 64, 128, True, False), (129, 16, 128, True, False),
                                            (513, 4, 128, True, False), (77, 64, 32, False, False),
                                            (5000, 20, 128, True, False)])
def test_linear_kernels_vs_torch(M, N, K, relu, res):
    from njf_b200 import train_trunk as TT

    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).requires_grad_(True)
    b = torch.randn(N, generator=g).to(DEV).requires_grad_(True)
    r = torch.randn(M, N, generator=g).to(DEV).requires_grad_(True) if res else None
    gy = torch.randn(M, N, generator=g).to(DEV)
    y = TT._Linear.apply(x, w, b, relu, r)
    got = torch.autograd.grad(y, [x, w, b] + ([r] if res else []), gy)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    rd = r.detach().double().requires_grad_(True) if res else None
    yr = F.linear(torch.relu(xd) if relu else xd, wd, bd) + (rd if res else 0)
    ref = torch.autograd.grad(yr, [xd, wd, bd] + ([rd] if res else []), gy.double())
    np.testing.assert_allclose(y.detach().cpu().numpy(), yr.detach().cpu().numpy(), atol=2e-5, rtol=1e-5)
    for a, c, name in zip(got, ref, "xwbr"):
        scale = float(c.abs().max()) + 1e-12
        np.testing.assert_allclose(a.cpu().numpy(), c.cpu().numpy(), atol=1e-5 * scale + 1e-6, rtol=1e-5, err_msg=name)


@pytest.mark.parametrize("M,N,K,relu,res", [(1, 4, 4, False, False), (300, 128, 64, False, False), (257, 128, 128, True, True),
                                            (1000, 64, 128, True, False), (129, 16, 128, True, False),
                                            (513, 4, 128, True, False), (77, 64, 32, False, False),
                                            (5000, 20, 128, True, False), (40000, 128, 128, True, True),
                                            (20011, 128, 20, True, False), (333, 128, 4, False, False)])
def test_linear_kernels_tf32_tensor_cores_vs_torch(M, N, K, relu, res):
    """The tcgen05 kind::tf32 variants (selected by torch.set_float32_matmul_precision("high"), the reference's
    train.py:64-65 setting) against float64: operands rounded to 10 mantissa bits, fp32 accumulation -> errors of
    ~1e-3 of the result's scale."""
    from njf_b200 import train_trunk as TT

    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).requires_grad_(True)
    b = torch.randn(N, generator=g).to(DEV).requires_grad_(True)
    r = torch.randn(M, N, generator=g).to(DEV).requires_grad_(True) if res else None
    gy = torch.randn(M, N, generator=g).to(DEV)
    prev = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("high")
    try:
        assert TT.tensor_cores()
        y = TT._Linear.apply(x, w, b, relu, r)
        got = torch.autograd.grad(y, [x, w, b] + ([r] if res else []), gy)
    finally:
        torch.set_float32_matmul_precision(prev)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    rd = r.detach().double().requires_grad_(True) if res else None
    yr = F.linear(torch.relu(xd) if relu else xd, wd, bd) + (rd if res else 0)
    ref = torch.autograd.grad(yr, [xd, wd, bd] + ([rd] if res else []), gy.double())
    np.testing.assert_allclose(y.detach().cpu().numpy(), yr.detach().cpu().numpy(), atol=6e-3, rtol=2e-3)
    for a, c, name in zip(got, ref, "xwbr"):
        scale = float(c.abs().max()) + 1e-12
        np.testing.assert_allclose(a.cpu().numpy(), c.cpu().numpy(), atol=3e-3 * scale, rtol=2e-3, err_msg=name)


def test_padded_linear_and_gather_scatter_vs_torch():
    from njf_b200 import train_trunk as TT

    g = torch.Generator().manual_seed(3)
    # widths that are not multiples of 4 (63-wide encoding, 31-wide colour-head input, 3 / 1 / 18 outputs)
    for M, N, K in [(200, 128, 63), (333, 64, 31), (100, 3, 64), (50, 1, 128), (90, 18, 128)]:
        x = torch.randn(M, K, generator=g).to(DEV).requires_grad_(True)
        w = torch.randn(N, K, generator=g).to(DEV).requires_grad_(True)
        b = torch.randn(N, generator=g).to(DEV).requires_grad_(True)
        y = TT.linear(x, w, b, True)
        yr = F.linear(torch.relu(x), w, b)
        assert y.shape == (M, N)
        np.testing.assert_allclose(y.detach().cpu().numpy(), yr.detach().cpu().numpy(), atol=2e-4, rtol=1e-4)
        gy = torch.randn(M, N, generator=g).to(DEV)
        for a, c in zip(torch.autograd.grad(y, [x, w, b], gy), torch.autograd.grad(yr, [x, w, b], gy)):
            np.testing.assert_allclose(a.cpu().numpy(), c.cpu().numpy(), atol=2e-4 * float(c.abs().max()) + 1e-6, rtol=1e-4)
    # gather / scatter
    P, CH, M = 7 * 9 * 2, 384, 1000
    maps = torch.randn(P, CH, generator=g).to(DEV).requires_grad_(True)
    pix = torch.randint(0, P, (M, 4), generator=g, dtype=torch.int32).to(DEV)
    tw = torch.rand(M, 4, generator=g).to(DEV)
    tw[::7, 1] = 0.0
    z = torch.cat(TT._GatherMaps.apply(maps, pix, tw), 1)
    zr = (maps[pix.long()] * tw[..., None]).sum(1)
    np.testing.assert_allclose(z.detach().cpu().numpy(), zr.detach().cpu().numpy(), atol=1e-5, rtol=1e-5)
    gz = torch.randn(M, CH, generator=g).to(DEV)
    (a,), (c,) = torch.autograd.grad(z, [maps], gz), torch.autograd.grad(zr, [maps], gz)
    np.testing.assert_allclose(a.cpu().numpy(), c.cpu().numpy(), atol=2e-4, rtol=1e-4)


def test_sample_setup_and_sh_vs_oracle():
    from njf_b200 import train_trunk as TT

    g = torch.Generator().manual_seed(5)
    B, N, Hf, Wf = 2, 700, 12, 16
    pts = torch.randn(B, N, 3, generator=g) * torch.tensor([0.6, 0.5, 0.4]) + torch.tensor([0.0, 0.0, 1.6])
    K = synth.normalized_intrinsics(**synth.ALLEGRO_INTRINSICS_PX)[None].repeat(B, 1, 1)
    c2w = torch.stack([torch.eye(4), synth.relative_target_pose(2)])
    feat = torch.randn(B, 8, Hf, Wf, generator=g)
    enc, pix, tw = TT.sample_setup(torch.inverse(c2w).to(DEV).contiguous(), K.to(DEV).contiguous(), pts.to(DEV), Hf, Wf)
    z_ref, cam = O.pixel_aligned(pts, c2w, K, feat)
    enc_ref = O.posenc(cam).reshape(B * N, 63)
    # the camera-space point differs by an ulp or two (fma chain vs torch's einsum); column (dim, k) sees that through
    # d sin(2 pi x 2^k) / dx = 2 pi 2^k
    tol = 2e-5 + 2 * np.pi * (2.0 ** np.tile(np.arange(10), 6)) * 1e-6
    err = np.abs(enc[:, :60].cpu().numpy() - enc_ref[:, :60].numpy())
    assert (err <= tol[None]).all(), float((err / tol[None]).max())
    np.testing.assert_allclose(enc[:, 60:63].cpu().numpy(), enc_ref[:, 60:].numpy(), atol=2e-6)
    assert float(enc[:, 63].abs().max()) == 0.0
    fmap = feat.permute(0, 2, 3, 1).reshape(B * Hf * Wf, 8).to(DEV)
    z = (fmap[pix.long()] * tw[..., None]).sum(1)
    np.testing.assert_allclose(z.cpu().numpy(), z_ref.reshape(B * N, 8).numpy(), atol=2e-5)
    d = F.normalize(torch.randn(500, 3, generator=g), dim=-1)
    for conv in ("tcnn", "nerfstudio_torch"):
        for rnd in (True, False):
            sh = TT.sh16(d.to(DEV), conv, rnd)
            np.testing.assert_allclose(sh.cpu().numpy(), O.sh4((d + 1) / 2, rnd, conv).numpy(), atol=1e-6 if not rnd else 1e-3)


# ----------------------------------------------------------------------------- training steps
def _perception_loss(rgb, weights_list, mids_list, target_rgb, target_depth):
    """An rgb MSE plus differentiable functions of every level's weights standing in for the DS-depth, interlevel and
    distortion terms (model_wrapper.py:116-140): any dependence on weights_list exercises the same backward path."""
    loss = F.mse_loss(rgb, target_rgb)
    for w, mid in zip(weights_list, mids_list):
        loss = loss + 0.08 * ((w * mid).sum(-2) - target_depth).pow(2).mean() / len(weights_list) + 0.01 * (w * w).sum(-2).mean()
    return loss


def _grad_report(got, ref, tol):
    worst = ("", 0.0)
    for n, gg in got.items():
        gr = ref[n]
        assert gr is not None, n
        denom = float(gr.norm())
        if denom < 1e-10:
            assert float(gg.norm()) < 1e-8, n
            continue
        rel = float((gg - gr).norm()) / denom
        if rel > worst[1]:
            worst = (n, rel)
        assert rel < tol, (n, rel)
    return worst


@pytest.fixture
def matmul_precision(request):
    prev = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision(request.param)
    yield request.param
    torch.set_float32_matmul_precision(prev)


@pytest.mark.parametrize("head,A,s_prop,s_nerf,B,R,matmul_precision",
                         [("jacobian_transformer", 8, (32,), 48, 2, 70, "highest"),
                          ("jacobian_mlp", 6, (24, 16), 40, 1, 50, "highest"),
                          ("jacobian_transformer", 8, (32,), 48, 2, 70, "high")], indirect=["matmul_precision"])
def test_perception_phase_gradients_vs_oracle_autograd(head, A, s_prop, s_nerf, B, R, matmul_precision):
    """matmul_precision "highest": fp32 SIMT kernels, tolerance 2e-3.  "high" (the reference's training setting):
    the trunk GEMMs run kind::tf32 on the tensor cores and torch's own GEMMs TF32 as well; the CPU oracle stays fp32,
    tolerance 5e-2 relative L2 per tensor (measured 2.5e-2 on lin_in: 10-bit operand mantissas through an 11-layer chain)."""
    from njf_b200 import train as T
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    tf32 = matmul_precision != "highest"
    m, sd = _model(head, A, s_prop, s_nerf)
    m.train()
    img, K, kpx, ctxt, trgt, o, d, zn, zf, act, _, _ = _inputs(A, B, R, 3 + s_nerf)
    g = torch.Generator().manual_seed(9)
    target_rgb, target_depth = torch.rand(B, R, 3, generator=g), 0.5 + 2.0 * torch.rand(B, R, 1, generator=g)
    m.jitter_generator = torch.Generator(device=DEV).manual_seed(123)
    out = m.forward(CameraInput(img, ctxt, K, trgt, kpx), RenderingInput(o, d, zn, zf), RobotInput(act))
    to = out.training_output
    assert out.standard_output.rgb.requires_grad and len(to.weights_list) == len(s_prop) + 1
    mids = [(s.starts + s.ends) / 2 for s in to.ray_samples_list]
    loss = _perception_loss(out.standard_output.rgb, to.weights_list, mids, target_rgb, target_depth)
    loss.backward()
    trained = lambda n: not (n.startswith("decoder.") and "jacobian" in n)
    got = {n: p.grad.detach().cpu().clone() for n, p in m.named_parameters() if p.grad is not None and trained(n)}
    assert any(n.startswith("encoder.") for n in got) and any(n.startswith("proposal_networks.") for n in got)
    assert any("color_head" in n for n in got) and any("density_head" in n for n in got)
    m.zero_grad(set_to_none=True)

    # the oracle on the same jitter tables: decoder / proposal parameters as CPU leaves, the encoder through the
    # model's own (cuDNN) encoder so that its parameter gradients come out of the same backward pass
    bins0, us = T.stratified_tables(s_prop, s_nerf, B, R, False, DEV, generator=torch.Generator(device=DEV).manual_seed(123))
    w = {k: v.detach().cpu().clone() for k, v in m.state_dict().items() if not k.startswith("encoder.")}
    for n in list(w):
        if n in got:
            w[n].requires_grad_(True)
    feat = m.encoder(img.to(DEV)).float().cpu()
    ref = O.render_forward(w, O.FieldSpec(head, A), feat, ctxt, K, trgt, kpx, o, d, zn, zf, act, s_prop, s_nerf,
                           bins0=bins0.cpu(), us=[u.cpu() for u in us])
    near, far = zn[:, None, None], zf[:, None, None]
    ref_w, ref_mid = [], []
    for lvl in range(len(s_prop) + 1):
        b = ref[f"prop_bins_{lvl}"] if lvl < len(s_prop) else ref["final_bins"]
        e = b * far + (1 - b) * near
        ref_mid.append(((e[..., :-1] + e[..., 1:]) / 2)[..., None])
        ref_w.append((ref[f"prop_weights_{lvl}"] if lvl < len(s_prop) else ref["weights"])[..., None])
    ref_loss = _perception_loss(ref["rgb"], ref_w, ref_mid, target_rgb, target_depth)
    ref_loss.backward()
    ref_g = {n: (w[n].grad if n in w else dict(m.named_parameters())[n].grad.detach().cpu()) for n in got}
    np.testing.assert_allclose(float(loss), float(ref_loss), rtol=5e-3 if tf32 else 2e-4)
    np.testing.assert_allclose(out.standard_output.rgb.detach().numpy(), ref["rgb"].detach().numpy(), atol=3e-3 if tf32 else 2e-5)
    np.testing.assert_allclose(to.weights_list[0][..., 0].detach().numpy(), ref["prop_weights_0"].detach().numpy(),
                               atol=2e-3 if tf32 else 2e-5)
    np.testing.assert_allclose(out.standard_output.optical_flow.detach().numpy(), ref["optical_flow"].detach().numpy(),
                               atol=4e-2 * float(ref["optical_flow"].abs().max()) + 1e-3)
    worst = _grad_report(got, ref_g, 5e-2 if tf32 else 2e-3)
    print(f"perception-phase gradients ({head}, matmul precision {matmul_precision}), {len(got)} tensors: worst relative L2 error {worst[1]:.2e} ({worst[0]})")


def test_mlp_head_action_phase_gradients_vs_oracle_autograd():
    from njf_b200 import train as T
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    head, A, s_prop, s_nerf, B, R = "jacobian_mlp", 6, (32,), 48, 2, 60
    m, sd = _model(head, A, s_prop, s_nerf)
    _freeze_like_model_wrapper(m)
    m.train()
    img, K, kpx, ctxt, trgt, o, d, zn, zf, act, target, mask = _inputs(A, B, R, 21)
    m.jitter_generator = torch.Generator(device=DEV).manual_seed(5)
    out = m.forward(CameraInput(img, ctxt, K, trgt, kpx), RenderingInput(o, d, zn, zf), RobotInput(act),
                    compute_vis_features=True)
    loss = _flow_loss(out.standard_output.optical_flow, target, mask)
    loss.backward()
    got = {n: p.grad.detach().cpu() for n, p in m.named_parameters() if p.requires_grad}
    assert got and all("jacobian_head" in n for n in got) and all(v is not None for v in got.values())
    bins0, us = T.stratified_tables(s_prop, s_nerf, B, R, False, DEV, generator=torch.Generator(device=DEV).manual_seed(5))
    w = {k: v.clone() for k, v in sd.items()}
    for n in got:
        w[n] = w[n].clone().requires_grad_(True)
    with torch.no_grad():
        feat = m.encoder(img.to(DEV)).float().cpu()
    ref = O.render_forward(w, O.FieldSpec(head, A), feat, ctxt, K, trgt, kpx, o, d, zn, zf, act, s_prop, s_nerf,
                           bins0=bins0.cpu(), us=[u.cpu() for u in us])
    ref_loss = _flow_loss(ref["optical_flow"], target, mask)
    ref_loss.backward()
    # sample placement, weights and colours of this phase come from the fused (fp16-operand) render, the Jacobian trunk
    # and its gradients from the fp32 layer kernels: tolerances as in the cross-attention head's action-phase test
    np.testing.assert_allclose(float(loss), float(ref_loss), rtol=5e-2)
    jm = float(ref["action_features"].abs().max())
    np.testing.assert_allclose(out.vis_output.action_features.detach().numpy(), ref["action_features"].detach().numpy(),
                               atol=2e-2 * jm)
    assert not out.standard_output.rgb.requires_grad and out.training_output.weights_list[-1].shape == (B, R, s_nerf, 1)
    # measured worst 6.2e-3 (lin_in.weight: its input is the positional encoding up to sin(2 pi 512 x), the tensor most
    # sensitive to the ~0.2 % of samples the fp16 proposal places one bin over)
    worst = _grad_report(got, {n: w[n].grad for n in got}, 1e-2)
    print(f"MLP-head action-phase gradients, {len(got)} tensors: worst relative L2 error {worst[1]:.2e} ({worst[0]})")


def test_perception_phase_training_steps_reduce_the_loss():
    """A few Adam steps with every parameter trainable (model_wrapper.py:87-106 optimiser)."""
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    head, A, s_prop, s_nerf, B, R = "jacobian_transformer", 8, (32,), 32, 2, 128
    m, _ = _model(head, A, s_prop, s_nerf)
    m.train()
    img, K, kpx, ctxt, trgt, o, d, zn, zf, act, _, _ = _inputs(A, B, R, 5)
    g = torch.Generator().manual_seed(2)
    target_rgb, target_depth = torch.rand(B, R, 3, generator=g), 0.5 + 2.0 * torch.rand(B, R, 1, generator=g)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)
    cam, rin, rob = CameraInput(img, ctxt, K, trgt, kpx), RenderingInput(o, d, zn, zf), RobotInput(act)
    losses = []
    for step in range(10):
        m.step_before_iter(step)
        out = m.forward(cam, rin, rob)
        to = out.training_output
        mids = [(s.starts + s.ends) / 2 for s in to.ray_samples_list]
        loss = _perception_loss(out.standard_output.rgb, to.weights_list, mids, target_rgb, target_depth)
        opt.zero_grad()
        loss.backward()
        opt.step()
        m.step_after_iter(step)
        losses.append(float(loss))
    print("perception loss per step:", [f"{l:.5f}" for l in losses])
    assert all(np.isfinite(losses)) and min(losses[-3:]) < losses[0]
    # back in eval mode the fused kernels render with the updated parameters
    m.eval()
    with torch.no_grad():
        ev = m.forward(cam, rin, rob)
    assert torch.isfinite(ev.standard_output.rgb).all()


@pytest.mark.parametrize("name,tol", [("train_transformer", 1e-2), ("train_mlp_2prop", 6e-3)])
def test_training_step_vs_reference_fixture(name, tol):
    """The kernels' training step against the UNMODIFIED REFERENCE's own autograd (tests/golden/train_*.npz, written by
    oracle/make_golden.py train_fixture): same synthetic weights, the reference's encoder output as the feature map,
    the jitter tables the reference drew from its seed (replayed through Model.jitter_tables), the same loss.  Bins,
    weights, rgb and the loss must agree to fp32 round-off, every parameter gradient and the gradient w.r.t. the
    feature map to `tol` of the tensor's RMS (fp32 layer kernels; for the cross-attention fixture the flow term of
    the loss reaches the trunks through Jacobians evaluated by the fused fp16 query kernel, hence the wider bound, and
    the head itself is not differentiated on this path)."""
    import os

    from helpers import GOLDEN
    from njf_b200.model import CameraInput, RenderingInput, RobotInput
    from njf_b200.train import stratified_tables

    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    t = lambda k: torch.from_numpy(np.ascontiguousarray(z[k]))
    head, A = str(z["head"]), int(z["action_dim"])
    s_prop, s_nerf = tuple(int(v) for v in z["s_prop"]), int(z["s_nerf"])
    m, _ = _model(head, A, s_prop, s_nerf)
    m.load_state_dict(synth.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, int(z["wseed"]), "trained"))
    m.train()
    feat = t("feat").to(DEV).requires_grad_(True)
    m.encoder.forward = lambda image: feat          # the reference's encoder output (an input of the fixture)
    o, d = t("origins"), t("dirs")
    B, R = o.shape[:2]
    torch.manual_seed(int(z["seed"]))
    m.jitter_tables = stratified_tables(s_prop, s_nerf, B, R, False, "cpu")
    out = m.forward(CameraInput(t("image"), t("ctxt_c2w"), t("ctxt_k"), t("trgt_c2w"), t("trgt_k_px")),
                    RenderingInput(o, d, t("z_near"), t("z_far")), RobotInput(t("action")))
    to = out.training_output
    for lvl in range(len(s_prop) + 1):
        sb = to.ray_samples_list[lvl]
        b = torch.cat([sb.spacing_starts[..., 0], sb.spacing_ends[..., -1:, 0]], -1)
        np.testing.assert_allclose(b.detach().numpy(), z[f"bins_{lvl}"], rtol=0, atol=2e-5, err_msg=f"bins level {lvl}")
        np.testing.assert_allclose(to.weights_list[lvl][..., 0].detach().numpy(), z[f"weights_{lvl}"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(out.standard_output.rgb.detach().numpy(), z["rgb"], rtol=0, atol=1e-4)
    mids = [(s.starts + s.ends) / 2 for s in to.ray_samples_list]
    loss = F.mse_loss(out.standard_output.rgb, t("target_rgb")) + 0.01 * F.mse_loss(out.standard_output.optical_flow, t("target_flow"))
    for w, mid in zip(to.weights_list, mids):
        loss = loss + 0.08 * ((w * mid).sum(-2) - t("target_depth")).pow(2).mean() / len(mids) + 0.01 * (w * w).sum(-2).mean()
    np.testing.assert_allclose(float(loss.detach()), float(z["loss"]), rtol=2e-3)
    loss.backward()
    params = dict(m.named_parameters())
    stride, worst, checked = int(z["stride"]), ("", 0.0), 0
    for n in z["grad_names"]:
        n = str(n)
        if head == "jacobian_transformer" and (n == "feat" or (n.startswith("decoder.") and "jacobian" in n)):
            # the cross-attention head trains through train._RenderJacobianHead (tests/test_gpu_train.py); on this path it
            # is evaluated without gradient, so the fixture's flow term reaches neither the head nor -- through the
            # head's query MLP -- the feature map (the MLP-head fixture checks d loss / d features in full)
            continue
        g = feat.grad if n == "feat" else params[n].grad
        assert g is not None, n
        g = g.detach().cpu()
        ref_norm = float(z["gnorm/" + n])
        if ref_norm < 1e-12:
            continue
        sub, ref = g.reshape(-1)[::(1 if g.numel() <= 8192 else stride)].numpy(), z["gsub/" + n]
        rel = float(np.linalg.norm(sub - ref)) / (ref_norm * (len(ref) / g.numel()) ** 0.5)
        checked += 1
        if rel > worst[1]:
            worst = (n, rel)
        assert rel < tol, (n, rel)
    print(f"{name}: {checked} gradient tensors vs the reference's autograd, worst error {worst[1]:.2e} of RMS ({worst[0]})")
