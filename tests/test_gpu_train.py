"""Action-phase training through the fused path (VERDICT N1): train-mode forward (stratified jitter,
ModelTrainingOutput), hand-written backward kernels for the cross-attention Jacobian head (csrc/xf_backward.cu), and
the CUDA-graph frame.

Gradient parity: d loss / d theta for every Jacobian-head parameter against torch autograd through the CPU oracle
(fp32 reference formulation, un-folded attention) on the same jitter tables.  Tolerance: relative L2 error per
parameter tensor <= 5e-3 (measured 6e-4 .. 1e-3: the forward that feeds the backward runs fp16 tensor-core operands,
the backward itself is fp32), plus cosine similarity >= 0.9999.
"""
import numpy as np
import pytest
import torch

from helpers import O, synth
from test_gpu_model import DEV, _model, _scene

pytestmark = pytest.mark.gpu


def _freeze_like_model_wrapper(m):
    """models/model_wrapper.py:75-85 (dataset.mode == "action")."""
    m.decoder.freeze_non_action_parameters()
    for name, p in m.named_parameters():
        if "decoder" not in name:
            p.requires_grad = False


def _inputs(A, B, R, seed):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, 24, 32, generator=g)
    K = synth.normalized_intrinsics(**synth.ALLEGRO_INTRINSICS_PX)[None].repeat(B, 1, 1)
    kpx = K.clone(); kpx[:, 0] *= 640; kpx[:, 1] *= 480
    ctxt = torch.eye(4)[None].repeat(B, 1, 1)
    trgt = torch.stack([synth.relative_target_pose(1 + b) for b in range(B)])
    coords = torch.rand(R, 2, generator=g)
    rays = [synth.world_rays(coords, K[b], trgt[b]) for b in range(B)]
    o = torch.stack([r[0] for r in rays]); d = torch.stack([r[1] for r in rays])
    zn, zf = torch.full((B,), 0.5), torch.full((B,), 3.0)
    act = 0.3 * torch.randn(B, A, generator=g)
    target = 2.0 * torch.randn(B, R, 2, generator=g)
    mask = (torch.rand(B, R, generator=g) > 0.3).float()
    return img, K, kpx, ctxt, trgt, o, d, zn, zf, act, target, mask


def _flow_loss(flow, target, mask):
    """model_wrapper.py:148-163."""
    l = 0.01 * torch.nn.functional.mse_loss(flow, target, reduction="none")
    return (l * mask.unsqueeze(-1)).sum() / mask.sum()


@pytest.mark.parametrize("s_prop,s_nerf,B,R", [((32,), 48, 2, 70), ((24,), 160, 1, 37), ((64,), 128, 1, 64)])
def test_action_phase_gradients_vs_oracle_autograd(s_prop, s_nerf, B, R):
    from njf_b200 import train as T
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    head, A = "jacobian_transformer", 8
    m, sd = _model(head, A, s_prop, s_nerf)
    _freeze_like_model_wrapper(m)
    m.train()
    img, K, kpx, ctxt, trgt, o, d, zn, zf, act, target, mask = _inputs(A, B, R, 11 + s_nerf)
    m.jitter_generator = torch.Generator(device=DEV).manual_seed(77)
    out = m.forward(CameraInput(img, ctxt, K, trgt, kpx), RenderingInput(o, d, zn, zf), RobotInput(act))
    flow = out.standard_output.optical_flow
    assert flow.requires_grad and out.training_output is not None
    loss = _flow_loss(flow, target, mask)
    loss.backward()
    got = {n: p.grad.detach().cpu() for n, p in m.named_parameters() if p.requires_grad}
    assert got and all("jacobian" in n for n in got)

    # the oracle on the same jitter tables, autograd w.r.t. the same parameters
    bins0, us = T.stratified_tables(s_prop, s_nerf, B, R, False, DEV, generator=torch.Generator(device=DEV).manual_seed(77))
    w = {k: v.clone() for k, v in sd.items()}
    for n in got:
        w[n] = w[n].clone().requires_grad_(True)
    with torch.no_grad():   # still in train mode: the encoder's BatchNorm uses batch statistics, like the forward above did
        feat = m.encoder(img.to(DEV)).float().cpu()
    ref = O.render_forward(w, O.FieldSpec(head, A), feat, ctxt, K, trgt, kpx, o, d, zn, zf, act, s_prop, s_nerf,
                           bins0=bins0.cpu(), us=[u.cpu() for u in us])
    ref_loss = _flow_loss(ref["optical_flow"], target, mask)
    ref_loss.backward()
    np.testing.assert_allclose(float(loss), float(ref_loss), rtol=5e-2)
    worst = 0.0
    for n, gg in got.items():
        gr = w[n].grad
        assert gr is not None, n
        rel = float((gg - gr).norm() / gr.norm().clamp_min(1e-12))
        cos = float((gg * gr).sum() / (gg.norm() * gr.norm()).clamp_min(1e-20))
        worst = max(worst, rel)
        assert rel < 5e-3 and cos > 0.9999, (n, rel, cos)
    print(f"action-phase gradients, {len(got)} tensors: worst relative L2 error {worst:.2e}")
    # training output: weights / bins of every level (ModelTrainingOutput, model.py:377-382)
    to = out.training_output
    assert len(to.weights_list) == len(s_prop) + 1 and len(to.ray_samples_list) == len(s_prop) + 1
    np.testing.assert_allclose(to.weights_list[0][..., 0].numpy(), ref["prop_weights_0"].detach().numpy(), atol=2e-3)
    np.testing.assert_allclose(to.weights_list[-1][..., 0].numpy(), ref["weights"].detach().numpy(), atol=2e-3)
    np.testing.assert_allclose(to.ray_samples_list[0].spacing_starts[..., 0].cpu().numpy(), bins0[..., :-1].cpu().numpy(), atol=0)
    np.testing.assert_allclose(to.ray_samples_list[-1].spacing_ends[..., 0].cpu().numpy(),
                               ref["final_bins"][..., 1:].detach().numpy(), atol=3e-3)


def test_action_phase_training_steps_reduce_the_flow_loss():
    """A few Adam steps of the reference's action phase (model_wrapper.py:92-96 optimiser, :148-163 loss)."""
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    head, A, s_prop, s_nerf, B, R = "jacobian_transformer", 8, (32,), 32, 2, 128
    m, _ = _model(head, A, s_prop, s_nerf)
    _freeze_like_model_wrapper(m)
    m.train()
    img, K, kpx, ctxt, trgt, o, d, zn, zf, act, target, mask = _inputs(A, B, R, 5)
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=2e-3, weight_decay=1e-5)
    cam, rin, rob = CameraInput(img, ctxt, K, trgt, kpx), RenderingInput(o, d, zn, zf), RobotInput(act)
    losses = []
    for step in range(12):
        m.step_before_iter(step)
        out = m.forward(cam, rin, rob)
        loss = _flow_loss(out.standard_output.optical_flow, target, mask)
        opt.zero_grad()
        loss.backward()
        opt.step()
        m.step_after_iter(step)
        losses.append(float(loss))
    print("flow loss per step:", [f"{l:.5f}" for l in losses])
    assert all(np.isfinite(losses)) and min(losses[-3:]) < losses[0]


def test_train_mode_without_grad_and_perception_phase():
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    m, _ = _model("jacobian_transformer", 8, (16,), 16)
    sc = _scene(8)
    cam = CameraInput(sc["img"], sc["ctxt"], sc["K"], sc["trgt"], sc["kpx"])
    rin, rob = RenderingInput(sc["o"], sc["d"], sc["zn"], sc["zf"]), RobotInput(sc["act"])
    m.train()
    with torch.no_grad():   # jittered forward, no autograd: works with everything trainable
        out = m.forward(cam, rin, rob, compute_vis_features=True)
    assert out.training_output is not None and torch.isfinite(out.standard_output.rgb).all()
    out = m.forward(cam, rin, rob)   # perception phase (all parameters trainable): the trunk-training path (test_gpu_train_trunks.py)
    assert out.standard_output.rgb.requires_grad and out.training_output.weights_list[0].requires_grad


def test_cuda_graph_frame_replays_bit_identically():
    """Model.forward with cuda_graph=True: encoder + hoist + render captured once per frame shape; replays with new
    inputs equal the eager path bit for bit (CUDA inputs: both paths invert the poses on the device)."""
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    m, _ = _model("jacobian_transformer", 8, (32,), 32)
    outs = {}
    for graph in (False, True):
        m.cuda_graph = graph
        res = []
        for seed in (1, 2, 3):
            sc = _scene(8)
            g = torch.Generator().manual_seed(seed)
            dv = lambda t: t.to(DEV)
            cam = CameraInput(dv(torch.rand(1, 3, 24, 32, generator=g)), dv(sc["ctxt"]), dv(sc["K"]),
                              dv(synth.relative_target_pose(seed)[None]), dv(sc["kpx"]))
            rin = RenderingInput(dv(sc["o"]), dv(sc["d"]), dv(sc["zn"]), dv(sc["zf"]))
            with torch.no_grad():
                out = m.forward(cam, rin, RobotInput(dv(0.1 * torch.randn(1, 8, generator=g))), compute_vis_features=True)
            torch.cuda.synchronize()
            res.append([out.standard_output.rgb.clone(), out.standard_output.depth.clone(),
                        out.standard_output.optical_flow.clone(), out.vis_output.action_features.clone()])
        outs[graph] = res
    assert len(m._graphs) == 1
    for a, b in zip(outs[False], outs[True]):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    assert not torch.equal(outs[True][0][0], outs[True][1][0])   # the replays really saw different inputs


def test_arm_mode_uses_the_arm_head():
    """decoder.switch_mode("arm") (action_decoder_jacobian.py:306-313, 331, 439): jacobian_head_arm, a ResnetFC with
    3 * arm_action_dim outputs, replaces the Jacobian head."""
    from njf_b200 import model as M, modules as mod
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    A, A_arm, s_prop, s_nerf = 8, 6, (32,), 32
    mlp = mod.MlpCfg()
    dec = mod.ActionDecoderJacobianTransformerCfg(name="jacobian_transformer", mlp=mlp, transformer=mod.TransformerCfg(),
                                                  use_arm_model=True, arm_action_dim=A_arm)
    cfg = M.ModelCfg(action_dim=A, rendering=M.RenderingCfg(s_prop, s_nerf), encoder=mod.EncoderResnetCfg(),
                     density_decoder=mod.DensityDecoderMlpCfg("density_mlp", mlp), action_decoder=dec)
    m = M.Model(cfg).eval()
    sd = synth.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 31)
    m.load_state_dict(sd)
    m = m.to(DEV)
    m.decoder.switch_mode("arm")
    sc = _scene(A_arm)
    cam = CameraInput(sc["img"], sc["ctxt"], sc["K"], sc["trgt"], sc["kpx"])
    rin = RenderingInput(sc["o"], sc["d"], sc["zn"], sc["zf"])
    with torch.no_grad():
        out = m.forward(cam, rin, RobotInput(sc["act"]), compute_vis_features=True)
        feat = O.encoder_resnet34(sd, sc["img"])
        # the oracle's MLP-head formulation on the arm head's weights
        w = {k: v for k, v in sd.items() if "jacobian" not in k}
        w.update({"decoder.jacobian_head." + k[len("decoder.jacobian_head_arm."):]: v for k, v in sd.items()
                  if k.startswith("decoder.jacobian_head_arm.")})
        ref = O.render_forward(w, O.FieldSpec("jacobian_mlp", A_arm), feat, sc["ctxt"], sc["K"], sc["trgt"], sc["kpx"],
                               sc["o"], sc["d"], sc["zn"], sc["zf"], sc["act"], s_prop, s_nerf)
    assert out.vis_output.action_features.shape[-1] == 3 * A_arm
    jm = float(ref["action_features"].abs().max())
    np.testing.assert_allclose(out.vis_output.action_features.numpy(), ref["action_features"].numpy(), atol=4e-2 * jm)
    np.testing.assert_allclose(out.standard_output.rgb.numpy(), ref["rgb"].numpy(), atol=4e-3)
    m.decoder.switch_mode("regular")
    with torch.no_grad():
        out2 = m.forward(cam, rin, RobotInput(torch.zeros(1, A)), compute_vis_features=True)
    assert out2.vis_output.action_features.shape[-1] == 3 * A


def test_partial_repack_equals_full_repack():
    """After a change of the Jacobian-head parameters only (an optimiser step of the action phase) Model.field() re-packs
    the head in place (njf_field_update_head); the render must equal that of a freshly packed field bit for bit."""
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    m, _ = _model("jacobian_transformer", 8, (32,), 32)
    sc = _scene(8)
    cam = CameraInput(sc["img"], sc["ctxt"], sc["K"], sc["trgt"], sc["kpx"])
    rin, rob = RenderingInput(sc["o"], sc["d"], sc["zn"], sc["zf"]), RobotInput(sc["act"])
    with torch.no_grad():
        before = m.forward(cam, rin, rob, compute_vis_features=True)
        fld = m._field
        g = torch.Generator(device=DEV).manual_seed(1)
        for n, p in m.named_parameters():
            if n.startswith("decoder.") and "jacobian" in n:
                p.add_(0.05 * p.abs().mean() * torch.randn(p.shape, generator=g, device=DEV))
        after = m.forward(cam, rin, rob, compute_vis_features=True)
        assert m._field is fld                                   # updated in place, not rebuilt
        fresh, _ = _model("jacobian_transformer", 8, (32,), 32)
        fresh.load_state_dict(m.state_dict())
        ref = fresh.forward(cam, rin, rob, compute_vis_features=True)
    assert not torch.equal(before.vis_output.action_features, after.vis_output.action_features)
    for a, b in ((after.standard_output.rgb, ref.standard_output.rgb), (after.standard_output.optical_flow, ref.standard_output.optical_flow),
                 (after.vis_output.action_features, ref.vis_output.action_features)):
        assert torch.equal(a, b)
