"""The host mirror end to end on a GPU: njf_b200.Model (encoder + fused render) against the oracle,
through the same calls the reference's notebooks make (forward, patch_render, encode_image,
infer_optical_flow with autograd w.r.t. the action, compute_density)."""
import numpy as np
import pytest
import torch

from helpers import O, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(head, A, s_prop, s_nerf):
    from njf_b200 import model as M, modules as mod

    mlp = mod.MlpCfg()
    dec = (mod.ActionDecoderJacobianTransformerCfg(name=head, mlp=mlp, transformer=mod.TransformerCfg())
           if head == "jacobian_transformer" else mod.ActionDecoderJacobianMlpCfg(name=head, mlp=mlp))
    cfg = M.ModelCfg(action_dim=A, rendering=M.RenderingCfg(tuple(s_prop), s_nerf), encoder=mod.EncoderResnetCfg(),
                     density_decoder=mod.DensityDecoderMlpCfg("density_mlp", mlp), action_decoder=dec)
    m = M.Model(cfg).eval()
    sd = synth.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 31)
    m.load_state_dict(sd)
    return m.to(DEV), sd


def _scene(A, H=24, W=32, rh=10, rw=12):
    g = torch.Generator().manual_seed(8)
    img = torch.rand(1, 3, H, W, generator=g)
    K = synth.normalized_intrinsics(**synth.ALLEGRO_INTRINSICS_PX)[None]
    kpx = K.clone(); kpx[:, 0] *= 640; kpx[:, 1] *= 480
    ctxt, trgt = torch.eye(4)[None], synth.relative_target_pose(1)[None]
    o, d = synth.world_rays(synth.pixel_grid(rh, rw), K[0], trgt[0])
    return dict(img=img, K=K, kpx=kpx, ctxt=ctxt, trgt=trgt, o=o[None], d=d[None], zn=torch.tensor([0.5]),
                zf=torch.tensor([3.0]), act=0.1 * torch.randn(1, A, generator=g))


@pytest.mark.parametrize("head,A", [("jacobian_transformer", 8), ("jacobian_mlp", 6)])
def test_model_api_matches_oracle(head, A):
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    s_prop, s_nerf = (32,), 32
    m, sd = _model(head, A, s_prop, s_nerf)
    sc = _scene(A)
    cam = CameraInput(sc["img"], sc["ctxt"], sc["K"], sc["trgt"], sc["kpx"])      # HOST tensors in
    rin = RenderingInput(sc["o"], sc["d"], sc["zn"], sc["zf"])
    with torch.no_grad():
        out = m.forward(cam, rin, RobotInput(sc["act"]), compute_vis_features=True)
        feat = O.encoder_resnet34(sd, sc["img"])
        ref = O.render_forward(sd, O.FieldSpec(head, A), feat, sc["ctxt"], sc["K"], sc["trgt"], sc["kpx"], sc["o"], sc["d"],
                               sc["zn"], sc["zf"], sc["act"], s_prop, s_nerf)
    so, vo = out.standard_output, out.vis_output
    assert so.rgb.device.type == "cpu"          # results come back where the inputs lived
    np.testing.assert_allclose(so.rgb.numpy(), ref["rgb"].numpy(), atol=4e-3)
    np.testing.assert_allclose(so.depth.numpy(), ref["depth"].numpy(), atol=5e-3)
    jm = float(ref["action_features"].abs().max())
    np.testing.assert_allclose(vo.action_features.numpy(), ref["action_features"].numpy(), atol=4e-2 * jm)
    fm = float(ref["optical_flow"].abs().max())
    np.testing.assert_allclose(so.optical_flow.numpy(), ref["optical_flow"].numpy(), atol=4e-2 * fm + 0.1)
    assert vo.weights.shape == (1, 120, s_nerf) and vo.steps.shape == (1, 120, s_nerf)

    # patch_render: same pixels, reshaped to the frame; the encoder runs once
    pr = m.patch_render(cam, rin, RobotInput(sc["act"]), patch_size=50, render_height=10, render_width=12)
    assert pr.rgb.shape == (1, 10, 12, 3) and pr.action_features.shape == (1, 10, 12, 3 * A)
    assert pr.depth_rgb.shape == (1, 10, 12, 3) and pr.flow_rgb.shape == (1, 10, 12, 3)
    np.testing.assert_allclose(pr.rgb.reshape(1, 120, 3).cpu().numpy(), ref["rgb"].numpy(), atol=4e-3)

    # inverse dynamics: encode once, then differentiate the flow w.r.t. the action
    enc = m.encode_image(cam, rin, RobotInput(sc["act"]))
    assert enc.density.shape == (1, 120, s_nerf, 1) and enc.action_features.shape == (1, 120, s_nerf, 3 * A)
    act = torch.nn.Parameter((0.5 * sc["act"] + 0.02).to(DEV))
    flow = m.infer_optical_flow(enc, cam, RobotInput(act))
    loss = torch.nn.functional.smooth_l1_loss(flow, torch.zeros_like(flow))
    loss.backward()
    a_ref = torch.nn.Parameter(act.detach().cpu().clone())
    flow_ref = O.infer_optical_flow(enc.action_features.cpu(), enc.weights.cpu(), enc.ray_samples_positions.cpu(), a_ref,
                                    sc["trgt"], sc["kpx"])
    torch.nn.functional.smooth_l1_loss(flow_ref, torch.zeros_like(flow_ref)).backward()
    np.testing.assert_allclose(flow.detach().cpu().numpy(), flow_ref.detach().numpy(), atol=2e-2 * float(flow_ref.abs().max()) + 1e-3)
    np.testing.assert_allclose(act.grad.cpu().numpy(), a_ref.grad.numpy(), rtol=5e-2, atol=1e-3 * float(a_ref.grad.abs().max()) + 1e-6)

    # point queries
    pe = m.compute_pixel_encoding(cam, rin, RobotInput(sc["act"]))
    pts = torch.rand(1, 50, 3) * torch.tensor([1.0, 0.8, 2.0]) + torch.tensor([-0.5, -0.4, 0.6])
    dho, extras = m.compute_density(pts, pe)
    with torch.no_grad():
        sig, geo, jac, z, encf = O.field_heads(sd, pts, feat, sc["ctxt"], sc["K"], O.FieldSpec(head, A))
    np.testing.assert_allclose(dho.density.cpu().numpy(), sig.numpy(), rtol=3e-2, atol=2e-3)
    np.testing.assert_allclose(extras["jacobian_head_output"].cpu().numpy(), jac.numpy(), atol=3e-2 * float(jac.abs().max()))
    np.testing.assert_allclose(dho.xyz_features.cpu().numpy(), encf.numpy(), atol=2e-5)


@pytest.mark.parametrize("head,A", [("jacobian_transformer", 8), ("jacobian_mlp", 6)])
def test_inverse_dynamics_gauss_newton(head, A):
    """the normal equations of the notebooks' inverse-dynamics objective from njf_flow_gn_terms
    (a) equal an fp64 autograd evaluation of the collapsed formula on the same (jbar, p) to 1e-4 (the kernel's
    analytic Jacobian is right), (b) agree with the oracle's per-sample reference formulation within the fp16
    Jacobian tolerance, and (c) Levenberg-Marquardt on them recovers a known action from its own flow."""
    from njf_b200 import inverse_dynamics as ID
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    s_prop, s_nerf = (32,), 32
    m, sd = _model(head, A, s_prop, s_nerf)
    sc = _scene(A)
    cam = CameraInput(sc["img"], sc["ctxt"], sc["K"], sc["trgt"], sc["kpx"])
    rin = RenderingInput(sc["o"], sc["d"], sc["zn"], sc["zf"])
    enc = m.encode_image(cam, rin, RobotInput(sc["act"]))
    g = torch.Generator().manual_seed(3)
    u_true = (0.3 * torch.randn(1, A, generator=g)).to(DEV)
    with torch.no_grad():
        target = m.infer_optical_flow(enc, cam, RobotInput(u_true))
    u = (0.5 * u_true + 0.05).contiguous()
    t = ID.gauss_newton_terms(enc, cam, u, target)
    # (a) fp64 autograd on the collapsed formula
    w2c = torch.inverse(sc["trgt"]).double()
    kpx = sc["kpx"].double()
    jb, pp = enc.jbar.cpu().double(), enc.p.cpu().double()

    def flow64(a):
        x = pp + torch.einsum("brad,a->brd", jb.reshape(1, -1, A, 3), a)
        def proj(x_):
            c = torch.einsum("bij,brj->bri", w2c[:, :3, :3], x_) + w2c[:, None, :3, 3]
            k = torch.einsum("bij,brj->bri", kpx, c)
            return k[..., :2] / (k[..., 2:3] + 1e-9)
        return (proj(x) - proj(pp))[0]

    a64 = u[0].cpu().double()
    G = torch.autograd.functional.jacobian(flow64, a64)
    r = flow64(a64) - target[0].cpu().double()
    H64, g64, l64 = torch.einsum("ria,rib->ab", G, G), torch.einsum("ria,ri->a", G, r), (r ** 2).sum()
    hs = float(H64.abs().max())
    np.testing.assert_allclose(t.H[0].cpu().numpy(), H64.numpy(), atol=1e-4 * hs, rtol=1e-4)
    np.testing.assert_allclose(t.g[0].cpu().numpy(), g64.numpy(), atol=1e-4 * float(g64.abs().max()), rtol=1e-3)
    np.testing.assert_allclose(float(t.loss[0]), float(l64), rtol=1e-3)
    # (b) the oracle's reference formulation (per-sample J, weights, positions from the CPU oracle)
    with torch.no_grad():
        feat = O.encoder_resnet34(sd, sc["img"])
        ref = O.render_forward(sd, O.FieldSpec(head, A), feat, sc["ctxt"], sc["K"], sc["trgt"], sc["kpx"], sc["o"], sc["d"],
                               sc["zn"], sc["zf"], sc["act"], s_prop, s_nerf)
    Ho, go, lo = O.flow_gn_terms(ref["jacobian"], ref["weights"][..., None], ref["positions"], u.cpu(), sc["trgt"], sc["kpx"],
                                 target.cpu())
    np.testing.assert_allclose(t.H[0].cpu().numpy(), Ho[0].numpy(), atol=8e-2 * float(Ho.abs().max()))
    # (c) recover the action
    sol, hist = ID.solve_action(enc, cam, target, torch.zeros(1, A), iters=8)
    assert hist[-1] < 1e-6 * max(hist[0], 1e-12) + 1e-8, hist
    np.testing.assert_allclose(sol.cpu().numpy(), u_true.cpu().numpy(), atol=2e-3 * float(u_true.abs().max()) + 1e-4)


def test_training_mode_forward_carries_gradients():
    """Model.forward in .train() mode with every parameter trainable (the perception phase): ModelTrainingOutput is
    filled and the outputs carry autograd history (parity of the gradients: tests/test_gpu_train_trunks.py)."""
    m, _ = _model("jacobian_transformer", 8, (16,), 16)
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    sc = _scene(8)
    m.train()
    out = m.forward(CameraInput(sc["img"], sc["ctxt"], sc["K"], sc["trgt"], sc["kpx"]),
                    RenderingInput(sc["o"], sc["d"], sc["zn"], sc["zf"]), RobotInput(sc["act"]))
    assert out.training_output is not None and len(out.training_output.weights_list) == 2
    assert out.standard_output.rgb.requires_grad and out.standard_output.depth.requires_grad
    out.standard_output.rgb.sum().backward()
    assert m.decoder.color_head[0].weight.grad is not None and m.encoder.model.conv1.weight.grad is not None


@pytest.mark.parametrize("head,A", [("jacobian_transformer", 8), ("jacobian_mlp", 6)])
def test_decoder_level_plugin_methods(head, A):
    """The per-point methods the reference's own Model calls on registry-built decoders (njf_b200.plugin):
    DensityDecoderMlp.get_density, ActionDecoderJacobian.forward / encode_image / compute_density, stand-alone modules
    with their own packed field, against the oracle's per-point formulation."""
    from types import SimpleNamespace

    from njf_b200 import modules as mod

    mlp = mod.MlpCfg()
    dcfg = (mod.ActionDecoderJacobianTransformerCfg(name=head, mlp=mlp, transformer=mod.TransformerCfg())
            if head == "jacobian_transformer" else mod.ActionDecoderJacobianMlpCfg(name=head, mlp=mlp))
    dec = mod.get_action_decoder(dcfg, action_dim=A, encoder_dim=512)
    prop = mod.get_density_decoder(mod.DensityDecoderMlpCfg("density_mlp", mlp), encoder_dim=512)
    w = synth.synth_state_dict(synth.field_shapes(head, A), 31)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in w.items() if k.startswith("decoder.")})
    prop.load_state_dict({k[len("proposal_networks.0."):]: v for k, v in w.items() if k.startswith("proposal_networks.0.")})
    dec, prop = dec.to(DEV).eval(), prop.to(DEV).eval()
    g = torch.Generator().manual_seed(12)
    B, R, S = 2, 9, 7
    feat = torch.randn(B, 512, 12, 16, generator=g).abs() * 0.7
    K = synth.normalized_intrinsics(**synth.ALLEGRO_INTRINSICS_PX)[None].repeat(B, 1, 1)
    ctxt = torch.eye(4)[None].repeat(B, 1, 1)
    xyz = torch.rand(B, R, S, 3, generator=g) * torch.tensor([1.0, 0.8, 2.0]) + torch.tensor([-0.5, -0.4, 0.6])
    dirs = torch.nn.functional.normalize(torch.randn(B, R, S, 3, generator=g), dim=-1)
    act = 0.1 * torch.randn(B, A, generator=g)
    pe = SimpleNamespace(features=feat.to(DEV), extrinsics=ctxt.to(DEV), intrinsics=K.to(DEV), action=act.to(DEV))
    spec = O.FieldSpec(head, A)
    with torch.no_grad():
        out = dec.forward(xyz.to(DEV), dirs.to(DEV), pe)
        enc = dec.encode_image(xyz.to(DEV), pe)
        dho = dec.compute_density(xyz.reshape(B, R * S, 3).to(DEV), pe)
        sig_p = prop.get_density(xyz.to(DEV), pe)
        sig, geo, jac, z, encf = O.field_heads(w, xyz.reshape(B, R * S, 3), feat, ctxt, K, spec)
        rgb = O.color_head(w, geo, O.sh4((dirs.reshape(B, R * S, 3) + 1.0) / 2.0))
        flow = torch.einsum("bnad,ba->bnd", jac.reshape(B, R * S, A, 3), act)
        sig_prop = O.proposal_density(w, 0, xyz.reshape(B, R * S, 3), feat, ctxt, K, spec)
    c = lambda t: t.reshape(B, R * S, -1).cpu().numpy()
    assert out.density.shape == (B, R, S, 1) and out.action_features.shape == (B, R, S, 3 * A)
    np.testing.assert_allclose(c(out.density), sig.numpy(), rtol=3e-2, atol=2e-3)
    np.testing.assert_allclose(c(out.color), rgb.numpy(), atol=4e-3)
    jm = float(jac.abs().max())
    np.testing.assert_allclose(c(out.action_features), jac.numpy(), atol=3e-2 * jm)
    np.testing.assert_allclose(c(out.flow), flow.numpy(), atol=3e-2 * float(flow.abs().max()) + 1e-6)
    np.testing.assert_allclose(c(enc.action_features), jac.numpy(), atol=3e-2 * jm)
    np.testing.assert_allclose(dho.density_features.cpu().numpy(), geo.numpy(), atol=2e-2 * float(geo.abs().max()))
    np.testing.assert_allclose(dho.xyz_features.cpu().numpy(), encf.numpy(), atol=2e-5)
    np.testing.assert_allclose(dho.pixel_aligned_features.cpu().numpy(), z.numpy(), atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(c(sig_p), sig_prop.numpy(), rtol=3e-2, atol=2e-3)


def test_validation_video_renderer():
    """njf_b200.video.render_interpolated_view (models/model_wrapper.py:213-327): the first frame is the target view, the
    last the context view; every frame equals a direct Model.forward at the interpolated camera."""
    from njf_b200 import geometry as G, video as V
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    A, H, W = 8, 12, 16
    m, _ = _model("jacobian_transformer", A, (32,), 32)
    sc = _scene(A)
    K = sc["K"]
    pair = V.CameraPair(ctxt_extrinsics=sc["ctxt"], ctxt_intrinsics=K, trgt_extrinsics=sc["trgt"], trgt_intrinsics=K,
                        trgt_intrinsics_px=V.denormalize_intrinsics(K, W, H), height=H, width=W)
    for graph in (False, True):
        m.cuda_graph = graph
        vid = V.render_interpolated_view(m, sc["img"], pair, sc["act"], sc["zn"], sc["zf"], H, W, num_frames=4)
        assert vid["rgb"].shape == (1, 4, 3, H, W) and vid["depth"].shape == (1, 4, 3, H, W)
        assert vid["optical_flow"].shape == (1, 4, 3, H, W) and vid["pred_flow_t0"].shape == (1, 2, H, W)
        xy, _ = G.get_pixel_coordinates(H, W, device=torch.device(DEV))
        for idx, c2w in ((0, sc["trgt"]), (3, sc["ctxt"])):
            o, d = G.get_world_rays(xy.reshape(1, -1, 2), K.to(DEV), c2w.to(DEV))
            with torch.no_grad():
                out = m.forward(CameraInput(sc["img"].to(DEV), sc["ctxt"].to(DEV), K.to(DEV), c2w.to(DEV),
                                            V.denormalize_intrinsics(K, W, H).to(DEV)),
                                RenderingInput(o, d, sc["zn"].to(DEV), sc["zf"].to(DEV)), RobotInput(sc["act"].to(DEV)))
            ref = out.standard_output.rgb.reshape(1, H, W, 3).permute(0, 3, 1, 2)
            np.testing.assert_allclose(vid["rgb"][:, idx].cpu().numpy(), ref.cpu().numpy(), atol=2e-4)


def test_half_precision_encoder_and_nhwc_hoist():
    """Model.encoder_half: the encoder under fp16 autocast emits the NHWC fp16 map that
    njf_hoist_features_nhwc16 copies into its operand tiles.  The NHWC kernel path is checked bit for bit against the
    NCHW fp32 path on the same (fp16-representable) features; the half encoder against the fp32 / TF32 one at the
    stated tolerance (features 1e-2 of their scale, rgb 5e-3)."""
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    s_prop, s_nerf = (32,), 32
    m, sd = _model("jacobian_transformer", 8, s_prop, s_nerf)
    sc = _scene(8, H=48, W=64)
    img = sc["img"].to(DEV)
    with torch.no_grad():
        f32 = m.encoder(img).float()
        fh = m.encoder.forward_nhwc_half(img)
        assert fh.dtype == torch.float16 and fh.shape == (1, 24, 32, 512) and fh.is_contiguous()
        scale = float(f32.abs().max())
        assert float((fh.permute(0, 3, 1, 2).float() - f32).abs().max()) < 1e-2 * scale
        fld = m.field()
        a = fld.hoist_nhwc16(fh)
        b = fld.hoist(fh.permute(0, 3, 1, 2).float().contiguous())
        assert torch.equal(a, b)
        # two views into slots of a four-view buffer
        fh2 = torch.cat([fh, fh.flip(1)], 0).contiguous()
        nbytes = fld.hoist_nhwc16(fh2, view0=1, n_views_total=4).numel()
        big = fld.hoist_nhwc16(fh2, view0=1, n_views_total=4, maps=torch.zeros(nbytes, dtype=torch.uint8, device=DEV))
        ref = fld.hoist_views(fh2.permute(0, 3, 1, 2).float().contiguous(), 1, 4,
                              maps=torch.zeros(nbytes, dtype=torch.uint8, device=DEV))
        assert torch.equal(big, ref) and int(big.count_nonzero()) > 0
        cam = CameraInput(sc["img"], sc["ctxt"], sc["K"], sc["trgt"], sc["kpx"])
        rin, rob = RenderingInput(sc["o"], sc["d"], sc["zn"], sc["zf"]), RobotInput(sc["act"])
        full = m.forward(cam, rin, rob, compute_vis_features=True)
        m.encoder_half = True
        half = m.forward(cam, rin, rob, compute_vis_features=True)
        m.cuda_graph = True
        graphed = m.forward(cam, rin, rob, compute_vis_features=True)
    np.testing.assert_allclose(half.standard_output.rgb.numpy(), full.standard_output.rgb.numpy(), atol=5e-3)
    np.testing.assert_allclose(half.standard_output.depth.numpy(), full.standard_output.depth.numpy(), atol=2e-2)
    jm = float(full.vis_output.action_features.abs().max())
    np.testing.assert_allclose(half.vis_output.action_features.numpy(), full.vis_output.action_features.numpy(), atol=3e-2 * jm)
    np.testing.assert_allclose(graphed.standard_output.rgb.numpy(), half.standard_output.rgb.numpy(), atol=1e-6)
