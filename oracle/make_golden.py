"""TEST INFRASTRUCTURE: generate tests/golden/*.npz by running the UNMODIFIED reference
(imported from /root/reference through oracle/ref_shim.py) on seeded synthetic inputs.

Run in the build container only:   python oracle/make_golden.py
Weights are not stored: they are re-created from ``njf_b200/synth.py`` (seed + state-dict key; re-exported by ``oracle/synth.py``).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def scene(action_dim: int, rays_hw=(6, 8), img_hw=(24, 32), view=1, near=0.5, far=3.0, seed=2, batch=1):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(batch, 3, *img_hw, generator=g)
    K = synth.normalized_intrinsics(**synth.ALLEGRO_INTRINSICS_PX)[None].repeat(batch, 1, 1)
    ctxt = torch.eye(4)[None].repeat(batch, 1, 1)
    trgt = torch.stack([synth.relative_target_pose(view + b) for b in range(batch)])
    coords = synth.pixel_grid(*rays_hw)
    rays = [synth.world_rays(coords, K[b], trgt[b]) for b in range(batch)]
    o = torch.stack([r[0] for r in rays])
    d = torch.stack([r[1] for r in rays])
    kpx = K.clone()
    kpx[:, 0, :] *= 640
    kpx[:, 1, :] *= 480
    act = 0.1 * torch.randn(batch, action_dim, generator=g)
    zn = torch.full((batch,), near) + 0.05 * torch.arange(batch)
    zf = torch.full((batch,), far) + 0.1 * torch.arange(batch)
    return dict(image=img, ctxt_c2w=ctxt, ctxt_k=K, trgt_c2w=trgt, trgt_k_px=kpx, origins=o, dirs=d,
                z_near=zn, z_far=zf, action=act)


def render_fixture(name, head, action_dim, s_prop, s_nerf, wseed, regime="trained", **scene_kw):
    m = ref_shim.reference_modules()
    cfg = ref_shim.build_reference_cfg(action_dim, head, s_prop, s_nerf)
    model = m.Model(cfg).eval()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synth.synth_state_dict(shapes, wseed, regime))
    sc = scene(action_dim, **scene_kw)
    cam = m.CameraInput(sc["image"], sc["ctxt_c2w"], sc["ctxt_k"], sc["trgt_c2w"], sc["trgt_k_px"])
    rin = m.RenderingInput(sc["origins"], sc["dirs"], sc["z_near"], sc["z_far"])
    rob = m.RobotInput(sc["action"])
    rec = {}
    real_ss = torch.searchsorted

    def spy(*a, **k):
        r = real_ss(*a, **k)
        rec.setdefault("inds", []).append(r.clone())
        return r

    with torch.no_grad():
        feat = model.encoder(sc["image"])
        torch.searchsorted = spy
        try:
            out = model.forward(cam, rin, rob, compute_vis_features=True)
        finally:
            torch.searchsorted = real_ss
        enc = model.encode_image(cam, rin, rob)
        flow2 = model.infer_optical_flow(enc, cam, m.RobotInput(sc["action"] * 0.5 + 0.02))
        # per-sample intermediates through the reference's own sub-calls (model.py:323-351)
        from neural_jacobian_field.models.decoder.action_decoder import PixelEncoding  # type: ignore

        pe = PixelEncoding(features=feat, extrinsics=sc["ctxt_c2w"], intrinsics=sc["ctxt_k"], action=sc["action"])
        rs, pos, dirs, wl, rsl = model.compute_proposal(model.compute_ray_bundle(rin), pe)
        dec = model.decoder.forward(world_space_xyz=pos, world_space_dir=dirs, pixel_encoding=pe)
    so, vo = out.standard_output, out.vis_output
    fix = {k: v.numpy() for k, v in sc.items()}
    fix.update(
        feat=feat.numpy(), head=head, action_dim=action_dim, s_prop=np.array(s_prop), s_nerf=s_nerf,
        wseed=wseed, regime=regime,
        rgb=so.rgb.numpy(), depth=so.depth.numpy(), optical_flow=so.optical_flow.numpy(),
        action_features=vo.action_features.numpy(), steps=vo.steps.numpy(), weights=vo.weights.numpy(),
        ray_positions=vo.ray_positions.numpy(), ray_positions_warped=vo.ray_positions_warped.numpy(),
        final_bins=torch.cat([rs.spacing_starts[..., 0], rs.spacing_ends[..., -1:, 0]], -1).numpy(),
        proposal_weights=wl[-1][..., 0].numpy(), sigma=dec.density.numpy(), jacobian=dec.action_features.numpy(),
        rgb_samples=dec.color.numpy(), positions=pos.numpy(),
        enc_density=enc.density.numpy(), enc_jacobian=enc.action_features.numpy(), enc_weights=enc.weights.numpy(),
        enc_positions=enc.ray_samples_positions.numpy(), action2=(sc["action"] * 0.5 + 0.02).numpy(),
        flow2=flow2.numpy(),
    )
    for i, t in enumerate(rec["inds"]):
        fix[f"inds_{i + 1}"] = t.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **fix)
    print(name, {k: getattr(v, "shape", v) for k, v in fix.items() if k in ("rgb", "feat", "sigma", "inds_1")})


def pdf_fixture():
    """Reference PDFSampler (rendering/ray_samplers.py:326-451) in eval mode on hand-made histograms."""
    ref_shim.install()
    from neural_jacobian_field.rendering import ray_samplers as rsm  # type: ignore

    g = torch.Generator().manual_seed(5)
    cases = {}
    for s_in, s_out in ((16, 24), (64, 64), (128, 128), (256, 256), (48, 32)):
        R = 64
        wts = torch.rand(R, s_in, generator=g) ** 4
        wts[0] = 0.0                      # zero-weight ray -> eps padding branch
        wts[1] = 0.0
        wts[1, s_in // 3] = 1.0           # single spike
        wts[2] = 1.0 / s_in               # flat
        wts[3, : s_in // 2] = 0.0         # half empty
        wts[4] = torch.linspace(0, 1, s_in) * 1e-4   # tiny but non-zero
        bundle = rsm.RayBundle(origins=torch.zeros(R, 3), directions=torch.ones(R, 3),
                               nears=torch.full((R, 1), 0.5), fars=torch.full((R, 1), 3.0))
        uni = rsm.UniformSampler().eval()
        rs0 = uni(bundle, num_samples=s_in)
        pdf = rsm.PDFSampler(include_original=False).eval()
        rec = []
        real_ss = torch.searchsorted

        def spy(*a, **k):
            r = real_ss(*a, **k)
            rec.append(r.clone())
            return r

        torch.searchsorted = spy
        try:
            rs1 = pdf(bundle, rs0, wts[..., None], num_samples=s_out)
        finally:
            torch.searchsorted = real_ss
        tag = f"{s_in}_{s_out}"
        cases[f"w_{tag}"] = wts.numpy()
        cases[f"bins_in_{tag}"] = torch.cat([rs0.spacing_starts[..., 0], rs0.spacing_ends[..., -1:, 0]], -1).numpy()
        cases[f"bins_out_{tag}"] = torch.cat([rs1.spacing_starts[..., 0], rs1.spacing_ends[..., -1:, 0]], -1).numpy()
        cases[f"inds_{tag}"] = rec[0].numpy()
        cases[f"starts_{tag}"] = rs1.starts[..., 0].numpy()
        cases[f"ends_{tag}"] = rs1.ends[..., 0].numpy()
        # transmittance weights of the reference for sigma = wts*20 on the input samples
        cases[f"tw_{tag}"] = rs0.get_weights(wts[..., None] * 20.0)[..., 0].numpy()
        cases[f"deltas_{tag}"] = rs0.deltas[..., 0].numpy()
    np.savez_compressed(os.path.join(OUT, "pdf_sampler.npz"), **cases)
    print("pdf_sampler", len(cases))


def rays_fixture():
    """Reference ray generation (rendering/geometry.py:117-134 get_pixel_coordinates, :170-203
    get_world_rays_with_z) for two cameras of the Allegro rig shape on a 9x13 grid and a 400x400 grid corner."""
    ref_shim.install()
    from neural_jacobian_field.rendering import geometry as geo  # type: ignore

    K = synth.normalized_intrinsics(**synth.ALLEGRO_INTRINSICS_PX)[None].repeat(2, 1, 1)
    K[1, 0, 0] *= 1.07   # a second, slightly different camera
    K[1, 1, 2] += 0.013
    c2w = torch.stack([synth.relative_target_pose(1), synth.relative_target_pose(3)])
    out = dict(k_norm=K.numpy(), c2w=c2w.numpy())
    for tag, (h, w) in (("small", (9, 13)), ("full", (400, 400))):
        xy, sel = geo.get_pixel_coordinates(h, w)
        coords = xy.reshape(1, -1, 2).repeat(2, 1, 1)
        o, d, z = geo.get_world_rays_with_z(coords, K, c2w)
        o2, d2 = geo.get_world_rays(coords, K, c2w)
        assert torch.equal(o, o2) and torch.equal(d, d2)
        keep = slice(None) if tag == "small" else slice(0, 4096)   # keep the fixture small
        out.update({f"xy_{tag}": xy.numpy() if tag == "small" else xy.reshape(-1, 2)[keep].numpy(),
                    f"sel_{tag}": sel.numpy() if tag == "small" else sel.reshape(-1, 2)[keep].numpy(),
                    f"origins_{tag}": o[:, keep].numpy(), f"dirs_{tag}": d[:, keep].numpy(), f"z_{tag}": z[:, keep].numpy()})
    np.savez_compressed(os.path.join(OUT, "rays.npz"), **out)
    print("rays", {k: v.shape for k, v in out.items()})


def joint_sensitivity_fixture():
    """Reference inference/jacobian_color_map.py:53-109 (compute_joint_sensitivity, visualize_joint_sensitivity).
    The module imports matplotlib / cv2 at the top without using them in these functions: empty stand-ins."""
    import types

    ref_shim.install()
    for name in ("matplotlib", "matplotlib.pyplot", "cv2"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    from neural_jacobian_field.inference import jacobian_color_map as ref  # type: ignore

    g = torch.Generator().manual_seed(0)
    J = torch.randn(2, 5, 7, 24, generator=g)
    Q, _ = torch.linalg.qr(torch.randn(2, 3, 3, generator=g))
    E = torch.eye(4)[None].repeat(2, 1, 1)
    E[:, :3, :3] = Q
    E[:, :3, 3] = torch.randn(2, 3, generator=g)
    s0 = ref.compute_joint_sensitivity(J, None, 0)
    s1 = ref.compute_joint_sensitivity(J, E[0], 1)
    s2 = ref.compute_joint_sensitivity(J, E[:, None, None, None], 0)
    cm = torch.tensor(ref.JACOBIAN_COLORMAP["model_allegro"]).T
    np.savez_compressed(os.path.join(OUT, "joint_sensitivity.npz"), J=J.numpy(), E=E.numpy(), s0=s0.numpy(), s1=s1.numpy(),
                        s2=s2.numpy(), img0=ref.visualize_joint_sensitivity(s0, cm), cm=cm.numpy())
    print("joint_sensitivity")


def train_loss(rgb, flow, weights_list, mids_list, target_rgb, target_depth, target_flow):
    """The scalar both sides differentiate: an rgb MSE, the flow MSE of the action phase (model_wrapper.py:148-163) and
    two functions of every level's weights standing in for the DS-depth / interlevel / distortion terms (:116-140;
    nerfstudio's losses are not installed here) -- every trainable tensor of the model receives a gradient."""
    loss = torch.nn.functional.mse_loss(rgb, target_rgb) + 0.01 * torch.nn.functional.mse_loss(flow, target_flow)
    for w, mid in zip(weights_list, mids_list):
        loss = loss + 0.08 * ((w * mid).sum(-2) - target_depth).pow(2).mean() / len(weights_list) + 0.01 * (w * w).sum(-2).mean()
    return loss


def train_fixture(name, head, action_dim, s_prop, s_nerf, wseed, seed=1234, stride=53, single_jitter=False, **scene_kw):
    """TRAIN-MODE forward + backward of the unmodified reference (Model.train(): stratified jitter from torch's global
    CPU generator after torch.manual_seed(seed), ray_samplers.py:219-233, 389-401): outputs, the jittered bins and the
    gradient of ``train_loss`` w.r.t. every decoder / proposal-network parameter and w.r.t. the encoder output (norm
    + the whole tensor up to 8192 elements, else every ``stride``-th element)."""
    m = ref_shim.reference_modules()
    cfg = ref_shim.build_reference_cfg(action_dim, head, s_prop, s_nerf, single_jitter)
    model = m.Model(cfg).train()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synth.synth_state_dict(shapes, wseed, "trained"))
    sc = scene(action_dim, **scene_kw)
    B, R = sc["origins"].shape[:2]
    g = torch.Generator().manual_seed(seed + 1)
    target_rgb, target_depth = torch.rand(B, R, 3, generator=g), 0.5 + 2.0 * torch.rand(B, R, 1, generator=g)
    target_flow = 2.0 * torch.randn(B, R, 2, generator=g)
    # the encoder runs once, outside the seeded region, in eval mode (its output is an input of the fixture)
    model.encoder.eval()
    with torch.no_grad():
        feat = model.encoder(sc["image"])
    feat = feat.clone().requires_grad_(True)
    model.encoder.forward = lambda image: feat          # the render path sees the stored features (a leaf with a gradient)
    cam = m.CameraInput(sc["image"], sc["ctxt_c2w"], sc["ctxt_k"], sc["trgt_c2w"], sc["trgt_k_px"])
    rin = m.RenderingInput(sc["origins"], sc["dirs"], sc["z_near"], sc["z_far"])
    torch.manual_seed(seed)
    out = model.forward(cam, rin, m.RobotInput(sc["action"]))
    to = out.training_output
    mids = [(rs.starts + rs.ends) / 2 for rs in to.ray_samples_list]
    loss = train_loss(out.standard_output.rgb, out.standard_output.optical_flow, to.weights_list, mids, target_rgb,
                      target_depth, target_flow)
    loss.backward()
    fix = {k: v.numpy() for k, v in sc.items()}
    fix.update(feat=feat.detach().numpy(), head=head, action_dim=action_dim, s_prop=np.array(s_prop), s_nerf=s_nerf, wseed=wseed,
               seed=seed, stride=stride, single_jitter=int(single_jitter), target_rgb=target_rgb.numpy(), target_depth=target_depth.numpy(),
               target_flow=target_flow.numpy(), loss=float(loss), rgb=out.standard_output.rgb.detach().numpy(),
               depth=out.standard_output.depth.detach().numpy(), optical_flow=out.standard_output.optical_flow.detach().numpy())
    for i, (w, rs) in enumerate(zip(to.weights_list, to.ray_samples_list)):
        fix[f"weights_{i}"] = w[..., 0].detach().numpy()
        fix[f"bins_{i}"] = torch.cat([rs.spacing_starts[..., 0], rs.spacing_ends[..., -1:, 0]], -1).detach().numpy()
    names = []
    grads = {"feat": feat.grad}
    if not single_jitter:   # the single-jitter fixture pins the sampling tables only
        grads.update({n: p.grad for n, p in model.named_parameters() if not n.startswith("encoder.")})
    for n, gr in grads.items():
        assert gr is not None, n
        names.append(n)
        fix["gnorm/" + n] = float(gr.norm())
        fix["gsub/" + n] = gr.reshape(-1)[::(1 if gr.numel() <= 8192 else stride)].numpy().copy()
    fix["grad_names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **fix)
    print(name, "loss", float(loss), len(names), "gradient tensors")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        train_fixture("train_transformer", "jacobian_transformer", 8, (16,), 24, wseed=11)
        train_fixture("train_mlp_2prop", "jacobian_mlp", 6, (16, 12), 20, wseed=5, batch=2, rays_hw=(4, 6))
        train_fixture("train_single_jitter", "jacobian_mlp", 6, (16, 12), 20, wseed=5, seed=77, single_jitter=True, rays_hw=(3, 4))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "rays":
        rays_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "joint_sensitivity":
        joint_sensitivity_fixture()
        sys.exit(0)
    pdf_fixture()
    rays_fixture()
    joint_sensitivity_fixture()
    render_fixture("render_transformer", "jacobian_transformer", 8, (16,), 24, wseed=11)
    render_fixture("render_mlp", "jacobian_mlp", 6, (16,), 24, wseed=12)
    render_fixture("render_transformer_2prop_b2", "jacobian_transformer", 8, (16, 12), 16, wseed=13, batch=2,
                   rays_hw=(4, 6))
    render_fixture("render_transformer_initlike", "jacobian_transformer", 8, (32,), 32, wseed=14,
                   regime="init_like", rays_hw=(4, 4), view=0)
    train_fixture("train_transformer", "jacobian_transformer", 8, (16,), 24, wseed=11)
    train_fixture("train_mlp_2prop", "jacobian_mlp", 6, (16, 12), 20, wseed=5, batch=2, rays_hw=(4, 6))
    train_fixture("train_single_jitter", "jacobian_mlp", 6, (16, 12), 20, wseed=5, seed=77, single_jitter=True, rays_hw=(3, 4))
