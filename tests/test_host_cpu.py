"""CPU-only checks of the host-side mirror and the C-ABI library (no compute calls)."""
import os
import re

import pytest
import torch

from helpers import ROOT, synth


def _cfg(head="jacobian_transformer", A=8, s_prop=(16,), s_nerf=24):
    from njf_b200 import model as M, modules as mod

    mlp = mod.MlpCfg()
    if head == "jacobian_transformer":
        dec = mod.ActionDecoderJacobianTransformerCfg(name=head, mlp=mlp, transformer=mod.TransformerCfg())
    else:
        dec = mod.ActionDecoderJacobianMlpCfg(name=head, mlp=mlp)
    return M.ModelCfg(action_dim=A, rendering=M.RenderingCfg(tuple(s_prop), s_nerf), encoder=mod.EncoderResnetCfg(),
                      density_decoder=mod.DensityDecoderMlpCfg("density_mlp", mlp), action_decoder=dec)


def test_library_exports_every_declared_symbol():
    from njf_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "njf_b200.h")).read()
    names = set(re.findall(r"\b(njf_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 14
    L = _lib.lib()
    for n in sorted(names):
        assert hasattr(L, n), f"libnjf_b200.so does not export {n}"
    assert L.njf_version() >= 100


@pytest.mark.parametrize("head,A", [("jacobian_transformer", 8), ("jacobian_mlp", 6)])
def test_state_dict_keys_are_the_reference_contract(head, A):
    from njf_b200.model import Model

    m = Model(_cfg(head, A))
    sd = m.state_dict()
    hot = {k: tuple(v.shape) for k, v in sd.items() if not k.startswith("encoder.")}
    assert hot == synth.field_shapes(head, A)
    assert "encoder.model.conv1.weight" in sd and "encoder.model.layer3.5.bn2.running_var" in sd
    if os.path.isdir("/root/reference/project"):  # build container only: compare with the real thing
        import sys
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_shim

        ref = ref_shim.reference_modules().Model(ref_shim.build_reference_cfg(A, head, (16,), 24))
        rsd = ref.state_dict()
        assert {k: tuple(v.shape) for k, v in rsd.items()} == {k: tuple(v.shape) for k, v in sd.items()}
        m.load_state_dict(rsd, strict=True)


def test_freeze_and_schedule_hooks():
    from njf_b200.model import Model

    m = Model(_cfg())
    n = m.decoder.freeze_non_action_parameters()
    assert n > 0
    assert all(p.requires_grad == ("jacobian" in k) for k, p in m.decoder.named_parameters())
    m.step_before_iter(500)
    b, f = 10.0, 0.5
    assert abs(m._anneal - (b * f) / ((b - 1) * f + 1)) < 1e-9
    m.step_after_iter(500)
    assert m._step == 500


def test_no_cpu_fallback():
    from njf_b200 import _lib
    from njf_b200.model import CameraInput, Model, RenderingInput, RobotInput

    m = Model(_cfg()).eval()
    cam = CameraInput(torch.rand(1, 3, 16, 24), torch.eye(4)[None], torch.eye(3)[None], torch.eye(4)[None], torch.eye(3)[None])
    with pytest.raises(_lib.NjfError):
        m.forward(cam, RenderingInput(torch.zeros(1, 4, 3), torch.ones(1, 4, 3), torch.tensor([0.5]), torch.tensor([2.0])),
                  RobotInput(torch.zeros(1, 8)))


def test_unsupported_configs_fail_loudly():
    from njf_b200 import modules as mod

    with pytest.raises(NotImplementedError):
        mod.DensityDecoderMlp(mod.DensityDecoderMlpCfg("density_mlp", mod.MlpCfg(n_blocks=4)), 512)
    with pytest.raises(NotImplementedError):
        mod.get_action_decoder(type("C", (), {"name": "flow_mlp"})(), 8, 512)


def test_gelu_fit():
    """The one-MUFU GELU of xf_kernel (csrc/xf_head.cu, coefficients from tools/fit_gelu.py): fp32 evaluation
    of the committed coefficients against the exact-erf nn.GELU in float64."""
    import numpy as np
    from scipy.special import erf

    src = open(os.path.join(ROOT, "neural-jacobian-field_b200", "csrc", "xf_head.cu")).read()
    m = re.search(r"constexpr float c0 = ([^;]+);", src)
    c = np.array([float(x.split("=")[-1].strip().rstrip("f")) for x in m.group(0)[len("constexpr float "):-1].split(",")],
                 dtype=np.float32)
    assert c.shape == (7,)
    v = np.linspace(-12, 12, 600001).astype(np.float32)
    a = np.minimum(np.abs(v), np.float32(6.0))
    q = np.full_like(a, c[6])
    for k in range(5, -1, -1):
        q = q * a + c[k]
    g = np.maximum(v, 0) - np.float32(0.5) * (np.abs(v) * np.exp2(-q))
    ref = 0.5 * v.astype(np.float64) * (1 + erf(v.astype(np.float64) / np.sqrt(2)))
    assert np.abs(g - ref).max() < 5e-7


def test_product_path_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing in the shipped package may import it, and bench.py reaches it
    only through oracle_rays_per_s (cpu-baseline / reference legs) and the reference-arm ray generation."""
    import ast

    pkg = os.path.join(ROOT, "neural-jacobian-field_b200", "njf_b200")
    for fn in sorted(os.listdir(pkg)):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n.split(".")[0] in ("njf_oracle", "ref_shim", "make_golden") for n in names), (fn, names)
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    for fn_node in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        uses = [n for n in ast.walk(fn_node) if isinstance(n, ast.Import) and any(a.name in ("njf_oracle", "synth") for a in n.names)]
        if uses:
            assert fn_node.name in ("oracle_rays_per_s", "scene"), fn_node.name


def test_levenberg_marquardt_driver_cpu():
    """The LM driver of njf_b200.inverse_dynamics on a CPU stand-in for the normal-equation kernel: a batch of
    small non-linear least-squares problems (perspective-like residuals) built with torch autograd."""
    from njf_b200.inverse_dynamics import levenberg_marquardt

    g = torch.Generator().manual_seed(0)
    B, N, A = 3, 40, 5
    M = torch.randn(B, N, 2, A, generator=g, dtype=torch.float64)
    c = 3.0 + torch.rand(B, N, 1, generator=g, dtype=torch.float64)
    q = 0.1 * torch.randn(B, N, A, generator=g, dtype=torch.float64)
    u_true = 0.3 * torch.randn(B, A, generator=g, dtype=torch.float64)

    def f(u):  # (B,A) -> (B,N,2): linear map divided by a depth-like affine term
        return torch.einsum("bnia,ba->bni", M, u) / (c + torch.einsum("bna,ba->bn", q, u)[..., None])

    target = f(u_true)

    def terms(u):
        Hs, gs, ls = [], [], []
        for b in range(B):
            fb = lambda ub: (torch.einsum("nia,a->ni", M[b], ub) / (c[b] + (q[b] @ ub)[:, None]))
            G = torch.autograd.functional.jacobian(fb, u[b])          # (N,2,A)
            r = fb(u[b]) - target[b]
            Hs.append(torch.einsum("nia,nib->ab", G, G)); gs.append(torch.einsum("nia,ni->a", G, r)); ls.append((r ** 2).sum())
        return torch.stack(Hs), torch.stack(gs), torch.stack(ls)

    u, hist = levenberg_marquardt(terms, torch.zeros(B, A), iters=10)
    assert hist.shape == (11, B)
    assert bool((hist[1:] <= hist[:-1] + 1e-18).all())                # accepted cost never increases
    assert float(hist[-1].max()) < 1e-16 * max(float(hist[0].max()), 1.0) + 1e-20
    assert float((u - u_true).abs().max()) < 1e-8
    # a prior pulls towards the start and is honoured in the reported cost
    u2, hist2 = levenberg_marquardt(terms, torch.zeros(B, A), iters=10, prior_weight=1e3)
    assert float(u2.abs().max()) < float(u.abs().max())


def test_joint_sensitivity_matches_reference():
    """njf_b200.visualization against outputs of the reference's inference/jacobian_color_map.py
    (tests/golden/joint_sensitivity.npz; generated with matplotlib / cv2 stubbed)."""
    import numpy as np
    from njf_b200 import visualization as V

    z = np.load(os.path.join(ROOT, "tests", "golden", "joint_sensitivity.npz"))
    J, E, cm = torch.from_numpy(z["J"]), torch.from_numpy(z["E"]), torch.from_numpy(z["cm"])
    np.testing.assert_allclose(V.compute_joint_sensitivity(J, None, 0).numpy(), z["s0"], atol=1e-6)
    np.testing.assert_allclose(V.compute_joint_sensitivity(J, E[0], 1).numpy(), z["s1"], atol=1e-6)
    np.testing.assert_allclose(V.compute_joint_sensitivity(J, E[:, None, None, None], 0).numpy(), z["s2"], atol=1e-6)
    img = V.visualize_joint_sensitivity(torch.from_numpy(z["s0"]), cm)
    assert img.dtype == np.uint8 and int(np.abs(img.astype(int) - z["img0"].astype(int)).max()) <= 1
    assert np.allclose(np.array(V.JACOBIAN_COLORMAP["model_allegro"]).T, z["cm"])


def test_c_abi_argument_errors_without_a_gpu():
    """Argument validation of the C ABI happens before any CUDA call: every bad call returns non-zero and leaves a
    message in njf_last_error() (the Python layer turns it into NjfError) -- checked here without a device."""
    import ctypes

    from njf_b200 import _lib, api

    L = api._declare()
    err = lambda: L.njf_last_error().decode()
    # unsupported encoder width / head / action_dim are rejected by njf_field_create before it touches the device
    h = ctypes.c_void_p()
    arr = (api.NjfTensor * 1)()
    for desc, frag in ((api.NjfFieldDesc(api.HEADS["jacobian_transformer"], 8, 1, 256, 1), "encoder_dim"),
                       (api.NjfFieldDesc(api.HEADS["jacobian_transformer"], 9, 1, 512, 1), "action_dim"),
                       (api.NjfFieldDesc(7, 8, 1, 512, 1), "head"),
                       (api.NjfFieldDesc(api.HEADS["jacobian_mlp"], 6, 9, 512, 1), "n_proposal")):
        assert L.njf_field_create(ctypes.byref(desc), arr, 0, ctypes.byref(h)) != 0
        assert frag in err(), err()
    # a missing tensor is named
    desc = api.NjfFieldDesc(api.HEADS["jacobian_mlp"], 6, 1, 512, 1)
    assert L.njf_field_create(ctypes.byref(desc), arr, 0, ctypes.byref(h)) != 0 and "missing tensor" in err()
    with pytest.raises(_lib.NjfError):
        _lib.check(L.njf_field_create(ctypes.byref(desc), arr, 0, ctypes.byref(h)))
    with pytest.raises(_lib.NjfError):
        api.Field("flow_mlp", 8, 1, {})                       # ablation decoder: no kernel, fails loudly
    # null / inconsistent arguments of the stand-alone entry points
    assert L.njf_make_rays(None, None, None, 1, 4, 2, 2, None, None, None, None) != 0 and "null" in err()
    assert L.njf_pdf_sample(None, None, 0, None, 0, 1, 8, 8, 1.0, 8, None, None, None) != 0
    assert L.njf_flow_gn_terms(None, None, None, None, None, None, None, 10, 10, 6, None, None, None, None, None) != 0
    assert L.njf_flow_gn_workspace_doubles(3) > 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/project"), reason="needs the reference checkout (build container)")
@pytest.mark.parametrize("head,A", [("jacobian_transformer", 8), ("jacobian_mlp", 6)])
def test_registration_into_the_reference_registries(head, A):
    """njf_b200.plugin.register_into_reference(): the reference's own factories / Model then build B200 decoders
    (models/decoder/__init__.py:11-44), with the reference's state-dict keys; unregister restores its classes."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shim
    from njf_b200 import modules as mod, plugin

    ref_model = ref_shim.reference_modules()
    import neural_jacobian_field.models.decoder as ref_dec

    cfg = ref_shim.build_reference_cfg(A, head, (16,), 24)
    plain = ref_model.Model(cfg)
    own_flow = ref_dec.ACTION_DECODERS["flow_mlp"]
    assert plugin.register_into_reference()
    try:
        assert ref_dec.ACTION_DECODERS[head] is mod.ACTION_DECODERS[head]
        assert ref_dec.DENSITY_DECODERS["density_mlp"] is mod.DensityDecoderMlp
        assert ref_dec.ACTION_DECODERS["flow_mlp"] is own_flow          # no B200 kernel: the reference keeps its own
        m = ref_model.Model(cfg)                                         # the REFERENCE's Model, built from B200 decoders
        assert isinstance(m.decoder, mod.ActionDecoderJacobian) and isinstance(m.proposal_networks[0], mod.DensityDecoderMlp)
        assert m.decoder.action_dim == A
        m.load_state_dict(plain.state_dict(), strict=True)               # same parameter names and shapes
        for name in ("forward", "encode_image", "compute_density", "freeze_non_action_parameters", "switch_mode"):
            assert callable(getattr(m.decoder, name))
        assert callable(m.proposal_networks[0].get_density)
        with pytest.raises(Exception):                                   # no CUDA device here: fails loudly, no CPU fallback
            m.proposal_networks[0].get_density(torch.zeros(1, 2, 3, 3), None)
    finally:
        plugin.unregister_from_reference()
    assert ref_dec.ACTION_DECODERS[head] is not mod.ACTION_DECODERS[head]


def test_pose_interpolation_and_camera_rig_conventions():
    """njf_b200.video: interpolate_pose against the reference's scipy-based one, and CameraRig against
    the dataset's conventions (convention.post_process_camera_to_world, DatasetCommon.get_relative_transform, intrinsics
    normalisation) on the real Allegro rig file -- when the reference checkout is present; analytic checks otherwise."""
    import json
    import math

    from njf_b200 import video as V

    g = torch.Generator().manual_seed(0)
    a, b = synth.relative_target_pose(1), synth.relative_target_pose(4)
    assert torch.allclose(V.interpolate_pose(a, b, 0.0), a, atol=1e-6) and torch.allclose(V.interpolate_pose(a, b, 1.0), b, atol=1e-5)
    mid = V.interpolate_pose(a, b, 0.5)
    assert torch.allclose(mid[:3, :3] @ mid[:3, :3].T, torch.eye(3), atol=1e-5) and torch.allclose(mid[:3, 3], (a[:3, 3] + b[:3, 3]) / 2, atol=1e-6)
    ref_root = "/root/reference"
    if not os.path.isdir(ref_root):
        return
    import sys
    import types
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shim
    ref_shim.install()
    from neural_jacobian_field.utils import convention as C
    from neural_jacobian_field.visualization.view_interpolation import interpolate_pose as ref_interp

    for t in (0.0, 0.13, 0.5, 0.87, 1.0):
        assert torch.allclose(V.interpolate_pose(a, b, t), ref_interp(a, b, t), atol=2e-6)
    cams = json.load(open(os.path.join(ref_root, "notebooks/real_world/dataset_configs/allegro_config.json")))["cameras"]
    rig = V.CameraRig(cams)
    assert len(rig) == 12
    raw = torch.tensor([c["transform_matrix"] for c in cams], dtype=torch.float32)
    for i in (0, 5, 11):
        assert torch.equal(rig.c2w[i], C.post_process_camera_to_world(raw[i]))
    pair = rig.pair(2, 7)
    inv = torch.inverse(rig.c2w[2])                      # dataset.py:321-327 get_relative_transform
    assert torch.allclose(pair.ctxt_extrinsics[0], torch.eye(4), atol=1e-5)
    assert torch.equal(pair.trgt_extrinsics[0], torch.einsum("ij,jk->ik", inv, rig.c2w[7]))
    k = torch.eye(3); k[0, 0], k[1, 1], k[0, 2], k[1, 2] = cams[7]["fl_x"], cams[7]["fl_y"], cams[7]["cx"], cams[7]["cy"]
    k[:2] /= torch.tensor([cams[7]["w"], cams[7]["h"]])[:, None].float()   # dataset.py:283-294
    assert torch.allclose(pair.trgt_intrinsics[0], k)
    assert torch.allclose(pair.trgt_intrinsics_px, C.denormalize_intrinsics(pair.trgt_intrinsics, width=640, height=480))


def test_trunk_training_host_pieces():
    """The torch-level pieces of njf_b200/train_trunk.py and precise.py that need no kernel (transmittance weights,
    projection, trunc_exp backward, state-dict trunk views) against the oracle, the argument validation of the
    training entry points, and the absence of a CPU path."""
    import ctypes

    from helpers import O
    from njf_b200 import _lib, precise, train_trunk as TT
    from njf_b200.model import CameraInput, Model, RenderingInput, RobotInput

    g = torch.Generator().manual_seed(0)
    deltas = torch.rand(2, 5, 9, 1, generator=g) * 0.1
    deltas[0, 0, 3] = 0.0
    sig = torch.rand(2, 5, 9, 1, generator=g) * 30
    torch.testing.assert_close(TT.transmittance_weights(deltas, sig), O.transmittance_weights(deltas, sig), rtol=0, atol=0)
    p = torch.randn(2, 7, 3, generator=g) + torch.tensor([0.0, 0.0, 3.0])
    c2w = torch.stack([synth.relative_target_pose(1), synth.relative_target_pose(2)])
    k = synth.normalized_intrinsics(**synth.ALLEGRO_INTRINSICS_PX)[None].repeat(2, 1, 1)
    torch.testing.assert_close(TT.project(p, torch.inverse(c2w), k), O.project_px(p, c2w, k), rtol=1e-5, atol=1e-6)
    x = torch.tensor([-20.0, -1.0, 0.5, 20.0], requires_grad=True)
    TT._TruncExp.apply(x).sum().backward()   # activations.py:23-26: gradient exp(clamp(x, -15, 15))
    torch.testing.assert_close(x.grad, torch.exp(torch.clamp(x.detach(), -15, 15)))
    sd = synth.synth_state_dict(synth.field_shapes("jacobian_mlp", 6, n_proposal=2), 3)
    t = precise.trunk_from_state_dict(sd, "proposal_networks.1.density_head", "cpu")
    assert len(t.blocks) == 5 and len(t.lin_z) == 3 and t.lin_in.weight.shape == (128, 63) and t.lin_out.weight.shape == (1, 128)
    assert TT.lin_z_maps(t, torch.zeros(1, 2, 3, 512)).shape == (1, 2, 3, 384)
    L = TT._declare()
    err = lambda: L.njf_last_error().decode()
    assert L.njf_train_linear(None, None, None, None, None, None, 8, 8, 8, 1, 0, 0, None) != 0 and "null" in err()
    one = ctypes.c_void_p(16)
    assert L.njf_train_linear(one, one, None, None, None, one, 8, 130, 8, 1, 0, 0, None) != 0 and "multiples of 4" in err()
    assert L.njf_train_linear_wgrad(one, one, 8, 6, 8, 0, one, None, 0, None) != 0 and "multiples of 4" in err()
    assert L.njf_train_gather(one, one, one, 8, 384, 128, 100, one, None) != 0 and "multiple of 128" in err()
    assert L.njf_train_scatter(one, one, one, 8, 384, 300, 128, one, None) != 0
    assert L.njf_train_sh16(one, 4, 7, 1, one, None) != 0 and "convention" in err()
    assert L.njf_train_sample_setup(one, one, one, 0, 4, 2, 2, one, one, one, None) != 0
    # perception-phase forward on a CPU model: no fallback
    m = Model(_cfg()).train()
    cam = CameraInput(torch.rand(1, 3, 16, 24), torch.eye(4)[None], torch.eye(3)[None], torch.eye(4)[None], torch.eye(3)[None])
    with pytest.raises(_lib.NjfError):
        m.forward(cam, RenderingInput(torch.zeros(1, 4, 3), torch.ones(1, 4, 3), torch.tensor([0.5]), torch.tensor([2.0])),
                  RobotInput(torch.zeros(1, 8)))
