"""Pin the CPU oracle (oracle/njf_oracle.py) to the golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  CPU-only."""
import numpy as np
import pytest
import torch

from helpers import O, RENDER_FIXTURES, load_fixture, oracle_render, GOLDEN
import os


@pytest.mark.parametrize("name", RENDER_FIXTURES)
def test_render_matches_reference(name):
    fxt = load_fixture(name)
    fx = fxt[0]
    out = oracle_render(fxt)
    # integer sample indexing: bit-exact
    lvl = 1
    while f"inds_{lvl}" in fx:
        assert np.array_equal(out[f"inds_{lvl}"].numpy(), fx[f"inds_{lvl}"]), f"inds level {lvl}"
        lvl += 1
    tol = dict(rgb=3e-5, depth=1e-4, action_features=2e-4, steps=1e-5, weights=5e-5, ray_positions=1e-4,
               ray_positions_warped=1e-4, final_bins=5e-6, proposal_weights=5e-5, sigma=5e-4,
               rgb_samples=5e-4, positions=1e-5)  # per-sample values sit on the 2*pi*512*x fp32 noise floor
    for k, atol in tol.items():
        ref = fx[k]
        got = out[k].numpy()
        scale = max(1.0, float(np.abs(ref).max()))
        assert got.shape == ref.shape, k
        np.testing.assert_allclose(got, ref, rtol=0, atol=atol * scale, err_msg=k)
    ref, got = fx["jacobian"], out["jacobian"].numpy()
    np.testing.assert_allclose(got, ref, rtol=0, atol=5e-4 * max(float(np.abs(ref).max()), 1e-6))
    ref, got = fx["optical_flow"], out["optical_flow"].numpy()
    np.testing.assert_allclose(got, ref, rtol=0, atol=3e-5 * max(1.0, float(np.abs(ref).max())) + 2e-2)


@pytest.mark.parametrize("name", ["render_transformer", "render_mlp"])
def test_encode_image_and_inverse_flow(name):
    fx, t, head, A, s_prop, s_nerf, w = load_fixture(name)
    out = oracle_render((fx, t, head, A, s_prop, s_nerf, w))
    np.testing.assert_allclose(out["sigma"].numpy(), fx["enc_density"], atol=2e-4 * float(np.abs(fx["enc_density"]).max()))
    np.testing.assert_allclose(out["positions"].numpy(), fx["enc_positions"], atol=1e-5)
    flow = O.infer_optical_flow(t("enc_jacobian"), t("enc_weights"), t("enc_positions"), t("action2"),
                                t("trgt_c2w"), t("trgt_k_px"))
    np.testing.assert_allclose(flow.numpy(), fx["flow2"], atol=2e-2)


@pytest.mark.parametrize("tag", ["16_24", "64_64", "128_128", "256_256", "48_32"])
def test_pdf_sampler_bit_exact(tag):
    z = np.load(os.path.join(GOLDEN, "pdf_sampler.npz"))
    s_out = int(tag.split("_")[1])
    bins, inds = O.pdf_resample(torch.from_numpy(z[f"w_{tag}"]), torch.from_numpy(z[f"bins_in_{tag}"]), s_out)
    assert np.array_equal(inds.numpy(), z[f"inds_{tag}"])
    assert np.array_equal(bins.numpy(), z[f"bins_out_{tag}"])
    e = O.spacing_to_euclid(bins, torch.tensor(0.5), torch.tensor(3.0))
    assert np.array_equal(e[..., :-1].numpy(), z[f"starts_{tag}"])
    tw = O.transmittance_weights(torch.from_numpy(z[f"deltas_{tag}"])[..., None],
                                 torch.from_numpy(z[f"w_{tag}"])[..., None] * 20.0)[..., 0]
    assert np.array_equal(tw.numpy(), z[f"tw_{tag}"])


def test_encodings_known_values():
    x = torch.tensor([[0.0, 0.25, -0.5]])
    e = O.posenc(x)
    assert e.shape == (1, 63)
    # dim-major, freq-minor: column 10 is sin(2*pi*0.25*2^0) = 1 ; column 30+10 its cosine = 0
    assert abs(float(e[0, 10]) - 1.0) < 1e-6 and abs(float(e[0, 40])) < 1e-6
    assert torch.equal(e[0, 60:], x[0])
    s = O.sh4(torch.tensor([[0.5, 0.5, 1.0]]), fp16_round=False)  # direction (0,0,1)
    assert abs(float(s[0, 0]) - 0.2820948) < 1e-6 and abs(float(s[0, 2]) - 0.4886025) < 1e-6
    assert abs(float(s[0, 6]) - (0.9461747 - 0.3153916)) < 1e-6 and abs(float(s[0, 12]) - 0.3731763 * 2.0) < 1e-6


def test_ray_generation_matches_reference():
    """oracle pixel_coordinates / world_rays_with_z against the reference's geometry.get_pixel_coordinates and
    get_world_rays_with_z (tests/golden/rays.npz, made by oracle/make_golden.py rays)."""
    z = np.load(os.path.join(GOLDEN, "rays.npz"))
    K, c2w = torch.from_numpy(z["k_norm"]), torch.from_numpy(z["c2w"])
    xy, sel = O.pixel_coordinates(9, 13)
    assert np.array_equal(xy.numpy(), z["xy_small"]) and np.array_equal(sel.numpy(), z["sel_small"])
    o, d, zz = O.world_rays_with_z(xy.reshape(1, -1, 2).repeat(2, 1, 1), K, c2w)
    np.testing.assert_allclose(o.numpy(), z["origins_small"], atol=0, rtol=0)
    np.testing.assert_allclose(d.numpy(), z["dirs_small"], atol=2e-7, rtol=0)
    np.testing.assert_allclose(zz.numpy(), z["z_small"], atol=2e-7, rtol=0)
    xyf, self_ = O.pixel_coordinates(400, 400)
    assert np.array_equal(xyf.reshape(-1, 2)[:4096].numpy(), z["xy_full"])
    assert np.array_equal(self_.reshape(-1, 2)[:4096].numpy(), z["sel_full"])
    o, d, zz = O.world_rays_with_z(xyf.reshape(1, -1, 2)[:, :4096].repeat(2, 1, 1), K, c2w)
    np.testing.assert_allclose(d.numpy(), z["dirs_full"], atol=2e-7, rtol=0)


@pytest.mark.parametrize("name", ["train_transformer", "train_mlp_2prop", "train_single_jitter"])
def test_train_mode_forward_and_gradients_match_reference(name):
    """Train-mode parity of the oracle AND of njf_b200.train.stratified_tables with the unmodified reference
    (oracle/make_golden.py train_fixture): the reference's Model.train() forward draws its stratified jitter from
    torch's global generator; the same seed through stratified_tables must give the same bins, and autograd through
    the oracle the same loss and the same gradient for every decoder / proposal-network parameter and for the
    encoder output.  (The GPU training tests compare the kernels' gradients with autograd through this oracle.)"""
    from helpers import synth
    from njf_b200.train import stratified_tables

    def train_loss(rgb, flow, weights_list, mids_list, target_rgb, target_depth, target_flow):   # oracle/make_golden.py
        loss = torch.nn.functional.mse_loss(rgb, target_rgb) + 0.01 * torch.nn.functional.mse_loss(flow, target_flow)
        for w_, mid in zip(weights_list, mids_list):
            loss = loss + 0.08 * ((w_ * mid).sum(-2) - target_depth).pow(2).mean() / len(weights_list) + 0.01 * (w_ * w_).sum(-2).mean()
        return loss

    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    t = lambda k: torch.from_numpy(np.ascontiguousarray(z[k]))
    head, A = str(z["head"]), int(z["action_dim"])
    s_prop, s_nerf = tuple(int(v) for v in z["s_prop"]), int(z["s_nerf"])
    w = synth.synth_state_dict(synth.field_shapes(head, A, n_proposal=len(s_prop)), int(z["wseed"]), "trained")
    w = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    feat = t("feat").clone().requires_grad_(True)
    o, d = t("origins"), t("dirs")
    B, R = o.shape[:2]
    torch.manual_seed(int(z["seed"]))
    bins0, us = stratified_tables(s_prop, s_nerf, B, R, bool(int(z["single_jitter"])) if "single_jitter" in z.files else False, "cpu")
    out = O.render_forward(w, O.FieldSpec(head, A), feat, t("ctxt_c2w"), t("ctxt_k"), t("trgt_c2w"), t("trgt_k_px"), o, d,
                           t("z_near"), t("z_far"), t("action"), s_prop, s_nerf, bins0=bins0, us=us)
    near, far = t("z_near")[:, None, None], t("z_far")[:, None, None]
    ws, mids = [], []
    for lvl in range(len(s_prop) + 1):
        b = out[f"prop_bins_{lvl}"] if lvl < len(s_prop) else out["final_bins"]
        np.testing.assert_allclose(b.detach().numpy(), z[f"bins_{lvl}"], rtol=0, atol=5e-6, err_msg=f"bins level {lvl}")
        wl = out[f"prop_weights_{lvl}"] if lvl < len(s_prop) else out["weights"]
        np.testing.assert_allclose(wl.detach().numpy(), z[f"weights_{lvl}"], rtol=0, atol=5e-5, err_msg=f"weights level {lvl}")
        e = b * far + (1 - b) * near
        mids.append(((e[..., :-1] + e[..., 1:]) / 2)[..., None])
        ws.append(wl[..., None])
    np.testing.assert_allclose(out["rgb"].detach().numpy(), z["rgb"], rtol=0, atol=3e-5)
    loss = train_loss(out["rgb"], out["optical_flow"], ws, mids, t("target_rgb"), t("target_depth"), t("target_flow"))
    np.testing.assert_allclose(float(loss), float(z["loss"]), rtol=2e-5)
    loss.backward()
    stride = int(z["stride"])
    worst = 0.0
    for n in z["grad_names"]:
        n = str(n)
        g = feat.grad if n == "feat" else w[n].grad
        assert g is not None, n
        ref_norm = float(z["gnorm/" + n])
        if ref_norm < 1e-12:
            assert float(g.norm()) < 1e-10, n
            continue
        assert abs(float(g.norm()) - ref_norm) <= 1e-3 * ref_norm, (n, float(g.norm()), ref_norm)
        sub, ref = g.reshape(-1)[::(1 if g.numel() <= 8192 else stride)].numpy(), z["gsub/" + n]
        # error of the stored elements relative to the tensor's RMS gradient (a sub-vector of small entries is not
        # judged against its own tiny norm)
        rel = float(np.linalg.norm(sub - ref)) / (ref_norm * (len(ref) / g.numel()) ** 0.5)
        worst = max(worst, rel)
        # fp32 both sides; what is left is summation order plus the odd ReLU whose pre-activation sits within round-off of
        # zero (one sample of a ~1 000-sample batch then enters or leaves a bias gradient): measured 2.4e-4 / 2.6e-3
        assert rel < 6e-3, (n, rel)
    print(f"{name}: {len(z['grad_names'])} gradient tensors vs the reference's autograd, worst relative L2 error {worst:.2e}")
