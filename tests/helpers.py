"""Shared test helpers: load golden fixtures, rebuild their synthetic weights, run the oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "neural-jacobian-field_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import njf_oracle as O  # noqa: E402
import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
RENDER_FIXTURES = ["render_transformer", "render_mlp", "render_transformer_2prop_b2", "render_transformer_initlike"]


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    fx = {k: z[k] for k in z.files}
    t = lambda k: torch.from_numpy(np.ascontiguousarray(fx[k]))
    head, A = str(fx["head"]), int(fx["action_dim"])
    s_prop = tuple(int(v) for v in fx["s_prop"])
    weights = synth.synth_state_dict(synth.field_shapes(head, A, n_proposal=len(s_prop)), int(fx["wseed"]),
                                     str(fx["regime"]))
    return fx, t, head, A, s_prop, int(fx["s_nerf"]), weights


def oracle_render(fx_tuple):
    fx, t, head, A, s_prop, s_nerf, w = fx_tuple
    spec = O.FieldSpec(head=head, action_dim=A)
    with torch.no_grad():
        return O.render_forward(w, spec, t("feat"), t("ctxt_c2w"), t("ctxt_k"), t("trgt_c2w"), t("trgt_k_px"),
                                t("origins"), t("dirs"), t("z_near"), t("z_far"), t("action"), s_prop, s_nerf)
