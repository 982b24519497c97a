"""Per-phase SM-cycle attribution of the epilogue warps (needs the -DNJF_PROFILE build:
NJF_LIB=neural-jacobian-field_b200/lib/libnjf_b200_prof.so python tools/phase_profile.py)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench  # noqa: E402  (sets sys.path for the package)
from njf_b200 import _lib, api

PH = ["setup", "gather", "wait_acc", "epilogue", "weights", "pdf", "head", "color", "composite", "barrier", "other",
      "xf_layernorm", "xf_softmax", "xf_gelu", "xf_residual"]


def main():
    L = api._declare()
    L.njf_prof_read.restype = ctypes.c_int
    L.njf_prof_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cfg = bench.CONFIGS["cfg3"]
    model = bench.build_model(cfg, dev)
    sc = bench.scene(cfg, 0, dev, rays_device=dev)
    with torch.no_grad():
        feat = model.encoder(sc["image"]).float().contiguous()
    fld = model.field()
    Hf, Wf = feat.shape[-2:]
    R = cfg["H"] * cfg["W"]
    cams, keep = api.make_cameras(sc["ctxt_c2w"], sc["ctxt_k"], sc["trgt_c2w"], sc["trgt_k_px"], dev)
    maps = fld.hoist(feat)
    from njf_b200.render import render

    buf = (ctypes.c_ulonglong * 16)()

    def read():
        _lib.check(L.njf_prof_read(buf, 1))
        return list(buf)[:len(PH)]

    # full render once for warm-up, then reset
    args = (fld, maps, Hf, Wf, cams, sc["origins"], sc["dirs"], sc["z_near"], sc["z_far"], sc["action"], cfg["s_prop"], cfg["s_nerf"])
    res = render(*args)
    read()
    out = {}
    # proposal + field in one call, then a field-only call at the same bins -> separate the two kernels
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    res = render(*args)
    e1.record()
    both = read()
    res2 = render(*args, final_bins=res.level_bins[-1])
    e2.record()
    field = read()
    torch.cuda.synchronize()
    prop = [b - f for b, f in zip(both, field)]
    n_warps = 16 * 148
    for name, cyc, ms in (("proposal_kernel", prop, None), ("field_kernel", field, e1.elapsed_time(e2))):
        tot = sum(cyc)
        out[name] = {"avg_kcycles_per_warp": {p: round(c / n_warps / 1e3, 1) for p, c in zip(PH, cyc)},
                     "share": {p: round(c / max(tot, 1), 3) for p, c in zip(PH, cyc)}, "ms": ms}
    if hasattr(L, "njf_xf_prof_read"):
        xb = (ctypes.c_ulonglong * 8)()
        L.njf_xf_prof_read.restype = ctypes.c_int
        L.njf_xf_prof_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _lib.check(L.njf_xf_prof_read(xb, 1))
        xph = ["load+ln0", "wait_acc", "layernorm", "softmax", "gelu", "jacobian+composite"]
        cyc = list(xb)[:len(xph)]
        tot = sum(cyc)
        out["xf_kernel"] = {"avg_kcycles_per_warp": {p: round(c / n_warps / 1e3 / 3, 1) for p, c in zip(xph, cyc)},
                            "share": {p: round(c / max(tot, 1), 3) for p, c in zip(xph, cyc)},
                            "note": "three renders accumulated; 16 row warps x 148 SMs"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
