"""Child process of test_chunked_field_pass_is_bit_identical: renders a seeded scene and stores the outputs
(argv: out.npz R s_prop s_nerf).  The parent runs it twice, with and without NJF_XF_MAX_TILES."""
import sys

import numpy as np
import torch

from helpers import synth  # noqa: F401  (sets sys.path for the package / oracle)


def main(out, R, s_prop, s_nerf):
    from njf_b200 import api
    from njf_b200.render import render

    dev = "cuda:0"
    head, A = "jacobian_transformer", 8
    g = torch.Generator().manual_seed(5)
    w = synth.synth_state_dict(synth.field_shapes(head, A), 21)
    feat = torch.randn(2, 512, 10, 14, generator=g).abs() * 0.7
    K = synth.normalized_intrinsics(**synth.ALLEGRO_INTRINSICS_PX)[None].repeat(2, 1, 1)
    kpx = K.clone(); kpx[:, 0] *= 640; kpx[:, 1] *= 480
    ctxt = torch.eye(4)[None].repeat(2, 1, 1)
    trgt = torch.stack([synth.relative_target_pose(1), synth.relative_target_pose(2)])
    rays = [synth.world_rays(torch.rand(R, 2, generator=g), K[b], trgt[b]) for b in range(2)]
    o = torch.stack([r[0] for r in rays]).to(dev)
    d = torch.stack([r[1] for r in rays]).to(dev)
    zn, zf = torch.tensor([0.4, 0.5], device=dev), torch.tensor([2.5, 3.0], device=dev)
    act = (0.1 * torch.randn(2, A, generator=g)).to(dev)
    fld = api.Field(head, A, 1, w)
    maps = fld.hoist(feat.to(dev))
    cams, keep = api.make_cameras(ctxt, K, trgt, kpx, dev)
    res = render(fld, maps, 10, 14, cams, o, d, zn, zf, act, (s_prop,), s_nerf, per_sample=True)
    torch.cuda.synchronize()
    np.savez(out, **{k: getattr(res, k).cpu().numpy() for k in ("rgb", "depth", "flow", "jbar", "p", "pw", "jac", "sigma", "weights")})


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
