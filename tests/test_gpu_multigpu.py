"""Ray-sharded multi-view call on two GPUs over NCCL (BASELINE config 4): every rank renders a contiguous
range of the flattened (view, ray) space, the (min, max) of the sample steps is all-reduced between njf_field_pass and
njf_finish_pass, and the packed per-ray struct is gathered to rank 0 with one collective.  The gathered frame must
equal the single-GPU render of the whole call BIT FOR BIT, depth clip included.  Needs >= 2 GPUs (skipped otherwise;
run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch

import helpers  # noqa: F401  (sys.path)

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from helpers import synth
    from njf_b200 import api, parallel as P
    from njf_b200.render import render

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        head, A, s_prop, s_nerf, V, R = "jacobian_transformer", 8, (32,), 48, 3, 50
        Hf, Wf = 12, 16
        g = torch.Generator().manual_seed(0)
        w = synth.synth_state_dict(synth.field_shapes(head, A), 11)
        feat = (torch.randn(V, 512, Hf, Wf, generator=g).abs() * 0.7).to(dev)
        K = synth.normalized_intrinsics(**synth.ALLEGRO_INTRINSICS_PX)[None].repeat(V, 1, 1)
        kpx = K.clone(); kpx[:, 0] *= 640; kpx[:, 1] *= 480
        ctxt = torch.eye(4)[None].repeat(V, 1, 1)
        trgt = torch.stack([synth.relative_target_pose(1 + v) for v in range(V)])
        coords = torch.rand(R, 2, generator=g)
        rays = [synth.world_rays(coords, K[v], trgt[v]) for v in range(V)]
        o = torch.stack([r[0] for r in rays]).to(dev); d = torch.stack([r[1] for r in rays]).to(dev)
        zn = torch.tensor([0.5, 0.7, 0.9]).to(dev); zf = torch.tensor([3.0, 2.6, 3.3]).to(dev)   # the extremes live on different ranks
        act = (0.1 * torch.randn(V, A, generator=g)).to(dev)
        fld = api.Field(head, A, 1, w)
        cams, keep = api.make_cameras(ctxt, K, trgt, kpx, dev)
        start, stop = P.ray_shard(V * R, rank, world)
        v0, v1 = P.shard_views(start, stop, R)
        L = api._declare()
        maps = torch.full((L.njf_hoisted_bytes(fld.handle, V, Hf, Wf),), 0x7f, dtype=torch.uint8, device=dev)  # other views: junk
        fld.hoist_views(feat[v0:v1].contiguous(), v0, V, maps)
        res, frame = P.render_sharded(fld, maps, Hf, Wf, cams, o.reshape(-1, 3)[start:stop].contiguous(),
                                      d.reshape(-1, 3)[start:stop].contiguous(), zn, zf, act, s_prop, s_nerf, V, R, rank, world,
                                      gather=False, vis=True)
        frame = P.gather_rendered(res.packed[0], V * R, dst=0)
        torch.cuda.synchronize()
        mm = res.minmax.cpu().numpy()
        if rank == 0:
            whole = render(fld, fld.hoist(feat), Hf, Wf, cams, o, d, zn, zf, act, s_prop, s_nerf, vis=True, packed=True)
            torch.cuda.synchronize()
            np.savez(os.path.join(out_dir, "r0.npz"), frame=frame.cpu().numpy(), whole=whole.packed.reshape(V * R, -1).cpu().numpy(),
                     mm=mm, mm_whole=whole.minmax.cpu().numpy())
        else:
            assert frame is None
            np.savez(os.path.join(out_dir, "r1.npz"), mm=mm)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_sharded_render_equals_single_gpu_render(tmp_path):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    assert np.array_equal(r0["frame"], r0["whole"])           # bit-exact, depth clip included
    assert np.array_equal(r0["mm"], r0["mm_whole"]) and np.array_equal(r1["mm"], r0["mm_whole"])   # NCCL min/max all-reduce
