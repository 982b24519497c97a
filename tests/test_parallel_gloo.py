"""world_size-2 gloo tests of the multi-GPU host logic (ray sharding, packed gather, depth-clip
all-reduce) on CPU tensors."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers  # noqa: F401  (sys.path)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rays, A, q):
    from njf_b200 import parallel as P

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full = {k: torch.randn(n_rays, w, generator=g) for k, w in P.pack_widths(A).items()}
        a, b = P.ray_shard(n_rays, rank, world)
        mine = {k: v[a:b] for k, v in full.items()}
        frame = P.gather_rendered(P.pack_outputs(mine), n_rays)
        got = P.unpack_outputs(frame, A)
        ok = all(torch.equal(got[k], full[k]) for k in full)
        steps = torch.arange(n_rays, dtype=torch.float32)[a:b] + 1.0
        mm = torch.stack([steps.min(), steps.max()])
        P.allreduce_minmax(mm)
        ok = ok and mm.tolist() == [1.0, float(n_rays)]
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rays,A", [(1001, 8), (64, 6)])
def test_ray_shard_gather_and_clip_allreduce(n_rays, A):
    from njf_b200 import parallel as P

    cover = []
    for r in range(3):
        a, b = P.ray_shard(10, r, 3)
        cover += list(range(a, b))
    assert cover == list(range(10))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rays, A, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
