import sys, time, os, torch
sys.path.insert(0, "/root/repo")
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
import __graft_entry__ as ge; ge.build()
cfg = dict(bench.CONFIGS["cfg3"], s_prop=(256,), s_nerf=256)
m = bench.build_model(cfg, dev)
m.field(); torch.cuda.synchronize()
ts = []
for i in range(5):
    with torch.no_grad():
        m.decoder.jacobian_head.weight.add_(1e-6)   # bumps the version -> re-pack
    torch.cuda.synchronize(); t0 = time.perf_counter(); m.field(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("re-pack ms:", [round(t * 1e3, 2) for t in ts])
img = torch.rand(7, 3, 480, 640, device=dev)
m.train()
with torch.no_grad():
    for _ in range(2): m.encoder(img)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): f = m.encoder(img)
    torch.cuda.synchronize(); print("encoder 7 images ms:", (time.perf_counter() - t0) / 5 * 1e3)
    fld = m.field(); f = f.float().contiguous()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fld.hoist(f)
    torch.cuda.synchronize(); print("hoist 7 maps ms:", (time.perf_counter() - t0) / 5 * 1e3)
