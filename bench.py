#!/usr/bin/env python
"""Benchmark of the B200-native neural-jacobian-field render path.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): Allegro Jacobian-field
render, one 400x400 target view per GPU, 128 proposal + 128 final samples per ray, cross-attention
Jacobian head (action_dim 8), 480x640 context image (feature map 512x240x320), novel target view.
A "step" = hoist the lin_z layers onto the feature map + proposal pass + field pass + finish (+ one
NCCL all_gather of the rendered buffers when N > 1).  The ResNet-34 image encoder runs once, outside
the timed region of `value` (excluded on both sides), and inside `e2e`.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                      # the reference algorithm on host cores (oracle port)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.join(ROOT, "oracle")   # test infrastructure: imported ONLY by the cpu-baseline / reference legs
for _p in (ROOT, os.path.join(ROOT, "neural-jacobian-field_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "rays/sec (sigma+rgb+Jacobian, 128 samples/ray)"
HEAD, A = "jacobian_transformer", 8
S_PROP, S_NERF = (128,), 128
IMG_H, IMG_W = 480, 640
RENDER_H, RENDER_W = 400, 400
# reference-formulation work per ray (SURVEY.md section 8d / BASELINE.md section 4)
FLOP_PROPOSAL_SAMPLE = 737_280
FLOP_FIELD_SAMPLE = 1_321_904
GATHER_BYTES_SAMPLE_F16 = 4 * 512 * 2   # 4 taps x 512 channels at the kernels' fp16 storage precision
EXEC_MAC_PROPOSAL = 128 * 64 + 10 * 128 * 128 + 16 * 128            # executed tensor-core MACs / sample
EXEC_MAC_FIELD = EXEC_MAC_PROPOSAL + 64 * 64 + 2 * 64 * 64           # + q_enc + colour head (field_kernel)
EXEC_MAC_XF = 12 * 64 * 64 + 32 * 64                                 # xf_kernel: 3 x (M1, M2, W1, W2) + jacobian_head
# reference-formulation split of the main sample (370 560 + 284 096 + 6 272 + 24 MAC):
# field_kernel = density trunk + colour + the 575->64 query MLP + J.u ; xf_kernel = the rest of the cross-attention head
FLOP_FIELD_KERNEL_SAMPLE = 2 * (370_560 + 6_272 + 575 * 64 + 24)
FLOP_XF_KERNEL_SAMPLE = FLOP_FIELD_SAMPLE - FLOP_FIELD_KERNEL_SAMPLE


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm_gbs=6650.0, tf=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def scene(view: int, device=None, pin=False, rays_device=None):
    """Synthetic view `view`: context image, cameras of the Allegro rig shape, the 400x400 target ray grid.
    `rays_device`: generate the rays with the library's own kernel on that GPU (njf_b200.geometry); None = the
    reference-side CPU restatement (oracle/synth.py), used by the reference arm which must not touch our kernels."""
    from njf_b200 import synth

    g = torch.Generator().manual_seed(2 + view)
    img = torch.rand(1, 3, IMG_H, IMG_W, generator=g)
    K = synth.normalized_intrinsics(**synth.ALLEGRO_INTRINSICS_PX)[None]
    kpx = K.clone(); kpx[:, 0] *= IMG_W; kpx[:, 1] *= IMG_H
    ctxt, trgt = torch.eye(4)[None], synth.relative_target_pose(1 + view % 5)[None]
    if rays_device is not None:
        from njf_b200 import geometry

        o, d = geometry.get_world_rays_grid(RENDER_H, RENDER_W, K.to(rays_device), trgt.to(rays_device))
        o, d = o.cpu(), d.cpu()
    else:
        if ORACLE not in sys.path:
            sys.path.insert(0, ORACLE)
        import synth as osynth

        o, d = osynth.world_rays(osynth.pixel_grid(RENDER_H, RENDER_W), K[0], trgt[0])
        o, d = o[None], d[None]
    sc = dict(image=img, ctxt_c2w=ctxt, ctxt_k=K, trgt_c2w=trgt, trgt_k_px=kpx, origins=o.contiguous(),
              dirs=d.contiguous(), z_near=torch.tensor([0.65]), z_far=torch.tensor([3.2]),
              action=0.1 * torch.randn(1, A, generator=g))
    if pin:
        sc = {k: v.pin_memory() for k, v in sc.items()}
    if device is not None:
        sc = {k: v.to(device) for k, v in sc.items()}
    return sc


def hot_weights():
    from njf_b200 import synth

    return synth.synth_state_dict(synth.field_shapes(HEAD, A, n_proposal=len(S_PROP)), 11)


def build_model(device):
    from njf_b200 import model as M, modules as mod, synth

    mlp = mod.MlpCfg()
    cfg = M.ModelCfg(action_dim=A, rendering=M.RenderingCfg(S_PROP, S_NERF), encoder=mod.EncoderResnetCfg(),
                     density_decoder=mod.DensityDecoderMlpCfg("density_mlp", mlp),
                     action_decoder=mod.ActionDecoderJacobianTransformerCfg(name=HEAD, mlp=mlp, transformer=mod.TransformerCfg()))
    m = M.Model(cfg).eval()
    sd = m.state_dict()
    enc = synth.synth_state_dict({k: tuple(v.shape) for k, v in sd.items() if k.startswith("encoder.")}, 11)
    m.load_state_dict({**enc, **hot_weights()})
    return m.to(device)


def oracle_rays_per_s(nrays: int, steps: int, warmup: int, want_outputs=False, device=None, threads=None, rays_device=None):
    """The reference algorithm (oracle port, fp32) on a bounded ray sample: on the host cores (all threads)
    or, with `device`, as eager PyTorch on the GPU (what the reference's own code path does on one GPU)."""
    if ORACLE not in sys.path:
        sys.path.insert(0, ORACLE)
    import njf_oracle as O

    if threads:
        torch.set_num_threads(threads)
    w = hot_weights()
    sc = scene(0, rays_device=rays_device)
    g = torch.Generator().manual_seed(9)
    feat = torch.randn(1, 512, IMG_H // 2, IMG_W // 2, generator=g).abs() * 0.7
    idx = torch.randperm(RENDER_H * RENDER_W, generator=g)[:nrays]
    spec = O.FieldSpec(HEAD, A)
    if device is not None:
        w = {k: v.to(device) for k, v in w.items()}
        sc = {k: v.to(device) for k, v in sc.items()}
        feat, idx = feat.to(device), idx.to(device)
        sync = torch.cuda.synchronize
    else:
        sync = lambda: None
    run = lambda: O.render_forward(w, spec, feat, sc["ctxt_c2w"], sc["ctxt_k"], sc["trgt_c2w"], sc["trgt_k_px"],
                                   sc["origins"][:, idx], sc["dirs"][:, idx], sc["z_near"], sc["z_far"], sc["action"],
                                   S_PROP, S_NERF)
    with torch.no_grad():
        for _ in range(warmup):
            run()
        sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = run()
        sync()
        dt = (time.perf_counter() - t0) / steps
    return nrays / dt, dt, (feat, idx, out) if want_outputs else None


def best_host_threads():
    """The CPU port is a chain of small eager torch ops; on a many-core host the full thread count is often slower
    than a few dozen threads.  Give the baseline its best configuration: time 32 rays at a few thread counts."""
    n = os.cpu_count() or 1
    best, best_rps = n, 0.0
    for t in sorted({n, min(n, 64), min(n, 32), min(n, 16), min(n, 8)}, reverse=True):
        rps, _, _ = oracle_rays_per_s(32, 1, 1, threads=t)
        if rps > best_rps:
            best, best_rps = t, rps
    torch.set_num_threads(best)
    return best, best_rps


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if "NJF_BENCH_REF_RAYS" in os.environ:
        nrays = int(os.environ["NJF_BENCH_REF_RAYS"])
        best_host_threads()
    else:
        # bounded sample: calibrate on 32 rays (untimed), then size a step to ~8 s of host time so that the
        # whole --steps K run ends within a few minutes whatever the box's core count
        _, r0 = best_host_threads()
        budget = min(8.0, 150.0 / max(args.steps + 1, 1))
        nrays = int(max(32, min(512, 32 * round(r0 * budget / 32))))
    rps, dt, _ = oracle_rays_per_s(nrays, args.steps, min(args.warmup, 1))
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": rps, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "allegro_jacobian_400x400_s128", "head": HEAD, "action_dim": A, "samples": "128+128",
                   "note": "reference algorithm (oracle/njf_oracle.py port; /root/reference is absent on the GPU box), "
                           "encoder excluded"},
        "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{nrays} random rays of the 400x400 frame per step, 128+128 samples"},
        "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)
    args.warmup = max(args.warmup, 3)

    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ge.build()
    from njf_b200 import _lib, api
    from njf_b200.model import CameraInput, RenderingInput, RobotInput
    import ctypes

    L = api._declare()
    model = build_model(dev)
    sc = scene(rank, dev, rays_device=dev)
    with torch.no_grad():
        feat = model.encoder(sc["image"]).float().contiguous()   # once, outside the timed region
    fld = model.field()
    Hf, Wf = feat.shape[-2:]
    R = RENDER_H * RENDER_W
    cams, keep = api.make_cameras(sc["ctxt_c2w"], sc["ctxt_k"], sc["trgt_c2w"], sc["trgt_k_px"], dev)
    bins0, us = api.eval_tables(S_PROP, S_NERF, dev)
    f32 = dict(device=dev, dtype=torch.float32)
    nbytes = L.njf_hoisted_bytes(fld.handle, 1, Hf, Wf)
    maps = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    packed = torch.empty(R, 3 + 1 + 2 + 3 * A + 3 + 3, **f32)   # rendered buffers of this rank (one gather)
    outs = dict(rgb=torch.empty(1, R, 3, **f32), depth=torch.empty(1, R, 1, **f32), flow=torch.empty(1, R, 2, **f32),
                jbar=torch.empty(1, R, 3 * A, **f32), p=torch.empty(1, R, 3, **f32), pw=torch.empty(1, R, 3, **f32))
    lb = torch.empty(1, R, S_NERF + 1, **f32)
    minmax = torch.empty(2, **f32)
    a = api.NjfRenderArgs()
    a.B, a.R, a.n_levels, a.s_nerf = 1, R, 1, S_NERF
    a.s_prop[0] = S_PROP[0]
    a.origins, a.dirs = api.dptr(sc["origins"]), api.dptr(sc["dirs"])
    a.z_near, a.z_far, a.action = api.dptr(sc["z_near"]), api.dptr(sc["z_far"]), api.dptr(sc["action"])
    h_nf = (sc["z_near"].cpu().contiguous(), sc["z_far"].cpu().contiguous())
    a.h_z_near, a.h_z_far = h_nf[0].data_ptr(), h_nf[1].data_ptr()
    a.bins0, a.bins0_stride = api.dptr(bins0), 0
    a.u[0], a.u_stride[0] = api.dptr(us[0]), 0
    a.anneal, a.sum_vec_width = 1.0, api.default_sum_vec_width()
    a.maps, a.Hf, a.Wf = api.dptr(maps), Hf, Wf
    for k, t in outs.items():
        setattr(a, k, api.dptr(t))
    a.level_bins[0] = api.dptr(lb)
    a.minmax = api.dptr(minmax)
    ws = torch.empty(fld.workspace_bytes(1, R, S_PROP, S_NERF), dtype=torch.uint8, device=dev)   # caller-owned scratch
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    a.packed = api.dptr(packed)
    gathered = torch.empty(world * packed.numel(), **f32) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    st = api.stream_ptr
    h = fld.handle

    def step(timers=None):
        e = [ev() for _ in range(5)] if timers is not None else None
        if e: e[0].record()
        _lib.check(L.njf_hoist_features(h, api.dptr(feat), 1, Hf, Wf, api.dptr(maps), st()))
        if e: e[1].record()
        _lib.check(L.njf_proposal_pass(h, ctypes.byref(cams), ctypes.byref(a), 0, api.dptr(bins0), 0, st()))
        if e: e[2].record()
        _lib.check(L.njf_field_pass(h, ctypes.byref(cams), ctypes.byref(a), api.dptr(lb), S_NERF + 1, st()))
        if e: e[3].record()
        _lib.check(L.njf_finish_pass(h, ctypes.byref(cams), ctypes.byref(a), st()))
        if world > 1:   # finish_kernel wrote the packed per-ray struct: one collective, no torch.cat
            dist.all_gather_into_tensor(gathered, packed.view(-1))
        if e:
            e[4].record()
            timers.append(e)

    for _ in range(args.warmup):
        step()
        flush.zero_()
    torch.cuda.synchronize()
    if dist: dist.barrier()
    clocks = ClockSampler(local)
    if rank == 0: clocks.start()
    timers = []
    torch.cuda.synchronize()
    _lib.check(L.njf_debug_field_timing(1, None, None))   # per-kernel CUDA events inside njf_field_pass
    for _ in range(args.steps):
        step(timers)
        flush.zero_()          # L2 flush between timed steps (outside the per-step event pairs)
    torch.cuda.synchronize()
    if dist: dist.barrier()
    clk = clocks.stop() if rank == 0 else None
    tot = sum(e[0].elapsed_time(e[4]) for e in timers)            # ms over K steps
    t_hoist = sum(e[0].elapsed_time(e[1]) for e in timers) / args.steps
    t_prop = sum(e[1].elapsed_time(e[2]) for e in timers) / args.steps
    t_field = sum(e[2].elapsed_time(e[3]) for e in timers) / args.steps
    c_fk, c_xf = ctypes.c_float(0), ctypes.c_float(0)
    _lib.check(L.njf_debug_field_timing(0, ctypes.byref(c_fk), ctypes.byref(c_xf)))
    t_fk, t_xf = c_fk.value / args.steps, c_xf.value / args.steps
    tt = torch.tensor([tot], device=dev)
    if dist: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    tot = float(tt.item())
    ms_step = tot / args.steps
    value = world * R * args.steps / (tot * 1e-3)

    # ---- e2e: the public API call (Model.forward) with HOST pinned inputs, encoder + copies inside
    hs = scene(rank, pin=True, rays_device=dev)
    cam = CameraInput(hs["image"], hs["ctxt_c2w"], hs["ctxt_k"], hs["trgt_c2w"], hs["trgt_k_px"])
    rin = RenderingInput(hs["origins"], hs["dirs"], hs["z_near"], hs["z_far"])
    rob = RobotInput(hs["action"])
    with torch.no_grad():
        for _ in range(2):
            out = model.forward(cam, rin, rob)
        torch.cuda.synchronize()
        if dist: dist.barrier()
        e2e_steps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out = model.forward(cam, rin, rob)   # returns host tensors (device->host read of the result)
            _ = float(out.standard_output.rgb[0, 0, 0])
        torch.cuda.synchronize()
        te = (time.perf_counter() - t0)
    te_t = torch.tensor([te], device=dev)
    if dist: dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
    e2e_value = world * R * e2e_steps / float(te_t.item())
    h2d = sum(hs[k].numel() * 4 for k in hs)
    d2h = R * (3 + 1 + 2) * 4

    if rank != 0:
        if dist: dist.destroy_process_group()
        return
    pk = peaks()
    # ---- roofline of the dominant kernel (reference-formulation algorithmic work)
    cands = [("field_kernel", t_fk, FLOP_FIELD_KERNEL_SAMPLE, EXEC_MAC_FIELD, True),
             ("proposal_kernel", t_prop, FLOP_PROPOSAL_SAMPLE, EXEC_MAC_PROPOSAL, True),
             ("xf_kernel", t_xf, FLOP_XF_KERNEL_SAMPLE, EXEC_MAC_XF, False)]
    dom, t_dom, flop_s, exec_mac, gathers = max(cands, key=lambda c: c[1])
    evals = R * S_NERF
    alg_flop = evals * flop_s
    # algorithmic bytes: the 4-tap x 512-channel gather (trunk kernels) / the 64-float query stream + weight in,
    # J-bar out (xf_kernel); per-ray inputs and outputs
    alg_bytes = evals * (GATHER_BYTES_SAMPLE_F16 if gathers else 65 * 4) + R * (24 + 144)
    exec_flop = 2 * evals * exec_mac
    t_flop_bound = alg_flop / (pk["tf_sustained"] * 1e12)
    t_byte_bound = alg_bytes / (pk["hbm_gbs"] * 1e9)
    if t_flop_bound >= t_byte_bound:
        roof = {"bound": "tensor", "achieved": alg_flop / (t_dom * 1e-3) / 1e12, "peak": pk["tf_sustained"], "unit": "TFLOP/s"}
    else:
        roof = {"bound": "hbm", "achieved": alg_bytes / (t_dom * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof.update(kernel=dom, ms_per_launch=t_dom, peak_source=pk["source"] + " (sustained bf16 cuBLAS / copy bandwidth)",
                traffic=None,
                executed_tensor_tflops=exec_flop / (t_dom * 1e-3) / 1e12,
                executed_tensor_frac=exec_flop / (t_dom * 1e-3) / 1e12 / pk["tf_sustained"],
                algorithmic_gather_gbs=alg_bytes / (t_dom * 1e-3) / 1e9,
                algorithmic_gather_frac_of_hbm=alg_bytes / (t_dom * 1e-3) / 1e9 / pk["hbm_gbs"],
                note="algorithmic work = reference formulation (un-hoisted lin_z, un-folded attention); the kernel "
                     "executes fewer FLOPs (hoist + fold) and gathers from L2-resident fp16 maps, so fractions may exceed 1")
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp)).get(dom)   # ncu --set full capture of this build (profiles/ncu_*_summary.json)
        if t:
            roof["traffic"] = t["bytes_per_launch"]          # dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch
            roof["traffic_detail"] = t
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": "allegro_jacobian_400x400_s128", "head": HEAD, "action_dim": A, "samples": "128+128",
                   "rays_per_gpu": R, "context_image": f"{IMG_H}x{IMG_W}", "feature_map": f"512x{Hf}x{Wf}",
                   "view": "novel target view", "parallelism": f"ray-shard x{world} (one view per GPU, weak)",
                   "l2": "flushed between timed steps (256 MiB memset outside the event pairs)",
                   "encoder": "excluded from value (once per image, cuDNN), included in e2e",
                   "weights": "synthetic seeded (njf_b200/synth.py), random-init architecture of model_allegro.yaml"},
        "breakdown_ms": {"hoist": t_hoist, "proposal_kernel": t_prop, "field_pass": t_field, "field_kernel": t_fk, "xf_kernel": t_xf,
                         "finish+gather": ms_step - t_hoist - t_prop - t_field if world == 1 else None},
        "roofline": roof,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": float(te_t.item()) / e2e_steps * 1e3},
        "gpu_launches": 8 * args.steps,
        "clocks": clk,
    }
    if not args.no_cpu_baseline:
        # bounded CPU sample of the same workload + quality vs the oracle on those rays
        nrays = 256
        best_host_threads()
        rps, dt, (ofeat, idx, oref) = oracle_rays_per_s(nrays, 1, 0, want_outputs=True, rays_device=dev)
        from njf_b200.render import render
        m2 = fld.hoist(ofeat.to(dev))
        res = render(fld, m2, Hf, Wf, cams, sc["origins"][:, idx.to(dev)].contiguous(), sc["dirs"][:, idx.to(dev)].contiguous(),
                     sc["z_near"], sc["z_far"], sc["action"], S_PROP, S_NERF)
        torch.cuda.synchronize()
        mse = float(((res.rgb.cpu() - oref["rgb"]) ** 2).mean())
        jr = float((res.jbar.cpu() - oref["action_features"]).norm() / oref["action_features"].norm())
        line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{nrays} random rays of the frame, 128+128 samples, oracle/njf_oracle.py (torch CPU fp32, best of 5 thread counts), {dt:.1f} s"}
        try:  # the same port as eager PyTorch on this GPU (reference-style single-GPU path), for scale only
            grays = 2048
            grps, gdt, _ = oracle_rays_per_s(grays, 3, 1, device=dev)
            line["cpu_baseline"]["torch_gpu_port"] = {"value": grps, "unit": "rays/s",
                                                      "sample": f"{grays} rays/step (the reference's patch size), eager torch fp32 on cuda:0, {gdt * 1e3:.0f} ms/step"}
        except Exception as ex:  # noqa: BLE001 -- a baseline leg must not break the bench line
            line["cpu_baseline"]["torch_gpu_port"] = {"error": repr(ex)[:200]}
        line["quality"] = {"psnr_rgb_vs_oracle_db": 10 * __import__("math").log10(1.0 / max(mse, 1e-12)),
                           "jacobian_rel_l2_vs_oracle": jr, "rays": nrays}
    print(json.dumps(line))
    if dist: dist.destroy_process_group()


if __name__ == "__main__":
    main()
