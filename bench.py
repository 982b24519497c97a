#!/usr/bin/env python
"""Benchmark of the B200-native neural-jacobian-field render path.

Workloads (BASELINE.json `configs`; --config picks one, the default and the N=1 headline is cfg3):
  cfg3  Allegro Jacobian-field render: one 400x400 target view per GPU, 128 proposal + 128 final samples per ray,
        cross-attention Jacobian head (action_dim 8), 480x640 context image (feature map 512x240x320).  Weak scaling.
  cfg2  Allegro perception render: 200x200, 64 + 64 samples, same model / context image.
  cfg4  12-camera multi-view batch: 12 views x 160 000 rays flattened and split contiguously over the N ranks (strong
        scaling), (min, max) all-reduce between the field and the finish pass, one gather of the packed buffers.
  cfg5  pneumatic-hand inverse dynamics: 10 000 Jacobian queries per step (A=6 MLP head, 256 + 256 samples):
        encode_image once + 100 infer_optical_flow forward/backward Adam iterations on the collapsed encoding.
A "step" = hoist the lin_z layers onto the feature map + proposal pass + field pass + finish (+ the collective when
N > 1).  The ResNet-34 image encoder runs once per image, outside the timed region of `value` (SURVEY.md section 8d:
excluded on both sides) and inside `e2e`.  In the default run with N > 1 the line also carries `cfg4_strong`, the
strong-scaling number of the 12-view ray-sharded call on the same N GPUs.

  python bench.py --gpus N --steps K --warmup W [--config cfgX]     # this repo's CUDA path
  python bench.py --impl reference ...                              # the reference algorithm on host cores (oracle port)
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
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
IMG_H, IMG_W = 480, 640
CONFIGS = {
    "cfg2": dict(workload="allegro_perception_200x200_s64", head="jacobian_transformer", A=8, s_prop=(64,), s_nerf=64,
                 H=200, W=200, views=1),
    "cfg3": dict(workload="allegro_jacobian_400x400_s128", head="jacobian_transformer", A=8, s_prop=(128,), s_nerf=128,
                 H=400, W=400, views=1),
    "cfg4": dict(workload="allegro_12view_400x400_s128_ray_sharded", head="jacobian_transformer", A=8, s_prop=(128,),
                 s_nerf=128, H=400, W=400, views=12),
    "cfg5": dict(workload="pneumatic_inverse_dynamics_10k_queries_s256", head="jacobian_mlp", A=6, s_prop=(256,),
                 s_nerf=256, H=480, W=640, views=1, queries=10000, iters=100),
}
# reference-formulation work per network evaluation (SURVEY.md section 8d / BASELINE.md section 4)
FLOP_PROPOSAL_SAMPLE = 737_280
FLOP_FIELD_SAMPLE = 1_321_904
GATHER_BYTES_SAMPLE_F16 = 4 * 512 * 2   # 4 taps x 512 channels at the kernels' fp16 storage precision
EXEC_MAC_PROPOSAL = 128 * 64 + 10 * 128 * 128 + 16 * 128            # executed tensor-core MACs / sample
EXEC_MAC_FIELD = EXEC_MAC_PROPOSAL + 64 * 64 + 2 * 64 * 64           # + q_enc + colour head (field_kernel)
EXEC_MAC_XF = 12 * 64 * 64 + 32 * 64                                 # xf_kernel: 3 x (M1, M2, W1, W2) + jacobian_head
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


def scene(cfg, view: int, device=None, pin=False, rays_device=None, coords=None):
    """Synthetic view `view`: context image, cameras of the Allegro rig shape, the target ray grid (or the rays of
    explicit normalised pixel `coords`).  `rays_device`: generate the rays with the library's own kernel on that GPU
    (njf_b200.geometry); None = the reference-side CPU restatement (oracle/synth.py), used by the reference arm,
    which must not touch our kernels."""
    from njf_b200 import synth

    A = cfg["A"]
    g = torch.Generator().manual_seed(2 + view)
    img = torch.rand(1, 3, IMG_H, IMG_W, generator=g)
    K = synth.normalized_intrinsics(**synth.ALLEGRO_INTRINSICS_PX)[None]
    kpx = K.clone(); kpx[:, 0] *= IMG_W; kpx[:, 1] *= IMG_H
    ctxt, trgt = torch.eye(4)[None], synth.relative_target_pose(1 + view % 5)[None]
    if rays_device is not None:
        from njf_b200 import geometry

        if coords is None:
            o, d = geometry.get_world_rays_grid(cfg["H"], cfg["W"], K.to(rays_device), trgt.to(rays_device))
        else:
            o, d = geometry.get_world_rays(coords[None].to(rays_device), K.to(rays_device), trgt.to(rays_device))
        o, d = o.cpu(), d.cpu()
    else:
        if ORACLE not in sys.path:
            sys.path.insert(0, ORACLE)
        import synth as osynth

        o, d = osynth.world_rays(osynth.pixel_grid(cfg["H"], cfg["W"]) if coords is None else coords, K[0], trgt[0])
        o, d = o[None], d[None]
    sc = dict(image=img, ctxt_c2w=ctxt, ctxt_k=K, trgt_c2w=trgt, trgt_k_px=kpx, origins=o.contiguous(),
              dirs=d.contiguous(), z_near=torch.tensor([0.65]), z_far=torch.tensor([3.2]),
              action=0.1 * torch.randn(1, A, generator=g))
    if pin:
        sc = {k: v.pin_memory() for k, v in sc.items()}
    if device is not None:
        sc = {k: v.to(device) for k, v in sc.items()}
    return sc


def hot_weights(cfg):
    from njf_b200 import synth

    return synth.synth_state_dict(synth.field_shapes(cfg["head"], cfg["A"], n_proposal=len(cfg["s_prop"])), 11)


def build_model(cfg, device):
    from njf_b200 import model as M, modules as mod, synth

    mlp = mod.MlpCfg()
    dec = (mod.ActionDecoderJacobianTransformerCfg(name=cfg["head"], mlp=mlp, transformer=mod.TransformerCfg())
           if cfg["head"] == "jacobian_transformer" else mod.ActionDecoderJacobianMlpCfg(name=cfg["head"], mlp=mlp))
    mc = M.ModelCfg(action_dim=cfg["A"], rendering=M.RenderingCfg(cfg["s_prop"], cfg["s_nerf"]), encoder=mod.EncoderResnetCfg(),
                    density_decoder=mod.DensityDecoderMlpCfg("density_mlp", mlp), action_decoder=dec)
    m = M.Model(mc).eval()
    sd = m.state_dict()
    enc = synth.synth_state_dict({k: tuple(v.shape) for k, v in sd.items() if k.startswith("encoder.")}, 11)
    m.load_state_dict({**enc, **hot_weights(cfg)})
    return m.to(device)


def oracle_rays_per_s(cfg, nrays: int, steps: int, warmup: int, want_outputs=False, device=None, threads=None, rays_device=None):
    """The reference algorithm (oracle port, fp32) on a bounded ray sample: on the host cores (all threads)
    or, with `device`, as eager PyTorch on the GPU (what the reference's own code path does on one GPU)."""
    if ORACLE not in sys.path:
        sys.path.insert(0, ORACLE)
    import njf_oracle as O

    if threads:
        torch.set_num_threads(threads)
    w = hot_weights(cfg)
    sc = scene(cfg, 0, rays_device=rays_device)
    g = torch.Generator().manual_seed(9)
    feat = torch.randn(1, 512, IMG_H // 2, IMG_W // 2, generator=g).abs() * 0.7
    idx = torch.randperm(cfg["H"] * cfg["W"], generator=g)[:nrays]
    spec = O.FieldSpec(cfg["head"], cfg["A"])
    if device is not None:
        w = {k: v.to(device) for k, v in w.items()}
        sc = {k: v.to(device) for k, v in sc.items()}
        feat, idx = feat.to(device), idx.to(device)
        sync = torch.cuda.synchronize
    else:
        sync = lambda: None
    run = lambda: O.render_forward(w, spec, feat, sc["ctxt_c2w"], sc["ctxt_k"], sc["trgt_c2w"], sc["trgt_k_px"],
                                   sc["origins"][:, idx], sc["dirs"][:, idx], sc["z_near"], sc["z_far"], sc["action"],
                                   cfg["s_prop"], cfg["s_nerf"])
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


def best_host_threads(cfg):
    """The CPU port is a chain of small eager torch ops; on a many-core host the full thread count is often slower
    than a few dozen threads.  Give the baseline its best configuration: time 32 rays at a few thread counts."""
    n = os.cpu_count() or 1
    best, best_rps = n, 0.0
    for t in sorted({n, min(n, 64), min(n, 32), min(n, 16), min(n, 8)}, reverse=True):
        rps, _, _ = oracle_rays_per_s(cfg, 32, 1, 1, threads=t)
        if rps > best_rps:
            best, best_rps = t, rps
    torch.set_num_threads(best)
    return best, best_rps


def config_dict(cfg, world, extra=None):
    """The `config` object both arms print (same keys, so the driver's same_config check holds)."""
    d = {"workload": cfg["workload"], "head": cfg["head"], "action_dim": cfg["A"],
         "samples": f"{cfg['s_prop'][0]}+{cfg['s_nerf']}", "rays_per_view": cfg["H"] * cfg["W"] if "queries" not in cfg else cfg["queries"],
         "views": cfg["views"], "context_image": f"{IMG_H}x{IMG_W}", "feature_map": f"512x{IMG_H // 2}x{IMG_W // 2}",
         "view": "novel target view",
         "parallelism": (f"ray-shard x{world} (one view per GPU, weak)" if cfg["views"] == 1 else
                         f"{cfg['views']} views flattened and ray-sharded over {world} GPU(s) (strong)"),
         "l2": "flushed between timed steps (256 MiB memset outside the event pairs)",
         "encoder": "excluded from value (once per image, cuDNN), included in e2e",
         "weights": "synthetic seeded (njf_b200/synth.py), random-init architecture of the shipped model yaml"}
    if extra:
        d.update(extra)
    return d


def main_reference(args):
    """The reference's own CPU implementation of the path: /root/reference cannot travel to the GPU box, so this is
    the oracle port (pinned to the unmodified reference by tests/test_oracle_golden.py), all host threads, a bounded
    ray sample of the SAME workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config if args.config in ("cfg2", "cfg3") else "cfg3"]
    if "NJF_BENCH_REF_RAYS" in os.environ:
        nrays = int(os.environ["NJF_BENCH_REF_RAYS"])
        best_host_threads(cfg)
    else:
        # bounded sample: calibrate on 32 rays (untimed), then size a step to a few seconds of host time so that the
        # whole --steps K --warmup W run ends within a few minutes whatever the box's core count
        _, r0 = best_host_threads(cfg)
        budget = min(8.0, 150.0 / max(args.steps + args.warmup, 1))
        nrays = int(max(32, min(512, 32 * round(r0 * budget / 32))))
    rps, dt, _ = oracle_rays_per_s(cfg, nrays, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": rps, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(cfg, args.gpus),   # the same object as this repo's arm prints
        "note": "reference algorithm on the host cores (oracle/njf_oracle.py port; /root/reference is absent on the GPU box); "
                "each step renders a bounded random-ray sample of the frame",
        "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{nrays} random rays of the {cfg['H']}x{cfg['W']} frame per step, {cfg['s_prop'][0]}+{cfg['s_nerf']} samples"},
        "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ============================================================================= this repo's arm
def timed_steps(step, steps, warmup, flush, dist, clocks=None):
    """W untimed steps, then K steps bracketed by barrier + synchronize; per-step CUDA events on the launching stream;
    L2 flushed between steps outside the event pairs.  Returns the summed device time (ms) of the K steps (this rank)
    and the per-step event lists."""
    for _ in range(warmup):
        step(None)
        flush.zero_()
    torch.cuda.synchronize()
    if dist: dist.barrier()
    if clocks: clocks.start()
    timers = []
    torch.cuda.synchronize()
    for _ in range(steps):
        step(timers)
        flush.zero_()
    torch.cuda.synchronize()
    if dist: dist.barrier()
    tot = sum(e[0].elapsed_time(e[-1]) for e in timers)
    return tot, timers


def max_over_ranks(x: float, dev, dist) -> float:
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    if dist: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_cfg4(model, dev, rank, world, dist, steps, warmup, flush):
    """12 views x 160 000 rays, flattened (view, ray) space split contiguously over the ranks."""
    from njf_b200 import api, geometry, parallel as P

    cfg = CONFIGS["cfg4"]
    V, R = cfg["views"], cfg["H"] * cfg["W"]
    start, stop = P.ray_shard(V * R, rank, world)
    v0, v1 = P.shard_views(start, stop, R)
    scs = [scene(cfg, v) for v in range(V)]          # per-view cameras / near / far / action for ALL views (tiny)
    cat = lambda k: torch.cat([s[k] for s in scs]).to(dev)
    ctxt, K, trgt, kpx, zn, zf, act = (cat(k) for k in ("ctxt_c2w", "ctxt_k", "trgt_c2w", "trgt_k_px", "z_near", "z_far", "action"))
    o, d = geometry.get_world_rays_grid(cfg["H"], cfg["W"], K[v0:v1].contiguous(), trgt[v0:v1].contiguous())
    o = o.reshape(-1, 3)[start - v0 * R: stop - v0 * R].contiguous()
    d = d.reshape(-1, 3)[start - v0 * R: stop - v0 * R].contiguous()
    with torch.no_grad():   # this rank encodes only the views its ray range touches (outside the timed region)
        feat = torch.cat([model.encoder(scs[v]["image"].to(dev)).float() for v in range(v0, v1)]).contiguous()
    fld = model.field()
    Hf, Wf = feat.shape[-2:]
    L = api._declare()
    maps = torch.empty(L.njf_hoisted_bytes(fld.handle, V, Hf, Wf), dtype=torch.uint8, device=dev)
    cams, keep = api.make_cameras(ctxt, K, trgt, kpx, dev)
    ws = torch.empty(fld.workspace_bytes(1, stop - start, cfg["s_prop"], cfg["s_nerf"]), dtype=torch.uint8, device=dev)
    bins0, us = api.eval_tables(cfg["s_prop"], cfg["s_nerf"], dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    frames = {}

    def step(timers):
        e = [ev(), ev(), ev()] if timers is not None else None
        if e: e[0].record()
        fld.hoist_views(feat, v0, V, maps)
        res, frame = P.render_sharded(fld, maps, Hf, Wf, cams, o, d, zn, zf, act, cfg["s_prop"], cfg["s_nerf"], V, R, rank,
                                      world, gather=False, vis=False, workspace=ws, bins0=bins0, us=us)
        if e: e[1].record()
        frame = P.gather_rendered(res.packed[0], V * R, dst=0) if world > 1 else res.packed[0]
        if e:
            e[2].record()
            timers.append(e)
        frames["last"] = (res, frame)

    tot, timers = timed_steps(step, steps, warmup, flush, dist)
    t_comm = sum(e[1].elapsed_time(e[2]) for e in timers) / steps
    tot = max_over_ranks(tot, dev, dist)
    res, frame = frames["last"]
    ok = bool(torch.isfinite(res.packed).all())
    return {"workload": cfg["workload"], "scaling": "strong", "value": V * R * steps / (tot * 1e-3), "unit": "rays/s",
            "ms_per_step": tot / steps, "gather_ms": t_comm, "rays_total": V * R, "rays_this_rank": stop - start,
            "views_encoded_this_rank": v1 - v0, "collectives_per_step": "2 x all_reduce(1 float) + 1 gather(packed, dst 0)" if world > 1 else "none",
            "finite": ok, "frame_rows_on_rank0": int(frame.shape[0]) if frame is not None else None}


def run_cfg5(dev, rank, world, dist, steps, warmup, flush):
    """Inverse dynamics: 10 000 Jacobian queries per step, sharded over the ranks; the collapsed encoding (Jbar, p)
    is gathered once per step and rank 0 runs the 100 Adam iterations (preferred variant)."""
    from njf_b200 import parallel as P
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    cfg = CONFIGS["cfg5"]
    model = build_model(cfg, dev)
    Q, A, iters = cfg["queries"], cfg["A"], cfg["iters"]
    g = torch.Generator().manual_seed(5)
    coords = torch.rand(Q, 2, generator=g)
    a, b = P.ray_shard(Q, rank, world)
    sc = scene(cfg, 0, dev, rays_device=dev, coords=coords[a:b])
    cam = CameraInput(sc["image"], sc["ctxt_c2w"], sc["ctxt_k"], sc["trgt_c2w"], sc["trgt_k_px"])
    rin = RenderingInput(sc["origins"], sc["dirs"], sc["z_near"], sc["z_far"])
    u_true = 0.3 * torch.randn(1, A, generator=g).to(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    state = {}

    def gather_enc(enc):
        if world == 1:
            return enc.jbar, enc.p
        packed = torch.cat([enc.jbar[0], enc.p[0]], dim=1).contiguous()
        full = P.gather_rows(packed, Q)
        return full[None, :, :3 * A].contiguous(), full[None, :, 3 * A:].contiguous()

    with torch.no_grad():   # target flow of a known action on all queries
        enc = model.encode_image(cam, rin, RobotInput(sc["action"]))
        jb, pp = gather_enc(enc)
        from njf_b200.model import ModelInferenceEncoding
        full_enc = ModelInferenceEncoding(density=None, action_features=None, weights=enc.weights, ray_samples_positions=None,
                                          jbar=jb, p=pp)
        target = model.infer_optical_flow(full_enc, cam, RobotInput(u_true)).detach()

    def step(timers):
        e = [ev(), ev(), ev()] if timers is not None else None
        if e: e[0].record()
        with torch.no_grad():
            enc = model.encode_image(cam, rin, RobotInput(sc["action"]))
            jb, pp = gather_enc(enc)
        if e: e[1].record()
        fe = ModelInferenceEncoding(density=None, action_features=None, weights=enc.weights, ray_samples_positions=None, jbar=jb, p=pp)
        u = torch.nn.Parameter(torch.zeros(1, A, device=dev))
        opt = torch.optim.Adam([u], lr=0.05)
        for _ in range(iters):      # notebooks/real_world/2_inverse_dynamics.ipynb cell 26
            flow = model.infer_optical_flow(fe, cam, RobotInput(u))
            loss = torch.nn.functional.smooth_l1_loss(flow, target)
            opt.zero_grad()
            loss.backward()
            opt.step()
        if e:
            e[2].record()
            timers.append(e)
        state["u"], state["loss"] = u.detach(), float(loss.detach()) if timers is not None else None

    tot, timers = timed_steps(step, steps, warmup, flush, dist)
    t_enc = sum(e[0].elapsed_time(e[1]) for e in timers) / steps
    tot = max_over_ranks(tot, dev, dist)
    err = float((state["u"] - u_true).abs().max())
    out = {"workload": cfg["workload"], "value": Q * steps / (tot * 1e-3), "unit": "queries/s", "ms_per_step": tot / steps,
           "encode_image_ms": t_enc, "adam_iterations_per_step": iters, "adam_it_per_s": iters / max((tot / steps - t_enc) * 1e-3, 1e-9),
           "action_abs_err_after_step": err, "final_loss": state["loss"]}

    # the same step with the closed-form solver of the fast path (damped Gauss-Newton on the collapsed
    # encoding, csrc/inverse_dynamics.cu) instead of the notebook's 100 Adam iterations
    from njf_b200 import inverse_dynamics as ID
    gn = {}

    def step_gn(timers):
        e = [ev(), ev()] if timers is not None else None
        if e: e[0].record()
        with torch.no_grad():
            enc = model.encode_image(cam, rin, RobotInput(sc["action"]))
            jb, pp = gather_enc(enc)
            fe = ModelInferenceEncoding(density=None, action_features=None, weights=enc.weights, ray_samples_positions=None, jbar=jb, p=pp)
            u, hist = ID.solve_action(fe, cam, target, torch.zeros(1, A, device=dev), iters=4)
        if e:
            e[1].record()
            timers.append(e)
        gn["u"] = u

    try:
        tot_gn, _ = timed_steps(step_gn, steps, warmup, flush, dist)
        tot_gn = max_over_ranks(tot_gn, dev, dist)
        out["gauss_newton_solver"] = {"value": Q * steps / (tot_gn * 1e-3), "unit": "queries/s", "ms_per_step": tot_gn / steps,
                                      "iterations_per_step": 4,
                                      "action_abs_err_after_step": float((gn["u"].float() - u_true).abs().max())}
    except Exception as ex:  # noqa: BLE001 -- an extra leg must not break the headline line
        out["gauss_newton_solver"] = {"error": repr(ex)[:300]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg4 strong-scaling leg of the default run")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs under ncu)")
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
    from njf_b200 import _lib, api, parallel as P
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    L = api._declare()
    L.njf_debug_launch_count.restype = ctypes.c_longlong
    L.njf_debug_launch_count.argtypes = [ctypes.c_int]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cfg = CONFIGS[args.config]

    if args.config in ("cfg4", "cfg5"):
        clocks = ClockSampler(local)
        if rank == 0: clocks.start()
        r = (run_cfg4(build_model(cfg, dev), dev, rank, world, dist, args.steps, args.warmup, flush) if args.config == "cfg4"
             else run_cfg5(dev, rank, world, dist, args.steps, args.warmup, flush))
        clk = clocks.stop() if rank == 0 else None
        if rank == 0:
            line = {"metric": METRIC if args.config == "cfg4" else "Jacobian queries/sec (inverse-dynamics step)",
                    "value": r["value"], "unit": r["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": "f16", "data": "synthetic", "config": config_dict(cfg, world), "detail": r, "clocks": clk}
            print(json.dumps(line))
        if dist: dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ cfg3 / cfg2: one view per GPU (weak scaling)
    A, S_PROP, S_NERF = cfg["A"], cfg["s_prop"], cfg["s_nerf"]
    model = build_model(cfg, dev)
    sc = scene(cfg, rank, dev, rays_device=dev)
    with torch.no_grad():
        feat = model.encoder(sc["image"]).float().contiguous()   # once, outside the timed region
    fld = model.field()
    Hf, Wf = feat.shape[-2:]
    R = cfg["H"] * cfg["W"]
    cams, keep = api.make_cameras(sc["ctxt_c2w"].cpu(), sc["ctxt_k"].cpu(), sc["trgt_c2w"].cpu(), sc["trgt_k_px"].cpu(), dev)
    bins0, us = api.eval_tables(S_PROP, S_NERF, dev)
    f32 = dict(device=dev, dtype=torch.float32)
    maps = torch.empty(L.njf_hoisted_bytes(fld.handle, 1, Hf, Wf), dtype=torch.uint8, device=dev)
    packed = torch.empty(R, 12 + 3 * A, **f32)   # rendered buffers of this rank (one collective)
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
    gathered = torch.empty(world * packed.numel(), **f32) if (world > 1 and rank == 0) else None
    ev = lambda: torch.cuda.Event(enable_timing=True)
    st = api.stream_ptr
    h = fld.handle

    def step(timers):
        e = [ev() for _ in range(5)] if timers is not None else None
        if e: e[0].record()
        _lib.check(L.njf_hoist_features(h, api.dptr(feat), 1, Hf, Wf, api.dptr(maps), st()))
        if e: e[1].record()
        _lib.check(L.njf_proposal_pass(h, ctypes.byref(cams), ctypes.byref(a), 0, api.dptr(bins0), 0, st()))
        if e: e[2].record()
        _lib.check(L.njf_field_pass(h, ctypes.byref(cams), ctypes.byref(a), api.dptr(lb), S_NERF + 1, st()))
        if e: e[3].record()
        _lib.check(L.njf_finish_pass(h, ctypes.byref(cams), ctypes.byref(a), st()))
        if world > 1:   # finish_kernel wrote the packed per-ray struct: ONE gather to rank 0, no torch.cat
            dist.gather(packed.view(-1), list(gathered.view(world, -1).unbind(0)) if rank == 0 else None, dst=0)
        if e:
            e[4].record()
            timers.append(e)

    clocks = ClockSampler(local)
    for _ in range(args.warmup):
        step(None)
        flush.zero_()
    torch.cuda.synchronize()
    if dist: dist.barrier()
    if rank == 0: clocks.start()
    timers = []
    torch.cuda.synchronize()
    _lib.check(L.njf_debug_field_timing(1, None, None))   # per-kernel CUDA events inside njf_field_pass
    L.njf_debug_launch_count(1)
    for _ in range(args.steps):
        step(timers)
        flush.zero_()          # L2 flush between timed steps (outside the per-step event pairs)
    torch.cuda.synchronize()
    launches = int(L.njf_debug_launch_count(1))
    if dist: dist.barrier()
    clk = clocks.stop() if rank == 0 else None
    tot = sum(e[0].elapsed_time(e[4]) for e in timers)            # ms over K steps
    t_hoist = sum(e[0].elapsed_time(e[1]) for e in timers) / args.steps
    t_prop = sum(e[1].elapsed_time(e[2]) for e in timers) / args.steps
    t_field = sum(e[2].elapsed_time(e[3]) for e in timers) / args.steps
    t_tail = sum(e[3].elapsed_time(e[4]) for e in timers) / args.steps
    c_fk, c_xf = ctypes.c_float(0), ctypes.c_float(0)
    _lib.check(L.njf_debug_field_timing(0, ctypes.byref(c_fk), ctypes.byref(c_xf)))
    t_fk, t_xf = c_fk.value / args.steps, c_xf.value / args.steps
    tot = max_over_ranks(tot, dev, dist)
    ms_step = tot / args.steps
    value = world * R * args.steps / (tot * 1e-3)

    # ---- e2e: the public API call (Model.forward, one CUDA-graph launch per frame) with HOST pinned inputs; the
    # encoder, every host->device copy and the device->host read of rgb / depth / flow / Jbar / p / p' are inside
    hs = scene(cfg, rank, pin=True, rays_device=dev)
    cam = CameraInput(hs["image"], hs["ctxt_c2w"], hs["ctxt_k"], hs["trgt_c2w"], hs["trgt_k_px"])
    rin = RenderingInput(hs["origins"], hs["dirs"], hs["z_near"], hs["z_far"])
    rob = RobotInput(hs["action"])
    model.cuda_graph = True
    model.output_device = dev
    host_out = {k: torch.empty(v.shape, dtype=torch.float32).pin_memory() for k, v in outs.items()} if rank == 0 or world == 1 else None
    e2e_gather = torch.empty(world, R, 12 + 3 * A, **f32) if (world > 1 and rank == 0) else None
    host_frame = torch.empty(world, R, 12 + 3 * A).pin_memory() if (world > 1 and rank == 0) else None

    def e2e_step():
        out = model.forward(cam, rin, rob, compute_vis_features=True)
        so, vo = out.standard_output, out.vis_output
        dev_out = dict(rgb=so.rgb, depth=so.depth, flow=so.optical_flow, jbar=vo.action_features, p=vo.ray_positions,
                       pw=vo.ray_positions_warped)
        if world > 1:
            pk = torch.cat([dev_out[k][0] for k in P.PACK_ORDER], dim=1).contiguous()
            dist.gather(pk, list(e2e_gather.unbind(0)) if rank == 0 else None, dst=0)
            if rank == 0:
                host_frame.copy_(e2e_gather, non_blocking=True)
        else:
            for k, t in dev_out.items():
                host_out[k].copy_(t, non_blocking=True)
        torch.cuda.synchronize()
        return float(host_frame[0, 0, 0]) if (world > 1 and rank == 0) else (float(host_out["rgb"][0, 0, 0]) if world == 1 else 0.0)

    e2e_steps = max(3, min(args.steps, 10))
    te = float("nan")
    t_enc = None
    half_leg = None
    with torch.no_grad():
        if not args.no_e2e:
            for _ in range(3):
                e2e_step()
            torch.cuda.synchronize()
            if dist: dist.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            torch.cuda.synchronize()
            te = (time.perf_counter() - t0)
            img_d = hs["image"].to(dev)      # the encoder alone (inside e2e, outside value), for the record
            e0, e1 = ev(), ev()
            model.encoder(img_d)
            e0.record()
            for _ in range(5):
                model.encoder(img_d)
            e1.record()
            torch.cuda.synchronize()
            t_enc = e0.elapsed_time(e1) / 5
            if world == 1 and not args.no_extras:
                # the same call with the opt-in fp16 NHWC encoder (Model.encoder_half), for the record
                try:
                    model.encoder_half = True
                    for _ in range(3):
                        e2e_step()
                    torch.cuda.synchronize()
                    t0h = time.perf_counter()
                    for _ in range(e2e_steps):
                        e2e_step()
                    torch.cuda.synchronize()
                    te_half = (time.perf_counter() - t0h) / e2e_steps * 1e3
                    e0.record()
                    for _ in range(5):
                        model.encoder.forward_nhwc_half(img_d)
                    e1.record()
                    torch.cuda.synchronize()
                    half_leg = {"ms_per_step": te_half, "value": R / (te_half * 1e-3), "encoder_ms": e0.elapsed_time(e1) / 5,
                                "note": "Model.encoder_half = True: encoder under fp16 autocast, njf_hoist_features_nhwc16"}
                except Exception as ex:  # noqa: BLE001 -- an extra leg must not break the headline line
                    half_leg = {"error": repr(ex)[:300]}
                finally:
                    model.encoder_half = False
    te = max_over_ranks(te, dev, dist)
    e2e_value = world * R * e2e_steps / te
    h2d = world * sum(hs[k].numel() * 4 for k in hs)      # every rank copies its view's inputs in
    d2h = world * R * (12 + 3 * A) * 4                     # every rendered frame is read back (on rank 0 when N > 1)

    # ---- strong scaling on the same GPUs: the 12-view ray-sharded call (cfg4)
    cfg4 = None
    if not args.no_extras and world > 1:
        try:
            cfg4 = run_cfg4(model, dev, rank, world, dist, max(2, min(args.steps, 3)), 3, flush)
        except Exception as ex:  # noqa: BLE001 -- an extra leg must not break the headline line
            cfg4 = {"error": repr(ex)[:300]}

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
    # algorithmic bytes: the 4-tap x 512-channel gather (trunk kernels) / the fp16 query stream + weight in,
    # J-bar out (xf_kernel); per-ray inputs and outputs
    alg_bytes = evals * (GATHER_BYTES_SAMPLE_F16 if gathers else 33 * 4) + R * (24 + 144)
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
                     "executes fewer FLOPs (hoist + fold) and gathers from L2-resident fp16 maps (see l2_frac), so "
                     "fractions may exceed 1")
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and args.config == "cfg3":
        tj = json.load(open(tp))
        t = tj.get(dom)   # ncu --set full capture of this build (profiles/ncu_*_summary.json), all launches of a frame summed
        if t:
            roof["traffic"] = t["bytes_per_launch"]          # dram__bytes_read.sum + dram__bytes_write.sum per frame
            roof["traffic_detail"] = t
            roof["traffic_source"] = tj.get("_source")
            if t.get("l2_bytes"):
                roof["l2_frac"] = t["l2_bytes"] / (t_dom * 1e-3) / 1e9 / tj.get("_l2_peak_gbs", 20000.0)
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": config_dict(cfg, world),
        "breakdown_ms": {"hoist": t_hoist, "proposal_kernel": t_prop, "field_pass": t_field, "field_kernel": t_fk, "xf_kernel": t_xf,
                         "finish+gather": t_tail},
        "roofline": roof,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h),
                "ms_per_step": te / e2e_steps * 1e3, "encoder_ms": t_enc,
                "path": "Model.forward(compute_vis_features=True), one CUDA-graph launch per frame (encoder + hoist + render), "
                        "pinned host inputs, rgb/depth/flow/Jbar/p/p' read back to pinned host memory"
                        + ("; per-rank frames gathered to rank 0 first" if world > 1 else ""),
                "half_encoder": half_leg},
        "gpu_launches": launches,
        "clocks": clk,
    }
    if cfg4 is not None:
        line["cfg4_strong"] = cfg4
    if not args.no_cpu_baseline:
        # bounded CPU sample of the same workload + quality vs the oracle on those rays
        nrays = 256
        best_host_threads(cfg)
        rps, dt, (ofeat, idx, oref) = oracle_rays_per_s(cfg, nrays, 1, 0, want_outputs=True, rays_device=dev)
        from njf_b200.render import render
        m2 = fld.hoist(ofeat.to(dev))
        res = render(fld, m2, Hf, Wf, cams, sc["origins"][:, idx.to(dev)].contiguous(), sc["dirs"][:, idx.to(dev)].contiguous(),
                     sc["z_near"], sc["z_far"], sc["action"], S_PROP, S_NERF, sampler_outputs=True)
        torch.cuda.synchronize()
        mse = float(((res.rgb.cpu() - oref["rgb"]) ** 2).mean())
        jr = float((res.jbar.cpu() - oref["action_features"]).norm() / oref["action_features"].norm())
        mism = float((res.level_inds[0].cpu() != oref["inds_1"]).float().mean())
        line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{nrays} random rays of the frame, {S_PROP[0]}+{S_NERF} samples, oracle/njf_oracle.py (torch CPU fp32, best of 5 thread counts), {dt:.1f} s"}
        try:  # the same port as eager PyTorch on this GPU (reference-style single-GPU path), for scale only
            grays = 2048
            grps, gdt, _ = oracle_rays_per_s(cfg, grays, 3, 1, device=dev)
            line["cpu_baseline"]["torch_gpu_port"] = {"value": grps, "unit": "rays/s",
                                                      "sample": f"{grays} rays/step (the reference's patch size), eager torch fp32 on cuda:0, {gdt * 1e3:.0f} ms/step"}
        except Exception as ex:  # noqa: BLE001 -- a baseline leg must not break the bench line
            line["cpu_baseline"]["torch_gpu_port"] = {"error": repr(ex)[:200]}
        # the opt-in fp32 proposal mode (njf_b200/precise.py) on the same rays, and its cost on the whole frame
        from njf_b200 import precise
        trunks = [n.density_head for n in model.proposal_networks]
        pidx = idx.to(dev)
        pargs = (sc["z_near"], sc["z_far"], S_PROP, S_NERF)
        _, _, pinds, _ = precise.proposal_bins_fp32(trunks, ofeat.to(dev), keep[0], keep[1], sc["origins"][:, pidx].contiguous(),
                                                    sc["dirs"][:, pidx].contiguous(), *pargs)
        mism_p = float((pinds[0].cpu() != oref["inds_1"]).float().mean())
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        precise.proposal_bins_fp32(trunks, ofeat.to(dev), keep[0], keep[1], sc["origins"], sc["dirs"], *pargs)
        pe1.record()
        torch.cuda.synchronize()
        line["quality"] = {"psnr_rgb_vs_oracle_db": 10 * math.log10(1.0 / max(mse, 1e-12)),
                           "precise_proposal": {"index_mismatch_rate": mism_p, "proposal_levels_ms_per_frame": pe0.elapsed_time(pe1),
                                                "note": "Model.precise_proposal = True: proposal levels in fp32 (csrc/trunk_train.cu "
                                                        "SIMT kernels), final level on the fused field pass"},
                           "jacobian_rel_l2_vs_oracle": jr, "index_mismatch_rate": mism,
                           "index_mismatch_note": "fraction of PDF-sampler searchsorted indices that differ from the fp32 "
                                                  "oracle end to end (fp16 sigma moves a few CDF ties; the sampler is "
                                                  "bit-exact given identical weights)",
                           "rays": nrays}
    print(json.dumps(line))
    if dist: dist.destroy_process_group()


if __name__ == "__main__":
    main()
