"""Training steps at the reference's training shape
(``--phase action``: cross-attention head through the fused path, csrc/xf_backward.cu; ``--phase perception``: every
parameter trainable, rgb + weights losses, csrc/trunk_train.cu; ``--phase action_mlp``: the MLP Jacobian head).
Action-phase training step through the fused path: the reference's training shape
(configurations/config.yaml:18-20: batch 7 views x 256 rays, model_allegro.yaml: 256 + 256 samples per ray, A = 8
cross-attention head), everything but the Jacobian head frozen (models/model_wrapper.py:75-85), masked flow loss
(:148-163), Adam.  Prints one JSON object: ms per step split into forward / backward / optimiser + re-pack."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench  # noqa: E402


def main():
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--phase", default="action", choices=["action", "perception", "action_mlp"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tf32", action="store_true", help='torch.set_float32_matmul_precision("high") like the reference train.py:64-65')
    ap.add_argument("--profile", action="store_true", help="print the kernels of one step by device time (torch.profiler)")
    args = ap.parse_args()
    if args.tf32:
        torch.set_float32_matmul_precision("high")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    import __graft_entry__ as ge

    ge.build()
    from njf_b200 import synth
    from njf_b200.model import CameraInput, RenderingInput, RobotInput

    cfg = dict(bench.CONFIGS["cfg5" if args.phase == "action_mlp" else "cfg3"], s_prop=(256,), s_nerf=256)
    model = bench.build_model(cfg, dev)
    if args.phase != "perception":
        model.decoder.freeze_non_action_parameters()
        for n, p in model.named_parameters():
            if "decoder" not in n:
                p.requires_grad = False
    model.train()
    B, R, A = 7, 256, cfg["A"]
    g = torch.Generator().manual_seed(0)
    img = torch.rand(B, 3, bench.IMG_H, bench.IMG_W, generator=g).to(dev)
    K = synth.normalized_intrinsics(**synth.ALLEGRO_INTRINSICS_PX)[None].repeat(B, 1, 1)
    kpx = K.clone(); kpx[:, 0] *= bench.IMG_W; kpx[:, 1] *= bench.IMG_H
    ctxt = torch.eye(4)[None].repeat(B, 1, 1)
    trgt = torch.stack([synth.relative_target_pose(1 + b % 5) for b in range(B)])
    from njf_b200 import geometry
    coords = torch.rand(B, R, 2, generator=g).to(dev)
    o, d = geometry.get_world_rays(coords, K.to(dev), trgt.to(dev))
    cam = CameraInput(img, ctxt.to(dev), K.to(dev), trgt.to(dev), kpx.to(dev))
    rin = RenderingInput(o, d, torch.full((B,), 0.65, device=dev), torch.full((B,), 3.2, device=dev))
    rob = RobotInput((0.3 * torch.randn(B, A, generator=g)).to(dev))
    target = (2.0 * torch.randn(B, R, 2, generator=g)).to(dev)
    target_rgb, target_depth = torch.rand(B, R, 3, generator=g).to(dev), (0.7 + 2.0 * torch.rand(B, R, 1, generator=g)).to(dev)

    def loss_of(out):
        if args.phase != "perception":
            return 0.01 * torch.nn.functional.mse_loss(out.standard_output.optical_flow, target)
        # model_wrapper.py:116-146: rgb MSE + depth / interlevel / distortion terms on every level's weights (stand-ins
        # of the same cost: reductions over (B,R,S))
        to = out.training_output
        l = torch.nn.functional.mse_loss(out.standard_output.rgb, target_rgb)
        for w, sb in zip(to.weights_list, to.ray_samples_list):
            mid = (sb.starts + sb.ends) / 2
            l = l + 0.08 * ((w * mid).sum(-2) - target_depth).pow(2).mean() / len(to.weights_list) + 0.01 * (w * w).sum(-2).mean()
        return l
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-5)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t_f = t_b = t_o = 0.0
    steps, warm = args.steps, args.warmup
    walls = []
    losses = []
    for it in range(steps + warm):
        e = [ev() for _ in range(4)]
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        e[0].record()
        out = model.forward(cam, rin, rob)
        loss = loss_of(out)
        e[1].record()
        opt.zero_grad()
        loss.backward()
        e[2].record()
        opt.step()
        e[3].record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        losses.append(float(loss))
        if it >= warm:
            t_f += e[0].elapsed_time(e[1]); t_b += e[1].elapsed_time(e[2]); t_o += e[2].elapsed_time(e[3])
            walls.append(wall)
    ms = (t_f + t_b + t_o) / steps
    if args.profile:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            out = model.forward(cam, rin, rob)
            loss = loss_of(out)
            opt.zero_grad()
            loss.backward()
            opt.step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70), file=sys.stderr)
    what = {"action": "action-phase training step, A=8 cross-attention head (fused forward, csrc/xf_backward.cu)",
            "perception": "perception-phase training step, every parameter trainable incl. the encoder (csrc/trunk_train.cu)",
            "action_mlp": "action-phase training step, A=6 MLP Jacobian head (csrc/trunk_train.cu)"}[args.phase]
    print(json.dumps({"workload": what + ", 7 views x 256 rays, 256+256 samples", "phase": args.phase,
                      "ms_per_step": ms, "forward_ms": t_f / steps, "backward_ms": t_b / steps, "optimizer_ms": t_o / steps,
                      "wall_ms_median": sorted(walls)[len(walls) // 2] * 1e3, "tf32_library_gemms": bool(args.tf32), "train_rays_per_s": B * R / (ms * 1e-3),
                      "note": "forward includes the encoder (7 images, BatchNorm in train mode), the per-step re-pack of the "
                              "changed head weights (njf_field_create) and the hoist",
                      "loss_first_last": [losses[0], losses[-1]],
                      "reference": "README.md:142-143: ~10.7 steps/s on one A40 (~19 k train-rays/s)"}, indent=1))


if __name__ == "__main__":
    main()
