#!/usr/bin/env python
"""Config 4 of BASELINE.json: batch_inference multi-shot — 4 shots of 480x832x81 frames (S = 8190 video tokens) on one
8-GPU box with shot-parallel x CFG-parallel x Ulysses groups (fairygen_b200.cfg_parallel.Layout).  Times K denoising
steps of every shot (the full job is 50) and prints one JSON line: aggregate denoise steps/s over all shots.
Launch: torchrun --nproc-per-node 8 tools/bench_multishot.py [--layout shots,cfg,sp]"""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--shots", type=int, default=4)
    ap.add_argument("--layout", default="auto", help="'auto' or 'shots,cfg,sp'")
    args = ap.parse_args()
    import torch.distributed as dist

    import fairygen_b200 as fg
    from fairygen_b200 import synthetic
    from fairygen_b200.cfg_parallel import Layout, ParallelContext

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", init_method="env://", device_id=dev)
    cfg = fg.TI2V_5B
    layout = Layout.auto(world, args.shots, True, cfg.num_heads) if args.layout == "auto" else Layout(world, *map(int, args.layout.split(",")))
    par = ParallelContext(layout)
    shape = synthetic.latent_shape(cfg, 480, 832, 81)
    tokens = shape[2] * (shape[3] // 2) * (shape[4] // 2)
    eng = fg.WanDiTEngine(cfg, dev, sp=par.sequence_parallel())
    eng.load_state_dict(synthetic.random_state_dict(cfg, seed=0, device=dev, dtype=torch.bfloat16, lora_rank=32))
    den = fg.WanDenoiser(eng, 50, 5.0, 5.0, cfg_group=par if layout.cfg > 1 else None)
    mine = layout.shots_of(rank, args.shots)
    data = {}
    for i in mine:   # different prompt / noise per shot
        lat, z0, cp, cn = synthetic.synthetic_inputs(cfg, shape, text_len=512, pin=False)
        g = torch.Generator().manual_seed(100 + i)
        data[i] = [(lat.float() + 0.01 * torch.randn(lat.shape, generator=g)).to(torch.bfloat16).to(dev), z0.to(dev), cp.to(dev), cn.to(dev)]

    def run(first, count):
        for i in mine:
            lat, z0, cp, cn = data[i]
            for s in range(first, first + count):
                den.step(s, lat, cp, cn, z0)

    run(0, args.warmup)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(args.warmup, args.steps)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    finite = all(bool(torch.isfinite(v[0].float()).all()) for v in data.values())
    if rank == 0:
        total_steps = args.shots * args.steps
        flops = 2 * fg.counted_flops(cfg, tokens, 512) * total_steps
        print(json.dumps({
            "metric": "multishot_denoise_steps_per_s", "value": total_steps / (ms * 1e-3), "unit": "steps/s (all shots)", "n_gpus": world,
            "ms_total": ms, "steps_per_shot": args.steps, "shots": args.shots, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.shots} shots x 480x832x81 (S={tokens}), CFG on, merged rank-32 LoRA",
                       "parallelism": f"shots{layout.shots}_x_cfg{layout.cfg}_x_ulysses_sp{layout.sp}"},
            "achieved_tflops": flops / (ms * 1e-3) / 1e12, "seconds_for_50_steps_all_shots": 50 * args.shots / (total_steps / (ms * 1e-3)),
            "output_finite": finite}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
