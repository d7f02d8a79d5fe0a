#!/usr/bin/env python
"""Config 5 of BASELINE.json: stage-2 motion-LoRA fine-tune step (rank 32, unmerged adapters), 480x832x49 frames
(S = 5070 video tokens, 512 text tokens), all 30 TI2V-5B blocks, on 1 GPU (N GPUs = N data-parallel replicas with an
all-reduce of the lora_B2 gradient).  Prints one JSON line; FLOP accounting follows SURVEY.md §8(d):
fwd = counted(S), bwd = x-GEMM FLOPs (dgrad only) + 2.5 x attention-forward FLOPs."""
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
    ap.add_argument("--recompute", action="store_true", help="the reference's per-block checkpointing schedule")
    ap.add_argument("--profile", action="store_true", help="print the per-kernel CUDA time of one step (torch.profiler / CUPTI)")
    ap.add_argument("--sp", action="store_true", help="sequence-parallel: the N ranks split the tokens of ONE video (strong scaling) "
                                                     "instead of running N data-parallel replicas")
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=832)
    ap.add_argument("--frames", type=int, default=49)
    args = ap.parse_args()
    import torch.distributed as dist

    import fairygen_b200 as fg
    from fairygen_b200 import synthetic
    from fairygen_b200.training import Stage2Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", init_method="env://", device_id=dev)
        group = dist.group.WORLD
    cfg = fg.TI2V_5B
    shape = synthetic.latent_shape(cfg, args.height, args.width, args.frames)
    tokens = shape[2] * (shape[3] // 2) * (shape[4] // 2)
    sp = fg.SequenceParallel(group, exchange="p2p") if args.sp and world > 1 else None
    eng = fg.WanDiTEngine(cfg, dev, sp=sp)
    eng.load_state_dict(synthetic.random_state_dict(cfg, seed=0, device=dev, dtype=torch.bfloat16))
    lora = synthetic.random_lora(cfg, rank=32, seed=2, device=dev)
    tr = Stage2Trainer(eng, lora, rank=32, recompute=args.recompute, dp_group=None if sp is not None else group)
    tr.b2_flat.normal_(0, 0.02, generator=torch.Generator(device=dev).manual_seed(6))
    x0, _, ctx, _ = synthetic.synthetic_inputs(cfg, shape, text_len=512, pin=False)
    noise = torch.randn(shape, generator=torch.Generator().manual_seed(9 + (0 if sp is not None else rank))).to(torch.bfloat16)
    x0, ctx, noise = x0.to(dev), ctx.to(dev), noise.to(dev)

    def one(i):
        tr.zero_grad()
        tr.step(x0, noise, 500, ctx, seed=1000 + i)
        tr.optimizer_step(lr=1e-4)

    for i in range(args.warmup):
        one(i)
    torch.cuda.synchronize()
    if args.profile and world == 1:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            one(99)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70), file=sys.stderr)
    if world > 1:
        dist.barrier()
    tr.kernel_launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        one(args.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        d, f, L = cfg.dim, cfg.ffn_dim, 512
        gemm = 30 * (12 * tokens * d * d + 4 * tokens * d * f)
        attn = 30 * (4 * tokens * tokens * d + 4 * tokens * L * d)
        fwd = fg.counted_flops(cfg, tokens, L)
        bwd = gemm + 2.5 * attn
        total = fwd + bwd + (fwd if args.recompute else 0)
        print(json.dumps({
            "metric": "stage2_lora_train_steps_per_s", "value": (1 if sp is not None else world) * 1e3 / ms, "unit": "steps/s", "n_gpus": world, "ms_per_step": ms,
            "scaling": "strong" if sp is not None else "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"Wan2.2-TI2V-5B stage-2 motion-LoRA fine-tune step, {args.height}x{args.width}x{args.frames} "
                                   f"(S={tokens}), rank 32 unmerged adapters, only lora_B2 trainable, AdamW", "recompute": args.recompute,
                       "parallelism": f"sp{world}" if sp is not None else f"dp{world}"},
            "counted_flops_per_step": total, "achieved_tflops_per_gpu": total / (ms * 1e-3) / 1e12 / (world if sp is not None else 1),
            "gpu_launches": tr.kernel_launches, "loss": float(tr.loss_buf), "grad_norm": float(tr.grad_flat.norm()),
            "finite": bool(torch.isfinite(tr.grad_flat).all()),
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
