#!/usr/bin/env python
"""Turns the raw ncu output of tools/profile_round.sh (gpurun_out/<tag>_*) into the small summaries kept under profiles/:

    python tools/ncu_summarise.py r02

  <tag>_launches_summary.csv      per-kernel launch count / total time / share of the LAST bench step in the launch list
  <tag>_ncu_<name>_summary.csv    the handful of `ncu --set full` metrics the roofline discussion uses, one launch each
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "gpurun_out")
PROF = os.path.join(REPO, "profiles")

METRICS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.per_cycle_active",
    "launch__registers_per_thread",
    "launch__block_size",
    "launch__grid_size",
    "launch__cluster_size",
    "launch__shared_mem_per_block_dynamic",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def short_name(kernel: str) -> str:
    k = re.sub(r"^void ", "", kernel)
    k = k.replace("fgb::", "")
    return re.sub(r"\(.*$", "", k)


def launches(tag: str):
    raw = os.path.join(OUT, f"{tag}_launches_raw.csv")
    if not os.path.exists(raw):
        return
    rows = []
    with open(raw) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r["Metric Name"] == "gpu__time_duration.sum":
            rows.append((short_name(r["Kernel Name"]), float(r["Metric Value"]) * (1e-6 if r["Metric Unit"] == "ns" else 1e-3)))
    # the last denoise step = everything after the second-to-last cfg_fm_step_kernel up to and including the last one
    ends = [i for i, (k, _) in enumerate(rows) if k.startswith("cfg_fm_step")]
    sel = rows[ends[-2] + 1:ends[-1] + 1] if len(ends) >= 2 else rows
    agg = {}
    for k, ms in sel:
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + ms)
    total = sum(t for _, t in agg.values())
    path = os.path.join(PROF, f"{tag}_launches_summary.csv")
    with open(path, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:fgb:: python bench.py --steps 1 "
                "--warmup 1 --no-cpu-baseline --no-extras  (tools/profile_round.sh; N=1, S=27280)\n")
        f.write(f"# {len(rows)} launches captured in all; below: the {len(sel)} launches of the LAST denoise step of the run (2 DiT forwards + the fused "
                "CFG/Euler kernel).\n# per-launch times are cold-cache, serialised and at unthrottled clocks: compare SHARES with the bench line's "
                "live CUDA-event shares (kernel_ms_per_step), not absolutes\n")
        f.write("kernel,launches,total_ms,share_pct\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{n},{t:.3f},{100 * t / total:.2f}\n")
    plain = os.path.join(OUT, f"{tag}_prof_plain.json")
    if os.path.exists(plain):
        with open(plain) as f:
            d = json.loads(f.read().strip().splitlines()[-1])
        live = d["kernel_ms_per_step"]
        tot = sum(live.values())
        with open(path, "a") as f:
            f.write("# live CUDA-event shares of the plain run of the same command (bench line kernel_ms_per_step, "
                    f"{d['ms_per_step']:.1f} ms/step):\n")
            for k, v in live.items():
                f.write(f"# live,{k},{v:.3f},{100 * v / tot:.2f}\n")
    print("wrote", path)


def full(tag: str, name: str, header: str):
    rep = os.path.join(OUT, f"{tag}_ncu_{name}.ncu-rep")
    if not os.path.exists(rep):
        return
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, vals = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(head)}
    path = os.path.join(PROF, f"{tag}_ncu_{name}_summary.csv")
    with open(path, "w") as f:
        f.write(f"# {header}\n")
        f.write(f"Kernel Name,,{short_name(vals[idx['Kernel Name']])}\n")
        for m in METRICS:
            if m in idx:
                f.write(f"{m},{units[idx[m]]},{vals[idx[m]]}\n")
    print("wrote", path)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    launches(tag)
    cmd = "ncu --set full --clock-control none --import-source on, one launch of `KCHECK_BOUNDED=1 tools/kcheck {} 3 0` (tools/profile_round.sh)"
    full(tag, "attn", cmd.format("attn 27280 27280 24") + " — self-attention of the headline shape, bounded-score softmax")
    full(tag, "attn_cross", cmd.format("attn 27280 512 24") + " — cross-attention of the headline shape")
    full(tag, "gemm_ffn1", cmd.format("gemm 27280 14336 3072 1") + " — FFN1 (bias + GELU-tanh epilogue), 2-CTA kernel")
    full(tag, "gemm_ffn2", cmd.format("gemm 27280 3072 14336 2") + " — FFN2 (gated residual epilogue), 2-CTA kernel")


if __name__ == "__main__":
    main()
