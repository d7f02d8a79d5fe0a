"""CUDA-event timers around kernel launches (on the launching stream = torch's current stream) and an
nvidia-smi clock sampler, used by bench.py for the live roofline numbers."""
from __future__ import annotations

import subprocess
import threading
from collections import defaultdict
from statistics import median

import torch


class KernelTimer:
    """Accumulates (start, stop) event pairs per kernel class; resolved after a synchronize."""

    def __init__(self):
        self.pairs = defaultdict(list)

    def call(self, name, fn, *args, **kwargs):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*args, **kwargs)
        e1.record()
        self.pairs[name].append((e0, e1))
        return out

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, pairs in self.pairs.items():
            ms = [a.elapsed_time(b) for a, b in pairs]
            out[name] = {"launches": len(ms), "total_ms": sum(ms), "avg_ms": sum(ms) / len(ms)}
        return out

    def reset(self):
        self.pairs.clear()


class ClockSampler:
    """Samples SM clock / throttle reasons with nvidia-smi while a timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0, period_ms: int = 100):
        self.gpu_index, self.period_ms = gpu_index, period_ms
        self.lines = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "reasons": sorted(reasons),
                "samples": len(sm)}
