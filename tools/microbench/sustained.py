"""Burst vs sustained: the order-4 apply kernel and a plain device copy, each timed in 50 ms windows for ~3 s after an idle
period, with SM clock / power sampled alongside.  Explains why the first measurement of a process is 10-14 % faster than the
following ones (profiles/r02_sweep_orders.txt) and what the long PCG legs of bench.py run against."""
import importlib
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

lpf = importlib.import_module("master-thesis-lpf-in-mfem_b200")
torch.cuda.set_device(0)
torch.cuda.set_stream(torch.cuda.Stream())
sp = lpf.Space(lpf.Mesh.wave_tank(128, 2, 16).refine(2), 4)
ctx = lpf.Context(sp, device=0, stream=torch.cuda.current_stream().cuda_stream)
ctx.pa_setup(); ctx.set_option("affine", 0)
x = torch.rand(sp.ndof, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty_like(x)
a = torch.empty(1 << 29, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)     # 2 GiB each
ab = sp.ne * (48 * 6 ** 3 + 4 * 5 ** 3) + 16 * sp.ndof
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,temperature.memory,clocks_event_reasons.sw_power_cap,clocks_event_reasons.sw_thermal_slowdown",
                        "--format=csv,noheader", "-lms", "50"], stdout=subprocess.PIPE, text=True)


def windows(label, fn, bytes_per_call, calls, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for w in range(n):
        e0.record()
        for _ in range(calls):
            fn()
        e1.record(); e1.synchronize()
        out.append(bytes_per_call * calls / e0.elapsed_time(e1) / 1e6)
    print(label, " ".join("%.0f" % v for v in out), "GB/s per window", flush=True)


for rnd in range(2):
    time.sleep(3.0)                                   # idle: the GPU cools down / clocks drop
    windows("copy 2 GiB (read+write)", lambda: b.copy_(a), 2 * a.numel() * 4, 70, 40)
    time.sleep(3.0)
    windows("apply kernel order 4   ", lambda: ctx.time_apply(x, y, 1), ab, 40, 40) if False else None
    e = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for w in range(40):
        _, ms_k, _ = ctx.time_apply(x, y, 45)
        e.append(ab * 45 / ms_k / 1e6)
    print("apply kernel order 4 (algorithmic)", " ".join("%.0f" % v for v in e), "GB/s per window of 45 launches (+ 45 whole applies in between)", flush=True)
smi.terminate()
lines = smi.stdout.read().strip().splitlines()
print("nvidia-smi (every 50 ms): sm clock / power / T gpu / T mem / power cap / thermal")
for l in lines[::6]:
    print("  ", l)
ctx.close()
