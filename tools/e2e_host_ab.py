"""A/B of the pipelined host-buffer apply (lpf_apply_T_host, bench.py `e2e`): number of element chunks (= copies per direction)
and resolution of the dependency tracking, against the un-pipelined H2D -> apply -> D2H sequence and the PCIe ceiling
(tools/microbench/pcie_bw.py).
    python tools/e2e_host_ab.py [--refine 2] [--order 4]
"""
import argparse
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--refine", type=int, default=2)
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    lpf = importlib.import_module("master-thesis-lpf-in-mfem_b200")
    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream())
    sp = lpf.Space(lpf.Mesh.wave_tank(128, 2, 16).refine(a.refine), a.order)
    ctx = lpf.Context(sp, device=0, stream=torch.cuda.current_stream().cuda_stream)
    ctx.pa_setup()
    ctx.set_option("affine", 0)
    n = sp.ndof
    xh = (torch.rand(n, dtype=torch.float64) - 0.5).pin_memory()
    yh = torch.empty(n, dtype=torch.float64).pin_memory()
    yref = None

    def run(label):
        nonlocal yref
        ctx.apply_T_host(xh, yh)
        if yref is None:
            yref = yh.clone()
        err = float((yh - yref).abs().max() / yref.abs().max())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            ctx.apply_T_host(xh, yh)
        ms = (time.perf_counter() - t0) / a.reps * 1e3
        print(f"{label:34s} {ms:7.3f} ms  {n / ms / 1e6:6.2f} GDOF/s  {2 * 8 * n / ms / 1e6:6.1f} GB/s over PCIe (both directions)  max rel diff {err:.1e}", flush=True)

    ctx.set_option("host_pipeline", 0)
    run("un-pipelined")
    ctx.set_option("host_pipeline", 1)
    for K in (2, 4, 6, 8, 12, 16, 32, 64):
        ctx.set_option("hp_chunks", K)
        run(f"pipelined, {K} chunks")
    ctx.close()


if __name__ == "__main__":
    main()
