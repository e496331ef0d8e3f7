"""Tuning sweep of the PA apply kernel: for every order and kernel variant prints the element-kernel time,
achieved algorithmic GB/s, fraction of the measured HBM peak and GDOF/s.  Run on the GPU box:
    python tools/sweep.py [--orders 1,2,..] [--variants 0,1,2] [--refine-low 2] [--refine-high 1]
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--orders", default="1,2,3,4,5,6,7,8")
    ap.add_argument("--variants", default="0,20")
    ap.add_argument("--refine-low", type=int, default=2, help="refinements for orders <= 4")
    ap.add_argument("--refine-high", type=int, default=1, help="refinements for orders >= 5")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default="")
    ap.add_argument("--verbose", action="store_true", help="print the launch geometry (threads, shared memory, CTAs/SM) of every kernel variant")
    a = ap.parse_args()
    lpf = importlib.import_module("master-thesis-lpf-in-mfem_b200")
    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream())
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    rows = []
    meshes = {}
    for p in [int(t) for t in a.orders.split(",")]:
        r = a.refine_low if p <= 4 else a.refine_high
        if r not in meshes:
            meshes[r] = lpf.Mesh.wave_tank(128, 2, 16).refine(r)
        sp = lpf.Space(meshes[r], p)
        ctx = lpf.Context(sp, device=0, stream=torch.cuda.current_stream().cuda_stream)
        ctx.pa_setup()
        ctx.set_option("affine", 0)          # the general stored-q-data kernels
        if a.verbose:
            ctx.set_option("verbose", 1)
        x = torch.rand(sp.ndof, dtype=torch.float64, device="cuda") - 0.5
        y = torch.empty_like(x)
        D, Q = p + 1, p + 2
        ab = sp.ne * (48 * Q ** 3 + 4 * D ** 3) + 16 * sp.ndof
        for v in [int(t) for t in a.variants.split(",")]:
            try:
                ctx.set_option("apply_variant", v)
                ctx.time_apply(x, y, 3)
                ms_t, ms_k, _ = ctx.time_apply(x, y, a.reps)
            except lpf.LpfError as e:
                print(f"p={p} v={v}: {e}", flush=True)
                continue
            ms_t /= a.reps; ms_k /= a.reps
            row = dict(order=p, variant=v, refine=r, hexes=sp.ne, dofs=sp.ndof, kernel_ms=ms_k, apply_ms=ms_t,
                       gbs=ab / ms_k / 1e6, frac=ab / ms_k / 1e6 / peak, gdofs_kernel=sp.ndof / ms_k / 1e6,
                       gdofs_apply=sp.ndof / ms_t / 1e6)
            rows.append(row)
            print("p=%d v=%-3d r=%d hexes=%d dofs=%d kernel=%.4f ms apply=%.4f ms  %.0f GB/s  frac=%.3f  %.2f GDOF/s (kernel) %.2f (apply)"
                  % (p, v, r, sp.ne, sp.ndof, ms_k, ms_t, row["gbs"], row["frac"], row["gdofs_kernel"], row["gdofs_apply"]), flush=True)
        ctx.close()
        del sp, ctx, x, y
        torch.cuda.empty_cache()
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
