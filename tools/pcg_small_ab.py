"""A/B helper: ms per CG iteration of the RK4 leg on a given tank, for the library found under <repo root> (argv[1]).
    python tools/pcg_small_ab.py <repo root> <refine> [name=value ...]"""
import importlib
import os
import sys

import numpy as np
import torch

root, refine = sys.argv[1], int(sys.argv[2])
sys.path.insert(0, root)
lpf = importlib.import_module("master-thesis-lpf-in-mfem_b200")
torch.cuda.set_device(0)
torch.cuda.set_stream(torch.cuda.Stream())
sp = lpf.Space(lpf.Mesh.wave_tank(128, 2, 16).refine(refine), 4)
ctx = lpf.Context(sp, device=0, stream=torch.cuda.current_stream().cuda_stream)
for kv in sys.argv[3:]:
    ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
ctx.pa_setup(); ctx.set_option("affine", 0); ctx.jacobi_setup()
w = lpf.wave_params()
dt = w["T"] / 150
ctx.rhs_setup(lpf.make_rhs_params(w, rel_tol=1e-12, max_iter=2000))
ph = -w["k"] * sp.surf_xy[:, 0]
st = np.concatenate([0.5 * w["H"] * np.cos(ph), -0.5 * w["H"] * w["cwave"] / np.tanh(w["kh"]) * np.sin(ph)])
sd = torch.from_numpy(st).cuda()
t = ctx.rk4_step(sd, 0.0, dt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    t = ctx.rk4_step(sd, t, dt)
e1.record(); torch.cuda.synchronize()
its = sum(i.iterations for i in ctx.last_solve_info())
print(f"{os.path.basename(os.path.abspath(root)):8s} r={refine} {' '.join(sys.argv[3:]):28s} {e0.elapsed_time(e1) / 3:9.3f} ms per RK4 step, {1e3 * e0.elapsed_time(e1) / 3 / its:7.2f} us per CG iteration ({its} its)")
ctx.close()
del ctx, sp
