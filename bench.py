#!/usr/bin/env python
"""bench.py -- PA Laplace matvec GDOF/s (+ PCG time per RK4 step) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference's CPU path (C restatement, all host threads)

A "step" is one constrained operator apply y = (P^T A P)_c x on the whole mesh (the operation CGSolver
calls once per iteration).  Workload at N=1: wave-tank-big8 (128x2x16 hexes, x-periodic)
uniformly refined twice, H1 order 4: 262 144 hexes, 17 369 088 dofs, 2.7 GB of q-data (>> 126 MB L2, so
no L2 flush is needed between iterations).  N>1: weak scaling, the tank is N times longer (128 N cells in
x) and is cut into N x-slabs, one per GPU; the halo-sum rides on the element kernel (NVLink peer memory, own kernels).
Extra objects on the line: pcg_per_rk4_step (weak), strong_scaling (same total mesh on every N), e2e_rk4 (host-API RK4 step),
parity (N > 1: N ranks against 1 rank), cpu_baseline (C restatement of MFEM's CPU PA path on the host cores).
Prints ONE JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pa_laplace_matvec_gdofs"
UNIT = "GDOF/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--refine", type=int, default=2, help="uniform refinements of the 128x2x16 tank")
    ap.add_argument("--variant", type=int, default=0, help="apply kernel variant (tuning)")
    ap.add_argument("--no-rk4", action="store_true", help="skip the PCG-per-RK-step measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--rk4-refine", type=int, default=1)
    ap.add_argument("--no-sustained", action="store_true", help="skip the burst-vs-sustained leg (about 1 s of back-to-back applies)")
    ap.add_argument("--opt", action="append", default=[], help="extra lpf_set_option name=value (tuning A/B), repeatable")
    ap.add_argument("--p2p-fuse", type=int, default=-1, help="N>1: option p2p_fuse (-1 = library default: halo exchange inside the apply kernel, overlapped)")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"], help="N>1: own NVLink peer-memory exchange or NCCL send/recv + all-reduce")
    return ap.parse_args()


def workload_name(a, world=1):
    tank = "128x2x16" if world == 1 else f"{128 * world}x2x16"
    return f"wave-tank-big8({tank},x-periodic) r={a.refine} order={a.order}"


def config_dict(a, world):
    """identical in both arms (`--impl ours` / `--impl reference`): it names the WORKLOAD, not the implementation"""
    f, p = 2 ** a.refine, a.order
    hexes = 128 * 2 * 16 * f ** 3
    dofs = (128 * world * f * p) * (2 * f * p + 1) * (16 * f * p + 1)
    return {"workload": workload_name(a, world), "hexes_per_gpu": hexes, "dofs_global": dofs, "order": p,
            "l2_policy": "inputs larger than L2 (q-data %.2f GB per GPU), no flush" % (hexes * 48 * (p + 2) ** 3 / 1e9),
            "parallelism": f"x-slab domain decomposition x{world}, one slab per GPU" if world > 1 else "single GPU"}


def algorithmic_bytes(ne, ndof, p):
    """NE (48 Q^3 + 4 D^3) + 16 N_L  (q-data + gather map + read x + write y)."""
    D, Q = p + 1, p + 2
    return ne * (48 * Q ** 3 + 4 * D ** 3) + 16 * ndof


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(p, ne):
    """dram bytes per launch of the apply kernel from the committed ncu capture (profiles/apply_traffic.json),
    scaled to this launch's element count (traffic is proportional to the number of hexes)."""
    f = os.path.join(ROOT, "profiles", "apply_traffic.json")
    if os.path.exists(f):
        d = json.load(open(f))
        if f"order{p}" in d:
            return int(d[f"order{p}"] * ne / d["hexes"][f"order{p}"])
    return None


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the C restatement of MFEM's CPU PA path on the host cores.  Mesh, numbering, basis tables
# and arithmetic all come from oracle/ (oracle/cpu_reference.py); the product library is NOT loaded on this path.
# ------------------------------------------------------------------------------------------------
def cpu_legs(a, steps, warmup, with_rk4=True):
    """(cpu_baseline object, seconds per apply).  Apply: the SAME tank as the GPU arm's per-GPU workload (N = 1: the whole
    workload).  PCG: one full RK4 step (4 Jacobi-PCG solves to rel 1e-12) on the tank of the GPU arm's RK4 leg."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_reference as cr                      # baseline only
    f = 2 ** a.refine
    tk = cr.CpuTank(128 * f, 2 * f, 16 * f, a.order)
    sec = tk.time_applies(steps, warmup)
    out = {"value": tk.ndof / sec / 1e9, "unit": UNIT, "cores": tk.threads, "kind": "port",
           "sample": f"{steps} applies (+{warmup} warm-up) on wave-tank-big8 r={a.refine} order={a.order}: {tk.describe()}"
                     + (f" = one of the {a.gpus} x-slabs of the workload" if a.gpus > 1 else " = the whole workload"),
           "ms_per_apply": sec * 1e3,
           "what": "C/OpenMP restatement of MFEM's CPU partial-assembly path (oracle/pa_oracle.c, -march=native); MFEM/hypre/MPI are not installable here"}
    del tk
    if with_rk4:
        try:
            f1 = 2 ** a.rk4_refine
            t1 = cr.CpuTank(128 * f1, 2 * f1, 16 * f1, a.order)
            wv = cr.orc.Wave()
            sec_rk4, its = t1.time_rk4_step(wv.T / 150, 1e-12, 2000)
            out["rk4_ms_per_step"] = sec_rk4 * 1e3
            out["pcg_ms_per_cg_iteration"] = sec_rk4 * 1e3 / max(1, sum(its))
            out["rk4_cg_iterations_per_stage"] = its
            out["rk4_sample"] = f"one full RK4 step (dt = T/150, 4 Jacobi-PCG solves, rel 1e-12) on wave-tank-big8 r={a.rk4_refine} order={a.order}: {t1.describe()}"
        except Exception as e:                      # the baseline must never take the other numbers down with it
            out["rk4_error"] = str(e)[:200]
    return out, sec


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, a.steps), max(1, a.warmup)
    cb, sec = cpu_legs(a, steps, warmup, with_rk4=not a.no_rk4)
    # whole-job value: throughput is per host, the workload is N times one slab -- the host needs N times as long
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": sec * 1e3 * a.gpus, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(a, a.gpus),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if "rk4_ms_per_step" in cb:
        line["e2e_rk4"] = {"value": cb["rk4_ms_per_step"], "unit": "ms per RK4 step", "higher_is_better": False,
                           "workload": cb["rk4_sample"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    lpf = importlib.import_module("master-thesis-lpf-in-mfem_b200")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    torch.cuda.set_stream(torch.cuda.Stream())      # explicit stream shared by torch events and the C-ABI calls
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = a.order
    mesh = lpf.Mesh.wave_tank(128 * world, 2, 16, Lx=1.0 * world).refine(a.refine)
    sp = lpf.Space(mesh, p, nranks=world, rank=rank)
    stream = torch.cuda.current_stream().cuda_stream
    ctx = lpf.Context(sp, device=local, stream=stream)
    ctx.set_option("apply_variant", a.variant)
    if a.p2p_fuse >= 0:
        ctx.set_option("p2p_fuse", a.p2p_fuse)
    for kv in a.opt:
        ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    if world > 1 and a.comm == "p2p":
        ctx.p2p_connect(dist)
    elif world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(lpf.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        ctx.comm_init(bytes(idt.cpu().numpy().tobytes()))
    ctx.pa_setup()
    # The graded number is the GENERAL operator (stored q-data, 48 Q^3 bytes per element).  The tank is made
    # of affine hexes, for which the library would by default switch to its affine fast path (6 doubles per element):
    # that is measured separately below ("affine_fastpath") and never mixed into value / roofline / e2e.
    ctx.set_option("affine", 0)
    n = sp.ndof
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    y = torch.empty_like(x)
    ndof_global = int(sp.n_true_global)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, a.warmup)):
        ctx.apply_T(x, y)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(a.steps):
        ctx.apply_T(x, y)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launches - l0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t[0])
    # element kernel alone (roofline numerator), CUDA events on the launching stream inside the library
    _, ms_k, _ = ctx.time_apply(x, y, a.steps)
    tk = torch.tensor([ms_k], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tk, op=dist.ReduceOp.MAX)
    ms_kernel = float(tk[0]) / a.steps

    # end to end through the C-ABI with HOST buffers (pinned), H2D + apply + D2H per step
    xh = torch.empty(n, dtype=torch.float64).pin_memory()
    yh = torch.empty(n, dtype=torch.float64).pin_memory()
    xh.copy_(x.cpu())
    e2e_steps = max(2, min(a.steps, 10))
    ctx.apply_T_host(xh, yh)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.apply_T_host(xh, yh)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = ndof_global * e2e_steps / float(te[0]) / 1e9
    # extra: affine fast path (same operator, q-data stream replaced by one 6-entry tensor per element)
    aff = None
    ctx.set_option("affine", 1)
    if ctx.affine_active:
        for _ in range(3):
            ctx.apply_T(x, y)
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for _ in range(a.steps):
            ctx.apply_T(x, y)
        eb.record()
        barrier()
        ta = torch.tensor([ea.elapsed_time(eb)], dtype=torch.float64, device="cuda")
        _, msk_a, _ = ctx.time_apply(x, y, a.steps)
        tka = torch.tensor([msk_a], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ta, op=dist.ReduceOp.MAX)
            dist.all_reduce(tka, op=dist.ReduceOp.MAX)
        aff = {"value": ndof_global * a.steps / (float(ta[0]) * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": float(ta[0]) / a.steps,
               "kernel_ms": float(tka[0]) / a.steps,
               "note": "extra, NOT the graded number: affine hexes only (all wave tanks); D(q) = w_q * element tensor, no q-data stream"}
    ctx.set_option("affine", 0)
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    # per-rank spread of the element kernel (which rank limits the weak-scaling number, and by how much)
    kernel_ranks = [ms_k / a.steps]
    if world > 1:
        allk = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(allk, torch.tensor([ms_k / a.steps], dtype=torch.float64, device="cuda"))
        kernel_ranks = [float(t[0]) for t in allk]

    # PCG time per RK4 step (strongscaling.cpp-like protocol: rel 1e-12, RK4, dt = T/150) on the r=1 tank, weak-scaled
    rk = strong = None
    if not a.no_rk4:
        rk = rk4_measure(lpf, torch, a, local, stream, world, rank, dist if world > 1 else None, weak=True, refine=a.rk4_refine)
        # STRONG scaling (strongscaling.cpp:119-125 protocol): the same total mesh on every N -- big8 once refined
        # (2.25 M dofs) and big8 itself (299 520 dofs, the config the north star names), cut into N x-slabs
        strong = {}
        for name, r in (("big8_r1", 1), ("big8", 0)):
            m = rk4_measure(lpf, torch, a, local, stream, world, rank, dist if world > 1 else None, weak=False, refine=r, affine_leg=False)
            strong[name] = {k: m[k] for k in ("workload", "ms_per_rk4_step", "ms_per_cg_iteration", "cg_iterations_per_stage", "gpu_launches_per_step")}

    # burst vs sustained (profiles/r02_sustained.txt): everything above is timed within milliseconds of an idle period; after ~1 s of
    # back-to-back applies the board sits at its power limit and the same kernel runs at lower SM clocks.  Reported next to the burst
    # figure, never instead of it (the copy bandwidth the roofline divides by is a burst figure too, and flat under load).
    sustained = None
    if not a.no_sustained:
        n_heat = max(a.steps, int(0.5 / max(ms_max / a.steps * 1e-3, 1e-6)))      # the same count on every rank: the exchange is collective
        s2 = ClockSampler(local) if rank == 0 else None
        if s2 is not None:
            s2.start()
        ctx.time_apply(x, y, n_heat)                  # n_heat whole applies + n_heat kernels: about one second
        _, ms_ks, _ = ctx.time_apply(x, y, a.steps)
        if s2 is not None:
            s2.stop_flag = True
            s2.join(timeout=2)
        tks = torch.tensor([ms_ks], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tks, op=dist.ReduceOp.MAX)
        sustained = {"kernel_ms": float(tks[0]) / a.steps, "after": "%d back-to-back applies + %d element kernels (about 1 s)" % (n_heat, n_heat),
                     "clocks": s2.summary() if s2 is not None else None}

    # multi-GPU correctness carried by the bench line itself: a small tank solved on N ranks against the single-rank run
    parity = None
    if world > 1:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import check_multi_gpu
        res, fails = check_multi_gpu.run_checks(lpf, world, rank, local, stream, comm=a.comm, order=p, mesh_kind="tank", verbose=False)
        parity = {"mesh": "wave tank 32x2x8, perturbed, order %d, %d ranks vs 1 rank (same library; tests/ pin the 1-rank run against the oracle)" % (p, world),
                  "ok": not fails, "failed": fails, **{k.replace(" ", "_"): v for k, v in res.items()}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = ndof_global * a.steps / (ms_max * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    ab = algorithmic_bytes(sp.ne, sp.ndof, p)
    achieved = ab / (ms_kernel * 1e-3) / 1e9
    cfg = config_dict(a, world)
    assert cfg["hexes_per_gpu"] == sp.ne and cfg["dofs_global"] == ndof_global, (cfg, sp.ne, ndof_global)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
        "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "exchange": ((("halo-sum inside the apply kernel over NVLink peer memory (own kernels), overlapped with interior elements" if a.p2p_fuse in (-1, 2) else f"NVLink peer memory (own kernels), p2p_fuse={a.p2p_fuse}") if a.comm == "p2p" else "NCCL send/recv + all-reduce") if world > 1 else None),
        "apply_variant": a.variant,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(p, sp.ne), "peak_source": peak_src, "algorithmic_bytes_per_launch": ab,
                     "kernel_ms": ms_kernel, "kernel_ms_per_rank": {"min": min(kernel_ranks), "median": float(np.median(kernel_ranks)), "max": max(kernel_ranks)},
                     "kernel": "pa_apply_eo_kernel" if p >= 3 else "pa_apply_tma_kernel",
                     # the whole constrained apply behind `value` (zeroing of y + element kernel + essential rows): its own
                     # bytes are the kernel's plus one 8-byte write per dof for the zeroing the scatter-add needs
                     "whole_apply": {"bytes_per_step": ab + 8 * n, "ms": ms_max / a.steps,
                                     "achieved": (ab + 8 * n) / (ms_max / a.steps) / 1e6, "frac": (ab + 8 * n) / (ms_max / a.steps) / 1e6 / peak}},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 8 * n * world, "d2h_bytes_per_step": 8 * n * world,
                "steps": e2e_steps, "api": "lpf_apply_T_host (pinned host x -> device -> apply -> host y)"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    if sustained is not None:
        sustained["achieved"] = ab / (sustained["kernel_ms"] * 1e-3) / 1e9
        sustained["frac"] = sustained["achieved"] / peak
        line["roofline"]["sustained"] = sustained
    if aff is not None:
        line["affine_fastpath"] = aff
    if rk is not None:
        line["pcg_per_rk4_step"] = rk
        line["strong_scaling"] = strong
        line["e2e_rk4"] = {"value": rk["ms_per_rk4_step_host_api"], "unit": "ms per RK4 step", "higher_is_better": False,
                           "workload": rk["workload"], "api": "lpf_rk4_step_host (host state -> device, 4 Laplace solves + surface work, -> host)",
                           "h2d_bytes_per_step": rk["h2d_d2h_bytes_per_step"] // 2, "d2h_bytes_per_step": rk["h2d_d2h_bytes_per_step"] // 2}
    if parity is not None:
        line["parity"] = parity
    if not a.no_cpu:
        line["cpu_baseline"], _ = cpu_legs(a, max(3, min(a.steps, 10)), 2, with_rk4=not a.no_rk4)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def rk4_measure(lpf, torch, a, local, stream, world=1, rank=0, dist=None, weak=True, refine=1, affine_leg=True):
    """PCG time per RK4 step: strongscaling.cpp protocol (rel 1e-12, dt = T/150, 1 warm-up step).  weak: the tank is N times
    longer, one x-slab of 128 x 2 x 16 (refined) per GPU; strong: the SAME tank cut into N x-slabs.  Device-event time, max over ranks."""
    p = a.order
    nxf = world if weak else 1
    mesh = lpf.Mesh.wave_tank(128 * nxf, 2, 16, Lx=1.0 * nxf).refine(refine)
    sp = lpf.Space(mesh, p, nranks=world, rank=rank)
    ctx = lpf.Context(sp, device=local, stream=stream)
    if a.p2p_fuse >= 0:
        ctx.set_option("p2p_fuse", a.p2p_fuse)
    for kv in a.opt:
        ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    if world > 1 and a.comm == "p2p":
        ctx.p2p_connect(dist)
    elif world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(lpf.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        ctx.comm_init(bytes(idt.cpu().numpy().tobytes()))
    ctx.pa_setup()
    ctx.set_option("affine", 0)                 # general stored-q-data operator first (the graded configuration)
    ctx.jacobi_setup()
    w = lpf.wave_params()
    dt = w["T"] / 150
    ctx.rhs_setup(lpf.make_rhs_params(w, rel_tol=1e-12, max_iter=2000))
    xs, ys = sp.surf_xy[:, 0], sp.surf_xy[:, 1]
    ph = -w["k"] * (w["kx_dir"] * xs + w["ky_dir"] * ys)
    st = np.concatenate([0.5 * w["H"] * np.cos(ph), -0.5 * w["H"] * w["cwave"] / np.tanh(w["kh"]) * np.sin(ph)])
    sd = torch.from_numpy(st).cuda() if len(st) else torch.zeros(2, dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t = ctx.rk4_step(sd, 0.0, dt)               # warm-up step (ss.cpp:253)
    barrier()
    l0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nst = 2
    ev0.record()
    for _ in range(nst):
        t = ctx.rk4_step(sd, t, dt)
    ev1.record()
    barrier()
    tm = torch.tensor([ev0.elapsed_time(ev1) / nst], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm[0])
    infos = ctx.last_solve_info()
    its = [i.iterations for i in infos]
    launches_per_step = int((ctx.launches - l0) / nst)
    # the same step through the host-buffer entry point (H2D of the surface state + step + D2H), wall clock
    hs = torch.from_numpy(np.ascontiguousarray(sd.cpu().numpy()[:max(2, len(st))])).pin_memory()
    barrier()
    t0 = time.perf_counter()
    ctx.rk4_step_host(hs, t, dt)
    barrier()
    th = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(th, op=dist.ReduceOp.MAX)
    out = {"workload": f"wave-tank-big8 x{nxf} r={refine} order={p} on {world} GPU(s) ({sp.ne} hexes on rank 0, {int(sp.n_true_global)} dofs), RK4 dt=T/150, Jacobi-PCG rel 1e-12",
           "ms_per_rk4_step": ms, "cg_iterations_per_stage": its, "converged": [int(i.converged) for i in infos],
           "ms_per_cg_iteration": ms / max(1, sum(its)), "gpu_launches_per_step": launches_per_step,
           "ms_per_rk4_step_host_api": float(th[0]), "h2d_d2h_bytes_per_step": int(16 * len(st))}
    # extra: the same steps with the affine fast path
    ctx.set_option("affine", 1)
    if affine_leg and ctx.affine_active:
        sd2 = torch.from_numpy(st).cuda() if len(st) else torch.zeros(2, dtype=torch.float64, device="cuda")
        t2 = ctx.rk4_step(sd2, 0.0, dt)
        barrier()
        ev0.record()
        for _ in range(nst):
            t2 = ctx.rk4_step(sd2, t2, dt)
        ev1.record()
        barrier()
        tm2 = torch.tensor([ev0.elapsed_time(ev1) / nst], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tm2, op=dist.ReduceOp.MAX)
        its2 = [i.iterations for i in ctx.last_solve_info()]
        out["affine_fastpath"] = {"ms_per_rk4_step": float(tm2[0]), "cg_iterations_per_stage": its2,
                                  "ms_per_cg_iteration": float(tm2[0]) / max(1, sum(its2))}
    if world > 1:
        out["p2p_flag_timeouts"] = int(ctx.p2p_error()) if a.comm == "p2p" else None
    ctx.close()
    return out


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
