"""Multi-GPU parity check (one process per GPU, NCCL halo-sum): run under torchrun on N GPUs of one node.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_multi_gpu.py [--order 4] [--mesh tank|cylinder]

Every rank builds its x-slab (RCB) partition, runs the constrained operator, a Laplace solve and three RK4
steps through the C-ABI, and the results are compared with the single-GPU run of the same library on rank 0's
device (which tests/ pins against the oracle): operator 1e-12, potentials / elevations 1e-10, CG iterations
within +-1; copies of shared dofs must be bit-identical across ranks.
"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def run_checks(lpf, world, rank, local, stream, comm="p2p", order=4, mesh_kind="tank", verbose=True, host_apply=True, options=None):
    """N-rank results against the single-rank run of the same library on this rank's GPU (which tests/ pins against the
    oracle).  Returns {check name: value} and the list of failed checks (identical on every rank).  `options`: extra
    (name, value) pairs for lpf_set_option on the multi-rank context (e.g. p2p_fuse)."""
    stream_ptr = stream
    if mesh_kind == "tank":
        mesh = lpf.Mesh.wave_tank(32, 2, 8).perturb(0.1)
    else:
        mesh = lpf.Mesh.read(os.path.join(ROOT, "tests", "meshes", "cylinder_half.mesh"))
    p = order
    sp = lpf.Space(mesh, p, nranks=world, rank=rank)
    ctx = lpf.Context(sp, device=local, stream=stream_ptr)
    for k, v in (options or []):
        ctx.set_option(k, v)

    def connect(c):
        if comm == "p2p":
            c.p2p_connect(dist)                       # our own NVLink peer-memory exchange, no NCCL inside the solver
        else:
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt = torch.frombuffer(bytearray(lpf.comm_unique_id()), dtype=torch.uint8).cuda()
            dist.broadcast(idt, 0)
            c.comm_init(idt.cpu().numpy().tobytes())

    connect(ctx)
    ctx.pa_setup()
    ctx.jacobi_setup()

    # serial twin on this rank's GPU (every rank builds it: cheap at this size, and no gather of inputs needed)
    ssp = lpf.Space(mesh, p)
    sctx = lpf.Context(ssp, device=local, stream=stream_ptr)
    sctx.pa_setup()
    sctx.jacobi_setup()
    l2g = torch.from_numpy(sp.l2g.astype(np.int64)).cuda()
    fails, results = [], {}

    def check(name, val, tol):
        t = torch.tensor([float(val), 0.0 if val < tol else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst, bad = float(t[0]), float(t[1]) != 0
        results[name] = worst
        if rank == 0 and verbose:
            print(f"  {name:34s} {worst:.3e}  (tol {tol:g})  {'FAIL' if bad else 'ok'}", flush=True)
        if bad:
            fails.append(name)

    gen = torch.Generator(device="cuda").manual_seed(7)
    xg = torch.rand(ssp.ndof, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    yg = torch.empty_like(xg)
    sctx.apply_T(xg, yg)
    xl = xg[l2g].contiguous()
    yl = torch.empty_like(xl)
    ctx.apply_T(xl, yl)
    check("constrained apply (halo-sum)", rel(yl.cpu().numpy(), yg[l2g].cpu().numpy()), 1e-12)
    dg, dl = torch.empty_like(xg), torch.empty_like(xl)
    sctx.diag(dg); ctx.diag(dl)
    check("diagonal", rel(dl.cpu().numpy(), dg[l2g].cpu().numpy()), 1e-12)
    # shared copies bit-identical: sum over ranks of (value * owned) scattered to global == every copy
    glob = torch.zeros(ssp.ndof, dtype=torch.float64, device="cuda")
    glob[l2g] = yl * torch.from_numpy(sp.owned.astype(np.float64)).cuda()
    dist.all_reduce(glob)
    check("shared copies bit-identical", float((glob[l2g] - yl).abs().max()), 1e-300)

    # T-vectors <-> L-vectors (lpf_prolong / lpf_restrict, what the MFEM adapter uses): the true dofs of all ranks partition the
    # global dofs; P of the owned values reproduces every copy; R P = identity
    nt = ctx.ntrue(0)
    cnt = torch.tensor([float(nt)], dtype=torch.float64, device="cuda")
    dist.all_reduce(cnt)
    check("sum of true dofs - global dofs", abs(float(cnt[0]) - ssp.ndof), 0.5)
    owned_idx = torch.from_numpy(np.nonzero(sp.owned)[0].astype(np.int64)).cuda()
    xT = xl[owned_idx].contiguous() if nt else torch.zeros(1, dtype=torch.float64, device="cuda")
    xL2 = torch.full_like(xl, float("nan"))
    ctx.prolong(xT, xL2)
    check("prolong: P x_T == consistent L-vector", float((xL2 - xl).abs().max()), 1e-300)
    xT2 = torch.full_like(xT, float("nan"))
    ctx.restrict(xL2, xT2)
    check("restrict: R P x_T == x_T", float((xT2 - xT).abs().max()) if nt else 0.0, 1e-300)

    # Laplace solve from Airy Dirichlet data
    w = lpf.wave_params()
    xyz = ssp.node_coordinates()
    lo, hi = mesh.bounding_box()
    ex = -0.5 * w["H"] * w["cwave"] * np.cosh(w["k"] * (xyz[:, 2] - lo[2])) / np.sinh(w["kh"]) * np.sin(-w["k"] * xyz[:, 0])
    phi0 = np.zeros(ssp.ndof); phi0[ssp.ess] = ex[ssp.ess]
    pg = torch.from_numpy(phi0).cuda()
    pl = pg[l2g].contiguous()
    si = sctx.laplace_solve(pg, rel_tol=1e-12, max_iter=2000)
    mi = ctx.laplace_solve(pl, rel_tol=1e-12, max_iter=2000)
    check("laplace solve potential", rel(pl.cpu().numpy(), pg[l2g].cpu().numpy()), 1e-10)
    check("CG iterations |multi - single|", abs(mi.iterations - si.iterations), 1.5)
    results["iterations_single"], results["iterations_multi"] = si.iterations, mi.iterations
    if rank == 0 and verbose:
        print(f"    iterations: single {si.iterations}, {world} GPUs {mi.iterations}", flush=True)

    # three RK4 steps of the ss.cpp RHS
    prm = lpf.make_rhs_params(w, rel_tol=1e-12, max_iter=2000)
    sctx.rhs_setup(prm); ctx.rhs_setup(prm)
    ph = -w["k"] * ssp.surf_xy[:, 0]
    st = np.concatenate([0.5 * w["H"] * np.cos(ph), -0.5 * w["H"] * w["cwave"] / np.tanh(w["kh"]) * np.sin(ph)])
    sg = torch.from_numpy(st).cuda()
    sg_idx = torch.from_numpy(sp.surf_g.astype(np.int64)).cuda()
    ns_g, ns_l = ssp.nsurf, sp.nsurf
    sl = torch.cat([sg[:ns_g][sg_idx], sg[ns_g:][sg_idx]]).contiguous() if ns_l else torch.zeros(2, dtype=torch.float64, device="cuda")
    dt = w["T"] / 60
    tg = tl = 0.0
    for _ in range(3):
        tg = sctx.rk4_step(sg, tg, dt)
        tl = ctx.rk4_step(sl, tl, dt)
    if ns_l:
        check("eta after 3 RK4 steps", rel(sl[:ns_l].cpu().numpy(), sg[:ns_g][sg_idx].cpu().numpy()), 1e-10)
        check("phi_fs after 3 RK4 steps", rel(sl[ns_l:].cpu().numpy(), sg[ns_g:][sg_idx].cpu().numpy()), 1e-10)
    else:
        check("eta after 3 RK4 steps", 0.0, 1e-10); check("phi_fs after 3 RK4 steps", 0.0, 1e-10)
    its_s = [i.iterations for i in sctx.last_solve_info()]
    its_m = [i.iterations for i in ctx.last_solve_info()]
    check("RK4 stage CG iterations", max(abs(x - y) for x, y in zip(its_s, its_m)), 1.5)
    # host-buffer entry point on a mesh large enough for its pipelined path (H2D ranges / element chunks / D2H ranges,
    # ranges holding shared dofs leave after the halo-sum): must equal the device-resident apply of the same rank
    if mesh_kind == "tank" and host_apply:
        big = lpf.Mesh.wave_tank(64 * world, 2, 16, Lx=1.0 * world).refine(1)
        bsp = lpf.Space(big, p, nranks=world, rank=rank)
        bctx = lpf.Context(bsp, device=local, stream=stream_ptr)
        for k, v in (options or []):
            bctx.set_option(k, v)
        connect(bctx)
        bctx.pa_setup()
        # x is an L-vector: copies of shared dofs must be consistent across ranks, so key the values by global id
        key = torch.from_numpy(bsp.l2g.astype(np.float64)).cuda()
        xb = torch.sin(key * 0.001) * 0.5
        yb = torch.empty_like(xb)
        bctx.apply_T(xb, yb)
        xh = xb.cpu().pin_memory()
        yh = torch.full((bsp.ndof,), float("nan"), dtype=torch.float64).pin_memory()
        bctx.apply_T_host(xh, yh)
        check(f"host apply, pipelined ({bsp.ndof} dofs/rank)", rel(yh.numpy(), yb.cpu().numpy()), 1e-12)
        bctx.close()
    if rank == 0 and verbose:
        print(f"    stage iterations: single {its_s}, {world} GPUs {its_m}")
    if comm == "p2p":
        check("p2p flag waits timed out", float(ctx.p2p_error()), 0.5)
    ctx.close(); sctx.close()
    return results, fails


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--mesh", default="tank")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--p2p-fuse", type=int, default=-1, help="option p2p_fuse of the multi-rank context (-1: library default)")
    a = ap.parse_args()
    lpf = importlib.import_module("master-thesis-lpf-in-mfem_b200")
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    torch.cuda.set_stream(torch.cuda.Stream())
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream().cuda_stream
    if rank == 0:
        print(f"communication: {a.comm}, p2p_fuse {a.p2p_fuse}", flush=True)
    opts = [("p2p_fuse", a.p2p_fuse)] if a.p2p_fuse >= 0 else None
    _, fails = run_checks(lpf, world, rank, local, stream, a.comm, a.order, a.mesh, options=opts)
    if rank == 0:
        print("MULTI-GPU PARITY: " + ("OK" if not fails else "FAILED " + str(fails)), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
