"""Compares a dump of tools/mfem_crosscheck.cpp (real MFEM, run elsewhere) with this repository on a B200 and with the
oracle: the step that would turn "parity unpinned" into a pinned statement.

    python tools/mfem_crosscheck_compare.py dump.txt --mesh tank:128,2,16 [--refine 0] --order 4
    python tools/mfem_crosscheck_compare.py dump.txt --mesh tests/meshes/cylinder_half.mesh --order 4

Nodes are matched by coordinates (x modulo the tank length), so no assumption is made about MFEM's numbering."""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]


def read_dump(path):
    sec = {}
    with open(path) as f:
        lines = f.read().split("\n")
    i = 0
    while i < len(lines):
        t = lines[i].split()
        if not t:
            i += 1
            continue
        if t[0] in ("OPERATOR", "SOLVE"):
            n = int(t[1])
            sec[t[0]] = np.array([[float(v) for v in l.split()] for l in lines[i + 1:i + 1 + n]])
            i += 1 + n
        elif t[0] == "INFO":
            sec["INFO"] = [float(v) for v in t[1:]]
            i += 1
        else:
            i += 1
    return sec


def match(xyz_a, xyz_b, Lx, periodic):
    """index array m with xyz_a[i] ~ xyz_b[m[i]]"""
    def key(c):
        c = c.copy()
        if periodic:
            c[:, 0] = np.mod(c[:, 0] + 1e-9 * Lx, Lx)
        return np.round(c / (1e-7 * Lx)).astype(np.int64)
    ka, kb = key(xyz_a), key(xyz_b)
    oa, ob = np.lexsort(ka.T[::-1]), np.lexsort(kb.T[::-1])
    assert (ka[oa] == kb[ob]).all(), "node sets differ"
    m = np.empty(len(oa), dtype=np.int64)
    m[oa] = ob
    return m


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("dump")
    ap.add_argument("--mesh", required=True)
    ap.add_argument("--refine", type=int, default=0)
    ap.add_argument("--order", type=int, required=True)
    a = ap.parse_args()
    lpf = importlib.import_module("master-thesis-lpf-in-mfem_b200")
    d = read_dump(a.dump)
    if a.mesh.startswith("tank:"):
        nx, ny, nz = (int(t) for t in a.mesh[5:].split(","))
        mesh, periodic = lpf.Mesh.wave_tank(nx, ny, nz).refine(a.refine), True
    else:
        mesh, periodic = lpf.Mesh.read(a.mesh).refine(a.refine), False
    sp = lpf.Space(mesh, a.order)
    Lx = d["INFO"][3]
    xyz = sp.node_coordinates()
    op = d["OPERATOR"]
    m = match(xyz, op[:, :3], Lx, periodic)                      # product dof i <-> MFEM dof m[i]
    torch.cuda.set_stream(torch.cuda.Stream())
    ctx = lpf.Context(sp, device=0, stream=torch.cuda.current_stream().cuda_stream)
    ctx.pa_setup(); ctx.set_option("affine", 0); ctx.set_option("deterministic", 1)
    rel = lambda x, y: float(np.abs(x - y).max() / np.abs(y).max())
    u = torch.from_numpy(op[m, 3].copy()).cuda()
    y = torch.empty_like(u)
    ctx.apply_L(u, y)
    print("operator vs MFEM PA  %.3e   vs MFEM FA  %.3e   (MFEM PA vs FA %.3e)" % (rel(y.cpu().numpy(), op[m, 4]), rel(y.cpu().numpy(), op[m, 5]), rel(op[:, 4], op[:, 5])))
    dg = torch.empty_like(u)
    ctx.diag(dg)
    print("diagonal vs MFEM     %.3e" % rel(dg.cpu().numpy(), op[m, 6]))
    ctx.jacobi_setup()
    so = d["SOLVE"]
    phi0 = np.zeros(sp.ndof)
    phi0[sp.ess] = so[m, 3][sp.ess]
    pd = torch.from_numpy(phi0).cuda()
    info = ctx.laplace_solve(pd, rel_tol=1e-12, max_iter=5000)
    print("potential vs MFEM    %.3e   CG iterations %d (MFEM %d)" % (rel(pd.cpu().numpy(), so[m, 3]), info.iterations, int(d["INFO"][0])))
    wt = torch.empty(sp.nsurf, dtype=torch.float64, device="cuda")
    ctx.surface_dz(pd, wt)
    print("w~ = GetDerivative(1,2) on the surface vs MFEM  %.3e" % rel(wt.cpu().numpy(), so[m, 4][sp.surf2vol]))


if __name__ == "__main__":
    main()
