"""Generates the committed fixtures under tests/golden/ and tests/meshes/ from the ORACLE (self-generated:
the reference ships no golden vectors) and from the reference's mesh data.

Run in the build container (needs /root/reference for the cylinder mesh):
    python tests/golden/make_golden.py
Outputs:
    tests/meshes/cylinder_half.mesh   the reference's Gmsh mesh (Meshes/mesh_cylinder_half.msh) re-written
                                      by oracle.write_mfem_mesh as MFEM mesh v1.0 (config 4 input)
    tests/golden/tank_p3.npz          inputs (corners, gather, ess, surf2vol, x) and oracle outputs
                                      (A x, diag, PCG iterations/solution, w~, two RK4 steps) on a perturbed
                                      periodic tank, order 3
    tests/golden/known_answers.json   scalar known answers + analytic p-convergence table
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lpf_oracle as orc  # noqa: E402


def main():
    os.makedirs(os.path.join(ROOT, "tests", "meshes"), exist_ok=True)
    ref = "/root/reference/Meshes/mesh_cylinder_half.msh"
    if os.path.exists(ref):
        m = orc.read_gmsh22(ref)
        orc.write_mfem_mesh(m, os.path.join(ROOT, "tests", "meshes", "cylinder_half.mesh"), with_nodes=False)

    # ---- golden vectors on a small perturbed periodic tank ----
    mesh = orc.perturb_mesh(orc.uniform_refine(orc.make_wave_tank(3, 1, 1, 1.0, 0.1, 1 / (2 * np.pi), True)), 0.15)
    p = 3
    sp = orc.build_h1_space(mesh, p)
    A = orc.PAOperator(sp)
    x = orc.hash_noise(sp.ndof)
    y = A.mult(x)
    yc = orc.ConstrainedOperator(A, sp.ess).mult(x)
    diag = A.diag()
    wv = orc.Wave()
    h = 1 / (2 * np.pi)
    f = orc.RhsLinear(sp, wv, rel_tol=1e-12, max_iter=1000)
    xs, ys = sp.surf_xy[:, 0], sp.surf_xy[:, 1]
    state0 = np.concatenate([wv.eta(0.0, xs, ys), wv.phi_fs(0.0, xs, ys)])
    phi = f.solve_laplace(state0[f.ns:]).copy()
    its0 = f.last.iters
    wt = orc.get_derivative_z(sp, phi, f.selems)[sp.surf2vol]
    dt = wv.T / 40
    st, t = state0.copy(), 0.0
    for _ in range(2):
        st, t = orc.rk4_step(f, st, t, dt)
    # relaxation-zone variant
    lo, hi = mesh.bounding_box()
    cgen = orc.relax_cgen(xs, lo[0], lo[0] + 0.4)
    cabs = orc.relax_cabs(xs, hi[0] - 0.3, hi[0])
    fr = orc.RhsLinear(sp, wv, rel_tol=1e-12, max_iter=1000, relax=orc.Relax(cgen, cabs, tau=dt))
    str_, tr = state0.copy(), 0.0
    for _ in range(2):
        str_, tr = orc.rk4_step(fr, str_, tr, dt)
    np.savez_compressed(
        os.path.join(HERE, "tank_p3.npz"), order=p, corners=mesh.corners, gather=sp.gather.astype(np.int32),
        ess=sp.ess.astype(np.int32), surf2vol=sp.surf2vol.astype(np.int32), surf_xy=sp.surf_xy[:, :2], x=x, y=y, yc=yc,
        diag=diag, state0=state0, phi=phi, its0=its0, wt=wt, dt=dt, state2=st, its_rk=np.array(f.iters[1:]),
        cgen=cgen, cabs=cabs, state2_relax=str_, qd=A.qd)

    # ---- scalar known answers + analytic p-convergence (laplace-parallel-pconv.cpp protocol) ----
    ka = dict(c=wv.c, T=wv.T, omega=wv.omega, h=h, kh_period_mode=orc.dispersion_kh(9.81, 1.13392 / 3, h, 40))
    m3 = orc.make_wave_tank(3, 1, 1, 1.0, 0.1, h, True)
    table = []
    for pp in range(1, 9):
        s = orc.build_h1_space(m3, pp)
        Ap = orc.PAOperator(s)
        dinv = orc.jacobi_dinv(Ap, s.ess)
        ex = wv.phi(0.0, s.xyz[:, 0], s.xyz[:, 1], s.xyz[:, 2], h)
        x0 = np.zeros(s.ndof); x0[s.ess] = ex[s.ess]
        Ac, X, B = orc.form_linear_system(Ap, s.ess, x0, np.zeros(s.ndof))
        X, info = orc.pcg(Ac, dinv, B, X, 1e-12, 0.0, 1000)
        w = orc.get_derivative_z(s, X)[s.surf2vol]
        wex = wv.w_surface(0.0, s.surf_xy[:, 0], s.surf_xy[:, 1])
        table.append(dict(p=pp, ndof=int(s.ndof), err_phi=float(np.abs(X - ex).max()), iters=int(info.iters),
                          err_w=float(np.abs(w - wex).max())))
    ka["laplace_pconv"] = table
    with open(os.path.join(HERE, "known_answers.json"), "w") as fjson:
        json.dump(ka, fjson, indent=1)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
