"""Config-scale GPU parity: every BASELINE.json config against an INDEPENDENT oracle.

The oracle side builds its own mesh (generator / reader / refinement), its own H1 numbering (vertex-weight
signatures) and its own basis tables (oracle/lpf_oracle.py) and does the element arithmetic in the C oracle
(oracle/pa_oracle.c); nothing of the product is reused.  The two numberings are related by geometry
(tests/util.py dof_map).  Tolerances are the north star's: operator 1e-12 relative (max norm), converged
potentials / surface state 1e-10 relative, CG iteration counts within +-1.

  C1  Solvers/laplace_solver.cpp:11-14          wave-tank.mesh r=2, order 3 (BASELINE) and 4 (as written)
  C2  Solvers/PF_linear_serial.cpp:269-272,320-344,396-416   wave-tank-finite.mesh r=1, order 5, serial constants, 2 RK4 steps
  C3  Convergence_and_Scaling/ws.cpp:91-94,123  wave-tank-big8.mesh order 4: operator, solve, one RK4 step
  C4  Solvers/cylinder-diffraction.cpp:136-141,225-226       mesh_cylinder_half order 4 (as written) and 6 (BASELINE)
  C5  ss.cpp / ws.cpp sweep                     big8 orders 1-8 (and the once-refined tank up to order 4): one apply each
"""
import math
import os

import numpy as np
import pytest

from util import IndependentOracle, dof_map, rel_err, surface_map

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
H0 = 1.0 / (2.0 * np.pi)
TOL_OP, TOL_SOL = 1e-12, 1e-10


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _refined(orc, m, r):
    for _ in range(r):
        m = orc.uniform_refine(m)
    return m


class Pair:
    """product context + independent oracle of the same discrete problem, and the maps between their numberings"""

    def __init__(self, lpf, orc, corc, torch, mesh, omesh, p, deterministic=0):
        self.lpf, self.orc, self.torch = lpf, orc, torch
        self.sp = lpf.Space(mesh, p)
        self.o = IndependentOracle(orc, corc, omesh, p)
        self.pm = dof_map(self.sp, self.o.sp)
        self.sm = surface_map(self.sp, self.o.sp, self.pm)
        assert set(self.pm[self.sp.ess].tolist()) == set(self.o.sp.ess.tolist()), "essential dof sets differ"
        self.ctx = lpf.Context(self.sp, device=0, stream=torch.cuda.current_stream().cuda_stream)
        self.ctx.pa_setup()
        self.ctx.set_option("affine", 0)            # the graded operator: stored q-data
        if deterministic:
            self.ctx.set_option("deterministic", 1)
        self.ctx.jacobi_setup()

    def to_product(self, vo):                        # oracle-numbered vector -> product numbering
        return vo[self.pm]

    def check_operator(self, seed=0x5EED1234):
        xo = self.orc.hash_noise(self.o.n, seed=seed)
        xd = _dev(self.torch, self.to_product(xo))
        yd = self.torch.empty_like(xd)
        self.ctx.apply_T(xd, yd)
        ref = self.o.constrained().mult(xo)
        err = rel_err(yd.cpu().numpy(), self.to_product(ref))
        assert err < TOL_OP, f"constrained operator: {err:.2e}"
        self.ctx.apply_L(xd, yd)
        err_l = rel_err(yd.cpu().numpy(), self.to_product(self.o.mult(xo)))
        assert err_l < TOL_OP, f"unconstrained operator: {err_l:.2e}"
        dg = self.torch.empty_like(xd)
        self.ctx.diag(dg)
        err_d = rel_err(dg.cpu().numpy(), self.to_product(self.o.diag()))
        assert err_d < TOL_OP, f"diagonal: {err_d:.2e}"
        return err

    def check_solve(self, phi_o, rel_tol, max_iter, its_tol=1, sol_tol=None):
        """Laplace solve on both sides.  Converged potentials (rel 1e-12 solves) must agree to 1e-10; a solve stopped at a
        looser tolerance (ws.cpp: 1e-8) is only defined up to that tolerance -- two runs whose counts differ by the
        allowed +-1 differ by one CG update -- so it is compared at 10 x rel_tol."""
        sol_tol = sol_tol if sol_tol is not None else max(TOL_SOL, 10.0 * rel_tol)
        pd = _dev(self.torch, self.to_product(phi_o))
        info = self.ctx.laplace_solve(pd, rel_tol=rel_tol, max_iter=max_iter)
        X, oi = self.o.solve(phi_o, rel_tol, max_iter)
        err = rel_err(pd.cpu().numpy(), self.to_product(X))
        assert info.converged == oi.converged
        assert abs(info.iterations - oi.iters) <= its_tol, (info.iterations, oi.iters)
        assert err < sol_tol, f"potential: {err:.2e}"
        return info, oi, X, pd

    def close(self):
        self.ctx.close()


def _airy_dirichlet(o, H, k, kh, cw, zmax, h):
    """phi_exact of laplace_solver.cpp:70-81 at the oracle's nodes (only the essential values are used)"""
    x, z = o.sp.xyz[:, 0], o.sp.xyz[:, 2]
    return -0.5 * H * cw * np.cosh(k * (z - zmax + h)) / math.sinh(kh) * np.sin(-k * x)


# ---- C1 -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("p", [3, 4])
def test_c1_laplace_solver_wave_tank_r2(lpf, orc, corc, cuda, p):
    """laplace_solver.cpp: wave-tank.mesh, 2 refinements (192 hexes; 6 084 dofs at order 3, 13 872 at order 4), Airy
    Dirichlet data of the mesh-mode wave (H = 0.005, k = 2 pi / Lx), Laplace solve, w = d(phi)/dz on the surface."""
    pr = Pair(lpf, orc, corc, cuda, lpf.Mesh.wave_tank(3, 1, 1).refine(2),
              _refined(orc, orc.make_wave_tank(3, 1, 1, 1.0, 0.1, H0, True), 2), p)
    assert pr.sp.ne == 192 and pr.sp.ndof == {3: 6084, 4: 13872}[p]
    pr.check_operator()
    lo, hi = pr.o.mesh.bounding_box()
    Hw, g = 0.005, 9.81
    k = 2 * math.pi / (hi[0] - lo[0]); h = hi[2] - lo[2]; kh = k * h
    cw = math.sqrt((g / k) * math.tanh(kh))
    phi_o = _airy_dirichlet(pr.o, Hw, k, kh, cw, hi[2], h)
    info, oi, X, pd = pr.check_solve(phi_o, 1e-12, 500)
    assert abs(oi.iters - {3: 68, 4: 105}[p]) <= 1
    # the discrete solution is the Airy potential to discretisation accuracy (known answer of the driver)
    assert rel_err(X, phi_o) < {3: 1e-6, 4: 3e-8}[p]      # 6.2e-7 / 1.6e-8 (tests/test_oracle.py pins the oracle's values)
    wt = cuda.empty(pr.sp.nsurf, dtype=cuda.float64, device="cuda")
    pr.ctx.surface_dz(pd, wt)
    assert rel_err(wt.cpu().numpy(), pr.o.surface_dz(X)[pr.sm]) < TOL_SOL
    pr.close()


# ---- C2 -----------------------------------------------------------------------------------------
def test_c2_pf_linear_serial_finite_tank_p5(lpf, orc, corc, cuda):
    """PF_linear_serial.cpp as written: wave-tank-finite.mesh (36x1x1 on 12 x 1 x h), 1 refinement, order 5 (43 681 dofs),
    H = 0.05, T = 1.13392/3 with kh from the 40-step dispersion iteration, N_g = N_s = 2, tau = dt = 8T/800: operator,
    then two RK4 steps with the relaxation zones against the oracle."""
    torch = cuda
    pr = Pair(lpf, orc, corc, torch, lpf.Mesh.wave_tank(36, 1, 1, 12.0, 1.0, H0, False).refine(1),
              _refined(orc, orc.make_wave_tank(36, 1, 1, 12.0, 1.0, H0, False), 1), 5)
    assert pr.sp.ne == 288 and pr.sp.ndof == 43681
    pr.check_operator()
    lo, hi = pr.o.mesh.bounding_box()
    Hw, g, T = 0.05, 9.81, 1.13392 / 3
    h = hi[2] - lo[2]
    kh = orc.dispersion_kh(g, T, h, 40)
    assert abs(kh - 4.48434627651864) < 1e-12                       # SURVEY 8c known answer
    k = kh / h; omega = 2 * math.pi / T; cw = omega / k; lam = 2 * math.pi / k
    dt = 8.0 * T / 800
    wv = orc.Wave(H=Hw, g=g, lam=lam, kh=kh)
    wv.c, wv.T, wv.omega = cw, T, omega                              # period mode: omega is the input, c = omega / k
    osp = pr.o.sp
    xs_o = osp.surf_xy[:, 0]
    rx = orc.Relax(orc.relax_cgen(xs_o, lo[0], lo[0] + 2.0 * lam), orc.relax_cabs(xs_o, hi[0] - 2.0 * lam, hi[0]), tau=dt)
    f = orc.RhsLinear(osp, wv, rel_tol=1e-12, max_iter=1000, relax=rx, operator=pr.o)
    st_o = np.concatenate([wv.eta(0.0, xs_o, osp.surf_xy[:, 1]), wv.phi_fs(0.0, xs_o, osp.surf_xy[:, 1])])
    w = dict(H=Hw, g=g, lam=lam, k=k, kh=kh, cwave=cw, T=T, omega=omega, kx_dir=1.0, ky_dir=0.0)
    xs_p = pr.sp.surf_xy[:, 0]
    pr.ctx.rhs_setup(lpf.make_rhs_params(w, tau=dt, use_relaxation=True, rel_tol=1e-12, max_iter=1000),
                     orc.relax_cgen(xs_p, lo[0], lo[0] + 2.0 * lam), orc.relax_cabs(xs_p, hi[0] - 2.0 * lam, hi[0]))
    ns = pr.sp.nsurf
    both = np.concatenate([pr.sm, ns + pr.sm])
    sd = _dev(torch, st_o[both])
    t_o = t_p = 0.0
    for step in range(2):
        st_o, t_o = orc.rk4_step(f, st_o, t_o, dt)
        t_p = pr.ctx.rk4_step(sd, t_p, dt)
        its_p = [i.iterations for i in pr.ctx.last_solve_info()]
        its_o = f.iters[-4:]
        assert all(abs(a - b) <= 1 for a, b in zip(its_p, its_o)), (step, its_p, its_o)
        assert rel_err(sd.cpu().numpy()[:ns], st_o[:ns][pr.sm]) < TOL_SOL
        assert rel_err(sd.cpu().numpy()[ns:], st_o[ns:][pr.sm]) < TOL_SOL
    pr.close()


# ---- C3 -----------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def big8(lpf, orc, corc, cuda):
    pr = Pair(lpf, orc, corc, cuda, lpf.Mesh.wave_tank(128, 2, 16), orc.make_wave_tank(128, 2, 16, 1.0, 0.1, H0, True), 4,
              deterministic=1)
    yield pr
    pr.close()


def test_c3_big8_p4_operator_and_solve(lpf, orc, big8, cuda):
    """wave-tank-big8.mesh, order 4: 4 096 hexes, 299 520 dofs, 4 608 surface dofs (ws.cpp:123).  Operator 1e-12; the
    Laplace solve of the t = 0 Airy surface potential at the tolerance ws.cpp uses (rel 1e-8, <= 300 its) and at the
    1e-12 of the solver drivers: potential 1e-10, iterations +-1 (196 / 247)."""
    pr = big8
    assert pr.sp.ne == 4096 and pr.sp.ndof == 299520 and pr.sp.nsurf == 4608
    pr.check_operator()
    pr.ctx.set_option("deterministic", 0)
    pr.check_operator(seed=77)                      # the default (atomic scatter) path on the same mesh
    pr.ctx.set_option("deterministic", 1)
    wv = orc.Wave()
    osp = pr.o.sp
    phi_o = np.zeros(pr.o.n)
    phi_o[osp.surf2vol] = wv.phi_fs(0.0, osp.surf_xy[:, 0], osp.surf_xy[:, 1])
    i8, o8, _, _ = pr.check_solve(phi_o, 1e-8, 300)
    assert abs(o8.iters - 196) <= 2
    i12, o12, _, _ = pr.check_solve(phi_o, 1e-12, 1000)
    assert abs(o12.iters - 247) <= 3
    # deterministic mode: the same solve again gives the same count and the same bits
    pd1, pd2 = _dev(cuda, phi_o[pr.pm]), _dev(cuda, phi_o[pr.pm])
    a = pr.ctx.laplace_solve(pd1, rel_tol=1e-12, max_iter=1000)
    b = pr.ctx.laplace_solve(pd2, rel_tol=1e-12, max_iter=1000)
    assert a.iterations == b.iterations == i12.iterations and cuda.equal(pd1, pd2)
    # default (atomic) mode: still within +-1 of the oracle at the ws.cpp tolerance
    pr.ctx.set_option("deterministic", 0)
    pr.ctx.jacobi_setup()
    pr.check_solve(phi_o, 1e-8, 300)
    pr.ctx.set_option("deterministic", 1)
    pr.ctx.jacobi_setup()


@pytest.mark.parametrize("rel_tol,max_iter,tol_eta,tol_phi", [(1e-13, 1000, TOL_SOL, TOL_SOL), (1e-8, 300, 1e-5, 1e-6)])
def test_c3_big8_p4_rk4_step(lpf, orc, big8, cuda, rel_tol, max_iter, tol_eta, tol_phi):
    """one RK4 step of the ws.cpp physics (no relaxation zones, dt = T/10) on big8, order 4: with solves converged to the
    round-off floor (rel 1e-13; at 1e-12 the two last iterates still differ by 3e-10 in eta after the d/dz and the large
    step dt = T/10) the state agrees to 1e-10; with ws.cpp's own solver settings (rel 1e-8, <= 300 its) the two sides stop at
    iterates that differ by the solver tolerance, amplified in eta by d/dz -- compared at that level."""
    pr = big8
    wv = orc.Wave()
    osp = pr.o.sp
    dt = wv.T / 10
    f = orc.RhsLinear(osp, wv, rel_tol=rel_tol, max_iter=max_iter, operator=pr.o)
    st_o = np.concatenate([wv.eta(0.0, osp.surf_xy[:, 0], osp.surf_xy[:, 1]), wv.phi_fs(0.0, osp.surf_xy[:, 0], osp.surf_xy[:, 1])])
    pr.ctx.rhs_setup(lpf.make_rhs_params(lpf.wave_params(), rel_tol=rel_tol, max_iter=max_iter))
    ns = pr.sp.nsurf
    sd = _dev(cuda, st_o[np.concatenate([pr.sm, ns + pr.sm])])
    st_o, _ = orc.rk4_step(f, st_o, 0.0, dt)
    pr.ctx.rk4_step(sd, 0.0, dt)
    its_p = [i.iterations for i in pr.ctx.last_solve_info()]
    assert all(abs(a - b) <= 1 for a, b in zip(its_p, f.iters)), (its_p, f.iters)
    assert rel_err(sd.cpu().numpy()[:ns], st_o[:ns][pr.sm]) < tol_eta
    assert rel_err(sd.cpu().numpy()[ns:], st_o[ns:][pr.sm]) < tol_phi


# ---- C4 -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("p", [4, 6])
def test_c4_cylinder_half_mesh(lpf, orc, corc, cuda, p):
    """cylinder-diffraction.cpp: mesh_cylinder_half (3 192 hexes, unstructured in x-y), order 4 as written (221 085 dofs)
    and order 6 as BASELINE names it (727 075 dofs): operator + the first Laplace solve of the driver (rel 1e-12, <= 2000)."""
    path = os.path.join(HERE, "meshes", "cylinder_half.mesh")
    pr = Pair(lpf, orc, corc, cuda, lpf.Mesh.read(path), orc.read_mfem_mesh(path), p, deterministic=1)
    assert pr.sp.ne == 3192 and pr.sp.ndof == {4: 221085, 6: 727075}[p]
    assert not pr.ctx.affine_active
    pr.check_operator()
    pr.ctx.set_option("deterministic", 0)
    pr.check_operator(seed=5)
    pr.ctx.set_option("deterministic", 1)
    wv = orc.Wave()
    osp = pr.o.sp
    phi_o = np.zeros(pr.o.n)
    phi_o[osp.surf2vol] = wv.phi_fs(0.0, osp.surf_xy[:, 0], osp.surf_xy[:, 1])
    info, oi, X, pd = pr.check_solve(phi_o, 1e-12, 2000)
    wt = cuda.empty(pr.sp.nsurf, dtype=cuda.float64, device="cuda")
    pr.ctx.surface_dz(pd, wt)
    assert rel_err(wt.cpu().numpy(), pr.o.surface_dz(X)[pr.sm]) < TOL_SOL
    pr.close()


# ---- C5 -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("p,refine", [(p, 0) for p in range(1, 9)] + [(2, 1), (3, 1), (4, 1)])
def test_c5_order_sweep_big8(lpf, orc, corc, cuda, p, refine):
    """ss.cpp / ws.cpp sweep: one constrained apply per order on big8 (4 096 hexes; 2.2 M dofs at order 8) and on the once
    refined tank (32 768 hexes; 2.2 M dofs at order 4) against the independent oracle."""
    pr = Pair(lpf, orc, corc, cuda, lpf.Mesh.wave_tank(128, 2, 16).refine(refine),
              _refined(orc, orc.make_wave_tank(128, 2, 16, 1.0, 0.1, H0, True), refine), p)
    assert pr.sp.ndof == (128 * p * 2 ** refine) * (2 * p * 2 ** refine + 1) * (16 * p * 2 ** refine + 1)
    pr.check_operator()
    pr.close()
