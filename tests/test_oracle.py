"""Pins the oracle (CPU, no GPU): two independent operators, analytic known answers, committed goldens.
Mirrors the reference's own verification drivers: laplace-parallel-pconv.cpp,
convergence-parallel-partial.cpp, Solvers/laplace_solver.cpp."""
import json
import os

import numpy as np
import pytest

from util import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
H = 1.0 / (2.0 * np.pi)


def test_basis_tables(orc):
    for p in range(1, 9):
        bs = orc.make_basis(p)
        assert abs(bs.qwts.sum() - 1.0) < 1e-15
        n = 2 * bs.Q - 1                       # Gauss rule with Q points is exact to degree 2Q-1
        assert abs((bs.qwts * bs.qpts ** n).sum() - 1.0 / (n + 1)) < 1e-15
        assert np.allclose(bs.B.sum(axis=1), 1.0, atol=1e-14)      # partition of unity
        assert np.allclose(bs.G.sum(axis=1), 0.0, atol=1e-11)
        assert np.allclose(bs.nodes, 1.0 - bs.nodes[::-1], atol=1e-16)
        # collocation derivative differentiates x^p exactly
        assert np.allclose(bs.Dhat @ bs.nodes ** p, p * bs.nodes ** (p - 1), atol=1e-11)


def test_scalar_known_answers(orc):
    """lambda-mode and period-mode parameters of the drivers."""
    wv = orc.Wave()
    assert abs(wv.c - 1.09045154194859) < 1e-14
    assert abs(wv.T - 0.917051296211702) < 1e-14
    assert abs(wv.omega - 6.85150910656268) < 1e-13
    assert abs(orc.dispersion_kh(9.81, 1.13392 / 3, H, 40) - 4.48434627651864) < 1e-13


def test_dof_counts(orc):
    m = orc.make_wave_tank(3, 1, 1, 1.0, 0.1, H, True)
    for p in range(1, 6):
        assert orc.build_h1_space(m, p).ndof == 3 * p * (p + 1) ** 2
    m2 = orc.uniform_refine(orc.uniform_refine(m))
    assert m2.ne == 192
    assert orc.build_h1_space(m2, 3).ndof == 6084        # BASELINE config 1
    assert orc.build_h1_space(m2, 4).ndof == 13872
    mf = orc.make_wave_tank(36, 1, 1, 12.0, 1.0, H, False)
    assert orc.build_h1_space(mf, 4).ndof == (36 * 4 + 1) * 25


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5])
def test_pa_equals_fa(orc, p):
    """PA (sum factorisation) and FA (dense element matrices -> CSR) are independent restatements of the
    same bilinear form and must agree to 1e-12 (north star: 'MFEM's own CPU PA and FA path')."""
    m = orc.perturb_mesh(orc.uniform_refine(orc.make_wave_tank(3, 1, 1, 1.0, 0.1, H, True)), 0.15)
    s = orc.build_h1_space(m, p)
    pa, fa = orc.PAOperator(s), orc.FAOperator(s)
    assert pa.detJ.min() > 0
    x = orc.hash_noise(s.ndof)
    assert rel_err(pa.mult(x), fa.mult(x)) < 1e-12
    assert rel_err(pa.diag(), fa.diag()) < 1e-12
    assert np.abs(pa.mult(np.ones(s.ndof))).max() < 1e-13          # constants are in the kernel
    y = orc.hash_noise(s.ndof, seed=7)
    assert abs(np.dot(y, pa.mult(x)) - np.dot(x, pa.mult(y))) < 1e-12 * np.abs(np.dot(y, pa.mult(x)))   # symmetry
    assert np.dot(x, pa.mult(x)) > 0


def test_c_oracle_matches_numpy(orc, corc):
    m = orc.perturb_mesh(orc.uniform_refine(orc.make_wave_tank(3, 1, 1, 1.0, 0.1, H, True)), 0.15)
    for p in (1, 2, 4, 6, 8):
        s = orc.build_h1_space(m, p)
        bs = s.basis
        A = orc.PAOperator(s)
        cop = corc.COperator(p, m.corners, s.gather, s.ndof, dict(B=bs.B, G=bs.G, Dhat=bs.Dhat, nodes=bs.nodes, qpts=bs.qpts, qwts=bs.qwts))
        assert rel_err(cop.qd, A.qd) < 1e-13
        x = orc.hash_noise(s.ndof)
        assert rel_err(cop.mult(x), A.mult(x)) < 1e-13
        assert rel_err(cop.diag(), A.diag()) < 1e-12
        phi = orc.hash_noise(s.ndof, seed=3)
        w, cnt = cop.deriv_z(phi)
        assert rel_err(w / cnt, orc.get_derivative_z(s, phi)) < 1e-12


def test_c_oracle_pcg_matches_numpy(orc, corc):
    m = orc.uniform_refine(orc.make_wave_tank(3, 1, 1, 1.0, 0.1, H, True))
    p = 3
    s = orc.build_h1_space(m, p)
    bs = s.basis
    A = orc.PAOperator(s)
    dinv = orc.jacobi_dinv(A, s.ess)
    wv = orc.Wave()
    x0 = np.zeros(s.ndof)
    x0[s.ess] = wv.phi(0.0, s.xyz[s.ess, 0], s.xyz[s.ess, 1], s.xyz[s.ess, 2], H)
    Ac, X, B = orc.form_linear_system(A, s.ess, x0, np.zeros(s.ndof))
    X, info = orc.pcg(Ac, dinv, B, X, 1e-12, 0.0, 1000)
    cop = corc.COperator(p, m.corners, s.gather, s.ndof, dict(B=bs.B, G=bs.G, Dhat=bs.Dhat, nodes=bs.nodes, qpts=bs.qpts, qwts=bs.qwts))
    Xc, ci = cop.pcg(s.ess, dinv, x0, 1e-12, 0.0, 1000)
    assert abs(ci["iterations"] - info.iters) <= 1 and ci["converged"]
    assert rel_err(Xc, X) < 1e-10
    assert ci["applies"] == ci["iterations"] + 2          # n + 2 applies per stage


def test_laplace_pconv_known_answers(orc):
    """laplace-parallel-pconv.cpp protocol on the 3-element periodic tank: spectral convergence towards the
    Airy potential; values frozen in known_answers.json (and equal to SURVEY App. E)."""
    ka = json.load(open(os.path.join(HERE, "golden", "known_answers.json")))
    wv = orc.Wave()
    m = orc.make_wave_tank(3, 1, 1, 1.0, 0.1, H, True)
    prev = None
    for row in ka["laplace_pconv"][:7]:
        p = row["p"]
        s = orc.build_h1_space(m, p)
        A = orc.PAOperator(s)
        dinv = orc.jacobi_dinv(A, s.ess)
        ex = wv.phi(0.0, s.xyz[:, 0], s.xyz[:, 1], s.xyz[:, 2], H)
        x0 = np.zeros(s.ndof); x0[s.ess] = ex[s.ess]
        Ac, X, B = orc.form_linear_system(A, s.ess, x0, np.zeros(s.ndof))
        X, info = orc.pcg(Ac, dinv, B, X, 1e-12, 0.0, 1000)
        err = np.abs(X - ex).max()
        assert abs(err - row["err_phi"]) < 1e-3 * row["err_phi"] + 1e-13
        assert abs(info.iters - row["iters"]) <= 1
        if prev is not None:
            assert err < 0.2 * prev            # at least ~one decade per order
        prev = err
    survey = {1: (7.3e-4, 1), 2: (4.4e-5, 8), 3: (3.4e-6, 19), 4: (3.3e-7, 44), 5: (2.7e-8, 65), 6: (1.7e-9, 86)}
    for row in ka["laplace_pconv"]:
        if row["p"] in survey:
            assert abs(row["err_phi"] - survey[row["p"]][0]) < 0.05 * survey[row["p"]][0]
            assert abs(row["iters"] - survey[row["p"]][1]) <= 1


def test_time_stepper_returns_after_one_period(orc):
    """convergence-parallel-partial.cpp idea, shortened: one period of the linear standing/progressive wave
    on the periodic tank brings eta back to eta(0) up to the discretisation error."""
    wv = orc.Wave()
    m = orc.make_wave_tank(3, 1, 1, 1.0, 0.1, H, True)
    s = orc.build_h1_space(m, 4)
    f = orc.RhsLinear(s, wv, rel_tol=1e-12, max_iter=2000)
    xs, ys = s.surf_xy[:, 0], s.surf_xy[:, 1]
    st = np.concatenate([wv.eta(0.0, xs, ys), wv.phi_fs(0.0, xs, ys)])
    nsteps = 40
    dt, t = wv.T / nsteps, 0.0
    for _ in range(nsteps):
        st, t = orc.rk4_step(f, st, t, dt)
    err = np.abs(st[:f.ns] - wv.eta(t, xs, ys)).max()
    assert err < 2e-5 * 1.0 and err < 0.01 * 0.5 * wv.H       # 5.2e-6 at dt = T/150


def test_golden_regression(orc):
    g = np.load(os.path.join(HERE, "golden", "tank_p3.npz"))
    p = int(g["order"])
    mesh = orc.HexMesh(np.zeros((g["gather"].shape[0], 8), dtype=np.int64), g["corners"], np.zeros((0, 4), int), np.zeros(0, int), 0)
    bs = orc.make_basis(p)
    s = orc.H1Space(p, bs, mesh, g["gather"].astype(np.int64), int(g["gather"].max()) + 1, None, g["ess"].astype(np.int64),
                    g["surf2vol"].astype(np.int64), g["surf_xy"])
    A = orc.PAOperator(s)
    assert rel_err(A.qd, g["qd"]) < 1e-14
    assert rel_err(A.mult(g["x"]), g["y"]) < 1e-13
    assert rel_err(A.diag(), g["diag"]) < 1e-13
    assert rel_err(orc.ConstrainedOperator(A, s.ess).mult(g["x"]), g["yc"]) < 1e-13
    fa = orc.FAOperator(s)
    assert rel_err(fa.mult(g["x"]), g["y"]) < 1e-12


def test_independent_oracle_pinned_on_config1(orc, corc):
    """The config-scale GPU tests (test_gpu_config_parity.py) check against IndependentOracle = oracle mesh + oracle
    numbering + oracle basis + C element kernels.  Pin it here, on the CPU: against the numpy PA operator, and against the
    known answers of config 1 (laplace_solver.cpp on wave-tank.mesh r=2: SURVEY App. E iteration counts 68 / 105)."""
    import math
    from util import IndependentOracle, rel_err
    H0 = 1.0 / (2.0 * np.pi)
    for p, its, err in ((3, 68, 6.3e-7), (4, 105, 1.7e-8)):
        m = orc.uniform_refine(orc.uniform_refine(orc.make_wave_tank(3, 1, 1, 1.0, 0.1, H0, True)))
        o = IndependentOracle(orc, corc, m, p)
        A = orc.PAOperator(o.sp)
        x = orc.hash_noise(o.n)
        assert rel_err(o.mult(x), A.mult(x)) < 1e-13 and rel_err(o.diag(), A.diag()) < 1e-13
        lo, hi = m.bounding_box()
        k = 2 * math.pi / (hi[0] - lo[0]); h = hi[2] - lo[2]; kh = k * h
        cw = math.sqrt((9.81 / k) * math.tanh(kh))
        phi = -0.5 * 0.005 * cw * np.cosh(k * (o.sp.xyz[:, 2] - hi[2] + h)) / math.sinh(kh) * np.sin(-k * o.sp.xyz[:, 0])
        X, info = o.solve(phi, 1e-12, 500)
        assert info.converged and info.iters == its
        assert rel_err(X, phi) < err
        w_ex = -0.5 * 0.005 * cw * k * np.sin(-k * o.sp.surf_xy[:, 0])
        assert rel_err(o.surface_dz(X), w_ex) < {3: 2e-4, 4: 3e-6}[p]
        # GetDerivative through the C oracle == the numpy restatement
        assert rel_err(o.surface_dz(X), orc.get_derivative_z(o.sp, X, orc.surface_elements(o.sp))[o.sp.surf2vol]) < 1e-12
