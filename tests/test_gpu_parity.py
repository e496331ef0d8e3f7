"""GPU parity tests: the CUDA path, called through the C-ABI, against the oracle on identical inputs.

Tolerances (north star): operator action 1e-12 relative (max norm), converged potentials / surface
elevations 1e-10 relative, CG iteration counts within +-1."""
import os

import numpy as np
import pytest

from util import oracle_space_from, rel_err

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
H = 1.0 / (2.0 * np.pi)
TOL_OP = 1e-12
TOL_SOL = 1e-10


def _ctx(lpf, torch, sp):
    ctx = lpf.Context(sp, device=0, stream=torch.cuda.current_stream().cuda_stream)
    ctx.pa_setup()
    return ctx


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


@pytest.fixture(scope="module")
def tank(lpf):
    """wave-tank.mesh topology (3 periodic hexes), refined once and perturbed: 24 non-affine hexes."""
    return lpf.Mesh.wave_tank(3, 1, 1).refine(1).perturb(0.15)


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
def test_qdata_apply_diag_all_orders(lpf, orc, cuda, tank, p):
    torch = cuda
    sp = lpf.Space(tank, p)
    osp = oracle_space_from(orc, sp)
    A = orc.PAOperator(osp)
    fa = orc.FAOperator(osp) if p <= 5 else None
    ctx = _ctx(lpf, torch, sp)
    Q3 = (p + 2) ** 3
    qd = torch.zeros(sp.ne * 6 * Q3, dtype=torch.float64, device="cuda")
    ctx.pa_qdata(qd)                                   # MFEM pa_data layout [Q^3][6][ne] with q fastest
    assert rel_err(qd.cpu().numpy().reshape(sp.ne, 6, Q3), A.qd) < 1e-13
    x = orc.hash_noise(sp.ndof)
    xd, yd = _dev(torch, x), torch.empty(sp.ndof, dtype=torch.float64, device="cuda")
    ctx.apply_L(xd, yd)
    y = yd.cpu().numpy()
    assert rel_err(y, A.mult(x)) < TOL_OP
    if fa is not None:
        assert rel_err(y, fa.mult(x)) < TOL_OP         # against the full-assembly twin as well
    ctx.set_option("max_ctas", 1)                      # one persistent CTA walks all batches
    ctx.apply_L(xd, yd)
    assert rel_err(yd.cpu().numpy(), A.mult(x)) < TOL_OP
    ctx.set_option("max_ctas", 0)
    # E-vector entry point: AddMultPA semantics (accumulates)
    xE = _dev(torch, x[sp.gather])
    yE = torch.ones_like(xE)
    ctx.pa_apply_E(xE, yE)
    assert rel_err(yE.cpu().numpy() - 1.0, orc.pa_apply_E(A.qd, osp.basis, x[sp.gather])) < TOL_OP
    # constrained operator
    ctx.apply_T(xd, yd)
    assert rel_err(yd.cpu().numpy(), orc.ConstrainedOperator(A, osp.ess).mult(x)) < TOL_OP
    dg = torch.empty_like(xd)
    ctx.diag(dg)
    assert rel_err(dg.cpu().numpy(), A.diag()) < TOL_OP
    # AssembleDiagonalPA: E-vector diagonal, accumulated
    dE = torch.ones(sp.ne * (p + 1) ** 3, dtype=torch.float64, device="cuda")
    ctx.pa_diag_E(dE)
    assert rel_err(dE.cpu().numpy().reshape(sp.ne, -1) - 1.0, orc.pa_diag_E(A.qd, osp.basis)) < TOL_OP
    ctx.jacobi_setup()
    ctx.jacobi_dinv(dg)
    assert rel_err(dg.cpu().numpy(), orc.jacobi_dinv(A, osp.ess)) < TOL_OP
    ctx.close()


@pytest.mark.parametrize("p,variant", [(p, v) for p in range(1, 11) for v in (20, 30)] + [(7, 33), (8, 33)]
                         + [(p, v) for p in range(5, 10) for v in (40, 41)])
def test_apply_kernel_alternative_variants(lpf, orc, cuda, p, variant):
    """The non-default (elements per CTA, CTAs per SM) instantiations of the persistent kernel, forced through
    several batches per CTA."""
    torch = cuda
    sp = lpf.Space(lpf.Mesh.wave_tank(5, 1, 3).perturb(0.12) if p >= 5 else lpf.Mesh.wave_tank(7, 1, 3).refine(1).perturb(0.12), p)
    A = orc.PAOperator(oracle_space_from(orc, sp))
    ctx = _ctx(lpf, torch, sp)
    ctx.set_option("apply_variant", variant)
    x = orc.hash_noise(sp.ndof, seed=11)
    xd, yd = _dev(torch, x), torch.empty(sp.ndof, dtype=torch.float64, device="cuda")
    ref = A.mult(x)
    for max_ctas in (0, 3, 1):
        ctx.set_option("max_ctas", max_ctas)
        ctx.apply_L(xd, yd)
        assert rel_err(yd.cpu().numpy(), ref) < TOL_OP, (p, variant, max_ctas)
    ctx.close()


@pytest.mark.parametrize("variant", [0, 20, 30])
def test_apply_kernel_variants_p4(lpf, orc, cuda, variant):
    """Every compiled (elements-per-CTA, pipelining) variant of the order-4 kernel, on a mesh whose element
    count (7x1x3 refined once = 168, perturbed) is ragged for every batch size and spans several batches
    per CTA for the persistent kernels."""
    torch = cuda
    tank = lpf.Mesh.wave_tank(7, 1, 3).refine(1).perturb(0.12)
    sp = lpf.Space(tank, 4)
    A = orc.PAOperator(oracle_space_from(orc, sp))
    ctx = _ctx(lpf, torch, sp)
    ctx.set_option("apply_variant", variant)
    x = orc.hash_noise(sp.ndof, seed=11)
    xd, yd = _dev(torch, x), torch.empty(sp.ndof, dtype=torch.float64, device="cuda")
    ref = A.mult(x)
    for max_ctas in (0, 5, 1):          # persistent kernels: many batches per CTA, odd/even iteration counts
        ctx.set_option("max_ctas", max_ctas)
        ctx.apply_L(xd, yd)
        assert rel_err(yd.cpu().numpy(), ref) < TOL_OP, (variant, max_ctas)
    # PCG denominator path (x . A x accumulated in the kernel) through a short constrained solve
    ctx.jacobi_setup()
    b = _dev(torch, orc.hash_noise(sp.ndof, seed=12))
    xs = torch.zeros_like(b)
    info = ctx.pcg(b, xs, rel_tol=1e-10, max_iter=400)
    ctx.set_option("apply_variant", 0); ctx.set_option("max_ctas", 0)
    xs0 = torch.zeros_like(b)
    info0 = ctx.pcg(b, xs0, rel_tol=1e-10, max_iter=400)
    assert info.converged and abs(info.iterations - info0.iterations) <= 1
    assert rel_err(xs.cpu().numpy(), xs0.cpu().numpy()) < 1e-8
    ctx.close()


def test_operator_properties_and_ragged_sizes(lpf, orc, cuda):
    """Element counts that are not multiples of the elements-per-CTA batch, linearity, symmetry, A 1 = 0."""
    torch = cuda
    for (nx, ny, nz, p) in [(3, 1, 1, 4), (5, 1, 1, 4), (7, 1, 3, 3), (3, 1, 1, 1), (13, 1, 1, 2)]:
        sp = lpf.Space(lpf.Mesh.wave_tank(nx, ny, nz).perturb(0.1), p)
        A = orc.PAOperator(oracle_space_from(orc, sp))
        ctx = _ctx(lpf, torch, sp)
        x, z = orc.hash_noise(sp.ndof), orc.hash_noise(sp.ndof, seed=99)
        xd, zd = _dev(torch, x), _dev(torch, z)
        yx, yz, ys = (torch.empty_like(xd) for _ in range(3))
        ctx.apply_L(xd, yx); ctx.apply_L(zd, yz); ctx.apply_L(2.0 * xd - 3.0 * zd, ys)
        assert rel_err(yx.cpu().numpy(), A.mult(x)) < TOL_OP
        assert rel_err(ys.cpu().numpy(), (2.0 * yx - 3.0 * yz).cpu().numpy()) < TOL_OP
        assert abs(float(torch.dot(zd, yx) - torch.dot(xd, yz))) < 1e-12 * abs(float(torch.dot(zd, yx)))
        ctx.apply_L(torch.ones_like(xd), ys)
        assert float(ys.abs().max()) < 1e-12 * float(yx.abs().max())
        ctx.close()


def test_cylinder_mesh_operator(lpf, orc, cuda):
    """BASELINE config 4 geometry (mesh_cylinder_half): non-affine hexes, unstructured numbering."""
    torch = cuda
    m = lpf.Mesh.read(os.path.join(HERE, "meshes", "cylinder_half.mesh"))
    for p in (2, 4):
        sp = lpf.Space(m, p)
        A = orc.PAOperator(oracle_space_from(orc, sp))
        ctx = _ctx(lpf, torch, sp)
        x = orc.hash_noise(sp.ndof)
        xd, yd = _dev(torch, x), torch.empty(sp.ndof, dtype=torch.float64, device="cuda")
        ctx.apply_T(xd, yd)
        assert rel_err(yd.cpu().numpy(), orc.ConstrainedOperator(A, np.sort(sp.ess)).mult(x)) < TOL_OP
        ctx.diag(yd)
        assert rel_err(yd.cpu().numpy(), A.diag()) < TOL_OP
        ctx.close()


def test_api_state_errors(lpf, cuda, tank):
    torch = cuda
    sp = lpf.Space(tank, 2)
    ctx = lpf.Context(sp, device=0)
    x = torch.zeros(sp.ndof, dtype=torch.float64, device="cuda")
    with pytest.raises(lpf.LpfError, match="before lpf_pa_setup"):
        ctx.apply_L(x, x.clone())
    ctx.pa_setup()
    with pytest.raises(lpf.LpfError, match="before lpf_jacobi_setup"):
        ctx.pcg(x, x.clone())
    with pytest.raises(lpf.LpfError, match="unknown option"):
        ctx.set_option("nonsense", 1)
    with pytest.raises(lpf.LpfError, match="no such CUDA device"):
        lpf.Context(sp, device=99)
    ctx.close()


@pytest.mark.parametrize("p,rel", [(2, 1e-12), (4, 1e-12), (4, 1e-8), (6, 1e-12), (9, 1e-12), (10, 1e-12)])
def test_laplace_solve_matches_oracle_pcg(lpf, orc, cuda, tank, p, rel):
    """FormLinearSystem + CGSolver(Jacobi) + RecoverFEMSolution: iterations +-1, potential 1e-10."""
    torch = cuda
    sp = lpf.Space(tank, p)
    osp = oracle_space_from(orc, sp)
    A = orc.PAOperator(osp)
    dinv = orc.jacobi_dinv(A, osp.ess)
    wv = orc.Wave()
    x0 = np.zeros(sp.ndof)
    x0[osp.ess] = wv.phi(0.0, osp.xyz[osp.ess, 0], osp.xyz[osp.ess, 1], osp.xyz[osp.ess, 2], H)
    Ac, X, B = orc.form_linear_system(A, osp.ess, x0, np.zeros(sp.ndof))
    X, info = orc.pcg(Ac, dinv, B, X, rel, 0.0, 1000)
    ctx = _ctx(lpf, torch, sp)
    ctx.jacobi_setup()
    phi = _dev(torch, x0 + 0.123 * (1 - np.isin(np.arange(sp.ndof), osp.ess)))   # interior garbage must be ignored
    gi = ctx.laplace_solve(phi, rel_tol=rel, max_iter=1000)
    assert gi.converged and abs(gi.iterations - info.iters) <= 1, (gi, info)
    tol = TOL_SOL if rel <= 1e-12 else 50 * rel
    assert rel_err(phi.cpu().numpy(), X) < tol
    assert gi.applies in (info.applies, info.applies - 1, info.applies + 1, info.applies - 2)
    # generic CGSolver::Mult entry point with a non-trivial initial guess
    b = orc.hash_noise(sp.ndof, seed=5); xg = orc.hash_noise(sp.ndof, seed=6)
    Xo, io = orc.pcg(orc.ConstrainedOperator(A, osp.ess), dinv, b, xg.copy(), rel, 0.0, 1000)
    bd, xd = _dev(torch, b), _dev(torch, xg)
    g2 = ctx.pcg(bd, xd, rel_tol=rel, max_iter=1000)
    assert g2.converged and abs(g2.iterations - io.iters) <= 1
    assert rel_err(xd.cpu().numpy(), Xo) < tol
    assert abs(g2.initial_norm - io.initial_norm) < 1e-10 * io.initial_norm
    # without CUDA graphs the result is the same
    ctx.set_option("use_graph", 0)
    xd2 = _dev(torch, xg)
    g3 = ctx.pcg(bd, xd2, rel_tol=rel, max_iter=1000)
    assert abs(g3.iterations - g2.iterations) <= 1 and rel_err(xd2.cpu().numpy(), Xo) < tol
    ctx.close()


def test_pcg_stopping_rules(lpf, orc, cuda, tank):
    """max_iter cap (rel_tol 1e-24 drivers always hit it), zero right-hand side, chunking."""
    torch = cuda
    sp = lpf.Space(tank, 3)
    osp = oracle_space_from(orc, sp)
    A = orc.PAOperator(osp)
    dinv = orc.jacobi_dinv(A, osp.ess)
    ctx = _ctx(lpf, torch, sp)
    ctx.jacobi_setup()
    b = orc.hash_noise(sp.ndof, seed=5)
    for chunk in (1, 7, 16):
        ctx.set_option("pcg_chunk", chunk)
        xd = torch.zeros(sp.ndof, dtype=torch.float64, device="cuda")
        gi = ctx.pcg(_dev(torch, b), xd, rel_tol=1e-24, max_iter=25)
        Xo, io = orc.pcg(orc.ConstrainedOperator(A, osp.ess), dinv, b, np.zeros(sp.ndof), 1e-24, 0.0, 25)
        assert (gi.iterations, gi.converged) == (25, 0) and io.iters == 25 and not io.converged
        assert rel_err(xd.cpu().numpy(), Xo) < 1e-9
    z = torch.zeros(sp.ndof, dtype=torch.float64, device="cuda")
    gi = ctx.pcg(z, z.clone(), rel_tol=1e-12, max_iter=100)
    assert gi.iterations == 0 and gi.converged == 1           # nom = 0 <= r0 = 0
    ctx.close()


@pytest.mark.parametrize("p", [2, 4])
def test_surface_dz(lpf, orc, cuda, tank, p):
    torch = cuda
    sp = lpf.Space(tank, p)
    osp = oracle_space_from(orc, sp)
    ctx = _ctx(lpf, torch, sp)
    phi = orc.hash_noise(sp.ndof, seed=21)
    wt = torch.empty(sp.nsurf, dtype=torch.float64, device="cuda")
    ctx.surface_dz(_dev(torch, phi), wt)
    ref = orc.get_derivative_z(osp, phi)[osp.surf2vol]        # full-volume GetDerivative, then the trace
    assert rel_err(wt.cpu().numpy(), ref) < TOL_OP
    ctx.close()


def _initial_state(orc, sp):
    wv = orc.Wave()
    xs, ys = sp.surf_xy[:, 0], sp.surf_xy[:, 1]
    return wv, np.concatenate([wv.eta(0.0, xs, ys), wv.phi_fs(0.0, xs, ys)])


@pytest.mark.parametrize("relax", [False, True])
def test_rhs_and_rk4_match_oracle(lpf, orc, cuda, tank, relax):
    """rhs_linear::Mult + RK4Solver::Step (ss.cpp RHS without zones; PF_linear_par_partial.cpp with zones)."""
    torch = cuda
    p = 4
    sp = lpf.Space(tank, p)
    osp = oracle_space_from(orc, sp)
    wv, st0 = _initial_state(orc, sp)
    dt = wv.T / 40
    xs = sp.surf_xy[:, 0]
    rx = None
    cgen = cabs = None
    if relax:
        cgen, cabs = orc.relax_cgen(xs, 0.0, 0.4), orc.relax_cabs(xs, 0.7, 1.0)
        rx = orc.Relax(cgen, cabs, tau=dt)
    f = orc.RhsLinear(osp, wv, rel_tol=1e-12, max_iter=1000, relax=rx)
    ctx = _ctx(lpf, torch, sp)
    ctx.jacobi_setup()
    ctx.rhs_setup(lpf.make_rhs_params(lpf.wave_params(), tau=dt, use_relaxation=relax, rel_tol=1e-12, max_iter=1000), cgen, cabs)
    # single RHS evaluation at a non-zero stage time
    f.set_time(0.37)
    ko = f.mult(st0)
    sd, kd = _dev(torch, st0), torch.empty(2 * sp.nsurf, dtype=torch.float64, device="cuda")
    ctx.rhs(0.37, sd, kd)
    assert rel_err(kd.cpu().numpy()[:sp.nsurf], ko[:sp.nsurf]) < TOL_SOL
    assert rel_err(kd.cpu().numpy()[sp.nsurf:], ko[sp.nsurf:]) < TOL_SOL
    # three RK4 steps, device state vs oracle state
    f.iters.clear()
    so, to, tg = st0.copy(), 0.0, 0.0
    for _ in range(3):
        so, to = orc.rk4_step(f, so, to, dt)
        tg = ctx.rk4_step(sd, tg, dt)
    assert abs(tg - to) < 1e-15
    assert rel_err(sd.cpu().numpy()[:sp.nsurf], so[:sp.nsurf]) < TOL_SOL      # eta
    assert rel_err(sd.cpu().numpy()[sp.nsurf:], so[sp.nsurf:]) < TOL_SOL      # phi_fs
    its = [i.iterations for i in ctx.last_solve_info()]
    assert len(its) == 4 and all(abs(a - b) <= 1 for a, b in zip(its, f.iters[-4:])), (its, f.iters[-4:])
    # host-buffer entry point gives the same step
    hs = torch.from_numpy(st0.copy()).pin_memory()
    t1 = ctx.rk4_step_host(hs, 0.0, dt)
    f2 = orc.RhsLinear(osp, wv, rel_tol=1e-12, max_iter=1000, relax=rx)
    s1, _ = orc.rk4_step(f2, st0.copy(), 0.0, dt)
    assert abs(t1 - dt) < 1e-15 and rel_err(hs.numpy(), s1) < TOL_SOL
    ctx.close()


def test_cylinder_rhs_third_weight_and_envelope(lpf, orc, cuda):
    """cylinder-diffraction.cpp: RHS with C_gen, C_abs and C_absy on the unstructured half-cylinder mesh, the eta
    envelope (max over steps, scaled by 2/H) and the rim sampling of it."""
    torch = cuda
    p = 2
    m = lpf.Mesh.read(os.path.join(HERE, "meshes", "cylinder_half.mesh"))
    sp = lpf.Space(m, p)
    osp = oracle_space_from(orc, sp)
    wv, st0 = _initial_state(orc, sp)
    dt = wv.T / 35
    lo, hi = m.bounding_box()
    xs, ys = sp.surf_xy[:, 0], sp.surf_xy[:, 1]
    cgen = orc.relax_cgen(xs, lo[0], lo[0] + 2.5)
    cabs = orc.relax_cabs(xs, hi[0] - 4.0, hi[0])
    cabsy = orc.relax_cabs(ys, hi[1] - 3.0, hi[1])
    assert cabsy.max() == 1.0 and (cabsy > 0).sum() > 10
    f = orc.RhsLinear(osp, wv, rel_tol=1e-12, max_iter=2000, relax=orc.Relax(cgen, cabs, tau=dt, cabsy=cabsy))
    ctx = _ctx(lpf, torch, sp)
    ctx.jacobi_setup()
    ctx.rhs_setup(lpf.make_rhs_params(lpf.wave_params(), tau=dt, use_relaxation=True, rel_tol=1e-12, max_iter=2000), cgen, cabs)
    ctx.rhs_set_cabsy(cabsy)
    ns = sp.nsurf
    f.set_time(0.21)
    ko = f.mult(st0)
    sd, kd = _dev(torch, st0), torch.empty(2 * ns, dtype=torch.float64, device="cuda")
    ctx.rhs(0.21, sd, kd)
    assert rel_err(kd.cpu().numpy()[:ns], ko[:ns]) < TOL_SOL and rel_err(kd.cpu().numpy()[ns:], ko[ns:]) < TOL_SOL
    # without the third weight the RHS differs where C_absy > 0 -- the term is really applied
    ctx.rhs_set_cabsy(None)
    ctx.rhs(0.21, sd, kd)
    assert rel_err(kd.cpu().numpy()[:ns], ko[:ns]) > 1e-6
    ctx.rhs_set_cabsy(cabsy)
    # one RK4 step with the envelope (the numpy oracle needs ~9 s per solve on this mesh)
    ctx.envelope_reset()
    so, to, tg = st0.copy(), 0.0, 0.0
    env = np.full(ns, -1e300)
    for _ in range(1):
        so, to = orc.rk4_step(f, so, to, dt)
        tg = ctx.rk4_step(sd, tg, dt)
        env = np.maximum(env, so[:ns])
        ctx.envelope_update(sd)
    assert rel_err(sd.cpu().numpy(), so) < TOL_SOL
    eg = ctx.envelope_get(2.0 / wv.H)
    assert rel_err(eg, env * (2.0 / wv.H)) < TOL_SOL
    th, idx = sp.rim()
    assert len(th) >= 8 and np.all(np.isfinite(eg[idx]))
    ctx.close()


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
def test_affine_fast_path(lpf, orc, cuda, tank, p):
    """Affine hexes: D(q) = w_q * (detJ J^-1 J^-T) from one 6-entry tensor per element instead of stored q-data.
    Same operator to 1e-12, same solve; switched off automatically on a non-affine mesh."""
    torch = cuda
    m = lpf.Mesh.wave_tank(5, 2, 3).refine(1) if p <= 4 else lpf.Mesh.wave_tank(4, 1, 2)
    sp = lpf.Space(m, p)
    osp = oracle_space_from(orc, sp)
    A = orc.PAOperator(osp)
    x = orc.hash_noise(sp.ndof)
    ref = orc.ConstrainedOperator(A, np.sort(sp.ess)).mult(x)
    xd, yd = _dev(torch, x), torch.empty(sp.ndof, dtype=torch.float64, device="cuda")
    ctx = _ctx(lpf, torch, sp)
    assert ctx.affine_active
    ctx.set_option("max_ctas", 7)                 # several batches per persistent CTA
    ctx.apply_T(xd, yd)
    y_aff = yd.cpu().numpy().copy()
    assert rel_err(y_aff, ref) < TOL_OP
    ctx.set_option("affine", 0)
    assert not ctx.affine_active
    ctx.apply_T(xd, yd)
    assert rel_err(yd.cpu().numpy(), ref) < TOL_OP and rel_err(yd.cpu().numpy(), y_aff) < TOL_OP
    # Laplace solve: same potential, iteration counts within +-1 of the stored-q-data path
    ctx.jacobi_setup()
    phi0 = np.zeros(sp.ndof)
    phi0[sp.surf2vol] = np.cos(2 * np.pi * sp.surf_xy[:, 0])
    pd0 = _dev(torch, phi0)
    i0 = ctx.laplace_solve(pd0, rel_tol=1e-10, max_iter=2000)
    ctx.set_option("affine", 1)
    pd1 = _dev(torch, phi0)
    i1 = ctx.laplace_solve(pd1, rel_tol=1e-10, max_iter=2000)
    assert i0.converged and i1.converged and abs(i0.iterations - i1.iterations) <= 1
    assert rel_err(pd1.cpu().numpy(), pd0.cpu().numpy()) < 1e-8          # both solved to rel 1e-10
    ctx.close()
    # a perturbed copy of the same mesh is not affine: the fast path must stay off by itself
    sp2 = lpf.Space(tank, min(p, 3))
    c2 = _ctx(lpf, torch, sp2)
    assert not c2.affine_active
    c2.close()


def _time_stepper_error(lpf, torch, mesh, p, nsteps=150, rel_tol=1e-13):
    """convergence-parallel-partial.cpp: one period in nsteps (+1) RK4 steps, nodal max error of eta / phi_fs at t = T + dt."""
    sp = lpf.Space(mesh, p)
    ctx = _ctx(lpf, torch, sp)
    ctx.jacobi_setup()
    w = lpf.wave_params()
    dt = w["T"] / nsteps
    ctx.rhs_setup(lpf.make_rhs_params(w, rel_tol=rel_tol, max_iter=2000))
    xs = sp.surf_xy[:, 0]
    coth = 1.0 / np.tanh(w["kh"])
    eta = lambda t: 0.5 * w["H"] * np.cos(w["omega"] * t - w["k"] * xs)
    pfs = lambda t: -0.5 * w["H"] * w["cwave"] * coth * np.sin(w["omega"] * t - w["k"] * xs)
    sd = _dev(torch, np.concatenate([eta(0.0), pfs(0.0)]))
    t = 0.0
    for _ in range(nsteps + 1):
        t = ctx.rk4_step(sd, t, dt)
    st = sd.cpu().numpy()
    ctx.close()
    return sp.ndof, np.abs(st[:sp.nsurf] - eta(t)).max(), np.abs(st[sp.nsurf:] - pfs(t)).max()


def test_time_stepper_p_convergence_known_answers(lpf, cuda):
    """The reference's own verification (Convergence_and_Scaling/convergence-parallel-partial.cpp:150,249-305): spectral
    p-convergence of eta after one wave period on the 3-element periodic wave-tank.mesh.  Known answers: SURVEY.md App. E
    (restated prototype, nodal max norms) -- same digits expected to ~10 %."""
    torch = cuda
    known = {1: (12, 5.3e-3, 9.2e-3), 2: (54, 4.6e-4, 7.3e-4), 3: (144, 1.3e-4, 1.8e-4), 4: (300, 5.2e-6, 7.0e-6),
             5: (540, 5.6e-7, 7.5e-7), 6: (882, 1.5e-8, 1.8e-8)}
    prev = None
    for p, (dofs, e_eta, e_phi) in known.items():
        n, ee, ep = _time_stepper_error(lpf, torch, lpf.Mesh.wave_tank(3, 1, 1), p)
        assert n == dofs
        assert 0.8 * e_eta < ee < 1.25 * e_eta, (p, ee, e_eta)
        assert 0.8 * e_phi < ep < 1.25 * e_phi, (p, ep, e_phi)
        assert prev is None or ee < prev            # monotone spectral decay
        prev = ee


def test_time_stepper_h_convergence_known_answers(lpf, cuda):
    """convergence-parallel-partial-hconv.cpp idea at order 4: eta error 5.2e-6 / 4.3e-7 / 3.3e-8 on 0 / 1 / 2 uniform
    refinements, i.e. a factor >= 10 per level."""
    torch = cuda
    known = [(300, 5.2e-6), (1944, 4.3e-7), (13872, 3.3e-8)]
    m = lpf.Mesh.wave_tank(3, 1, 1)
    errs = []
    for lvl, (dofs, e_eta) in enumerate(known):
        n, ee, _ = _time_stepper_error(lpf, torch, m, 4)
        assert n == dofs and 0.8 * e_eta < ee < 1.25 * e_eta, (lvl, n, ee)
        errs.append(ee)
        m = m.refine(1)
    assert errs[0] / errs[1] > 10 and errs[1] / errs[2] > 10


def test_golden_vectors(lpf, cuda):
    """Committed oracle outputs (tests/golden/tank_p3.npz, made by make_golden.py) through the
    arrays-only descriptor route an MFEM adapter would take."""
    torch = cuda
    g = np.load(os.path.join(HERE, "golden", "tank_p3.npz"))
    p = int(g["order"])
    sp = lpf.Space.from_arrays(p, g["corners"], g["gather"], g["ess"], g["surf2vol"], g["surf_xy"])
    ctx = _ctx(lpf, torch, sp)
    ctx.jacobi_setup()
    xd, yd = _dev(torch, g["x"]), torch.empty(sp.ndof, dtype=torch.float64, device="cuda")
    ctx.apply_L(xd, yd); assert rel_err(yd.cpu().numpy(), g["y"]) < TOL_OP
    ctx.apply_T(xd, yd); assert rel_err(yd.cpu().numpy(), g["yc"]) < TOL_OP
    ctx.diag(yd); assert rel_err(yd.cpu().numpy(), g["diag"]) < TOL_OP
    ns = sp.nsurf
    phi = torch.zeros(sp.ndof, dtype=torch.float64, device="cuda")
    phi[torch.from_numpy(g["surf2vol"].astype(np.int64)).cuda()] = _dev(torch, g["state0"][ns:])
    info = ctx.laplace_solve(phi, rel_tol=1e-12, max_iter=1000)
    assert abs(info.iterations - int(g["its0"])) <= 1 and rel_err(phi.cpu().numpy(), g["phi"]) < TOL_SOL
    wt = torch.empty(ns, dtype=torch.float64, device="cuda")
    ctx.surface_dz(phi, wt); assert rel_err(wt.cpu().numpy(), g["wt"]) < TOL_SOL
    w = lpf.wave_params()
    dt = float(g["dt"])
    for relax, key in ((False, "state2"), (True, "state2_relax")):
        ctx.rhs_setup(lpf.make_rhs_params(w, tau=dt, use_relaxation=relax), g["cgen"], g["cabs"])
        sd, t = _dev(torch, g["state0"]), 0.0
        for _ in range(2):
            t = ctx.rk4_step(sd, t, dt)
        assert rel_err(sd.cpu().numpy()[:ns], g[key][:ns]) < TOL_SOL
        assert rel_err(sd.cpu().numpy()[ns:], g[key][ns:]) < TOL_SOL
    ctx.close()


def test_mfem_adapter_descriptor_route(lpf, orc, cuda, tank):
    """The route include/lpf_mfem_adapter.hpp takes: geometry as a Jacobian array in MFEM's
    GeometricFactors::J layout [Q^3][3][3][ne] plus nodal J^-1 columns, no corner coordinates."""
    torch = cuda
    p = 3
    sp0 = lpf.Space(tank, p)
    osp = oracle_space_from(orc, sp0)
    bs = osp.basis
    J = orc.jacobians_trilinear(osp.mesh.corners, bs.qpts)                  # [e, q, c, k]
    jac = np.ascontiguousarray(np.transpose(J, (0, 3, 2, 1)))               # [e][k][c][q]  == J[q + Q3*(c + 3*(k + 3e))]
    Jn = orc.jacobians_trilinear(osp.mesh.corners, bs.nodes)
    jinv_z = np.ascontiguousarray(np.linalg.inv(Jn)[:, :, :, 2])            # [e][n][k]: Jinv(k, z)
    sp = lpf.Space.from_arrays(p, sp0.corners, sp0.gather, sp0.ess, sp0.surf2vol, sp0.surf_xy, jac=jac, jinv_z=jinv_z)
    ctx = _ctx(lpf, torch, sp)
    ctx.jacobi_setup()
    A = orc.PAOperator(osp)
    x = orc.hash_noise(sp.ndof)
    xd, yd = _dev(torch, x), torch.empty(sp.ndof, dtype=torch.float64, device="cuda")
    ctx.apply_T(xd, yd)
    assert rel_err(yd.cpu().numpy(), orc.ConstrainedOperator(A, osp.ess).mult(x)) < TOL_OP
    wv, st0 = _initial_state(orc, sp0)
    dt = wv.T / 40
    ctx.rhs_setup(lpf.make_rhs_params(lpf.wave_params(), rel_tol=1e-12, max_iter=1000))
    sd = _dev(torch, st0)
    ctx.rk4_step(sd, 0.0, dt)
    so, _ = orc.rk4_step(orc.RhsLinear(osp, wv, rel_tol=1e-12, max_iter=1000, operator=A), st0.copy(), 0.0, dt)
    assert rel_err(sd.cpu().numpy(), so) < TOL_SOL
    ctx.close()


def test_full_size_properties_big8(lpf, cuda):
    """wave-tank-big8 (4096 hexes, 299 520 dofs at p=4; BASELINE config 3) and one refinement of it:
    size-independent checks -- symmetry, null space, positivity, E-vector vs L-vector consistency,
    analytic Airy solve (the reference's own known answer)."""
    torch = cuda
    m = lpf.Mesh.wave_tank(128, 2, 16)
    sp = lpf.Space(m, 4)
    assert sp.ndof == 299520
    ctx = _ctx(lpf, torch, sp)
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(sp.ndof, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    z = torch.rand(sp.ndof, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    yx, yz = torch.empty_like(x), torch.empty_like(x)
    ctx.apply_L(x, yx); ctx.apply_L(z, yz)
    assert abs(float(torch.dot(z, yx) - torch.dot(x, yz))) < 1e-11 * abs(float(torch.dot(z, yx)))
    assert float(torch.dot(x, yx)) > 0
    ctx.apply_L(torch.ones_like(x), yz)
    assert float(yz.abs().max()) < 1e-11 * float(yx.abs().max())
    # gather -> AddMultPA -> scatter by hand equals the fused path
    gm = torch.from_numpy(sp.gather.astype(np.int64)).cuda()
    xE = x[gm].contiguous()
    yE = torch.zeros_like(xE)
    ctx.pa_apply_E(xE, yE)
    y2 = torch.zeros_like(x).index_add_(0, gm.reshape(-1), yE.reshape(-1))
    assert float((y2 - yx).abs().max()) < 1e-12 * float(yx.abs().max())
    # Airy known answer: Dirichlet data from the exact potential on the free surface
    ctx.jacobi_setup()
    w = lpf.wave_params()
    xyz = sp.node_coordinates()
    ex = -0.5 * w["H"] * w["cwave"] * np.cosh(w["k"] * xyz[:, 2]) / np.sinh(w["kh"]) * np.sin(-w["k"] * xyz[:, 0])
    phi0 = np.zeros(sp.ndof); phi0[sp.ess] = ex[sp.ess]
    phi = _dev(torch, phi0)
    info = ctx.laplace_solve(phi, rel_tol=1e-12, max_iter=1000)
    assert info.converged and 200 <= info.iterations <= 300          # 247 at rel 1e-12
    assert np.abs(phi.cpu().numpy() - ex).max() < 1e-9 * np.abs(ex).max() * 1e3
    # host-buffer entry point at a size where it pipelines H2D / element chunks / D2H over three streams: same result as
    # the device-resident apply, with and without the pipeline, general and affine kernels, repeated calls
    yd = torch.empty_like(x)
    xh = x.cpu().pin_memory()
    for aff in (1, 0):
        ctx.set_option("affine", aff)
        ctx.apply_T(x, yd)
        ref = yd.cpu().numpy()
        for pipe in (1, 0, 1):
            ctx.set_option("host_pipeline", pipe)
            yh = torch.full((sp.ndof,), np.nan, dtype=torch.float64).pin_memory()
            ctx.apply_T_host(xh, yh)
            assert rel_err(yh.numpy(), ref) < TOL_OP, (aff, pipe)
    ctx.close()


def test_host_pipeline_on_unstructured_mesh(lpf, orc, cuda):
    """Pipelined host apply on the cylinder mesh (order 6: 727 k dofs): chunk / range dependencies come from the gather
    map, not from any structure of the mesh."""
    torch = cuda
    m = lpf.Mesh.read(os.path.join(HERE, "meshes", "cylinder_half.mesh"))
    sp = lpf.Space(m, 6)
    assert sp.ndof > (1 << 18)
    ctx = _ctx(lpf, torch, sp)
    x = _dev(torch, orc.hash_noise(sp.ndof))
    yd = torch.empty_like(x)
    ctx.apply_T(x, yd)
    xh = x.cpu().pin_memory()
    yh = torch.full((sp.ndof,), np.nan, dtype=torch.float64).pin_memory()
    ctx.apply_T_host(xh, yh)
    assert rel_err(yh.numpy(), yd.cpu().numpy()) < TOL_OP
    yp = torch.full((sp.ndof,), np.nan, dtype=torch.float64)          # pageable host memory also works (staged copies)
    ctx.apply_T_host(x.cpu(), yp)
    assert rel_err(yp.numpy(), yd.cpu().numpy()) < TOL_OP
    ctx.close()
