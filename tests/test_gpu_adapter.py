"""The MFEM adapter (include/lpf_mfem_adapter.hpp) EXECUTED on the GPU against the functional MFEM stand-in
(drivers/stub/mfem.hpp) by drivers/bin/adapter_check, compared with direct C-ABI calls on the same problem.

The adapter builds its descriptor the way a maintainer would from MFEM objects (GeometricFactors Jacobians at the
DiffusionIntegrator rule, ElementRestriction gather map, essential true dofs, free-surface maps) -- so the two sides differ
in how the geometry reaches the library (Jacobian array vs trilinear corners) and must agree to the operator tolerance."""
import os
import struct
import subprocess

import numpy as np
import pytest

from util import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "master-thesis-lpf-in-mfem_b200", "drivers", "bin", "adapter_check")


def _read(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    o = 0
    while o < len(data):
        (ln,) = struct.unpack_from("i", data, o); o += 4
        name = data[o:o + ln].decode(); o += ln
        (n,) = struct.unpack_from("i", data, o); o += 4
        out[name] = np.frombuffer(data, dtype=np.float64, count=n, offset=o).copy(); o += 8 * n
    return out


def _noise(i):
    z = ((i.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)) * np.uint64(0xBF58476D1CE4E5B9))
    z ^= z >> np.uint64(29)
    return (z % np.uint64(2000001)).astype(np.float64) / 1e6 - 1.0


def _ensure_built():
    if not os.path.exists(BIN):
        subprocess.check_call(["bash", os.path.join(ROOT, "master-thesis-lpf-in-mfem_b200", "drivers", "build.sh")])


@pytest.mark.parametrize("nranks,mesh,order", [(4, "tank:16,2,4", 3), (8, "tank:32,2,8", 4), (3, "cyl", 2)])
def test_parallel_descriptor_from_group_communicator(nranks, mesh, order):
    """CPU: the halo plan SpaceDescBuilder derives from a ParFiniteElementSpace's GroupCommunicator (groups -> neighbour
    lists, ownership, rank-ordered reduction sources; T-dof numbering) equals the library's own plan on every rank."""
    _ensure_built()
    spec = os.path.join(ROOT, "tests", "meshes", "cylinder_half.mesh") if mesh == "cyl" else mesh
    r = subprocess.run([BIN, "host-par", str(nranks), spec, str(order)], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mesh,order", [("tank:6,2,3", 4), ("cyl", 2)])
def test_adapter_classes_run_on_gpu(lpf, cuda, tmp_path, mesh, order):
    torch = cuda
    _ensure_built()
    spec = os.path.join(ROOT, "tests", "meshes", "cylinder_half.mesh") if mesh == "cyl" else mesh
    out = str(tmp_path / "adapter.bin")
    r = subprocess.run([BIN, "gpu", spec, str(order), out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = _read(out)
    # the same problem through the C-ABI (corners -> q-data inside the library)
    m = lpf.Mesh.read(spec) if mesh == "cyl" else lpf.Mesh.wave_tank(6, 2, 3)
    sp = lpf.Space(m, order)
    ctx = lpf.Context(sp, device=0, stream=torch.cuda.current_stream().cuda_stream)
    ctx.pa_setup()
    ctx.set_option("affine", 0)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    D3 = (order + 1) ** 3
    xE = dev(_noise(np.arange(sp.ne * D3)))
    yE = torch.ones_like(xE)
    ctx.pa_apply_E(xE, yE)
    assert rel_err(got["AddMultPA"], yE.cpu().numpy()) < 1e-12
    dE = torch.zeros_like(xE)
    ctx.pa_diag_E(dE)
    assert rel_err(got["AssembleDiagonalPA"], dE.cpu().numpy()) < 1e-12
    x = dev(_noise(np.arange(sp.ndof) + 17))
    y = torch.empty_like(x)
    ctx.apply_T(x, y)
    assert rel_err(got["Mult"], y.cpu().numpy()) < 1e-12
    dg = torch.empty_like(x)
    ctx.diag(dg)
    assert rel_err(got["AssembleDiagonal"], dg.cpu().numpy()) < 1e-12
    # the integrator's E-vector diagonal, restricted (G^T), is the operator's diagonal: what OperatorJacobiSmoother builds
    dsum = np.zeros(sp.ndof)
    np.add.at(dsum, sp.gather.reshape(-1), got["AssembleDiagonalPA"])
    assert rel_err(dsum, got["AssembleDiagonal"]) < 1e-12
    # CG with the adapter's right-hand side
    ctx.jacobi_setup()
    X0 = np.zeros(sp.ndof)
    ess = np.asarray(sp.ess)
    X0[ess] = np.cos(6.0 * sp.surf_xy[np.arange(len(ess)), 0])
    B = torch.empty_like(x)
    ctx.apply_T(dev(X0), B)
    Bn = 0.5 * B.cpu().numpy() + 1e-3 * _noise(np.arange(sp.ndof) + 5)
    Bn[ess] = X0[ess]
    Xd = dev(X0)
    info = ctx.pcg(dev(Bn), Xd, rel_tol=1e-12, max_iter=2000)
    assert abs(info.iterations - int(got["CG_info"][0])) <= 1 and int(got["CG_info"][1]) == 1
    # both are solutions of A_c X = B ...
    res = torch.empty_like(x)
    for name, sol in (("adapter", got["CG_X"]), ("c-abi", Xd.cpu().numpy())):
        ctx.apply_T(dev(sol), res)
        r = np.abs(res.cpu().numpy() - Bn).max() / np.abs(Bn).max()
        assert r < 1e-9, (name, r)
    # ... and the same one
    assert rel_err(got["CG_X"], Xd.cpu().numpy()) < 1e-8
    # rhs_linear + RK4
    w = lpf.wave_params()
    ctx.rhs_setup(lpf.make_rhs_params(w, tau=w["T"] / 150, rel_tol=1e-12, max_iter=2000))
    ph = -w["k"] * sp.surf_xy[:, 0]
    st = np.concatenate([0.5 * w["H"] * np.cos(ph), -0.5 * w["H"] * w["cwave"] / np.tanh(w["kh"]) * np.sin(ph)])
    sd, kd = dev(st), torch.empty(2 * sp.nsurf, dtype=torch.float64, device="cuda")
    ctx.rhs(0.0, sd, kd)
    assert rel_err(got["rhs"], kd.cpu().numpy()) < 1e-10
    t = 0.0
    for _ in range(2):
        t = ctx.rk4_step(sd, t, w["T"] / 150)
    assert rel_err(got["state_after_2_steps"], sd.cpu().numpy()) < 1e-10 and abs(got["t"][0] - t) < 1e-15
    ctx.close()
