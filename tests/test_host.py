"""Host-side logic and the C-ABI surface, no GPU needed: the library loads, exports every declared symbol,
its mini-FEM numbering is equivalent to the oracle's independent algorithm, and the partition / halo plan
reproduces the serial operator when ranks exchange partial sums (gloo, world_size 2)."""
import os
import re
import socket
import sys

import numpy as np
import pytest

from util import oracle_space_from, rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
H = 1.0 / (2.0 * np.pi)
REF_MESHES = "/root/reference/Meshes"


def test_library_exports_every_declared_symbol(lpf):
    hdr = open(os.path.join(ROOT, "include", "lpf_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lpf_[A-Za-z0-9_]+)\s*\(", hdr))
    assert len(declared) > 50
    for name in sorted(declared):
        assert hasattr(lpf.lib, name), f"liblpf_b200.so does not export {name}"
    assert declared == set(lpf.SIGNATURES), declared ^ set(lpf.SIGNATURES)
    assert lpf.lib.lpf_version() == 200


def test_error_reporting_without_aborting(lpf):
    with pytest.raises(lpf.LpfError, match="cannot open"):
        lpf.Mesh.read("/nonexistent/file.mesh")
    with pytest.raises(lpf.LpfError):
        lpf.Mesh.wave_tank(0, 1, 1)
    with pytest.raises(lpf.LpfError, match="nx >= 3"):
        lpf.Mesh.wave_tank(2, 1, 1, periodic_x=True)
    m = lpf.Mesh.wave_tank(3, 1, 1)
    with pytest.raises(lpf.LpfError, match="order"):
        lpf.Space(m, 11)
    with pytest.raises(lpf.LpfError):
        lpf.Space(m, 2, nranks=2, rank=5)


def test_device_entry_points_fail_loudly_without_gpu(lpf):
    """No CPU fallback: with no CUDA device lpf_create must return NULL and say why."""
    if lpf.lib.lpf_device_count() > 0:
        pytest.skip("a GPU is present")
    sp = lpf.Space(lpf.Mesh.wave_tank(3, 1, 1), 2)
    with pytest.raises(lpf.LpfError, match="(?i)cuda|device"):
        lpf.Context(sp)


def test_basis_tables_match_oracle(lpf, orc):
    for p in range(1, 11):
        t = lpf.basis_tables(p)
        bs = orc.make_basis(p)
        for k, ref in (("nodes", bs.nodes), ("qpts", bs.qpts), ("qwts", bs.qwts), ("B", bs.B), ("G", bs.G), ("Dhat", bs.Dhat)):
            assert np.abs(t[k] - ref).max() < 2e-13 * max(1.0, np.abs(ref).max()), (p, k)


def _check_numbering(lpf, orc, mesh, omesh, p):
    """Same partition of (element, local node) pairs into dofs as the oracle's vertex-weight signatures."""
    sp = lpf.Space(mesh, p)
    os_ = orc.build_h1_space(omesh, p)
    assert (sp.ne, sp.ndof, len(sp.ess)) == (omesh.ne, os_.ndof, len(os_.ess))
    key = lambda c: np.round(c.mean(axis=1) * 1e9).astype(np.int64)
    ka, kb = key(sp.corners), key(omesh.corners)
    ia, ib = np.lexsort(ka.T), np.lexsort(kb.T)
    assert np.all(ka[ia] == kb[ib])
    perm = np.empty(sp.ne, dtype=int); perm[ia] = ib
    assert np.allclose(sp.corners, omesh.corners[perm], atol=1e-13)
    go = os_.gather[perm]
    h2o = -np.ones(sp.ndof, dtype=np.int64)
    h2o[sp.gather.reshape(-1)] = go.reshape(-1)
    assert np.all(h2o[sp.gather] == go), "host numbering merges dofs the oracle keeps apart"
    assert len(np.unique(h2o)) == sp.ndof, "host numbering splits dofs the oracle merges"
    assert set(h2o[sp.ess].tolist()) == set(os_.ess.tolist())
    assert set(h2o[sp.surf2vol].tolist()) == set(os_.surf2vol.tolist())
    assert set(perm[sp.surf_elems].tolist()) == set(orc.surface_elements(os_).tolist())
    lo, hi = omesh.bounding_box()
    dd = np.abs(sp.node_coordinates() - os_.xyz[h2o])
    dd[:, 0] = np.minimum(dd[:, 0], np.abs(dd[:, 0] - (hi[0] - lo[0])))     # periodic seam copies
    assert dd.max() < 1e-12
    dd = np.abs(sp.surf_xy - os_.xyz[h2o[sp.surf2vol]][:, :2])
    dd[:, 0] = np.minimum(dd[:, 0], np.abs(dd[:, 0] - (hi[0] - lo[0])))
    assert dd.max() < 1e-12
    return sp


@pytest.mark.parametrize("p", [1, 2, 3, 4, 6])
def test_numbering_periodic_tank(lpf, orc, p):
    _check_numbering(lpf, orc, lpf.Mesh.wave_tank(3, 1, 1), orc.make_wave_tank(3, 1, 1, 1.0, 0.1, H, True), p)


def test_numbering_refined_and_finite(lpf, orc):
    om = orc.uniform_refine(orc.uniform_refine(orc.make_wave_tank(3, 1, 1, 1.0, 0.1, H, True)))
    sp = _check_numbering(lpf, orc, lpf.Mesh.wave_tank(3, 1, 1).refine(2), om, 3)
    assert sp.ndof == 6084
    om = orc.uniform_refine(orc.make_wave_tank(6, 1, 2, 12.0, 1.0, H, False))
    _check_numbering(lpf, orc, lpf.Mesh.wave_tank(6, 1, 2, 12.0, 1.0, H, periodic_x=False).refine(1), om, 2)


def test_big8_counts(lpf):
    """wave-tank-big8.mesh family: global dofs (128p)(2p+1)(16p+1), surface (128p)(2p+1)  (SURVEY 8)."""
    m = lpf.Mesh.wave_tank(128, 2, 16)
    assert m.ne == 4096 and m.nv == 128 * 3 * 17
    for p in (1, 2, 4):
        sp = lpf.Space(m, p)
        assert sp.ndof == 128 * p * (2 * p + 1) * (16 * p + 1)
        assert sp.nsurf == 128 * p * (2 * p + 1)
    assert lpf.Space(m, 4).ndof == 299520


def test_cylinder_mesh_reader(lpf, orc, tmp_path):
    path = os.path.join(HERE, "meshes", "cylinder_half.mesh")
    m = lpf.Mesh.read(path)
    om = orc.read_mfem_mesh(path)
    assert m.ne == 3192 and m.nv == 4290
    sp = _check_numbering(lpf, orc, m, om, 2)
    assert sp.nsurf == len(np.unique(sp.surf2vol))
    A = orc.PAOperator(oracle_space_from(orc, sp))
    assert A.detJ.min() > 0                                    # all hexes positively oriented


@pytest.mark.skipif(not os.path.isdir(REF_MESHES), reason="reference meshes not on this machine")
@pytest.mark.parametrize("name,p", [("wave-tank.mesh", 4), ("wave-tank-finite.mesh", 3), ("wave-tank-big.mesh", 2)])
def test_reads_reference_mfem_meshes(lpf, orc, name, p):
    path = os.path.join(REF_MESHES, name)
    _check_numbering(lpf, orc, lpf.Mesh.read(path), orc.read_mfem_mesh(path), p)


@pytest.mark.skipif(not os.path.isdir(REF_MESHES), reason="reference meshes not on this machine")
def test_reads_reference_gmsh_mesh_and_generator_equals_file(lpf, orc):
    m = lpf.Mesh.read(os.path.join(REF_MESHES, "mesh_cylinder_half.msh"))
    assert (m.ne, m.nv) == (3192, 4290)
    bdr, attr = m.boundary()
    assert (attr == 2).sum() == 798 and (attr == 3).sum() == 112
    # our generator reproduces wave-tank-big.mesh (Meshes/wave_tank.cpp with 32x2x8)
    f = lpf.Mesh.read(os.path.join(REF_MESHES, "wave-tank-big.mesh"))
    g = lpf.Mesh.wave_tank(32, 2, 8)
    assert f.ne == g.ne == 512 and f.nv == g.nv
    kf = np.sort(np.round(f.corners().mean(axis=1) * 1e9).astype(np.int64), axis=0)
    kg = np.sort(np.round(g.corners().mean(axis=1) * 1e9).astype(np.int64), axis=0)
    assert np.all(kf == kg)
    assert lpf.Space(f, 3).ndof == lpf.Space(g, 3).ndof == 16800


def test_mfem_writer_reader_round_trip(lpf, orc, tmp_path):
    om = orc.perturb_mesh(orc.make_wave_tank(4, 2, 2, 1.0, 0.1, H, True), 0.1)
    path = str(tmp_path / "t.mesh")
    orc.write_mfem_mesh(om, path)                     # periodic -> nodes block
    assert "L2_T1_3D_P1" in open(path).read()
    _check_numbering(lpf, orc, lpf.Mesh.read(path), om, 3)
    of = orc.make_wave_tank(3, 2, 2, 2.0, 1.0, 0.5, False)
    orc.write_mfem_mesh(of, path)                     # plain vertices
    assert "L2_T1_3D_P1" not in open(path).read()
    _check_numbering(lpf, orc, lpf.Mesh.read(path), of, 2)


@pytest.mark.parametrize("nranks", [2, 3, 4, 8])
def test_partition_invariants(lpf, nranks):
    m = lpf.Mesh.wave_tank(16, 2, 4)
    p = 2
    serial = lpf.Space(m, p)
    parts = [lpf.Space(m, p, nranks=nranks, rank=r) for r in range(nranks)]
    assert sum(s.ne for s in parts) == serial.ne
    assert sum(int(s.owned.sum()) for s in parts) == serial.ndof          # every true dof owned exactly once
    assert sum(int(s.surf_owned.sum()) for s in parts) == serial.nsurf
    assert all(s.n_true_global == serial.ndof for s in parts)
    # x-slabs of a periodic ring: 2 neighbours each (1 if only two ranks)
    for s in parts:
        assert len(s.nbr_rank) == (1 if nranks == 2 else 2)
    # pairwise send lists describe the same global dofs in the same order
    for a in range(nranks):
        for ia, b in enumerate(parts[a].nbr_rank):
            sa = parts[a].send_dofs[parts[a].nbr_offset[ia]:parts[a].nbr_offset[ia + 1]]
            ib = list(parts[b].nbr_rank).index(a)
            sb = parts[b].send_dofs[parts[b].nbr_offset[ib]:parts[b].nbr_offset[ib + 1]]
            assert np.array_equal(parts[a].l2g[sa], parts[b].l2g[sb])
    # global multiplicities of surface dofs
    cnt = np.bincount(serial.gather.reshape(-1), minlength=serial.ndof)
    for s in parts:
        assert np.array_equal(s.surf_mult, cnt[s.l2g[s.surf2vol]])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, q):
    import importlib
    import torch.distributed as dist
    import torch
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, HERE)
    lpf = importlib.import_module("master-thesis-lpf-in-mfem_b200")
    import lpf_oracle as orc
    from util import oracle_space_from as osf
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        m = lpf.Mesh.wave_tank(8, 2, 3).perturb(0.1)
        p = 3
        sp = lpf.Space(m, p, nranks=world, rank=rank)
        A = orc.PAOperator(osf(orc, sp))
        xg = orc.hash_noise(sp.n_true_global)
        y = A.mult(xg[sp.l2g])                                  # local partial sums
        # halo-sum exactly as the device path does it: pack, exchange with each neighbour, rank-ordered add
        send = torch.from_numpy(y[sp.send_dofs].copy())
        recv = torch.zeros_like(send)
        reqs = []
        for i, nb in enumerate(sp.nbr_rank):
            a, b = int(sp.nbr_offset[i]), int(sp.nbr_offset[i + 1])
            reqs.append(dist.isend(send[a:b], int(nb)))
            reqs.append(dist.irecv(recv[a:b], int(nb)))
        for r in reqs:
            r.wait()
        recv = recv.numpy()
        for i, dof in enumerate(sp.shared_dofs):
            s = 0.0
            own = y[dof]
            for j in range(sp.red_off[i], sp.red_off[i + 1]):
                s += own if sp.red_src[j] < 0 else recv[sp.red_src[j]]
            y[dof] = s
        # owner-masked dot product + all-reduce == serial dot product
        loc = torch.tensor([float(np.dot(y * sp.owned, xg[sp.l2g]))], dtype=torch.float64)
        dist.all_reduce(loc)
        q.put((rank, sp.l2g.copy(), y, float(loc[0])))
    finally:
        dist.destroy_process_group()


def test_two_rank_halo_sum_reproduces_serial_operator(lpf, orc):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    m = lpf.Mesh.wave_tank(8, 2, 3).perturb(0.1)
    sp = lpf.Space(m, 3)
    A = orc.PAOperator(oracle_space_from(orc, sp))
    xg = orc.hash_noise(sp.ndof)
    yref = A.mult(xg)
    copies = {}
    for rank, l2g, y, dotv in res:
        assert rel_err(y, yref[l2g]) < 1e-13                     # every sharer ends with the full sum
        assert abs(dotv - float(np.dot(yref, xg))) < 1e-12 * abs(float(np.dot(yref, xg)))
        copies[rank] = dict(zip(l2g.tolist(), y.tolist()))
    shared = set(copies[0]) & set(copies[1])
    assert len(shared) > 0
    assert all(copies[0][g] == copies[1][g] for g in shared)     # bit-identical copies on both sharers


@pytest.mark.parametrize("nranks", [1, 3])
def test_cylinder_rim_extraction(lpf, nranks):
    """cylinder-diffraction.cpp:449-505: vertices shared by the wall (attr 3) and the free surface (attr 2), r = a,
    theta = atan2 >= 0; gathered over the ranks, sorted, duplicates dropped."""
    m = lpf.Mesh.read(os.path.join(HERE, "meshes", "cylinder_half.mesh"))
    bdr, attr = m.boundary()
    wall = set(bdr[attr == 3].reshape(-1).tolist()) & set(bdr[attr == 2].reshape(-1).tolist())
    pts = []
    for r in range(nranks):
        sp = lpf.Space(m, 2, nranks=nranks, rank=r)
        th, idx = sp.rim(3, 4.0, 4.0, 0.5, 5e-3)
        assert np.all(th >= 0) and np.all(th <= np.pi + 1e-12)
        if len(th) == 0:                       # a part may hold no rim vertex at all
            continue
        xy = sp.surf_xy[idx]
        assert np.abs(np.hypot(xy[:, 0] - 4.0, xy[:, 1] - 4.0) - 0.5).max() < 5e-3
        assert np.allclose(np.arctan2(xy[:, 1] - 4.0, xy[:, 0] - 4.0), th)
        pts += list(th)
    th = np.sort(np.array(pts))
    th = th[np.concatenate([[True], np.diff(th) > 1e-10])]
    assert len(th) == len(wall) >= 8           # every wall/free-surface vertex of the half cylinder, once
    assert th[0] < 1e-9 and abs(th[-1] - np.pi) < 1e-9


@pytest.mark.parametrize("nranks", [1, 2])
def test_surface_quads_cover_the_surface_space(lpf, nranks):
    m = lpf.Mesh.wave_tank(8, 2, 2)
    p = 3
    nf = 0
    for r in range(nranks):
        sp = lpf.Space(m, p, nranks=nranks, rank=r)
        q = sp.surface_quads()
        nf += len(q)
        assert q.shape[1] == (p + 1) ** 2 and q.min() >= 0 and q.max() < sp.nsurf
        assert set(q.reshape(-1).tolist()) == set(range(sp.nsurf))
        # each quad is a tensor lattice: its nodes sit on the GLL points of its bounding rectangle
        xy = sp.surf_xy[q]                               # [nf][D*D][2]
        t = lpf.basis_tables(p)["nodes"]
        for f in range(len(q)):
            lo, hi = xy[f].min(axis=0), xy[f].max(axis=0)
            ex = lo[0] + (hi[0] - lo[0]) * t
            assert np.allclose(np.sort(np.unique(np.round(xy[f][:, 0], 12))), np.round(ex, 12), atol=1e-10) or hi[0] - lo[0] > 0.5
    assert nf == 16


def test_surface_vtu_writer(lpf, tmp_path):
    """f3: ParaView piece of the free-surface fields: Lagrange quads of order p in VTK node order, point data;
    cells own their nodes, so the periodic seam cell keeps its true coordinates."""
    import xml.etree.ElementTree as ET
    m = lpf.Mesh.wave_tank(4, 2, 2)
    for p, ho in ((3, True), (2, False)):
        sp = lpf.Space(m, p)
        eta = np.cos(2 * np.pi * sp.surf_xy[:, 0])
        path = tmp_path / f"s{p}.vtu"
        sp.write_surface_vtu(path, {"eta": eta, "phi_fs": 2 * eta}, high_order=ho)
        piece = ET.parse(path).getroot().find("UnstructuredGrid/Piece")
        ncell = 8 if ho else 8 * p * p
        assert int(piece.get("NumberOfPoints")) == 8 * (p + 1) ** 2 and int(piece.get("NumberOfCells")) == ncell
        pts = np.array(piece.find("Points/DataArray").text.split(), dtype=float).reshape(-1, 3)
        assert np.allclose(pts[:, 2], H) and pts[:, 0].max() == 1.0 and pts[:, 0].min() == 0.0
        arr = {d.get("Name"): d for d in piece.findall("Cells/DataArray")}
        conn = np.array(arr["connectivity"].text.split(), dtype=int).reshape(ncell, -1)
        types = np.array(arr["types"].text.split(), dtype=int)
        assert np.all(types == (70 if ho else 9)) and conn.shape[1] == ((p + 1) ** 2 if ho else 4)
        c = pts[conn[:, :4], :2]                       # the first four nodes of a cell are its corners, in cyclic order
        assert np.allclose(c[:, 0] + c[:, 2], c[:, 1] + c[:, 3])
        d1, d2 = c[:, 2] - c[:, 0], c[:, 3] - c[:, 1]
        area = 0.5 * np.abs(d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0])
        assert np.isclose(area.sum(), 1.0 * 0.1)       # seam cell included: total = Lx * Ly
        if ho:                                         # edge nodes 4 .. 4+p-2 lie on the edge between corners 0 and 1
            e = pts[conn[:, 4:4 + p - 1], :2]
            for k in range(p - 1):
                a0, a1 = e[:, k] - c[:, 0], c[:, 1] - c[:, 0]
                assert np.allclose(a0[:, 0] * a1[:, 1] - a0[:, 1] * a1[:, 0], 0.0, atol=1e-14)
                assert np.all((a0 * a1).sum(axis=1) > 0) and np.all((a0 * a1).sum(axis=1) < (a1 * a1).sum(axis=1))
        pd = {d.get("Name"): np.array(d.text.split(), dtype=float) for d in piece.findall("PointData/DataArray")}
        # periodic field sampled at the cell's own coordinates
        assert np.allclose(pd["eta"], np.cos(2 * np.pi * pts[:, 0])) and np.allclose(pd["phi_fs"], 2 * pd["eta"])


@pytest.mark.parametrize("family", ["tank_r2_p3", "finite_r1_p5", "big8_p2", "cylinder_p3"])
def test_geometric_dof_map_between_product_and_oracle_numberings(lpf, orc, family):
    """tests/util.py dof_map relates the product's numbering to the oracle's by geometry alone; it must be a bijection that
    maps essential dofs onto essential dofs and surface dofs onto surface dofs on every mesh family of BASELINE.json."""
    from util import dof_map, surface_map
    H0 = 1.0 / (2.0 * np.pi)
    cyl = os.path.join(os.path.dirname(os.path.abspath(__file__)), "meshes", "cylinder_half.mesh")
    mk = {"tank_r2_p3": (lambda: lpf.Mesh.wave_tank(3, 1, 1).refine(2), lambda: orc.uniform_refine(orc.uniform_refine(orc.make_wave_tank(3, 1, 1, 1.0, 0.1, H0, True))), 3),
          "finite_r1_p5": (lambda: lpf.Mesh.wave_tank(36, 1, 1, 12.0, 1.0, H0, False).refine(1), lambda: orc.uniform_refine(orc.make_wave_tank(36, 1, 1, 12.0, 1.0, H0, False)), 5),
          "big8_p2": (lambda: lpf.Mesh.wave_tank(128, 2, 16), lambda: orc.make_wave_tank(128, 2, 16, 1.0, 0.1, H0, True), 2),
          "cylinder_p3": (lambda: lpf.Mesh.read(cyl), lambda: orc.read_mfem_mesh(cyl), 3)}[family]
    sp, osp = lpf.Space(mk[0](), mk[2]), orc.build_h1_space(mk[1](), mk[2])
    pm = dof_map(sp, osp)
    sm = surface_map(sp, osp, pm)
    assert set(pm[sp.ess].tolist()) == set(osp.ess.tolist())
    # same physical points (x compared on the circle: a seam dof of the x-periodic tanks may be represented by x = 0 or x = Lx)
    Lx = osp.mesh.bounding_box()[1][0] - osp.mesh.bounding_box()[0][0]
    a, b = sp.surf_xy, osp.surf_xy[sm][:, :2]
    assert np.abs(a[:, 1] - b[:, 1]).max() < 1e-12
    assert np.abs(np.exp(2j * np.pi * a[:, 0] / Lx) - np.exp(2j * np.pi * b[:, 0] / Lx)).max() < 1e-10
