"""CPU oracle for the LPF Laplace hot path -- TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the algorithm the reference
(hirschjulien/Master-Thesis-LPF-in-MFEM) executes on its hot path.  It is the
checker for the CUDA product: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  The product
package never imports anything from oracle/.

PARITY UNPINNED at the MFEM boundary: the arithmetic of the path lives in MFEM
(mfem.org, version not pinned by the reference; "2010-2025" copyright in
Solvers/makefile:1 => MFEM >= 4.8), which is absent from /root/reference and
from this image, and the reference ships no golden vectors.
The oracle is therefore pinned by (i) two independent operator
implementations (sum-factorised partial assembly vs dense element-matrix full
assembly) that must agree to 1e-12, (ii) the reference's own analytic
known-answer drivers (Airy wave; Convergence_and_Scaling/laplace-parallel-pconv.cpp:80-87,197-213,
convergence-parallel-partial.cpp:219-227,288-305, Solvers/laplace_solver.cpp:70-81,126-138),
and (iii) the scalar known answers of SURVEY.md 8c.

Reference call sites restated here (file:line relative to /root/reference):
  Solvers/PF_linear_par_partial.cpp:118-124   PA assemble + Jacobi   -> pa_setup, pa_diag
  Solvers/PF_linear_par_partial.cpp:130-244   rhs_linear::Mult       -> RhsLinear.mult
  Solvers/PF_linear_par_partial.cpp:155       FormLinearSystem       -> form_linear_system
  Solvers/PF_linear_par_partial.cpp:157-164   CGSolver               -> pcg
  Solvers/PF_linear_par_partial.cpp:169       GetDerivative(1,2,w)   -> get_derivative_z
  Solvers/PF_linear_par_partial.cpp:415-447   C_gen / C_abs          -> relax_cgen, relax_cabs
  Solvers/PF_linear_par_partial.cpp:472-494   RK4Solver              -> rk4_step
  Convergence_and_Scaling/ss.cpp:62-105       RHS w/o relaxation     -> RhsLinear(relax=None)
  Meshes/wave_tank.cpp:13-47, Meshes/wave-tank-finite.cpp:10-45      -> make_wave_tank
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

# lexicographic corner c = cx + 2 cy + 4 cz  <->  MFEM/Gmsh hex vertex number
LEX2MFEM = np.array([0, 1, 3, 2, 4, 5, 7, 6])
# the 6 faces of a hex as lexicographic corners in cyclic order
HEX_FACES_LEX = np.array([
    [0, 1, 3, 2],  # z = 0
    [4, 5, 7, 6],  # z = 1
    [0, 1, 5, 4],  # y = 0
    [2, 3, 7, 6],  # y = 1
    [0, 2, 6, 4],  # x = 0
    [1, 3, 7, 5],  # x = 1
])


# --------------------------------------------------------------------------
# 1-D bases and quadrature ([MFEM] fem/intrules.cpp,
# fem/fe/fe_h1.cpp: Gauss-Lobatto nodal basis, Gauss-Legendre rule p+2 points)
# --------------------------------------------------------------------------
def _legendre(n, x):
    """P_n(x), P_n'(x) on [-1,1] by the three-term recurrence."""
    x = np.asarray(x, dtype=np.float64)
    p0 = np.ones_like(x)
    if n == 0:
        return p0, np.zeros_like(x)
    p1 = x.copy()
    for k in range(2, n + 1):
        p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
    dp = n * (x * p1 - p0) / (x * x - 1.0) if n > 0 else np.zeros_like(x)
    return p1, dp


def gauss_legendre(n):
    """n-point Gauss-Legendre rule on [0,1]; weights sum to 1."""
    k = np.arange(1, n + 1)
    x = -np.cos((2 * k - 1) * np.pi / (2 * n))
    for _ in range(100):
        p, dp = _legendre(n, x)
        dx = p / dp
        x = x - dx
        if np.max(np.abs(dx)) < 1e-16:
            break
    p, dp = _legendre(n, x)
    w = 2.0 / ((1.0 - x * x) * dp * dp)
    x = 0.5 * (x + 1.0)
    w = 0.5 * w
    # symmetrise
    x = 0.5 * (x + (1.0 - x[::-1]))
    w = 0.5 * (w + w[::-1])
    return x, w


def gll_points(p):
    """p+1 Gauss-Lobatto-Legendre points on [0,1] (H1_FECollection default)."""
    if p == 0:
        return np.array([0.5])
    if p == 1:
        return np.array([0.0, 1.0])
    n = p
    # interior nodes: roots of P_n'(x); Chebyshev-Gauss-Lobatto initial guess
    k = np.arange(1, n)
    x = -np.cos(np.pi * k / n)
    for _ in range(100):
        pn, dpn = _legendre(n, x)
        # (1-x^2) P_n'' = 2x P_n' - n(n+1) P_n
        d2 = (2 * x * dpn - n * (n + 1) * pn) / (1.0 - x * x)
        dx = dpn / d2
        x = x - dx
        if np.max(np.abs(dx)) < 1e-16:
            break
    x = np.concatenate([[-1.0], x, [1.0]])
    x = 0.5 * (x + 1.0)
    x = 0.5 * (x + (1.0 - x[::-1]))
    return x


def lagrange_eval(nodes, pts):
    """B[q,d] = l_d(pts[q]), G[q,d] = l_d'(pts[q]) for the Lagrange basis on `nodes`."""
    nodes = np.asarray(nodes, dtype=np.float64)
    pts = np.asarray(pts, dtype=np.float64)
    n = len(nodes)
    B = np.ones((len(pts), n))
    G = np.zeros((len(pts), n))
    for d in range(n):
        others = [m for m in range(n) if m != d]
        denom = np.prod([nodes[d] - nodes[m] for m in others]) if others else 1.0
        for qi, xq in enumerate(pts):
            B[qi, d] = np.prod([xq - nodes[m] for m in others]) / denom if others else 1.0
            s = 0.0
            for a in others:
                s += np.prod([xq - nodes[m] for m in others if m != a]) if len(others) > 1 else 1.0
            G[qi, d] = s / denom
    return B, G


@dataclass
class Basis:
    """1-D tables of an H1 order-p hex element with the DiffusionIntegrator rule."""
    p: int
    D: int
    Q: int
    nodes: np.ndarray   # GLL nodes [D]
    qpts: np.ndarray    # Gauss-Legendre points [Q]
    qwts: np.ndarray    # weights [Q]
    B: np.ndarray       # [Q, D]
    G: np.ndarray       # [Q, D]
    Dhat: np.ndarray    # collocation derivative [D, D]: Dhat[a, m] = l_m'(x_a)


def make_basis(p, q_extra=2):
    D, Q = p + 1, p + q_extra
    nodes = gll_points(p)
    qp, qw = gauss_legendre(Q)
    B, G = lagrange_eval(nodes, qp)
    _, Dhat = lagrange_eval(nodes, nodes)
    return Basis(p, D, Q, nodes, qp, qw, B, G, Dhat)


# --------------------------------------------------------------------------
# Mesh (hexes only), readers, generators, refinement  
# --------------------------------------------------------------------------
@dataclass
class HexMesh:
    """Conforming hex mesh.  Topology through vertex ids (MFEM order per
    element); geometry as per-element trilinear corners in LEXICOGRAPHIC corner
    order (covers both plain-vertex meshes and periodic meshes that carry an
    L2_T1_3D_P1 `nodes` block)."""
    elems: np.ndarray          # [NE, 8] vertex ids, MFEM ordering
    corners: np.ndarray        # [NE, 8, 3] corner coordinates, lexicographic
    bdr: np.ndarray            # [NB, 4] vertex ids (cyclic)
    bdr_attr: np.ndarray       # [NB]
    nv: int

    @property
    def ne(self):
        return self.elems.shape[0]

    def bounding_box(self):
        c = self.corners.reshape(-1, 3)
        return c.min(axis=0), c.max(axis=0)


def read_mfem_mesh(path):
    """MFEM mesh v1.0 reader (hex + quad boundary; vertices or L2_T1_3D_P1 nodes)."""
    with open(path) as f:
        lines = [ln.strip() for ln in f]
    lines = [ln for ln in lines if ln and not ln.startswith('#')]
    assert lines[0].startswith('MFEM mesh v1.0'), lines[0]
    i = 1
    elems = bdr = battr = None
    nv = 0
    vcoords = None
    nodes = None
    while i < len(lines):
        key = lines[i]
        if key == 'dimension':
            assert int(lines[i + 1]) == 3
            i += 2
        elif key == 'elements':
            n = int(lines[i + 1])
            rows = np.array([[int(t) for t in lines[i + 2 + j].split()] for j in range(n)])
            assert np.all(rows[:, 1] == 5), 'hexahedra only'
            elems = rows[:, 2:10]
            i += 2 + n
        elif key == 'boundary':
            n = int(lines[i + 1])
            rows = np.array([[int(t) for t in lines[i + 2 + j].split()] for j in range(n)])
            assert np.all(rows[:, 1] == 3), 'quad boundary only'
            battr = rows[:, 0]
            bdr = rows[:, 2:6]
            i += 2 + n
        elif key == 'vertices':
            nv = int(lines[i + 1])
            i += 2
            if i < len(lines) and re.fullmatch(r'\d+', lines[i]):
                vdim = int(lines[i])
                vcoords = np.array([[float(t) for t in lines[i + 1 + j].split()] for j in range(nv)])
                assert vdim == 3
                i += 1 + nv
        elif key == 'nodes':
            assert lines[i + 1] == 'FiniteElementSpace'
            fec = lines[i + 2].split(':')[1].strip()
            vdim = int(lines[i + 3].split(':')[1])
            ordering = int(lines[i + 4].split(':')[1])
            assert fec == 'L2_T1_3D_P1' and vdim == 3 and ordering == 1, (fec, vdim, ordering)
            vals = np.array([float(t) for ln in lines[i + 5:] for t in ln.split()])
            nodes = vals.reshape(elems.shape[0], 8, 3)   # lexicographic, xyz interleaved
            i = len(lines)
        else:
            raise ValueError('unknown section ' + key)
    if nodes is None:
        nodes = vcoords[elems[:, LEX2MFEM]]
    return HexMesh(elems, nodes, bdr, battr, nv)


def read_gmsh22(path):
    """Gmsh 2.2 ASCII reader: hexes (type 5), quads (type 3, attr = physical tag)."""
    with open(path) as f:
        lines = [ln.strip() for ln in f]
    i = lines.index('$Nodes')
    nn = int(lines[i + 1])
    ids = np.zeros(nn, dtype=np.int64)
    xyz = np.zeros((nn, 3))
    for j in range(nn):
        t = lines[i + 2 + j].split()
        ids[j] = int(t[0])
        xyz[j] = [float(t[1]), float(t[2]), float(t[3])]
    remap = {int(g): j for j, g in enumerate(ids)}
    i = lines.index('$Elements')
    nel = int(lines[i + 1])
    hexes, quads, qattr = [], [], []
    for j in range(nel):
        t = [int(s) for s in lines[i + 2 + j].split()]
        etype, ntags = t[1], t[2]
        phys = t[3]
        nod = t[3 + ntags:]
        if etype == 5:
            hexes.append([remap[n] for n in nod])
        elif etype == 3:
            quads.append([remap[n] for n in nod])
            qattr.append(phys)
    elems = np.array(hexes)
    corners = xyz[elems[:, LEX2MFEM]]
    return HexMesh(elems, corners, np.array(quads), np.array(qattr), nn)


def make_wave_tank(nx, ny, nz, Lx, Ly, H, periodic_x):
    """Cartesian box of hexes with the reference's boundary attributes
    (Meshes/wave_tank.cpp:13-47: bottom=1, top=2, ymin=3, ymax=4, x periodic;
    Meshes/wave-tank-finite.cpp:10-45: additionally x=xmax ->5, x=xmin ->6)."""
    nvx = nx if periodic_x else nx + 1

    def vid(i, j, k):
        return (i % nvx if periodic_x else i) + nvx * (j + (ny + 1) * k)

    elems, corners = [], []
    xs = np.linspace(0.0, Lx, nx + 1)
    ys = np.linspace(0.0, Ly, ny + 1)
    zs = np.linspace(0.0, H, nz + 1)
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                lex = [vid(i + a, j + b, k + c) for c in (0, 1) for b in (0, 1) for a in (0, 1)]
                mf = [0] * 8
                for l in range(8):
                    mf[LEX2MFEM[l]] = lex[l]
                elems.append(mf)
                corners.append([[xs[i + a], ys[j + b], zs[k + c]]
                                for c in (0, 1) for b in (0, 1) for a in (0, 1)])
    bdr, battr = [], []
    for j in range(ny):
        for i in range(nx):
            bdr.append([vid(i, j, 0), vid(i, j + 1, 0), vid(i + 1, j + 1, 0), vid(i + 1, j, 0)]); battr.append(1)
            bdr.append([vid(i, j, nz), vid(i + 1, j, nz), vid(i + 1, j + 1, nz), vid(i, j + 1, nz)]); battr.append(2)
    for k in range(nz):
        for i in range(nx):
            bdr.append([vid(i, 0, k), vid(i + 1, 0, k), vid(i + 1, 0, k + 1), vid(i, 0, k + 1)]); battr.append(3)
            bdr.append([vid(i, ny, k), vid(i, ny, k + 1), vid(i + 1, ny, k + 1), vid(i + 1, ny, k)]); battr.append(4)
    if not periodic_x:
        for k in range(nz):
            for j in range(ny):
                bdr.append([vid(nx, j, k), vid(nx, j + 1, k), vid(nx, j + 1, k + 1), vid(nx, j, k + 1)]); battr.append(5)
                bdr.append([vid(0, j, k), vid(0, j, k + 1), vid(0, j + 1, k + 1), vid(0, j + 1, k)]); battr.append(6)
    nv = nvx * (ny + 1) * (nz + 1)
    return HexMesh(np.array(elems), np.array(corners), np.array(bdr), np.array(battr), nv)


def _trilinear(corners, xi):
    """corners [...,8,3] lexicographic, xi [n,3] in [0,1]^3 -> [..., n, 3]."""
    x, y, z = xi[:, 0], xi[:, 1], xi[:, 2]
    N = np.stack([(1 - x) * (1 - y) * (1 - z), x * (1 - y) * (1 - z), (1 - x) * y * (1 - z), x * y * (1 - z),
                  (1 - x) * (1 - y) * z, x * (1 - y) * z, (1 - x) * y * z, x * y * z], axis=1)  # [n,8]
    return np.einsum('nc,...cd->...nd', N, corners)


def uniform_refine(mesh: HexMesh) -> HexMesh:
    """Topological uniform refinement (each hex -> 8), geometry by the trilinear
    map of the parent ([MFEM] Mesh::UniformRefinement; used at
    Solvers/PF_linear_par_partial.cpp:266,271)."""
    ne = mesh.ne
    ev = mesh.elems[:, LEX2MFEM]          # [NE,8] vertex ids in lexicographic corner order
    new_id = {}
    nv = mesh.nv

    def get(key):
        nonlocal nv
        r = new_id.get(key)
        if r is None:
            r = nv
            new_id[key] = r
            nv += 1
        return r

    elems, corners = [], []
    lat = np.array([[a, b, c] for c in range(3) for b in range(3) for a in range(3)], dtype=np.float64) / 2.0
    for e in range(ne):
        v = ev[e]
        # 27 lattice points: id by the set of parent corners whose "span" contains the point
        pid = np.zeros((3, 3, 3), dtype=np.int64)
        for c in range(3):
            for b in range(3):
                for a in range(3):
                    sel = [v[(ca) + 2 * (cb) + 4 * (cc)]
                           for cc in ((0, 1) if c == 1 else (c // 2,))
                           for cb in ((0, 1) if b == 1 else (b // 2,))
                           for ca in ((0, 1) if a == 1 else (a // 2,))]
                    if len(sel) == 1:
                        pid[a, b, c] = sel[0]
                    elif len(sel) == 8:
                        pid[a, b, c] = get(('c', e))
                    else:
                        pid[a, b, c] = get(tuple(sorted(sel)))
        X = _trilinear(mesh.corners[e], lat).reshape(3, 3, 3, 3)  # [c][b][a][xyz]
        for c in range(2):
            for b in range(2):
                for a in range(2):
                    lex = [pid[a + da, b + db, c + dc] for dc in (0, 1) for db in (0, 1) for da in (0, 1)]
                    mf = [0] * 8
                    for l in range(8):
                        mf[LEX2MFEM[l]] = lex[l]
                    elems.append(mf)
                    corners.append([X[c + dc, b + db, a + da] for dc in (0, 1) for db in (0, 1) for da in (0, 1)])
    bdr, battr = [], []
    for f in range(mesh.bdr.shape[0]):
        q = mesh.bdr[f]
        m = [get(tuple(sorted((int(q[i]), int(q[(i + 1) % 4]))))) for i in range(4)]
        ctr = get(tuple(sorted(int(t) for t in q)))
        for i in range(4):
            bdr.append([int(q[i]), m[i], ctr, m[(i + 3) % 4]])
            battr.append(mesh.bdr_attr[f])
    return HexMesh(np.array(elems), np.array(corners), np.array(bdr), np.array(battr), nv)


def perturb_mesh(mesh: HexMesh, amp=0.1):
    """Smooth, x-seam-periodic displacement of all corners so q-data is non-affine
    (SURVEY 8d 'perturbed copy')."""
    lo, hi = mesh.bounding_box()
    L = hi - lo
    hmin = np.min(np.linalg.norm(mesh.corners[:, 1] - mesh.corners[:, 0], axis=1))
    X = (mesh.corners - lo) / L
    s = amp * hmin * np.sin(2 * np.pi * X[..., 0]) * np.sin(np.pi * X[..., 1]) * np.sin(np.pi * X[..., 2])
    out = mesh.corners.copy()
    out[..., 0] += s
    out[..., 1] += 0.5 * s * (L[1] / L[0])
    out[..., 2] += 0.7 * s * (L[2] / L[0])
    return HexMesh(mesh.elems, out, mesh.bdr, mesh.bdr_attr, mesh.nv)


# --------------------------------------------------------------------------
# H1 space: global numbering by (vertex, integer-weight) signatures
# --------------------------------------------------------------------------
@dataclass
class H1Space:
    p: int
    basis: Basis
    mesh: HexMesh
    gather: np.ndarray        # [NE, D^3] element-lexicographic -> global dof (ElementRestriction)
    ndof: int
    xyz: np.ndarray           # [ndof, 3] physical node coordinates (one representative per dof)
    ess: np.ndarray           # essential dofs (closure of boundary attr 2)
    surf2vol: np.ndarray      # surface dof s -> volume dof
    surf_xy: np.ndarray       # [ns, 3] coordinates of the surface dofs


def build_h1_space(mesh: HexMesh, p: int, ess_attr=2) -> H1Space:
    """H1_FECollection(order,3) on hexes ([MFEM] fem/fe_coll.cpp, fespace.cpp):
    one dof per vertex, p-1 per edge, (p-1)^2 per face, (p-1)^3 per interior.
    A node is identified by the multiset {(vertex id, trilinear integer weight)}
    of the corners with non-zero weight -- orientation independent, never by
    coordinates (periodic seam copies coincide topologically, not geometrically)."""
    bs = make_basis(p)
    D = bs.D
    ne = mesh.ne
    ev = mesh.elems[:, LEX2MFEM].astype(np.int64)       # [NE,8] lexicographic
    ii = np.arange(D)
    wI = np.stack([p - ii, ii], axis=0)                  # weight of corner bit 0/1 at lattice i
    # W[c, k, j, i] for corner c=(cx,cy,cz)
    W = np.zeros((8, D, D, D), dtype=np.int64)
    for c in range(8):
        cx, cy, cz = c & 1, (c >> 1) & 1, (c >> 2) & 1
        W[c] = wI[cz][:, None, None] * wI[cy][None, :, None] * wI[cx][None, None, :]
    W = W.reshape(8, D ** 3).T                           # [D^3, 8]
    base = p ** 3 + 1
    sig = np.where(W[None, :, :] > 0, ev[:, None, :] * base + W[None, :, :], -1)   # [NE, D^3, 8]
    # interior nodes (all 8 weights > 0) are private to the element even if two
    # elements shared all 8 vertices; add the element id to be safe
    sig = np.sort(sig, axis=2)
    interior = np.all(W > 0, axis=1)
    eid = np.broadcast_to(np.arange(ne)[:, None], (ne, D ** 3))
    extra = np.where(interior[None, :], eid, -1)
    keys = np.concatenate([sig, extra[:, :, None]], axis=2).reshape(ne * D ** 3, 9)
    _, first, inv = np.unique(keys, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    # renumber by first appearance (element-major) for locality
    order = np.argsort(first, kind='stable')
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    gather = rank[inv].reshape(ne, D ** 3)
    ndof = len(order)
    # coordinates
    lat = np.array([[bs.nodes[i], bs.nodes[j], bs.nodes[k]] for k in range(D) for j in range(D) for i in range(D)])
    X = _trilinear(mesh.corners, lat)                     # [NE, D^3, 3]
    xyz = np.zeros((ndof, 3))
    xyz[gather.reshape(-1)] = X.reshape(-1, 3)
    # boundary faces with ess_attr -> (element, local face) by sorted vertex set
    fmap = {}
    for e in range(ne):
        for f in range(6):
            fmap.setdefault(tuple(sorted(ev[e, HEX_FACES_LEX[f]])), []).append((e, f))
    ess_list, seen = [], set()
    surf_faces = []
    for b in range(mesh.bdr.shape[0]):
        if mesh.bdr_attr[b] != ess_attr:
            continue
        for (e, f) in fmap[tuple(sorted(int(t) for t in mesh.bdr[b]))]:
            surf_faces.append((e, f))
            for d in _face_local_dofs(f, D):
                g = int(gather[e, d])
                if g not in seen:
                    seen.add(g)
                    ess_list.append(g)
    surf2vol = np.array(ess_list, dtype=np.int64)         # first-appearance order
    ess = np.sort(surf2vol)
    sp_ = H1Space(p, bs, mesh, gather, ndof, xyz, ess, surf2vol, xyz[surf2vol])
    sp_.surf_faces = surf_faces
    return sp_


def build_tank_space(nx, ny, nz, Lx, Ly, H, p, periodic_x=True) -> H1Space:
    """H1 space of the wave-tank generator's box mesh (Meshes/wave_tank.cpp:13-47, Meshes/wave-tank-finite.cpp:10-45) by
    LATTICE numbering: dof (ix, iy, iz) of the (nx p [+1]) x (ny p + 1) x (nz p + 1) node lattice, x wrapped when
    periodic.  Same dof identification as build_h1_space(make_wave_tank(...)) (tests/test_oracle.py checks that the two
    numberings induce the same partition) at a cost linear in the mesh size -- this is what lets the CPU reference arm of
    bench.py run the full 262 144-hex workload.  A uniformly refined tank IS the tank with 2^r times the cells."""
    bs = make_basis(p)
    D = bs.D
    NX = nx * p if periodic_x else nx * p + 1
    NY, NZ = ny * p + 1, nz * p + 1
    ex, ey, ez = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing='ij')
    ex, ey, ez = (a.transpose(2, 1, 0).reshape(-1) for a in (ex, ey, ez))          # element index: x fastest
    i = np.arange(D)
    gx = (ex[:, None] * p + i[None, :]) % NX if periodic_x else ex[:, None] * p + i[None, :]
    gy = ey[:, None] * p + i[None, :]
    gz = ez[:, None] * p + i[None, :]
    gather = (gx[:, None, None, :] + NX * (gy[:, None, :, None] + NY * gz[:, :, None, None])).reshape(len(ex), D ** 3)
    ndof = NX * NY * NZ
    hx, hy, hz = Lx / nx, Ly / ny, H / nz
    corners = np.zeros((len(ex), 8, 3))
    for c in range(8):
        corners[:, c, 0] = (ex + (c & 1)) * hx
        corners[:, c, 1] = (ey + ((c >> 1) & 1)) * hy
        corners[:, c, 2] = (ez + ((c >> 2) & 1)) * hz
    mesh = HexMesh(elems=np.zeros((len(ex), 8), dtype=np.int64), corners=corners, bdr=np.zeros((0, 4), dtype=np.int64),
                   bdr_attr=np.zeros(0, dtype=np.int64), nv=0)
    nodes = bs.nodes
    X1 = ((np.arange(nx)[:, None] + nodes[None, :p]).reshape(-1) * hx) if periodic_x else \
        np.concatenate([(np.arange(nx)[:, None] + nodes[None, :p]).reshape(-1), [nx]]) * hx
    Y1 = np.concatenate([(np.arange(ny)[:, None] + nodes[None, :p]).reshape(-1), [ny]]) * hy
    Z1 = np.concatenate([(np.arange(nz)[:, None] + nodes[None, :p]).reshape(-1), [nz]]) * hz
    xyz = np.stack([np.tile(X1, NY * NZ), np.tile(np.repeat(Y1, NX), NZ), np.repeat(Z1, NX * NY)], axis=1)
    surf2vol = NX * NY * (NZ - 1) + np.arange(NX * NY, dtype=np.int64)             # top lattice plane = boundary attribute 2
    return H1Space(p, bs, mesh, gather, ndof, xyz, surf2vol.copy(), surf2vol, xyz[surf2vol])


def _face_local_dofs(f, D):
    idx = np.arange(D ** 3).reshape(D, D, D)   # [k][j][i]
    if f == 0: return idx[0, :, :].reshape(-1)
    if f == 1: return idx[D - 1, :, :].reshape(-1)
    if f == 2: return idx[:, 0, :].reshape(-1)
    if f == 3: return idx[:, D - 1, :].reshape(-1)
    if f == 4: return idx[:, :, 0].reshape(-1)
    return idx[:, :, D - 1].reshape(-1)


# --------------------------------------------------------------------------
# Partial assembly: q-data, apply, diagonal   
# --------------------------------------------------------------------------
def jacobians_trilinear(corners, qpts):
    """J[e, q, c, k] = d x_c / d xi_k at tensor Gauss points, q = qx + Q(qy + Q qz)."""
    Q = len(qpts)
    xi = np.array([[qpts[i], qpts[j], qpts[k]] for k in range(Q) for j in range(Q) for i in range(Q)])
    x, y, z = xi[:, 0], xi[:, 1], xi[:, 2]
    dN = np.zeros((len(xi), 8, 3))
    for c in range(8):
        cx, cy, cz = c & 1, (c >> 1) & 1, (c >> 2) & 1
        fx = x if cx else 1 - x
        fy = y if cy else 1 - y
        fz = z if cz else 1 - z
        sx = 1.0 if cx else -1.0
        sy = 1.0 if cy else -1.0
        sz = 1.0 if cz else -1.0
        dN[:, c, 0] = sx * fy * fz
        dN[:, c, 1] = fx * sy * fz
        dN[:, c, 2] = fx * fy * sz
    return np.einsum('qck,ecd->eqdk', dN, corners)


def pa_setup(J, qwts):
    """DiffusionIntegrator::AssemblePA -> PADiffusionSetup3D restated
    ([MFEM] fem/integ/bilininteg_diffusion_pa.cpp): D = w/detJ * adj(J) adj(J)^T,
    6 symmetric entries in the order 11,21,31,22,32,33.  Returns [NE, 6, Q^3]."""
    Q = len(qwts)
    W = np.array([qwts[i] * qwts[j] * qwts[k] for k in range(Q) for j in range(Q) for i in range(Q)])
    J11, J12, J13 = J[..., 0, 0], J[..., 0, 1], J[..., 0, 2]
    J21, J22, J23 = J[..., 1, 0], J[..., 1, 1], J[..., 1, 2]
    J31, J32, J33 = J[..., 2, 0], J[..., 2, 1], J[..., 2, 2]
    detJ = J11 * (J22 * J33 - J32 * J23) - J21 * (J12 * J33 - J32 * J13) + J31 * (J12 * J23 - J22 * J13)
    cw = W[None, :] / detJ
    A11 = (J22 * J33) - (J23 * J32)
    A12 = (J32 * J13) - (J12 * J33)
    A13 = (J12 * J23) - (J22 * J13)
    A21 = (J31 * J23) - (J21 * J33)
    A22 = (J11 * J33) - (J13 * J31)
    A23 = (J21 * J13) - (J11 * J23)
    A31 = (J21 * J32) - (J31 * J22)
    A32 = (J31 * J12) - (J11 * J32)
    A33 = (J11 * J22) - (J12 * J21)
    qd = np.stack([
        cw * (A11 * A11 + A12 * A12 + A13 * A13),
        cw * (A11 * A21 + A12 * A22 + A13 * A23),
        cw * (A11 * A31 + A12 * A32 + A13 * A33),
        cw * (A21 * A21 + A22 * A22 + A23 * A23),
        cw * (A21 * A31 + A22 * A32 + A23 * A33),
        cw * (A31 * A31 + A32 * A32 + A33 * A33)], axis=1)
    return qd, detJ


def pa_apply_E(qd, bs: Basis, xE):
    """DiffusionIntegrator::AddMultPA -> PADiffusionApply3D restated: y_E = B^T G^T D (G,B) x_E.
    xE [NE, D^3] lexicographic (x fastest); returns yE (NOT accumulated)."""
    D, Q = bs.D, bs.Q
    B, G = bs.B, bs.G
    ne = xE.shape[0]
    u = xE.reshape(ne, D, D, D)                      # [e, dz, dy, dx]
    bx = np.einsum('qd,ezyd->ezyq', B, u)
    gx = np.einsum('qd,ezyd->ezyq', G, u)
    bb = np.einsum('ry,ezyq->ezrq', B, bx)           # B_y B_x
    gb = np.einsum('ry,ezyq->ezrq', B, gx)           # B_y G_x
    bg = np.einsum('ry,ezyq->ezrq', G, bx)           # G_y B_x
    g0 = np.einsum('sz,ezrq->esrq', B, gb).reshape(ne, Q ** 3)
    g1 = np.einsum('sz,ezrq->esrq', B, bg).reshape(ne, Q ** 3)
    g2 = np.einsum('sz,ezrq->esrq', G, bb).reshape(ne, Q ** 3)
    f0 = qd[:, 0] * g0 + qd[:, 1] * g1 + qd[:, 2] * g2
    f1 = qd[:, 1] * g0 + qd[:, 3] * g1 + qd[:, 4] * g2
    f2 = qd[:, 2] * g0 + qd[:, 4] * g1 + qd[:, 5] * g2
    f0 = f0.reshape(ne, Q, Q, Q); f1 = f1.reshape(ne, Q, Q, Q); f2 = f2.reshape(ne, Q, Q, Q)
    # transposed contractions z -> y -> x
    c_gb = np.einsum('sz,esrq->ezrq', B, f0)
    c_bg = np.einsum('sz,esrq->ezrq', B, f1)
    c_bb = np.einsum('sz,esrq->ezrq', G, f2)
    ta = np.einsum('ry,ezrq->ezyq', B, c_bb) + np.einsum('ry,ezrq->ezyq', G, c_bg)
    tb = np.einsum('ry,ezrq->ezyq', B, c_gb)
    y = np.einsum('qd,ezyq->ezyd', B, ta) + np.einsum('qd,ezyq->ezyd', G, tb)
    return y.reshape(ne, D ** 3)


def pa_diag_E(qd, bs: Basis):
    """DiffusionIntegrator::AssembleDiagonalPA restated: per-element diagonal [NE, D^3]."""
    D, Q = bs.D, bs.Q
    B, G = bs.B, bs.G
    ne = qd.shape[0]
    q6 = qd.reshape(ne, 6, Q, Q, Q)                  # [e, c, qz, qy, qx]
    BB, GG, BG = B * B, G * G, B * G
    idx = {(0, 0): 0, (1, 0): 1, (0, 1): 1, (2, 0): 2, (0, 2): 2, (1, 1): 3, (2, 1): 4, (1, 2): 4, (2, 2): 5}
    out = np.zeros((ne, D, D, D))
    for a in range(3):
        for b in range(3):
            # psi_a psi_b factorises per direction: direction t uses G if t==a or t==b
            def fac(t):
                ga, gb_ = (t == a), (t == b)
                return GG if (ga and gb_) else (BG if (ga or gb_) else BB)
            Fx, Fy, Fz = fac(0), fac(1), fac(2)
            out += np.einsum('esrq,qi,rj,sk->ekji', q6[:, idx[(a, b)]], Fx, Fy, Fz, optimize=True)
    return out.reshape(ne, D ** 3)


def element_matrices(qd, bs: Basis):
    """Full-assembly twin: K_e[i,j] = sum_q grad phi_i^T D grad phi_j, dense [NE, D^3, D^3].
    Restates DiffusionIntegrator::AssembleElementMatrix with the same rule
    (Solvers/laplace_solver.cpp:99-103)."""
    D, Q = bs.D, bs.Q
    B, G = bs.B, bs.G
    # grad tables [3, Q^3, D^3]
    def kron3(Z, Y, X):
        return np.einsum('sk,rj,qi->srqkji', Z, Y, X).reshape(Q ** 3, D ** 3)
    Gr = np.stack([kron3(B, B, G), kron3(B, G, B), kron3(G, B, B)])
    sym = [[0, 1, 2], [1, 3, 4], [2, 4, 5]]
    ne = qd.shape[0]
    K = np.zeros((ne, D ** 3, D ** 3))
    for a in range(3):
        for b in range(3):
            K += np.einsum('qi,eq,qj->eij', Gr[a], qd[:, sym[a][b]], Gr[b], optimize=True)
    return K


def assemble_csr(K, gather, ndof):
    ne, n, _ = K.shape
    rows = np.repeat(gather, n, axis=1).reshape(-1)
    cols = np.tile(gather, (1, n)).reshape(-1)
    return sp.csr_matrix((K.reshape(-1), (rows, cols)), shape=(ndof, ndof))


class PAOperator:
    """P^T A P with A = G^T (B^T D B) G for one rank (serial: P = I).
    Mirrors PABilinearFormExtension::Mult ([MFEM] fem/bilinearform_ext.cpp)."""

    def __init__(self, space: H1Space, J=None):
        self.space = space
        bs = space.basis
        if J is None:
            J = jacobians_trilinear(space.mesh.corners, bs.qpts)
        self.qd, self.detJ = pa_setup(J, bs.qwts)
        self.n = space.ndof

    def mult(self, x):
        g = self.space.gather
        yE = pa_apply_E(self.qd, self.space.basis, x[g])
        y = np.zeros(self.n)
        np.add.at(y, g.reshape(-1), yE.reshape(-1))
        return y

    def diag(self):
        g = self.space.gather
        dE = pa_diag_E(self.qd, self.space.basis)
        d = np.zeros(self.n)
        np.add.at(d, g.reshape(-1), dE.reshape(-1))
        return d


class FAOperator:
    def __init__(self, space: H1Space, J=None):
        bs = space.basis
        if J is None:
            J = jacobians_trilinear(space.mesh.corners, bs.qpts)
        qd, _ = pa_setup(J, bs.qwts)
        self.A = assemble_csr(element_matrices(qd, bs), space.gather, space.ndof)
        self.n = space.ndof

    def mult(self, x):
        return self.A @ x

    def diag(self):
        return self.A.diagonal()


# --------------------------------------------------------------------------
# Constrained system + Jacobi + PCG   
# --------------------------------------------------------------------------
class ConstrainedOperator:
    """[MFEM] linalg/operator.cpp ConstrainedOperator (DIAG_ONE)."""

    def __init__(self, A, ess):
        self.A, self.ess = A, np.asarray(ess)

    def mult(self, x):
        z = x.copy()
        z[self.ess] = 0.0
        y = self.A.mult(z)
        y[self.ess] = x[self.ess]
        return y

    def eliminate_rhs(self, x, b):
        w = np.zeros_like(x)
        w[self.ess] = x[self.ess]
        b = b - self.A.mult(w)
        b[self.ess] = x[self.ess]
        return b


def form_linear_system(A, ess, x, b):
    """Operator::FormLinearSystem with copy_interior=0: X = x on ess, 0 elsewhere."""
    Ac = ConstrainedOperator(A, ess)
    X = np.zeros_like(x)
    X[ess] = x[ess]
    B = Ac.eliminate_rhs(X, b.copy())
    return Ac, X, B


def jacobi_dinv(A, ess):
    """OperatorJacobiSmoother(a, ess_tdof): dinv = 1/diag, dinv[ess] = 1."""
    d = A.diag()
    assert np.all(d > 0)
    dinv = 1.0 / d
    dinv[ess] = 1.0
    return dinv


@dataclass
class PCGInfo:
    iters: int
    converged: bool
    final_norm: float
    initial_norm: float
    applies: int


def pcg(Ac, dinv, b, x, rel_tol, abs_tol, max_iter, dot=None):
    """CGSolver::Mult restated ([MFEM] linalg/solvers.cpp), iterative_mode = true.
    Stops on (r, M r) <= max(rel^2 (r0, M r0), abs^2)."""
    if dot is None:
        dot = lambda a, c: float(np.dot(a, c))
    applies = 0
    r = b - Ac.mult(x); applies += 1
    z = dinv * r
    d = z.copy()
    nom0 = nom = dot(d, r)
    if nom < 0.0:
        return x, PCGInfo(0, False, nom, nom, applies)
    r0 = max(nom * rel_tol * rel_tol, abs_tol * abs_tol)
    if nom <= r0:
        return x, PCGInfo(0, True, math.sqrt(nom), math.sqrt(nom0), applies)
    z = Ac.mult(d); applies += 1
    den = dot(z, d)
    if den <= 0.0 and den == 0.0:
        return x, PCGInfo(0, False, math.sqrt(nom), math.sqrt(nom0), applies)
    converged, final_iter = False, max_iter
    i = 1
    betanom = nom
    while True:
        alpha = nom / den
        x = x + alpha * d
        r = r - alpha * z
        z = dinv * r
        betanom = dot(r, z)
        if betanom < 0.0:
            converged, final_iter = False, i
            break
        if betanom <= r0:
            converged, final_iter = True, i
            break
        i += 1
        if i > max_iter:
            break
        beta = betanom / nom
        d = z + beta * d
        z = Ac.mult(d); applies += 1
        den = dot(d, z)
        if den <= 0.0 and den == 0.0:
            final_iter = i
            break
        nom = betanom
    return x, PCGInfo(final_iter, converged, math.sqrt(max(betanom, 0.0)), math.sqrt(nom0), applies)


# --------------------------------------------------------------------------
# GetDerivative(1, 2, w) restated   
# --------------------------------------------------------------------------
def nodal_jacobians(space: H1Space):
    bs = space.basis
    return jacobians_trilinear(space.mesh.corners, bs.nodes)   # [NE, D^3, 3, 3]


def get_derivative_z(space: H1Space, phi, elems=None):
    """w(i) = (1/m_i) sum_{e ni i} sum_k Jinv(k, z) dphi_e/dxi_k at node i
    (GridFunction::AccumulateAndCountDerivativeValues + divide, [MFEM] fem/gridfunc.cpp).
    `elems` restricts the element loop (valid for dofs all of whose elements are listed)."""
    bs = space.basis
    D = bs.D
    g = space.gather if elems is None else space.gather[elems]
    corners = space.mesh.corners if elems is None else space.mesh.corners[elems]
    ne = g.shape[0]
    u = phi[g].reshape(ne, D, D, D)
    g0 = np.einsum('am,ezym->ezya', bs.Dhat, u)
    g1 = np.einsum('am,ezmx->ezax', bs.Dhat, u)
    g2 = np.einsum('am,emyx->eayx', bs.Dhat, u)
    Jn = jacobians_trilinear(corners, bs.nodes)
    Jinv = np.linalg.inv(Jn)                                   # [e, n, k(ref), c(phys)]
    a = (Jinv[:, :, 0, 2] * g0.reshape(ne, -1) + Jinv[:, :, 1, 2] * g1.reshape(ne, -1)
         + Jinv[:, :, 2, 2] * g2.reshape(ne, -1))
    w = np.zeros(space.ndof)
    cnt = np.zeros(space.ndof)
    np.add.at(w, g.reshape(-1), a.reshape(-1))
    np.add.at(cnt, g.reshape(-1), 1.0)
    cnt[cnt == 0] = 1.0
    return w / cnt


def surface_elements(space: H1Space):
    """Elements holding at least one free-surface dof (all that GetDerivative needs for w~)."""
    mask = np.zeros(space.ndof, dtype=bool)
    mask[space.surf2vol] = True
    return np.nonzero(mask[space.gather].any(axis=1))[0]


# --------------------------------------------------------------------------
# Wave parameters, relaxation weights, RHS, RK4   
# --------------------------------------------------------------------------
@dataclass
class Wave:
    H: float = 0.01
    g: float = 9.81
    lam: float = 1.0
    kh: float = 1.0
    theta: float = 0.0

    def __post_init__(self):
        self.k = 2.0 * math.pi / self.lam
        self.c = math.sqrt((self.g / self.k) * math.tanh(self.kh))
        self.T = self.lam / self.c
        self.omega = 2.0 * math.pi / self.T
        self.kx = math.cos(self.theta)
        self.ky = math.sin(self.theta)

    def phase(self, t, x, y):
        return self.omega * t - self.k * (self.kx * x + self.ky * y)

    def eta(self, t, x, y):
        return 0.5 * self.H * np.cos(self.phase(t, x, y))

    def phi_fs(self, t, x, y):
        return -0.5 * self.H * self.c * (math.cosh(self.kh) / math.sinh(self.kh)) * np.sin(self.phase(t, x, y))

    def phi(self, t, x, y, z, h):
        """Airy potential with z measured from the bottom (z_bottom = 0, surface z = h):
        Solvers/laplace_solver.cpp:70-81."""
        return (-0.5 * self.H * self.c * np.cosh(self.k * z) / math.sinh(self.kh)
                * np.sin(self.phase(t, x, y)))

    def w_surface(self, t, x, y):
        return -0.5 * self.H * self.c * self.k * np.sin(self.phase(t, x, y))


def dispersion_kh(g, T, h, n):
    """Solvers/PF_linear_par_partial.cpp:20-32 fixed-point iteration."""
    w = 2.0 * math.pi / T
    kh = max((w * w) * h / g, 1e-8)
    for _ in range(n):
        x = max(kh, 1e-12)
        kh = math.sqrt((w * w / g) * h * kh * (math.cosh(x) / math.sinh(x)))
        kh = max(kh, 1e-8)
    return kh


def relax_cgen(x, xg0, xg1):
    """Solvers/PF_linear_par_partial.cpp:419-426."""
    xi = (x - xg0) / (xg1 - xg0)
    v = 1 - (-2.0 * xi ** 3 + 3.0 * xi ** 2)
    return np.where(x <= xg0, 1.0, np.where(x >= xg1, 0.0, v))


def relax_cabs(x, x0, x1, pw=5.0):
    """Solvers/PF_linear_par_partial.cpp:436-444."""
    xi = np.clip((x - x0) / (x1 - x0), 0.0, 1.0)
    return np.where(x <= x0, 0.0, np.where(x >= x1, 1.0, xi ** pw))


@dataclass
class Relax:
    cgen: np.ndarray
    cabs: np.ndarray       # C_abs (absorption towards x_max)
    tau: float
    n_ramp: float = 3.0
    cabsy: np.ndarray | None = None   # third weight of the cylinder driver, added after C_abs (cylinder-diffraction.cpp:199-210)


class RhsLinear:
    """rhs_linear::Mult (Solvers/PF_linear_par_partial.cpp:130-244; ss.cpp:62-105 without relaxation).
    State = [eta; phi_fs] on the surface dofs (serial: true dofs == local dofs)."""

    def __init__(self, space: H1Space, wave: Wave, rel_tol=1e-12, max_iter=1000, relax: Relax | None = None,
                 operator=None):
        self.space, self.wave = space, wave
        self.A = operator if operator is not None else PAOperator(space)
        self.dinv = jacobi_dinv(self.A, space.ess)
        self.rel_tol, self.max_iter = rel_tol, max_iter
        self.relax = relax
        self.selems = surface_elements(space)
        self.ns = len(space.surf2vol)
        self.t = 0.0
        self.last = None
        self.iters = []
        self.phi = np.zeros(space.ndof)

    def set_time(self, t):
        self.t = t

    def solve_laplace(self, phi_fs):
        sp_ = self.space
        phi = self.phi
        phi[sp_.surf2vol] = phi_fs                      # ParSubMesh::Transfer (:147)
        b = np.zeros(sp_.ndof)
        Ac, X, B = form_linear_system(self.A, sp_.ess, phi, b)
        X, info = pcg(Ac, self.dinv, B, X, self.rel_tol, 0.0, self.max_iter)
        self.phi = X.copy()                              # RecoverFEMSolution (:166)
        self.last = info
        self.iters.append(info.iters)
        return self.phi

    def mult(self, state):
        ns = self.ns
        eta, phi_fs = state[:ns], state[ns:]
        phi = self.solve_laplace(phi_fs)
        w = get_derivative_z(self.space, phi, self.selems)
        deta = w[self.space.surf2vol].copy()
        dphi = -self.wave.g * eta
        if self.relax is not None:
            rx, wv = self.relax, self.wave
            x, y = self.space.surf_xy[:, 0], self.space.surf_xy[:, 1]
            eta_e = wv.eta(self.t, x, y)
            phi_e = wv.phi_fs(self.t, x, y)
            alpha = min(1.0, max(0.0, self.t / (rx.n_ramp * wv.T)))
            inv_tau = 1.0 / rx.tau
            gw = alpha * rx.cgen
            deta = deta + (gw * inv_tau) * (eta_e - eta)
            dphi = dphi + (gw * inv_tau) * (phi_e - phi_fs)
            deta = deta + (rx.cabs * inv_tau) * (0.0 - eta)
            dphi = dphi + (rx.cabs * inv_tau) * (0.0 - phi_fs)
            if rx.cabsy is not None:
                deta = deta + (rx.cabsy * inv_tau) * (0.0 - eta)
                dphi = dphi + (rx.cabsy * inv_tau) * (0.0 - phi_fs)
        return np.concatenate([deta, dphi])


def rk4_step(f: RhsLinear, x, t, dt):
    """RK4Solver::Step ([MFEM] linalg/ode.cpp; used at Solvers/PF_linear_par_partial.cpp:494)."""
    f.set_time(t)
    k = f.mult(x)
    y = x + (dt / 2) * k
    z = x + (dt / 6) * k
    f.set_time(t + dt / 2)
    k = f.mult(y)
    y = x + (dt / 2) * k
    z = z + (dt / 3) * k
    k = f.mult(y)
    y = x + dt * k
    z = z + (dt / 3) * k
    f.set_time(t + dt)
    k = f.mult(y)
    x = z + (dt / 6) * k
    return x, t + dt


# --------------------------------------------------------------------------
# partition-independent pseudo-random vectors
# --------------------------------------------------------------------------
def splitmix64(x):
    x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def hash_noise(n, seed=0x5EED1234):
    with np.errstate(over='ignore'):
        h = splitmix64(np.arange(n, dtype=np.uint64) ^ np.uint64(seed))
    return (h >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0


def write_mfem_mesh(mesh: HexMesh, path, with_nodes=None):
    """Writes MFEM mesh v1.0 (own writer; used to produce self-contained test inputs).  with_nodes=True
    emits the discontinuous L2_T1_3D_P1 `nodes` block (needed for periodic meshes), False plain vertices."""
    if with_nodes is None:
        # periodic meshes cannot carry geometry in the vertices
        vx = np.full((mesh.nv, 3), np.nan)
        ok = True
        ev = mesh.elems[:, LEX2MFEM]
        for c in range(8):
            cur = vx[ev[:, c]]
            new = mesh.corners[:, c]
            bad = ~np.isnan(cur[:, 0]) & (np.abs(cur - new).max(axis=1) > 1e-13)
            if bad.any():
                ok = False
                break
            vx[ev[:, c]] = new
        with_nodes = not ok
    with open(path, 'w') as f:
        f.write('MFEM mesh v1.0\n\ndimension\n3\n\nelements\n%d\n' % mesh.ne)
        for e in range(mesh.ne):
            f.write('1 5 ' + ' '.join(str(int(v)) for v in mesh.elems[e]) + '\n')
        f.write('\nboundary\n%d\n' % len(mesh.bdr_attr))
        for b in range(len(mesh.bdr_attr)):
            f.write('%d 3 ' % mesh.bdr_attr[b] + ' '.join(str(int(v)) for v in mesh.bdr[b]) + '\n')
        f.write('\nvertices\n%d\n' % mesh.nv)
        if with_nodes:
            f.write('\nnodes\nFiniteElementSpace\nFiniteElementCollection: L2_T1_3D_P1\nVDim: 3\nOrdering: 1\n\n')
            for e in range(mesh.ne):
                for c in range(8):
                    f.write('%.17g %.17g %.17g\n' % tuple(mesh.corners[e, c]))
        else:
            vx = np.zeros((mesh.nv, 3))
            vx[mesh.elems[:, LEX2MFEM].reshape(-1)] = mesh.corners.reshape(-1, 3)
            f.write('3\n')
            for v in range(mesh.nv):
                f.write('%.17g %.17g %.17g\n' % tuple(vx[v]))
