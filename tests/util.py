"""Helpers shared by the tests: run the oracle on exactly the inputs the product sees."""
import numpy as np


def oracle_space_from(orc, space):
    """Builds an oracle H1Space that reuses the product's gather map / corners / essential list, so both
    sides compute on identical inputs (numbering equivalence itself is tested in test_host.py)."""
    p = space.order
    bs = orc.make_basis(p)
    mesh = orc.HexMesh(elems=np.zeros((space.ne, 8), dtype=np.int64), corners=np.array(space.corners),
                       bdr=np.zeros((0, 4), dtype=np.int64), bdr_attr=np.zeros(0, dtype=np.int64), nv=0)
    xyz = space.node_coordinates()
    s2v = np.array(space.surf2vol, dtype=np.int64)
    sp = orc.H1Space(p, bs, mesh, np.array(space.gather, dtype=np.int64), space.ndof, xyz,
                     np.sort(np.array(space.ess, dtype=np.int64)), s2v, np.array(xyz[s2v]))
    return sp


def rel_err(a, b):
    a = np.asarray(a); b = np.asarray(b)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


# ------------------------------------------------------------------------------------------------
# Independent oracle side for the config-scale parity tests: mesh, refinement, H1 numbering and basis tables
# all come from oracle/ (lpf_oracle.py), the element arithmetic from the C oracle (pa_oracle.c).  Nothing of
# the product is reused; the two numberings are related afterwards by geometry (dof_map).
# ------------------------------------------------------------------------------------------------
class IndependentOracle:
    def __init__(self, orc, corc, omesh, p, threads=None):
        import os
        self.orc, self.mesh, self.p = orc, omesh, p
        self.sp = orc.build_h1_space(omesh, p)
        bs = self.sp.basis
        corc.set_threads(threads or os.cpu_count())
        self.cop = corc.COperator(p, omesh.corners, self.sp.gather.astype(np.int32), self.sp.ndof,
                                  dict(B=bs.B, G=bs.G, Dhat=bs.Dhat, nodes=bs.nodes, qpts=bs.qpts, qwts=bs.qwts))
        self.n = self.sp.ndof

    # PAOperator interface of lpf_oracle (mult / diag), so ConstrainedOperator, pcg, RhsLinear accept it
    def mult(self, x):
        return self.cop.mult(x)

    def diag(self):
        return self.cop.diag()

    def constrained(self):
        return self.orc.ConstrainedOperator(self, self.sp.ess)

    def dinv(self):
        return self.orc.jacobi_dinv(self, self.sp.ess)

    def solve(self, phi_ess, rel_tol, max_iter):
        """FormLinearSystem + Jacobi-PCG + RecoverFEMSolution with essential data phi_ess [ndof] (interior ignored)."""
        Ac, X, B = self.orc.form_linear_system(self, self.sp.ess, phi_ess, np.zeros(self.n))
        return self.orc.pcg(Ac, self.dinv(), B, X, rel_tol, 0.0, max_iter)

    def surface_dz(self, phi):
        el = self.orc.surface_elements(self.sp)
        w, cnt = self.cop.deriv_z(phi, el)
        cnt[cnt == 0] = 1.0
        return (w / cnt)[self.sp.surf2vol]


def _trilinear_nodes(corners, nodes1d):
    D = len(nodes1d)
    lat = np.array([[nodes1d[i], nodes1d[j], nodes1d[k]] for k in range(D) for j in range(D) for i in range(D)])
    x, y, z = lat[:, 0], lat[:, 1], lat[:, 2]
    N = np.stack([(1 - x) * (1 - y) * (1 - z), x * (1 - y) * (1 - z), (1 - x) * y * (1 - z), x * y * (1 - z),
                  (1 - x) * (1 - y) * z, x * (1 - y) * z, (1 - x) * y * z, x * y * z], axis=1)
    return np.einsum("nc,ecd->end", N, corners)


def dof_map(space, osp):
    """product L-dof -> oracle dof, found by GEOMETRY only: elements are matched by centroid, element-local nodes by
    their physical coordinates (computed from each side's own corners; inside one element coordinates are unique even
    on periodic meshes).  Asserts that the result is a consistent bijection."""
    p = space.order
    D3 = (p + 1) ** 3
    cp, co = np.asarray(space.corners), np.asarray(osp.mesh.corners)
    assert cp.shape == co.shape, (cp.shape, co.shape)
    lo, hi = co.reshape(-1, 3).min(axis=0), co.reshape(-1, 3).max(axis=0)
    scale = 1e-7 * (hi - lo).max()

    def order(c):
        k = np.round(c / scale).astype(np.int64)
        return np.lexsort((k[:, 2], k[:, 1], k[:, 0]))

    ip, io = order(cp.mean(axis=1)), order(co.mean(axis=1))
    assert np.abs(cp.mean(axis=1)[ip] - co.mean(axis=1)[io]).max() < 1e-9 * (hi - lo).max(), "element centroids do not match"
    eo = np.empty(len(ip), dtype=np.int64)
    eo[ip] = io                                            # product element e <-> oracle element eo[e]
    Xp = _trilinear_nodes(cp, osp.basis.nodes)
    Xo = _trilinear_nodes(co[eo], osp.basis.nodes)
    go = np.asarray(osp.gather)[eo]
    if np.abs(Xp - Xo).max() > 1e-9 * (hi - lo).max():     # local axes differ: match nodes per element by coordinates
        ne = Xp.shape[0]

        def local_order(X):
            k = np.round(X / scale).astype(np.int64)
            e = np.repeat(np.arange(ne), D3)
            idx = np.lexsort((k[:, :, 2].reshape(-1), k[:, :, 1].reshape(-1), k[:, :, 0].reshape(-1), e))
            return idx.reshape(ne, D3) - (np.arange(ne) * D3)[:, None]

        lp, lo_ = local_order(Xp), local_order(Xo)
        perm = np.empty_like(lp)
        np.put_along_axis(perm, lp, lo_, axis=1)           # product local node k <-> oracle local node perm[e, k]
        Xo = np.take_along_axis(Xo, perm[:, :, None], axis=1)
        go = np.take_along_axis(go, perm, axis=1)
        assert np.abs(Xp - Xo).max() < 1e-9 * (hi - lo).max(), "element-local nodes do not match"
    gp = np.asarray(space.gather)
    pm = np.full(space.ndof, -1, dtype=np.int64)
    pm[gp.reshape(-1)] = go.reshape(-1)
    assert (pm >= 0).all() and (pm[gp] == go).all(), "the two numberings induce different dof identifications"
    assert len(np.unique(pm)) == space.ndof == osp.ndof, "dof map is not a bijection"
    return pm


def surface_map(space, osp, pm):
    """product surface dof s -> index into the oracle's surface list"""
    inv = np.full(osp.ndof, -1, dtype=np.int64)
    inv[np.asarray(osp.surf2vol)] = np.arange(len(osp.surf2vol))
    sm = inv[pm[np.asarray(space.surf2vol)]]
    assert (sm >= 0).all() and len(np.unique(sm)) == len(osp.surf2vol) == space.nsurf, "free-surface dof sets differ"
    return sm
