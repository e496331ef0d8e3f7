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
