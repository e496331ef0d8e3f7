"""lpf_b200 -- thin ctypes binding over liblpf_b200.so (include/lpf_b200.h).

The product is the C-ABI shared library (C++ host mini-FEM + sm_100a CUDA kernels); this module only
loads it and mirrors the reference's object names (Mesh, H1 space, rhs_linear-style context) so tests
and bench.py read like the reference drivers.  There is NO CPU fallback: if the library is missing or
no GPU is present, the device entry points raise LpfError.

The directory name contains '-' (it follows the reference repo name), so import it with
    importlib.import_module("master-thesis-lpf-in-mfem_b200")
tests/conftest.py and bench.py alias it as `lpf_b200`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblpf_b200.so")


class LpfError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        f"or {_HERE}/build.sh -- there is no CPU fallback for the hot path")

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_u8p = C.POINTER(C.c_uint8)


class SpaceDesc(C.Structure):
    _fields_ = [
        ("order", C.c_int), ("ne", C.c_int), ("ndof", C.c_int),
        ("corners", c_dp), ("jac", c_dp), ("jinv_z", c_dp), ("gather", c_ip),
        ("n_ess", C.c_int), ("ess", c_ip),
        ("n_surf", C.c_int), ("surf2vol", c_ip), ("surf_xy", c_dp),
        ("n_surf_elems", C.c_int), ("surf_elems", c_ip), ("surf_mult", c_ip),
        ("nranks", C.c_int), ("rank", C.c_int), ("owned", c_u8p), ("surf_owned", c_u8p),
        ("n_nbr", C.c_int), ("nbr_rank", c_ip), ("nbr_offset", c_ip), ("send_dofs", c_ip),
        ("n_shared", C.c_int), ("shared_dofs", c_ip), ("red_off", c_ip), ("red_src", c_ip),
        ("s_n_nbr", C.c_int), ("s_nbr_rank", c_ip), ("s_nbr_offset", c_ip), ("s_send", c_ip),
        ("s_n_shared", C.c_int), ("s_shared", c_ip), ("s_red_off", c_ip), ("s_red_src", c_ip),
        ("n_true_global", C.c_long), ("n_surf_global", C.c_long), ("l2g", c_ip), ("surf_g", c_ip),
    ]


class PcgInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("final_norm", C.c_double),
                ("initial_norm", C.c_double), ("applies", C.c_int)]

    def __repr__(self):
        return (f"PcgInfo(iterations={self.iterations}, converged={self.converged}, "
                f"final_norm={self.final_norm:.3e}, initial_norm={self.initial_norm:.3e}, applies={self.applies})")


class RhsParams(C.Structure):
    _fields_ = [("g", C.c_double), ("H", C.c_double), ("omega", C.c_double), ("k", C.c_double),
                ("kx_dir", C.c_double), ("ky_dir", C.c_double), ("cwave", C.c_double), ("kh", C.c_double),
                ("T", C.c_double), ("tau", C.c_double), ("n_ramp", C.c_double), ("use_relaxation", C.c_int),
                ("rel_tol", C.c_double), ("abs_tol", C.c_double), ("max_iter", C.c_int)]


# every symbol include/lpf_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SIGNATURES = {
    "lpf_last_error": (C.c_char_p, []),
    "lpf_version": (C.c_int, []),
    "lpf_mesh_read": (_VP, [C.c_char_p]),
    "lpf_mesh_make_wave_tank": (_VP, [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]),
    "lpf_mesh_refine": (C.c_int, [_VP, C.c_int]),
    "lpf_mesh_perturb": (C.c_int, [_VP, C.c_double]),
    "lpf_mesh_num_elements": (C.c_int, [_VP]),
    "lpf_mesh_num_vertices": (C.c_int, [_VP]),
    "lpf_mesh_num_bdr": (C.c_int, [_VP]),
    "lpf_mesh_bounding_box": (C.c_int, [_VP, c_dp, c_dp]),
    "lpf_mesh_corners": (c_dp, [_VP]),
    "lpf_mesh_elements": (c_ip, [_VP]),
    "lpf_mesh_bdr": (c_ip, [_VP]),
    "lpf_mesh_bdr_attr": (c_ip, [_VP]),
    "lpf_mesh_destroy": (None, [_VP]),
    "lpf_space_create": (_VP, [_VP, C.c_int, C.c_int, C.c_int, C.c_int]),
    "lpf_space_destroy": (None, [_VP]),
    "lpf_basis_tables": (C.c_int, [C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "lpf_space_desc_get": (C.c_int, [_VP, C.POINTER(SpaceDesc)]),
    "lpf_space_node_coordinates": (C.c_int, [_VP, c_dp]),
    "lpf_space_rim": (C.c_int, [_VP, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, c_ip, c_dp, C.c_int]),
    "lpf_space_surface_quads": (C.c_int, [_VP, c_ip, C.c_int]),
    "lpf_write_surface_vtu": (C.c_int, [_VP, C.c_char_p, C.c_double, C.c_int, C.POINTER(C.c_char_p), C.POINTER(c_dp), C.c_int]),
    "lpf_maccamy_fuchs": (C.c_double, [C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]),
    "lpf_create": (_VP, [C.POINTER(SpaceDesc), C.c_int, _VP]),
    "lpf_destroy": (None, [_VP]),
    "lpf_stream": (_VP, [_VP]),
    "lpf_sync": (C.c_int, [_VP]),
    "lpf_ndof": (C.c_int, [_VP]),
    "lpf_nsurf": (C.c_int, [_VP]),
    "lpf_ntrue": (C.c_int, [_VP, C.c_int]),
    "lpf_prolong": (C.c_int, [_VP, C.c_int, _VP, _VP]),
    "lpf_restrict": (C.c_int, [_VP, C.c_int, _VP, _VP]),
    "lpf_comm_unique_id": (C.c_int, [_VP]),
    "lpf_comm_init": (C.c_int, [_VP, _VP]),
    "lpf_p2p_export": (C.c_int, [_VP, _VP, C.POINTER(C.c_uint64), c_ip]),
    "lpf_p2p_connect": (C.c_int, [_VP, _VP, C.POINTER(C.c_uint64), c_ip, C.c_int]),
    "lpf_p2p_error": (C.c_int, [_VP]),
    "lpf_pa_setup": (C.c_int, [_VP]),
    "lpf_pa_qdata": (C.c_int, [_VP, _VP]),
    "lpf_pa_apply_E": (C.c_int, [_VP, _VP, _VP]),
    "lpf_apply_L": (C.c_int, [_VP, _VP, _VP]),
    "lpf_apply_T": (C.c_int, [_VP, _VP, _VP]),
    "lpf_apply_T_host": (C.c_int, [_VP, _VP, _VP]),
    "lpf_diag": (C.c_int, [_VP, _VP]),
    "lpf_pa_diag_E": (C.c_int, [_VP, _VP]),
    "lpf_jacobi_setup": (C.c_int, [_VP]),
    "lpf_jacobi_dinv": (C.c_int, [_VP, _VP]),
    "lpf_pcg": (C.c_int, [_VP, _VP, _VP, C.c_double, C.c_double, C.c_int, C.POINTER(PcgInfo)]),
    "lpf_laplace_solve": (C.c_int, [_VP, _VP, C.c_double, C.c_double, C.c_int, C.POINTER(PcgInfo)]),
    "lpf_surface_dz": (C.c_int, [_VP, _VP, _VP]),
    "lpf_rhs_setup": (C.c_int, [_VP, C.POINTER(RhsParams), c_dp, c_dp]),
    "lpf_rhs_set_cabsy": (C.c_int, [_VP, c_dp]),
    "lpf_envelope_reset": (C.c_int, [_VP]),
    "lpf_envelope_update": (C.c_int, [_VP, _VP]),
    "lpf_envelope_get": (C.c_int, [_VP, c_dp, C.c_double]),
    "lpf_rhs": (C.c_int, [_VP, C.c_double, _VP, _VP]),
    "lpf_rk4_step": (C.c_int, [_VP, _VP, c_dp, C.c_double]),
    "lpf_rk4_step_host": (C.c_int, [_VP, _VP, c_dp, C.c_double]),
    "lpf_last_solve_info": (C.c_int, [_VP, C.POINTER(PcgInfo), c_ip]),
    "lpf_phi_dev": (_VP, [_VP]),
    "lpf_time_apply": (C.c_int, [_VP, _VP, _VP, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_long)]),
    "lpf_launch_count": (C.c_long, [_VP]),
    "lpf_set_option": (C.c_int, [_VP, C.c_char_p, C.c_long]),
    "lpf_device_bytes": (C.c_size_t, [_VP]),
    "lpf_affine_active": (C.c_int, [_VP]),
    "lpf_dev_alloc": (_VP, [C.c_size_t]),
    "lpf_dev_free": (C.c_int, [_VP]),
    "lpf_memcpy_h2d": (C.c_int, [_VP, _VP, C.c_size_t]),
    "lpf_memcpy_d2h": (C.c_int, [_VP, _VP, C.c_size_t]),
    "lpf_memset_dev": (C.c_int, [_VP, C.c_int, C.c_size_t]),
    "lpf_host_alloc_pinned": (_VP, [C.c_size_t]),
    "lpf_host_free_pinned": (C.c_int, [_VP]),
    "lpf_device_count": (C.c_int, []),
}
for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)          # AttributeError here == the .so does not export a declared symbol
    _f.restype = _res
    _f.argtypes = _args


def last_error() -> str:
    return lib.lpf_last_error().decode()


def _check(rc, what):
    if rc != 0:
        raise LpfError(f"{what} failed ({rc}): {last_error()}")


def _np(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,))


def basis_tables(order):
    """GLL nodes, Gauss-Legendre (p+2) points/weights, B, G [Q,D], Dhat [D,D] as the library computes them."""
    D, Q = order + 1, order + 2
    nodes, qp, qw = np.zeros(D), np.zeros(Q), np.zeros(Q)
    B, G, Dh = np.zeros((Q, D)), np.zeros((Q, D)), np.zeros((D, D))
    dp = lambda a: a.ctypes.data_as(c_dp)
    _check(lib.lpf_basis_tables(order, dp(nodes), dp(qp), dp(qw), dp(B), dp(G), dp(Dh)), "lpf_basis_tables")
    return dict(nodes=nodes, qpts=qp, qwts=qw, B=B, G=G, Dhat=Dh)


class Mesh:
    """mfem::Mesh stand-in (Solvers/PF_linear_par_partial.cpp:263-266)."""

    def __init__(self, handle):
        if not handle:
            raise LpfError("mesh creation failed: " + last_error())
        self.h = handle

    @classmethod
    def read(cls, path):
        return cls(lib.lpf_mesh_read(str(path).encode()))

    @classmethod
    def wave_tank(cls, nx, ny, nz, Lx=1.0, Ly=0.1, H=1.0 / (2 * np.pi), periodic_x=True):
        return cls(lib.lpf_mesh_make_wave_tank(nx, ny, nz, Lx, Ly, H, int(periodic_x)))

    def refine(self, levels=1):
        _check(lib.lpf_mesh_refine(self.h, levels), "lpf_mesh_refine")
        return self

    def perturb(self, amp):
        _check(lib.lpf_mesh_perturb(self.h, amp), "lpf_mesh_perturb")
        return self

    @property
    def ne(self):
        return lib.lpf_mesh_num_elements(self.h)

    @property
    def nv(self):
        return lib.lpf_mesh_num_vertices(self.h)

    def bounding_box(self):
        lo, hi = np.zeros(3), np.zeros(3)
        _check(lib.lpf_mesh_bounding_box(self.h, lo.ctypes.data_as(c_dp), hi.ctypes.data_as(c_dp)), "bbox")
        return lo, hi

    def corners(self):
        return _np(lib.lpf_mesh_corners(self.h), self.ne * 24, np.float64).reshape(self.ne, 8, 3)

    def elements(self):
        return _np(lib.lpf_mesh_elements(self.h), self.ne * 8, np.int32).reshape(self.ne, 8)

    def boundary(self):
        nb = lib.lpf_mesh_num_bdr(self.h)
        return (_np(lib.lpf_mesh_bdr(self.h), nb * 4, np.int32).reshape(nb, 4),
                _np(lib.lpf_mesh_bdr_attr(self.h), nb, np.int32))

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:      # lib is None while the interpreter shuts down
            lib.lpf_mesh_destroy(self.h)
            self.h = None


class Space:
    """H1 (Par)FiniteElementSpace + free-surface trace space of one rank (:276-285, :407-412)."""

    def __init__(self, mesh: Mesh, order, ess_attr=2, nranks=1, rank=0):
        self.mesh = mesh
        self.order = order
        self.h = lib.lpf_space_create(mesh.h, order, ess_attr, nranks, rank)
        if not self.h:
            raise LpfError("lpf_space_create failed: " + last_error())
        self.desc = SpaceDesc()
        _check(lib.lpf_space_desc_get(self.h, C.byref(self.desc)), "lpf_space_desc_get")
        d = self.desc
        D3 = (order + 1) ** 3
        self.ne, self.ndof, self.nsurf = d.ne, d.ndof, d.n_surf
        self.gather = _np(d.gather, d.ne * D3, np.int32).reshape(d.ne, D3)
        self.corners = _np(d.corners, d.ne * 24, np.float64).reshape(d.ne, 8, 3)
        self.ess = _np(d.ess, d.n_ess, np.int32)
        self.surf2vol = _np(d.surf2vol, d.n_surf, np.int32)
        self.surf_xy = _np(d.surf_xy, 2 * d.n_surf, np.float64).reshape(-1, 2)
        self.surf_elems = _np(d.surf_elems, d.n_surf_elems, np.int32)
        self.surf_mult = _np(d.surf_mult, d.n_surf, np.int32)
        self.owned = _np(d.owned, d.ndof, np.uint8)
        self.surf_owned = _np(d.surf_owned, d.n_surf, np.uint8)
        self.l2g = _np(d.l2g, d.ndof, np.int32)
        self.surf_g = _np(d.surf_g, d.n_surf, np.int32)
        self.nbr_rank = _np(d.nbr_rank, d.n_nbr, np.int32)
        self.nbr_offset = _np(d.nbr_offset, d.n_nbr + 1, np.int32)
        self.send_dofs = _np(d.send_dofs, int(self.nbr_offset[-1]) if d.n_nbr else 0, np.int32)
        self.shared_dofs = _np(d.shared_dofs, d.n_shared, np.int32)
        self.red_off = _np(d.red_off, d.n_shared + 1, np.int32)
        self.red_src = _np(d.red_src, int(self.red_off[-1]) if d.n_shared else 0, np.int32)
        self.n_true_global = d.n_true_global
        self.n_surf_global = d.n_surf_global

    @classmethod
    def from_arrays(cls, order, corners, gather, ess, surf2vol=None, surf_xy=None, jac=None, jinv_z=None):
        """Serial space from plain arrays -- the route an MFEM adapter takes (gather map from
        ElementRestriction, geometry, ess_tdof_list; include/lpf_b200.h lpf_space_desc)."""
        self = cls.__new__(cls)
        self.mesh, self.h, self.order = None, None, order
        D3 = (order + 1) ** 3
        self.corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(-1, 8, 3)
        self.gather = np.ascontiguousarray(gather, dtype=np.int32).reshape(-1, D3)
        self.ess = np.ascontiguousarray(ess, dtype=np.int32)
        self.ne, self.ndof = self.gather.shape[0], int(self.gather.max()) + 1
        self.surf2vol = np.ascontiguousarray(surf2vol if surf2vol is not None else [], dtype=np.int32)
        self.surf_xy = np.ascontiguousarray(surf_xy if surf_xy is not None else np.zeros((0, 2)), dtype=np.float64).reshape(-1, 2)
        self.nsurf = len(self.surf2vol)
        mask = np.zeros(self.ndof, dtype=bool)
        mask[self.surf2vol] = True
        self.surf_elems = np.ascontiguousarray(np.nonzero(mask[self.gather].any(axis=1))[0], dtype=np.int32)
        cnt = np.bincount(self.gather.reshape(-1), minlength=self.ndof)
        self.surf_mult = np.ascontiguousarray(cnt[self.surf2vol], dtype=np.int32)
        self.owned = np.zeros(0, np.uint8); self.surf_owned = np.zeros(0, np.uint8)
        self.l2g = np.arange(self.ndof, dtype=np.int32); self.surf_g = np.arange(self.nsurf, dtype=np.int32)
        self.n_true_global, self.n_surf_global = self.ndof, self.nsurf
        d = SpaceDesc()
        d.order, d.ne, d.ndof = order, self.ne, self.ndof
        d.corners = self.corners.ctypes.data_as(c_dp)
        d.gather = self.gather.ctypes.data_as(c_ip)
        d.n_ess, d.ess = len(self.ess), self.ess.ctypes.data_as(c_ip)
        d.n_surf = self.nsurf
        d.surf2vol = self.surf2vol.ctypes.data_as(c_ip)
        d.surf_xy = self.surf_xy.ctypes.data_as(c_dp)
        d.n_surf_elems, d.surf_elems = len(self.surf_elems), self.surf_elems.ctypes.data_as(c_ip)
        d.surf_mult = self.surf_mult.ctypes.data_as(c_ip)
        d.nranks, d.rank = 1, 0
        d.n_true_global, d.n_surf_global = self.ndof, self.nsurf
        if jac is not None:
            # MFEM GeometricFactors::J layout [Q^3][3][3][ne] (+ nodal J^-1 column for GetDerivative): the
            # kernels then never look at `corners`
            self.jac = np.ascontiguousarray(jac, dtype=np.float64)
            self.jinv_z = np.ascontiguousarray(jinv_z, dtype=np.float64) if jinv_z is not None else None
            d.jac = self.jac.ctypes.data_as(c_dp)
            d.corners = None
            if self.jinv_z is not None:
                d.jinv_z = self.jinv_z.ctypes.data_as(c_dp)
        self.desc = d
        return self

    def rim(self, wall_attr=3, cx=4.0, cy=4.0, a=0.5, tol=5e-3):
        """(theta, local surface dof) of the mesh vertices on the cylinder rim (cylinder-diffraction.cpp:476-496)."""
        n = lib.lpf_space_rim(self.h, wall_attr, cx, cy, a, tol, None, None, 0)
        if n < 0:
            raise LpfError("lpf_space_rim: " + last_error())
        idx, th = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1))
        lib.lpf_space_rim(self.h, wall_attr, cx, cy, a, tol, idx.ctypes.data_as(c_ip), th.ctypes.data_as(c_dp), n)
        return th[:n], idx[:n]

    def surface_quads(self):
        """[nf][(p+1)^2] local surface dofs of the free-surface faces (order-p quads, lexicographic)."""
        nf = lib.lpf_space_surface_quads(self.h, None, 0)
        if nf < 0:
            raise LpfError("lpf_space_surface_quads: " + last_error())
        conn = np.zeros((max(nf, 1), (self.order + 1) ** 2), np.int32)
        lib.lpf_space_surface_quads(self.h, conn.ctypes.data_as(c_ip), nf)
        return conn[:nf]

    def write_surface_vtu(self, path, fields, z=0.0, high_order=True):
        """One ParaView piece (.vtu) of the free-surface fields {name: array[n_surf]} (pv_fs.Save(), :505-514)."""
        names = (C.c_char_p * max(1, len(fields)))(*[k.encode() for k in fields])
        arrs = [np.ascontiguousarray(v, dtype=np.float64) for v in fields.values()]
        ptrs = (c_dp * max(1, len(fields)))(*[a.ctypes.data_as(c_dp) for a in arrs])
        _check(lib.lpf_write_surface_vtu(self.h, str(path).encode(), float(z), len(fields), names, ptrs, int(high_order)), "lpf_write_surface_vtu")

    def node_coordinates(self):
        xyz = np.zeros((self.ndof, 3))
        if self.h is None:
            import ctypes as _C  # arrays-only space: evaluate the trilinear map at the GLL lattice here
            t = basis_tables(self.order)["nodes"]
            D = self.order + 1
            lat = np.array([[t[i], t[j], t[k]] for k in range(D) for j in range(D) for i in range(D)])
            x, y, z = lat[:, 0], lat[:, 1], lat[:, 2]
            N = np.stack([(1 - x) * (1 - y) * (1 - z), x * (1 - y) * (1 - z), (1 - x) * y * (1 - z), x * y * (1 - z),
                          (1 - x) * (1 - y) * z, x * (1 - y) * z, (1 - x) * y * z, x * y * z], axis=1)
            X = np.einsum("nc,ecd->end", N, self.corners)
            xyz[self.gather.reshape(-1)] = X.reshape(-1, 3)
            return xyz
        _check(lib.lpf_space_node_coordinates(self.h, xyz.ctypes.data_as(c_dp)), "node_coordinates")
        return xyz

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib.lpf_space_destroy(self.h)
            self.h = None


def _ptr(t):
    """Device pointer of a torch CUDA tensor (float64, contiguous) or a raw int."""
    if isinstance(t, int):
        return C.c_void_p(t)
    assert t.is_cuda and t.is_contiguous() and t.dtype.is_floating_point and t.element_size() == 8, "need contiguous cuda f64"
    return C.c_void_p(t.data_ptr())


class Context:
    """Device-resident operator + solver of one rank: what `rhs_linear` owns in the reference
    (a_loc_cach, jacobi, CGSolver; Solvers/PF_linear_par_partial.cpp:36-128)."""

    def __init__(self, space: Space, device=0, stream=None):
        self.space = space
        self.h = lib.lpf_create(C.byref(space.desc), device, C.c_void_p(stream) if stream else None)
        if not self.h:
            raise LpfError("lpf_create failed: " + last_error())
        self.ndof, self.nsurf = space.ndof, space.nsurf

    def set_option(self, name, value):
        _check(lib.lpf_set_option(self.h, name.encode(), int(value)), "lpf_set_option")

    def comm_init(self, id128: bytes):
        buf = C.create_string_buffer(id128, 128)
        _check(lib.lpf_comm_init(self.h, buf), "lpf_comm_init")

    def p2p_connect(self, dist=None):
        """Peer-memory exchange between the ranks of a torch.distributed job (one process per GPU): gathers the
        64-byte CUDA IPC handles of all mailboxes and connects.  After this no NCCL call is made by the solver."""
        import torch
        if dist is None:
            import torch.distributed as dist
        h = C.create_string_buffer(64)
        _check(lib.lpf_p2p_export(self.h, h, None, None), "lpf_p2p_export")
        mine = torch.frombuffer(bytearray(h.raw), dtype=torch.uint8).cuda()
        world = dist.get_world_size()
        allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(allh, mine)
        buf = b"".join(t.cpu().numpy().tobytes() for t in allh)
        _check(lib.lpf_p2p_connect(self.h, C.create_string_buffer(buf, len(buf)), None, None, 0), "lpf_p2p_connect")
        dist.barrier()

    def p2p_error(self):
        return lib.lpf_p2p_error(self.h)

    def pa_setup(self):
        _check(lib.lpf_pa_setup(self.h), "lpf_pa_setup")

    def pa_qdata(self, out):
        _check(lib.lpf_pa_qdata(self.h, _ptr(out)), "lpf_pa_qdata")

    def pa_apply_E(self, xE, yE):
        _check(lib.lpf_pa_apply_E(self.h, _ptr(xE), _ptr(yE)), "lpf_pa_apply_E")

    def apply_L(self, x, y):
        _check(lib.lpf_apply_L(self.h, _ptr(x), _ptr(y)), "lpf_apply_L")

    def apply_T(self, x, y):
        _check(lib.lpf_apply_T(self.h, _ptr(x), _ptr(y)), "lpf_apply_T")

    def apply_T_host(self, x_host, y_host):
        """x_host/y_host: pinned torch CPU tensors or numpy arrays (float64)."""
        _check(lib.lpf_apply_T_host(self.h, C.c_void_p(_hostptr(x_host)), C.c_void_p(_hostptr(y_host))), "lpf_apply_T_host")

    def diag(self, out):
        _check(lib.lpf_diag(self.h, _ptr(out)), "lpf_diag")

    def ntrue(self, which=0):
        return lib.lpf_ntrue(self.h, which)

    def prolong(self, xT, xL, which=0):
        _check(lib.lpf_prolong(self.h, which, _ptr(xT), _ptr(xL)), "lpf_prolong")

    def restrict(self, xL, xT, which=0):
        _check(lib.lpf_restrict(self.h, which, _ptr(xL), _ptr(xT)), "lpf_restrict")

    def pa_diag_E(self, out):
        _check(lib.lpf_pa_diag_E(self.h, _ptr(out)), "lpf_pa_diag_E")

    def jacobi_setup(self):
        _check(lib.lpf_jacobi_setup(self.h), "lpf_jacobi_setup")

    def jacobi_dinv(self, out):
        _check(lib.lpf_jacobi_dinv(self.h, _ptr(out)), "lpf_jacobi_dinv")

    def pcg(self, B, X, rel_tol=1e-12, abs_tol=0.0, max_iter=1000):
        info = PcgInfo()
        _check(lib.lpf_pcg(self.h, _ptr(B), _ptr(X), rel_tol, abs_tol, max_iter, C.byref(info)), "lpf_pcg")
        return info

    def laplace_solve(self, phi, rel_tol=1e-12, abs_tol=0.0, max_iter=1000):
        info = PcgInfo()
        _check(lib.lpf_laplace_solve(self.h, _ptr(phi), rel_tol, abs_tol, max_iter, C.byref(info)), "lpf_laplace_solve")
        return info

    def surface_dz(self, phi, wt):
        _check(lib.lpf_surface_dz(self.h, _ptr(phi), _ptr(wt)), "lpf_surface_dz")

    def rhs_setup(self, prm: RhsParams, cgen=None, cabs=None):
        cg = np.ascontiguousarray(cgen, dtype=np.float64) if cgen is not None else None
        ca = np.ascontiguousarray(cabs, dtype=np.float64) if cabs is not None else None
        _check(lib.lpf_rhs_setup(self.h, C.byref(prm), cg.ctypes.data_as(c_dp) if cg is not None else None,
                                 ca.ctypes.data_as(c_dp) if ca is not None else None), "lpf_rhs_setup")

    def rhs_set_cabsy(self, cabsy):
        ca = np.ascontiguousarray(cabsy, dtype=np.float64) if cabsy is not None else None
        _check(lib.lpf_rhs_set_cabsy(self.h, ca.ctypes.data_as(c_dp) if ca is not None else None), "lpf_rhs_set_cabsy")

    def envelope_reset(self):
        _check(lib.lpf_envelope_reset(self.h), "lpf_envelope_reset")

    def envelope_update(self, state):
        _check(lib.lpf_envelope_update(self.h, _ptr(state)), "lpf_envelope_update")

    def envelope_get(self, scale=1.0):
        env = np.zeros(max(1, lib.lpf_nsurf(self.h)))
        _check(lib.lpf_envelope_get(self.h, env.ctypes.data_as(c_dp), float(scale)), "lpf_envelope_get")
        return env[:lib.lpf_nsurf(self.h)]

    def rhs(self, t, state, dstate):
        _check(lib.lpf_rhs(self.h, float(t), _ptr(state), _ptr(dstate)), "lpf_rhs")

    def rk4_step(self, state, t, dt):
        tt = C.c_double(t)
        _check(lib.lpf_rk4_step(self.h, _ptr(state), C.byref(tt), float(dt)), "lpf_rk4_step")
        return tt.value

    def rk4_step_host(self, state_host, t, dt):
        tt = C.c_double(t)
        _check(lib.lpf_rk4_step_host(self.h, C.c_void_p(_hostptr(state_host)), C.byref(tt), float(dt)), "lpf_rk4_step_host")
        return tt.value

    def last_solve_info(self):
        arr = (PcgInfo * 4)()
        n = C.c_int(0)
        _check(lib.lpf_last_solve_info(self.h, arr, C.byref(n)), "lpf_last_solve_info")
        return [arr[i] for i in range(n.value)]

    def time_apply(self, x, y, reps):
        ms_t, ms_k, nl = C.c_float(0), C.c_float(0), C.c_long(0)
        _check(lib.lpf_time_apply(self.h, _ptr(x), _ptr(y), reps, C.byref(ms_t), C.byref(ms_k), C.byref(nl)), "lpf_time_apply")
        return ms_t.value, ms_k.value, nl.value

    def sync(self):
        _check(lib.lpf_sync(self.h), "lpf_sync")

    @property
    def launches(self):
        return lib.lpf_launch_count(self.h)

    @property
    def affine_active(self):
        return bool(lib.lpf_affine_active(self.h))

    @property
    def device_bytes(self):
        return lib.lpf_device_bytes(self.h)

    def close(self):
        if getattr(self, "h", None) and lib is not None:
            lib.lpf_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


def _hostptr(a):
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    assert not a.is_cuda and a.is_contiguous() and a.element_size() == 8
    return a.data_ptr()


def maccamy_fuchs(k, a, r, phi, tol=1e-10, max_iter=400, robust=False):
    """MacCamy-Fuchs envelope |eta|_max / (H/2) (cylinder-exact.cpp:53-115) through the C-ABI.  Default: the reference's
    stopping rule verbatim; robust=True: the phi-independent rule (see lpf_b200.h)."""
    r, phi = np.broadcast_arrays(np.asarray(r, dtype=np.float64), np.asarray(phi, dtype=np.float64))
    tol = -abs(tol) if robust else abs(tol)
    return np.array([lib.lpf_maccamy_fuchs(float(k), float(a), float(ri), float(pi), float(tol), int(max_iter))
                     for ri, pi in zip(r.ravel(), phi.ravel())]).reshape(r.shape)


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib.lpf_comm_unique_id(buf), "lpf_comm_unique_id")
    return buf.raw


def wave_params(H=0.01, g=9.81, lam=1.0, kh=1.0, theta=0.0):
    """'lambda-mode' wave parameters of the drivers (Solvers/PF_linear_par_partial.cpp:298-306)."""
    k = 2.0 * np.pi / lam
    c = float(np.sqrt((g / k) * np.tanh(kh)))
    T = lam / c
    return dict(H=H, g=g, lam=lam, k=k, kh=kh, cwave=c, T=T, omega=2.0 * np.pi / T,
                kx_dir=float(np.cos(theta)), ky_dir=float(np.sin(theta)))


def make_rhs_params(w, tau=0.0, use_relaxation=False, rel_tol=1e-12, abs_tol=0.0, max_iter=1000, n_ramp=3.0):
    return RhsParams(g=w["g"], H=w["H"], omega=w["omega"], k=w["k"], kx_dir=w["kx_dir"], ky_dir=w["ky_dir"],
                     cwave=w["cwave"], kh=w["kh"], T=w["T"], tau=tau, n_ramp=n_ramp,
                     use_relaxation=int(use_relaxation), rel_tol=rel_tol, abs_tol=abs_tol, max_iter=max_iter)
