"""ctypes loader of the C oracle (oracle/pa_oracle.c) -- TEST INFRASTRUCTURE ONLY (see pa_oracle.c header)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblpf_oracle.so")


def _cpu_sig():
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return hashlib.md5(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def _built_for_this_cpu():
    try:
        return open(_SO + ".cpu").read().strip() == _cpu_sig()
    except OSError:
        return False


# compiled -march=native: rebuild when missing or built on another CPU (the .so travels from the build container to the GPU box)
if not os.path.exists(_SO) or not _built_for_this_cpu():
    subprocess.check_call(["make", "-B", "-C", _HERE], stdout=subprocess.DEVNULL)
lib = C.CDLL(_SO)
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
lib.lpf_or_max_threads.restype = C.c_int


def dp(a):
    return a.ctypes.data_as(_dp)


def ip(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def set_threads(n):
    lib.lpf_or_set_threads(int(n))


def max_threads():
    return lib.lpf_or_max_threads()


class COperator:
    """One-rank PA operator on plain arrays (gather [ne,D^3] int32, corners [ne,8,3])."""

    def __init__(self, p, corners, gather, ndof, basis):
        self.p, self.ne, self.ndof = p, gather.shape[0], ndof
        self.D3 = (p + 1) ** 3
        self.Q = p + 2
        self.B = np.ascontiguousarray(basis["B"], dtype=np.float64)
        self.G = np.ascontiguousarray(basis["G"], dtype=np.float64)
        self.Dhat = np.ascontiguousarray(basis["Dhat"], dtype=np.float64)
        self.nodes = np.ascontiguousarray(basis["nodes"], dtype=np.float64)
        self.corners = np.ascontiguousarray(corners, dtype=np.float64)
        self.gather = np.ascontiguousarray(gather, dtype=np.int32)
        self.qd = np.zeros((self.ne, 6, self.Q ** 3))
        qp = np.ascontiguousarray(basis["qpts"], dtype=np.float64)
        qw = np.ascontiguousarray(basis["qwts"], dtype=np.float64)
        lib.lpf_or_setup(self.ne, self.Q, dp(self.corners), dp(qp), dp(qw), dp(self.qd))
        # transposed map (ElementRestriction offsets/indices)
        flat = self.gather.reshape(-1)
        order = np.argsort(flat, kind="stable").astype(np.int32)
        counts = np.bincount(flat, minlength=ndof)
        self.offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        self.indices = np.ascontiguousarray(order)

    def mult(self, x):
        y = np.zeros(self.ndof)
        x = np.ascontiguousarray(x, dtype=np.float64)
        lib.lpf_or_mult(self.ne, self.p, self.ndof, dp(self.B), dp(self.G), dp(self.qd), ip(self.gather),
                        ip(self.offsets), ip(self.indices), dp(x), dp(y))
        return y

    def mult_n(self, x, reps):
        """reps applies with the work buffers allocated once (timing entry of the CPU baseline); returns the last y."""
        y = np.zeros(self.ndof)
        x = np.ascontiguousarray(x, dtype=np.float64)
        lib.lpf_or_mult_n(self.ne, self.p, self.ndof, dp(self.B), dp(self.G), dp(self.qd), ip(self.gather),
                          ip(self.offsets), ip(self.indices), dp(x), dp(y), int(reps))
        return y

    def apply_E(self, xE):
        xE = np.ascontiguousarray(xE, dtype=np.float64)
        yE = np.zeros_like(xE)
        lib.lpf_or_apply_E(self.ne, self.p, dp(self.B), dp(self.G), dp(self.qd), dp(xE), dp(yE))
        return yE

    def diag(self):
        dE = np.zeros((self.ne, self.D3))
        lib.lpf_or_diag_E(self.ne, self.p, dp(self.B), dp(self.G), dp(self.qd), dp(dE))
        d = np.zeros(self.ndof)
        np.add.at(d, self.gather.reshape(-1), dE.reshape(-1))
        return d

    def pcg(self, ess, dinv, x, rel_tol, abs_tol, max_iter):
        ess = np.ascontiguousarray(ess, dtype=np.int32)
        dinv = np.ascontiguousarray(dinv, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64).copy()
        info = np.zeros(5)
        lib.lpf_or_pcg(self.ne, self.p, self.ndof, dp(self.B), dp(self.G), dp(self.qd), ip(self.gather), ip(self.offsets),
                       ip(self.indices), len(ess), ip(ess), dp(dinv), dp(x), C.c_double(rel_tol), C.c_double(abs_tol),
                       int(max_iter), dp(info))
        return x, dict(iterations=int(info[0]), converged=bool(info[1]), final_norm=info[2], initial_norm=info[3],
                       applies=int(info[4]))

    def deriv_z(self, phi, elems=None):
        w = np.zeros(self.ndof)
        cnt = np.zeros(self.ndof)
        phi = np.ascontiguousarray(phi, dtype=np.float64)
        el = np.ascontiguousarray(elems, dtype=np.int32) if elems is not None else None
        lib.lpf_or_deriv_z(len(el) if el is not None else self.ne, ip(el), self.p, dp(self.nodes), dp(self.Dhat),
                           dp(self.corners), ip(self.gather), dp(phi), dp(w), dp(cnt))
        return w, cnt
