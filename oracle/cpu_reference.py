"""CPU reference timings of the hot path -- TEST / BASELINE INFRASTRUCTURE ONLY.

Used by bench.py's `cpu_baseline` leg and `--impl reference` arm, nowhere else.  Everything here comes from oracle/:
the tank mesh and its H1 numbering from lpf_oracle.build_tank_space, the basis tables from lpf_oracle.make_basis, the
element arithmetic, Jacobi-PCG and GetDerivative from the C restatement of MFEM's CPU partial-assembly path
(pa_oracle.c, OpenMP over elements, compiled -march=native on the machine that runs it).  The product library
(liblpf_b200.so) is never loaded by this module.

MFEM / hypre / MPI are not installable in this image, so this is labelled "CPU restatement of MFEM PA, N cores"
(kind "port"), never "MFEM".
"""
import math
import os
import time

import numpy as np

import lpf_oracle as orc
import oracle_c

H0 = 1.0 / (2.0 * math.pi)


class CpuTank:
    """x-periodic wave tank (Meshes/wave_tank.cpp) of nx x ny x nz hexes on Lx x 0.1 x 1/(2 pi), order p"""

    def __init__(self, nx, ny, nz, p, Lx=1.0, threads=None):
        self.threads = threads or os.cpu_count()
        oracle_c.set_threads(self.threads)
        self.p = p
        self.sp = orc.build_tank_space(nx, ny, nz, Lx, 0.1, H0, p, True)
        bs = self.sp.basis
        self.cop = oracle_c.COperator(p, self.sp.mesh.corners, self.sp.gather.astype(np.int32), self.sp.ndof,
                                      dict(B=bs.B, G=bs.G, Dhat=bs.Dhat, nodes=bs.nodes, qpts=bs.qpts, qwts=bs.qwts))
        self.ne, self.ndof = self.sp.gather.shape[0], self.sp.ndof
        self.ess = np.sort(self.sp.ess).astype(np.int32)
        self._dinv = None

    def describe(self):
        return f"{self.ne} hexes, {self.ndof} dofs, order {self.p}"

    def dinv(self):
        if self._dinv is None:
            d = 1.0 / self.cop.diag()
            d[self.ess] = 1.0
            self._dinv = d
        return self._dinv

    # ---- operator ----
    def time_applies(self, steps, warmup, min_warm_s=0.5):
        """`steps` timed operator applies after `warmup` untimed ones (work buffers allocated once, as in MFEM).  The warm-up
        also lasts at least `min_warm_s` seconds: the OpenMP team and the host's clocks need a few hundred milliseconds of load
        to reach their steady state (seen as a 10x slower first call on small tanks), and the baseline is quoted at that state."""
        x = np.random.default_rng(0).random(self.ndof) - 0.5
        self.cop.mult_n(x, max(1, warmup))
        t_w = time.perf_counter()
        while time.perf_counter() - t_w < min_warm_s:
            self.cop.mult_n(x, 1)
        t0 = time.perf_counter()
        self.cop.mult_n(x, steps)
        return (time.perf_counter() - t0) / steps

    # ---- Jacobi-PCG ----
    def time_pcg(self, rel_tol=1e-12, max_iter=2000):
        """one full Laplace solve of the t = 0 Airy surface potential: (seconds, iterations, applies)"""
        wv = orc.Wave()
        x = np.zeros(self.ndof)
        x[self.sp.surf2vol] = wv.phi_fs(0.0, self.sp.surf_xy[:, 0], self.sp.surf_xy[:, 1])
        dinv = self.dinv()
        t0 = time.perf_counter()
        _, info = self.cop.pcg(self.ess, dinv, x, rel_tol, 0.0, max_iter)
        return time.perf_counter() - t0, info["iterations"], info["applies"]

    # ---- one RK4 step of the ss.cpp / strongscaling.cpp physics (no relaxation zones) ----
    def time_rk4_step(self, dt, rel_tol=1e-12, max_iter=2000):
        wv = orc.Wave()
        sp = self.sp
        ns = len(sp.surf2vol)
        sel = orc.surface_elements(sp).astype(np.int32)
        dinv = self.dinv()
        its = []

        def rhs(state):
            phi = np.zeros(self.ndof)
            phi[sp.surf2vol] = state[ns:]
            X, info = self.cop.pcg(self.ess, dinv, phi, rel_tol, 0.0, max_iter)
            its.append(info["iterations"])
            w, cnt = self.cop.deriv_z(X, sel)
            cnt[cnt == 0] = 1.0
            return np.concatenate([(w / cnt)[sp.surf2vol], -wv.g * state[:ns]])

        x = np.concatenate([wv.eta(0.0, sp.surf_xy[:, 0], sp.surf_xy[:, 1]), wv.phi_fs(0.0, sp.surf_xy[:, 0], sp.surf_xy[:, 1])])
        t0 = time.perf_counter()
        k = rhs(x); y = x + (dt / 2) * k; z = x + (dt / 6) * k
        k = rhs(y); y = x + (dt / 2) * k; z = z + (dt / 3) * k
        k = rhs(y); y = x + dt * k; z = z + (dt / 3) * k
        k = rhs(y); x = z + (dt / 6) * k
        return time.perf_counter() - t0, its
