"""MacCamy-Fuchs linear diffraction of a plane wave by a vertical circular cylinder -- TEST INFRASTRUCTURE.

Restates the analytic envelope of the reference's Solvers/cylinder-exact.cpp:53-115 (Boost.Math Bessel /
Neumann series, tolerance 1e-10, <= 400 terms) with scipy.special; Boost is absent from this image
(SURVEY.md 8f rank 2).  envelope(k, a, r, phi) = |eta|_max / (H/2), the quantity the reference plots against
the run-up extracted by Solvers/cylinder-diffraction.cpp:410-444,449-593 ("2 eta_max / H").
"""
import numpy as np
from scipy.special import jv, yv


def envelope_reference_rule(k, a, r, phi, tol=1e-10, max_iter=400):
    """Point-by-point, with the reference's stopping rule exactly as written (cylinder-exact.cpp:104-110): stop when
    |Re(term)| of two consecutive terms is below tol, oldterm starting at 0.  NOTE the quirk this documents: at
    phi = pi/2 the m = 1 term vanishes (cos(phi) = 0) and the series stops after its first term."""
    r, phi = np.broadcast_arrays(np.asarray(r, dtype=np.float64), np.asarray(phi, dtype=np.float64))
    out = np.zeros(r.shape)
    ka = k * a
    for idx in np.ndindex(r.shape):
        kr, ph = k * r[idx], phi[idx]
        E = jv(0, kr) - (jv(0, kr) + 1j * yv(0, kr)) * (-jv(1, ka) / complex(-jv(1, ka), -yv(1, ka)))
        old = 0.0
        for m in range(1, max_iter + 1):
            Jmp = 0.5 * (jv(m - 1, ka) - jv(m + 1, ka))
            Hmp = complex(Jmp, 0.5 * (yv(m - 1, ka) - yv(m + 1, ka)))
            if abs(Hmp) < 1e-14:
                continue
            term = 2.0 * np.exp(1j * m * np.pi / 2.0) * (jv(m, kr) - (jv(m, kr) + 1j * yv(m, kr)) * (Jmp / Hmp)) * np.cos(m * ph)
            nxt = term.real
            if np.isnan(nxt):
                break
            E += term
            if abs(nxt) < tol and abs(old) < tol:
                break
            old = nxt
        out[idx] = abs(E)
    return out


def envelope(k, a, r, phi, tol=1e-10, max_iter=400):
    """Vectorised series; stops when the terms are below tol at ALL points (robust form of the rule above)."""
    r = np.asarray(r, dtype=np.float64)
    phi = np.asarray(phi, dtype=np.float64)
    ka, kr = k * a, k * r
    # m = 0 (cylinder-exact.cpp:68-76)
    J0p = -jv(1, ka)
    H0p = complex(-jv(1, ka), -yv(1, ka))
    H0r = jv(0, kr) + 1j * yv(0, kr)
    E = jv(0, kr) - H0r * (J0p / H0p)
    old = np.zeros_like(np.real(E))
    for m in range(1, max_iter + 1):                 # cylinder-exact.cpp:78-111
        Jmp = 0.5 * (jv(m - 1, ka) - jv(m + 1, ka))
        Hmp = complex(Jmp, 0.5 * (yv(m - 1, ka) - yv(m + 1, ka)))
        if abs(Hmp) < 1e-14:
            continue
        Hmr = jv(m, kr) + 1j * yv(m, kr)
        im = np.exp(1j * m * np.pi / 2.0)
        term = 2.0 * im * (jv(m, kr) - Hmr * (Jmp / Hmp)) * np.cos(m * phi)
        nxt = np.real(term)
        if np.any(np.isnan(nxt)):
            break
        E = E + term
        if np.all(np.abs(nxt) < tol) and np.all(np.abs(old) < tol):
            break
        old = nxt
    return np.abs(E)


def envelope_on_cylinder_wronskian(k, a, phi, nterms=60):
    """Independent closed form on r = a: J_m H_m' - J_m' H_m = 2i / (pi ka)  =>
    E(a, phi) = sum_m eps_m i^m  2i / (pi ka H_m'(ka)) cos(m phi)  (textbook MacCamy-Fuchs run-up)."""
    ka = k * a
    phi = np.asarray(phi, dtype=np.float64)
    E = np.zeros_like(phi, dtype=np.complex128)
    for m in range(nterms):
        if m == 0:
            Hp = complex(-jv(1, ka), -yv(1, ka))
        else:
            Hp = complex(0.5 * (jv(m - 1, ka) - jv(m + 1, ka)), 0.5 * (yv(m - 1, ka) - yv(m + 1, ka)))
        eps = 1.0 if m == 0 else 2.0
        E = E + eps * np.exp(1j * m * np.pi / 2.0) * (2j / (np.pi * ka * Hp)) * np.cos(m * phi)
    return np.abs(E)
