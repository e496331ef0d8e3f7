"""The analytic reference curve of BASELINE config 4 (Solvers/cylinder-exact.cpp restated with scipy)."""
import numpy as np


def test_series_matches_wronskian_closed_form_on_the_cylinder():
    import maccamy_fuchs as mf
    k, a = 2.0 * np.pi, 0.5                       # lambda = 1, a = 0.5 (cylinder-diffraction.cpp:229-238)
    phi = np.linspace(0.0, np.pi, 181)
    e1 = mf.envelope(k, a, np.full_like(phi, a), phi)
    e2 = mf.envelope_on_cylinder_wronskian(k, a, phi)
    assert np.abs(e1 - e2).max() < 1e-9
    assert 1.5 < e1.max() < 2.1 and e1.min() > 0.2           # ka = pi: strong run-up on the weather side
    # symmetric about the wave direction, and the incident wave is recovered far from a thin cylinder
    assert np.allclose(mf.envelope(k, a, np.full(5, a), np.array([0.3, 1.0, 2.0, 2.5, 3.0])),
                       mf.envelope(k, a, np.full(5, a), -np.array([0.3, 1.0, 2.0, 2.5, 3.0])), atol=1e-12)
    far = mf.envelope(k, 1e-3, np.array([5.0, 7.0]), np.array([0.4, 2.0]))
    assert np.abs(far - 1.0).max() < 1e-4


def test_library_series_matches_scipy_restatement(lpf):
    """lpf_maccamy_fuchs (std::cyl_bessel_j / cyl_neumann, product) vs the scipy restatement (oracle)."""
    import maccamy_fuchs as mf
    k, a = 2.0 * np.pi, 0.5
    phi = np.linspace(0.0, np.pi, 37)
    for r in (a, 0.75, 2.0):
        e_lib = lpf.maccamy_fuchs(k, a, np.full_like(phi, r), phi, robust=True)
        e_ref = mf.envelope(k, a, np.full_like(phi, r), phi)
        assert np.abs(e_lib - e_ref).max() < 1e-9
        # default = the reference's own stopping rule, term for term
        assert np.abs(lpf.maccamy_fuchs(k, a, np.full_like(phi, r), phi) - mf.envelope_reference_rule(k, a, np.full_like(phi, r), phi)).max() < 1e-12
    # other ka
    for ka in (0.3, 1.0, 6.0):
        e_lib = lpf.maccamy_fuchs(ka / a, a, np.full_like(phi, a), phi, robust=True)
        assert np.abs(e_lib - mf.envelope_on_cylinder_wronskian(ka / a, a, phi, nterms=80)).max() < 1e-8


def test_reference_stopping_rule_quirk_at_right_angles(lpf):
    """cylinder-exact.cpp's rule |Re(term)| < tol twice in a row fires at m = 1 for phi = pi/2.  The library's DEFAULT is that
    rule (identical results to the reference, quirk included); robust=True is the corrected rule."""
    import maccamy_fuchs as mf
    k, a = 2.0 * np.pi, 0.5
    phi = np.array([0.0, 0.4, 1.0, 2.0, 2.7, np.pi])
    ref = mf.envelope_reference_rule(k, a, np.full_like(phi, a), phi)
    assert np.abs(lpf.maccamy_fuchs(k, a, np.full_like(phi, a), phi) - ref).max() < 1e-12
    assert np.abs(lpf.maccamy_fuchs(k, a, np.full_like(phi, a), phi, robust=True) - ref).max() < 1e-8
    bad = mf.envelope_reference_rule(k, a, np.array([a]), np.array([np.pi / 2]))[0]
    assert abs(lpf.maccamy_fuchs(k, a, np.array([a]), np.array([np.pi / 2]))[0] - bad) < 1e-12      # default reproduces the reference
    good = lpf.maccamy_fuchs(k, a, np.array([a]), np.array([np.pi / 2]), robust=True)[0]
    assert abs(good - mf.envelope_on_cylinder_wronskian(k, a, np.array([np.pi / 2]))[0]) < 1e-9
    assert abs(bad - good) > 1e-2
