"""Race-freedom of the apply kernels' shared-memory stage buffers, checked on the CPU from the stride tables in
csrc/apply_cfg.cuh (compute-sanitizer's racecheck is not available on the GPU pool).

The kernels (csrc/pa_apply_eo.cuh, pa_apply_tma.cuh) run five stages per element batch with CTA-wide barriers after the X, Y, Z
and Yt stages and NONE after the Xt stage.  Between two barriers no thread may write a word that another thread reads or
writes.  With separate A / B buffers that is the classic argument; with the aliased layout (LAY = 1: the A buffer lives inside
the B buffer, orders 7-9 by default) it rests on "a Y-stage thread overwrites only words it has read itself" -- which depends on
the strides, so it is enumerated here word by word for every order, layout and elements-per-CTA the library instantiates."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "master-thesis-lpf-in-mfem_b200", "csrc", "apply_cfg.cuh")


def _table(name, cols):
    src = open(CFG).read()
    body = src[src.index(name):]
    m = re.search(r"constexpr int T\[11\]\[%d\] = \{(.*?)\};" % cols, body, re.S)
    rows = re.findall(r"\{([^{}]*)\}", m.group(1))
    return [[int(v) for v in r.split(",")] for r in rows]


STRIDE = _table("constexpr int lpf_smem_stride", 4)      # {SAY, SAZ, SBZ, PAD}
ALIAS = _table("constexpr int lpf_smem_alias", 3)        # {SBY, SBZ, PAD}


def layout(p, lay):
    D, Q = p + 1, p + 2
    if lay == 1:
        SBY, SBZ, pad = ALIAS[p]
        SBA = D * SBZ
        return dict(D=D, Q=Q, SAY=SBY, SAZ=SBZ, SAA=SBA, SBY=SBY, SBZ=SBZ, SBA=SBA, OFFB=0, ES=3 * SBA + pad)
    SAY, SAZ, SBZ, pad = STRIDE[p]
    SAA, SBA = D * SAZ, D * SBZ
    return dict(D=D, Q=Q, SAY=SAY, SAZ=SAZ, SAA=SAA, SBY=Q, SBZ=SBZ, SBA=SBA, OFFB=2 * SAA, ES=2 * SAA + 3 * SBA + pad)


def accesses(p, E, lay):
    """per phase (= code between two CTA barriers): list of (thread, reads, writes) word sets"""
    L = layout(p, lay)
    D, Q = L["D"], L["Q"]
    LX, LY, LZ = D * D, D * Q, Q * Q
    NT = E * LZ

    def a_row(e, dz, dy):        # the X / Xt thread's line: A[k][dz][dy][0..Q)
        base = e * L["ES"] + dz * L["SAZ"] + dy * L["SAY"]
        return {base + k * L["SAA"] + q for k in range(2) for q in range(Q)}

    def a_col(e, dz, qx):        # the Y / Yt thread's words of A: A[k][dz][0..D)[qx]
        base = e * L["ES"] + dz * L["SAZ"] + qx
        return {base + k * L["SAA"] + i * L["SAY"] for k in range(2) for i in range(D)}

    def b_col_y(e, dz, qx):      # the Y / Yt thread's words of B: B[k][dz][0..Q)[qx]
        base = e * L["ES"] + L["OFFB"] + dz * L["SBZ"] + qx
        return {base + k * L["SBA"] + q * L["SBY"] for k in range(3) for q in range(Q)}

    def b_col_z(e, q2):          # the Z thread's words of B: B[k][0..D)[qy][qx]
        qy, qx = divmod(q2, Q)
        base = e * L["ES"] + L["OFFB"] + qy * L["SBY"] + qx
        return {base + k * L["SBA"] + i * L["SBZ"] for k in range(3) for i in range(D)}

    xt_x, y, z, yt = [], [], [], []
    for t in range(NT):
        if t < E * LX:
            e, l = divmod(t, LX); dz, dy = divmod(l, D)
            w = a_row(e, dz, dy)
            xt_x.append((t, w, w))                       # Xt(b) reads the line, X(b + 1) writes it: no barrier in between
        if t < E * LY:
            e, l = divmod(t, LY); dz, qx = divmod(l, Q)
            y.append((t, a_col(e, dz, qx), b_col_y(e, dz, qx)))
            yt.append((t, b_col_y(e, dz, qx), a_col(e, dz, qx)))
        e, q2 = divmod(t, LZ)
        w = b_col_z(e, q2)
        z.append((t, w, w))
    return {"Xt+X": xt_x, "Y": y, "Z": z, "Yt": yt}, L, NT


# (order, elements per CTA, layout) the library instantiates: defaults (apply_order.cu launch_default), the alternative
# pairs of the tuning variants and the layout experiment (variants 40 / 41)
CASES = ([(1, 16, 0), (2, 8, 0), (3, 8, 0), (3, 5, 0), (4, 3, 0), (4, 2, 0), (5, 3, 0), (5, 2, 0), (6, 2, 0), (7, 1, 0), (8, 1, 0),
          (8, 2, 0), (9, 1, 0), (10, 1, 0)]
         + [(5, 3, 1), (6, 2, 1), (7, 1, 1), (8, 1, 1), (9, 1, 1)])


@pytest.mark.parametrize("p,E,lay", CASES)
def test_stage_buffers_are_race_free(p, E, lay):
    phases, L, NT = accesses(p, E, lay)
    for name, acc in phases.items():
        owner_w, owner_r = {}, {}
        for t, reads, writes in acc:
            for w in writes:
                assert owner_w.setdefault(w, t) == t, f"p={p} E={E} lay={lay} {name}: word {w} written by threads {owner_w[w]} and {t}"
            for w in reads:
                owner_r.setdefault(w, set()).add(t)
        for w, t in owner_w.items():
            others = owner_r.get(w, set()) - {t}
            assert not others, f"p={p} E={E} lay={lay} {name}: word {w} written by thread {t}, read by {sorted(others)[:4]}"
    # everything stays inside the element's slice of the work buffer
    top = max(max(w for _, r, wr in acc for w in (r | wr)) for acc in phases.values())
    assert top < E * L["ES"]


@pytest.mark.parametrize("p", range(1, 11))
def test_aliased_layout_keeps_a_inside_the_y_threads_own_words(p):
    """LAY = 1: the words a Y-stage thread reads from A are a subset of the words it writes to B (and vice versa in Yt)."""
    phases, _, _ = accesses(p, 1, 1)
    for t, reads, writes in phases["Y"]:
        assert reads <= writes
    for t, reads, writes in phases["Yt"]:
        assert writes <= reads
