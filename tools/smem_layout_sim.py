"""Offline shared-memory bank-conflict model of pa_apply_kernel's five stages.

For every unrolled LDS/STS instruction of a stage it lists the 8-byte word index each active lane touches,
splits the warp into two half-warps (64-bit accesses are served 16 lanes at a time) and counts wavefronts as
max over the 16 bank pairs of the number of distinct words.  Used to pick the per-order stride table
lpf_smem_stride() in csrc/pa_kernels.cuh; `python tools/smem_layout_sim.py 4 2` prints the wavefronts per element of the OLD
formula-padded layout ("current") and the best layouts of the searched family.
"""
import itertools
import sys


def wavefronts(addrs):
    """addrs: list of (lane, word) for active lanes of one warp; returns wavefront count."""
    tot = 0
    for half in (0, 1):
        banks = {}
        for lane, w in addrs:
            if lane // 16 == half:
                banks.setdefault(w % 16, set()).add(w)
        if banks:
            tot += max(len(v) for v in banks.values())
    return tot


def simulate(P, E, SAY, SAZ, SBZ, ES_pad, alias=False):
    """alias: the A buffer lives INSIDE the B buffer (A[k][dz][dy][qx] = B[k][dz][qy = dy][qx], round 2): SAY = Q, SAZ = SBZ."""
    D, Q = P + 1, P + 2
    LX, LY, LZ = D * D, D * Q, Q * Q
    NT = E * LZ
    SAA, SBA = D * SAZ, D * SBZ
    ES = (2 * SAA + 3 * SBA) + ES_pad
    OFFB = 2 * SAA
    if alias:
        assert SAY == Q and SAZ == SBZ
        ES = 3 * SBA + ES_pad
        OFFB = 0
    total, ideal = 0, 0

    def run(instrs):
        nonlocal total, ideal
        for ins in instrs:                      # ins: dict tid -> word
            for w0 in range(0, NT, 32):
                a = [(t - w0, ins[t]) for t in range(w0, min(w0 + 32, NT)) if t in ins]
                if a:
                    total += wavefronts(a)
                    ideal += (1 if all(l < 16 for l, _ in a) or all(l >= 16 for l, _ in a) else 2)

    # X stage stores: thread (e, dz, dy) writes a[q], a[SAA+q]
    ins = []
    for arr in range(2):
        for q in range(Q):
            d = {}
            for t in range(E * LX):
                e, l = divmod(t, LX); dz, dy = divmod(l, D)
                d[t] = e * ES + arr * SAA + dz * SAZ + dy * SAY + q
            ins.append(d)
    run(ins)
    sx = total
    # Y stage loads (dy = i) and stores (qy)
    ins = []
    for arr in range(2):
        for i in range(D):
            d = {}
            for t in range(E * LY):
                e, l = divmod(t, LY); dz, qx = divmod(l, Q)
                d[t] = e * ES + arr * SAA + dz * SAZ + i * SAY + qx
            ins.append(d)
    for arr in range(3):
        for q in range(Q):
            d = {}
            for t in range(E * LY):
                e, l = divmod(t, LY); dz, qx = divmod(l, Q)
                d[t] = e * ES + OFFB + arr * SBA + dz * SBZ + q * Q + qx
            ins.append(d)
    run(ins)
    sy = total - sx
    # Z stage loads + stores
    ins = []
    for rep in range(2):
        for arr in range(3):
            for i in range(D):
                d = {}
                for t in range(E * LZ):
                    e, q2 = divmod(t, LZ)
                    d[t] = e * ES + OFFB + arr * SBA + i * SBZ + q2
                ins.append(d)
    run(ins)
    sz = total - sx - sy
    # Yt loads = Y stores pattern, Yt stores = Y loads pattern; Xt loads = X stores pattern
    return dict(total=total + sy + sx, ideal=ideal + 0, x=sx, y=sy, z=sz, ES=ES,
                per_elem=(total + sy + sx) / E, smem_bytes=E * ES * 8)


if __name__ == "__main__":
    P, E = int(sys.argv[1]), int(sys.argv[2])
    D, Q = P + 1, P + 2
    if len(sys.argv) > 3 and sys.argv[3] == "alias":      # aliased layout: search (SBZ, pad) only
        best = []
        for SBZ in range(Q * Q, Q * Q + 33):
            for pad in range(0, 16):
                r = simulate(P, E, Q, SBZ, SBZ, pad, alias=True)
                best.append((r["per_elem"], r["smem_bytes"], Q, SBZ, SBZ, pad, r["x"], r["y"], r["z"]))
        best.sort()
        for b in best[:10]:
            print(b)
        sys.exit(0)
    def pad_to(v, m):
        while v % 16 != m % 16: v += 1
        return v
    SAY = Q if Q % 2 else Q + 1
    cur = simulate(P, E, SAY, pad_to(D * SAY, Q), pad_to(Q * Q, Q), (2 * D * pad_to(D * SAY, Q) + 3 * D * pad_to(Q * Q, Q) + 1) % 2 == 0 and 1 or 0)
    print("current", cur)
    best = []
    for SAY in range(Q, Q + 4):
        for SAZ in range(D * SAY, D * SAY + 17):
            for SBZ in range(Q * Q, Q * Q + 17):
                for pad in range(0, 16):
                    r = simulate(P, E, SAY, SAZ, SBZ, pad)
                    best.append((r["per_elem"], r["smem_bytes"], SAY, SAZ, SBZ, pad))
    best.sort()
    for b in best[:8]:
        print(b)
