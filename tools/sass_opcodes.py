"""Opcode table of the built apply kernels: `cuobjdump -sass` of every per-order object, counted per kernel.

    python tools/sass_opcodes.py [--all] > profiles/r02_sass_opcodes.md

What to look for: UBLKCP + SYNCS = TMA 1-D bulk copies with mbarrier completion; REDG = fire-and-forget scatter-add (ATOMG
must be 0); DMMA = FP64 tensor cores (0: profiles/r02_fp64_pipes.txt shows why); LDCU vs LDC = coefficients through the
uniform datapath (good) or as per-thread indexed constant loads (the cliff orders 7, 8 fell off in round 1, and what every
default kernel is checked for here).  Default kernels of every order are listed; --all lists every instantiation."""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "master-thesis-lpf-in-mfem_b200", "build")
OPS = ["DFMA", "DADD", "DMUL", "DMMA", "LDCU", "LDC", "R2UR", "UBLKCP", "SYNCS", "REDG", "ATOMG", "LDS", "STS", "BAR", "LDL", "STL"]
# (E, MINB) of the default stored-q-data kernel per order (apply_order.cu launch_default)
DEFAULT = {1: (16, 3), 2: (8, 3), 3: (8, 2), 4: (3, 4), 5: (3, 2), 6: (2, 3), 7: (1, 4), 8: (1, 3), 9: (1, 2), 10: (1, 1)}


def kernels(obj):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", obj], capture_output=True, text=True).stdout
    regs = {}
    for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+) STACK:(\d+)", res):
        regs[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    out, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            if op == "ATOMG" and "F64" not in line:      # integer atomics on the exchange counters are fine
                out[cur]["total"] += 1
                continue
            for o in OPS:
                if op == o or (o == "LDC" and op == "LDC") or (o != "LDC" and op.startswith(o)):
                    out[cur][o] += 1
                    break
            out[cur]["total"] += 1
    return out, regs


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, r))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--all", action="store_true")
    a = ap.parse_args()
    print("# SASS opcode counts of the PA apply kernels (static, per kernel; `python tools/sass_opcodes.py`)\n")
    print("Template arguments: `pa_apply_eo_kernel<P, E, DEN, MINB, AFF, DET, OVL, TABS, LAY, EQ>`, `pa_apply_tma_kernel<P, E, DEN, MINB, DET, OVL>`.")
    print("The batch loop is fully unrolled, so the counts are (prologue/epilogue aside) the instructions one warp issues per batch of E elements.\n")
    print("| kernel | regs | stack | " + " | ".join(OPS) + " | total |")
    print("|---|---|---|" + "---|" * (len(OPS) + 1))
    bad = []
    for p in range(1, 11):
        obj = os.path.join(BUILD, f"apply_p{p}.o")
        if not os.path.exists(obj):
            continue
        ks, regs = kernels(obj)
        dm = demangle(list(ks))
        for k in sorted(ks, key=lambda n: dm[n]):
            name = dm[k].replace("void ", "").replace("(ApplyKArgs)", "")
            name = re.sub(r"\(double const\*.*", "", name)
            E, MINB = DEFAULT[p]
            tabs, tabs_ovl = (1 if p >= 7 else 0), (1 if p >= 6 else 0)
            lay, eq = (1 if 7 <= p <= 9 else 0), ("true" if p == 7 else "false")      # apply_order.cu launch_default
            pat = (rf"pa_apply_eo_kernel<{p}, {E}, (true|false), {MINB}, false, (true|false), (false, {tabs}|true, {tabs_ovl}), {lay}, {eq}>" if p >= 3
                   else rf"pa_apply_tma_kernel<{p}, {E}, (true|false), {MINB}, (true|false), (true|false)>")
            is_default = re.fullmatch(pat, name) is not None
            if not (a.all or is_default or "evec" in name):
                continue
            c = ks[k]
            r = regs.get(k, (0, 0))
            print(f"| `{name}` | {r[0]} | {r[1]} | " + " | ".join(str(c[o]) for o in OPS) + f" | {c['total']} |")
            if is_default and "evec" not in name and p >= 3 and c["LDC"] > c["LDCU"]:
                bad.append(name)
            if c["ATOMG"] or c["DMMA"]:
                bad.append(name + " (ATOMG.F64 / DMMA)")
    print()
    if bad:
        print("**Kernels whose coefficients fell back to per-thread LDC (or that contain ATOMG / DMMA):** " + ", ".join(f"`{b}`" for b in bad))
        sys.exit(1)
    print("All default kernels from order 3 up load their coefficients through the uniform datapath (LDCU > LDC); no ATOMG, no DMMA anywhere.")


if __name__ == "__main__":
    main()
