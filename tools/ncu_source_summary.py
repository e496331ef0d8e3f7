"""Summarises an `ncu --page source --csv` dump: warp-instructions by opcode, stall samples by opcode,
shared-memory wavefronts.  Usage: python tools/ncu_source_summary.py src.csv [n_elements]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nel = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
c = {h: i for i, h in enumerate(hdr)}
inst = collections.Counter(); samp = collections.Counter(); thr = collections.Counter(); wf = collections.Counter(); wfi = collections.Counter()
tot_i = tot_s = 0
for r in rows[2:]:
    if len(r) < len(hdr) - 20:
        continue
    src = r[c['Source']].split()
    if not src:
        continue
    op = src[0] if not src[0].startswith('@') else src[1]
    op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LDS', 'STS', 'LDG', 'RED', 'BAR', 'SYNCS')) and '.' in op else '')
    n = float(r[c['Instructions Executed']] or 0); s = float(r[c['# Samples']] or 0)
    inst[op] += n; samp[op] += s; thr[op] += float(r[c['Thread Instructions Executed']] or 0)
    wf[op] += float(r[c['L1 Wavefronts Shared']] or 0); wfi[op] += float(r[c['L1 Wavefronts Shared Ideal']] or 0)
    tot_i += n; tot_s += s
print('total warp-inst %.0f (%.1f per element), samples %.0f' % (tot_i, tot_i / nel, tot_s))
print('%-14s %12s %8s %10s %8s %8s %10s %10s' % ('opcode', 'warp-inst', '%inst', 'per-elem', 'lanes', '%stall', 'smem-wf', 'wf-ideal'))
for op, n in inst.most_common(28):
    print('%-14s %12.0f %7.1f%% %10.1f %8.1f %7.1f%% %10.0f %10.0f' % (op, n, 100 * n / tot_i, n / nel, thr[op] / max(n, 1), 100 * samp[op] / max(tot_s, 1), wf[op], wfi[op]))
