#!/bin/bash
# round 2, call 24: where a CG iteration of the SMALL meshes goes (big8 itself, 299 520 dofs): apply variants / grid sizes, and the
# per-kernel durations of one RK4 step under ncu (serialised launches)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for o in "" "apply_variant=30" "apply_variant=20" "max_ctas=456" "max_ctas=342" "max_ctas=296" "max_ctas=148" "pcg_chunk=64"; do
  python tools/pcg_small_ab.py . 0 $o
done > gpurun_out/r02_pcg_small.txt 2>&1
cat gpurun_out/r02_pcg_small.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pa_apply|pcg_' -s 2000 -c 600 --csv --log-file gpurun_out/r02_pcg_small_launches.csv python tools/pcg_small_ab.py . 0 > gpurun_out/r02_pcg_small_ncu.log 2>&1
python - <<'PY'
import csv,collections,re
rows=[r for r in csv.reader(open('gpurun_out/r02_pcg_small_launches.csv')) if len(r)>=15 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    agg[re.sub(r'\(.*','',r[4])[:60]][0]+=1; agg[re.sub(r'\(.*','',r[4])[:60]][1]+=float(r[14])/1e3
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-62s %6d launches  avg %7.2f us"%(k,v[0],v[1]/v[0]))
PY
