#!/bin/bash
# round 2, call 14: ncu --set full of the order-7 / order-8 kernels with the aliased stage buffers (4 / 3 CTAs per SM), 262 144 hexes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for cfg in "7 41" "8 40"; do
  set -- $cfg
  python tools/sweep.py --orders $1 --variants $2 --refine-high 2 --reps 5 > gpurun_out/r02b_ncu_plain_p$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:pa_apply -s 3 -c 1 -o gpurun_out/r02b_eo_p$1 \
      python tools/sweep.py --orders $1 --variants $2 --refine-high 2 --reps 5 > gpurun_out/r02b_ncu_p$1.log 2>&1
  ncu -i gpurun_out/r02b_eo_p$1.ncu-rep --page raw --csv > gpurun_out/r02b_eo_p$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02b_eo_p$1.ncu-rep --page source --csv > gpurun_out/r02b_eo_p$1_source.csv 2>/dev/null
  tail -2 gpurun_out/r02b_ncu_plain_p$1.log
done
ls -la gpurun_out/*.ncu-rep
