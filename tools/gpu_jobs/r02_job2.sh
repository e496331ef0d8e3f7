#!/bin/bash
# round 2, second GPU call: the whole -m gpu suite (no -x), then ncu --set full of the order-7 and order-8 kernels with one
# coefficient-table copy per stage (the variants job 1 measured fastest)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_job2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job2_pytest.log
tail -15 gpurun_out/r02_job2_pytest.log
for cfg in "7 32" "8 0"; do
  set -- $cfg
  python tools/sweep.py --orders $1 --variants $2 --refine-high 1 --reps 5 > gpurun_out/r02_ncu_plain_p$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:pa_apply -s 3 -c 1 -o gpurun_out/r02_eo_p$1 \
      python tools/sweep.py --orders $1 --variants $2 --refine-high 1 --reps 5 > gpurun_out/r02_ncu_p$1.log 2>&1
  ncu -i gpurun_out/r02_eo_p$1.ncu-rep --page raw --csv > gpurun_out/r02_eo_p$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_eo_p$1.ncu-rep --page source --csv > gpurun_out/r02_eo_p$1_source.csv 2>/dev/null
  tail -2 gpurun_out/r02_ncu_plain_p$1.log
done
ls -la gpurun_out/*.ncu-rep
