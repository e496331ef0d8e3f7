#!/bin/bash
# round 2, ninth GPU call: the whole -m gpu suite + smoke on the final library, DRAM traffic of the apply kernel for every
# order (ncu, dram__bytes_read/write per launch), ncu --set full of the order-4 kernel, bench + reference arm
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_job9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job9_pytest.log
tail -6 gpurun_out/r02_job9_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_job9_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_job9_smoke.log
for p in 1 2 3 4 5 6 7 8; do
  r=2
  python tools/sweep.py --orders $p --variants 0 --refine-low $r --refine-high $r --reps 3 > gpurun_out/r02_traffic_plain_p$p.log 2>&1 &&
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pa_apply -s 4 -c 1 --csv \
      --log-file gpurun_out/r02_traffic_p$p.csv python tools/sweep.py --orders $p --variants 0 --refine-low $r --refine-high $r --reps 3 > gpurun_out/r02_traffic_ncu_p$p.log 2>&1
  tail -1 gpurun_out/r02_traffic_plain_p$p.log
done
python tools/sweep.py --orders 4 --variants 0 --reps 5 > gpurun_out/r02_ncu_plain_p4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pa_apply -s 3 -c 1 -o gpurun_out/r02_eo_p4 python tools/sweep.py --orders 4 --variants 0 --reps 5 > gpurun_out/r02_ncu_p4.log 2>&1
ncu -i gpurun_out/r02_eo_p4.ncu-rep --page raw --csv > gpurun_out/r02_eo_p4_raw.csv 2>/dev/null
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "reference rc=$?"; cut -c1-400 gpurun_out/r02_bench_reference_arm.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_1gpu.json
