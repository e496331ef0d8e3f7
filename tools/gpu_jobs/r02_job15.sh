#!/bin/bash
# round 2, call 15: new defaults (orders 7-9 with aliased stage buffers): whole -m gpu suite, PCIe ceiling, bench line, sweep of
# every order on 262 144 hexes with the SM clock / power logged next to it
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_job15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job15_pytest.log
tail -6 gpurun_out/r02_job15_pytest.log
python tools/microbench/pcie_bw.py > gpurun_out/r02_pcie_bw.txt 2>&1; cat gpurun_out/r02_pcie_bw.txt
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,power.limit,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown --format=csv,noheader -lms 100 > gpurun_out/r02_job15_smi.csv 2>&1 &
SMI=$!
timeout 900 python tools/sweep.py --orders 1,2,3,4,5,6,7,8 --variants 0 --refine-low 2 --refine-high 2 > gpurun_out/r02_sweep_final_r2.txt 2>&1; cat gpurun_out/r02_sweep_final_r2.txt
timeout 600 python tools/sweep.py --orders 9,10 --variants 0,30 --refine-high 1 > gpurun_out/r02_sweep_final_p9.txt 2>&1; cat gpurun_out/r02_sweep_final_p9.txt
kill $SMI
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu_b.json 2> gpurun_out/r02_bench_1gpu_b.err; tail -c 3000 gpurun_out/r02_bench_1gpu_b.json
