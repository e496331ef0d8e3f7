#!/bin/bash
# round 2, fifth GPU call: adapter executed on the GPU, per-kernel launch list of the bench command (PCG iteration
# breakdown), final order sweep at r=2, DRAM traffic per order
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_adapter.py tests/test_gpu_config_parity.py -m gpu -q -k "adapter or c3" > gpurun_out/r02_job5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job5_pytest.log
tail -6 gpurun_out/r02_job5_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_job5_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_job5_bench_ncu.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/r02_bench_launches.csv > gpurun_out/r02_bench_launches_ncu.txt 2>&1; cat gpurun_out/r02_bench_launches_ncu.txt
timeout 1200 python tools/sweep.py --orders 5,6,7,8 --variants 0 --refine-high 2 > gpurun_out/r02_sweep_high_r2.txt 2>&1; cat gpurun_out/r02_sweep_high_r2.txt
