#!/bin/bash
# round 2, first GPU call: the whole -m gpu suite on the refactored library, the FP64 pipe microbenchmark and the
# kernel-variant sweep of orders 5-8 (one table copy per stage vs the round-1 kernels).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_job1_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_job1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job1_pytest.log
tail -5 gpurun_out/r02_job1_pytest.log
timeout 120 tools/microbench/fp64_pipes > gpurun_out/r02_fp64_pipes.txt 2>&1; cat gpurun_out/r02_fp64_pipes.txt
timeout 600 python tools/sweep.py --orders 1,2,3,4 --variants 0 --refine-low 2 > gpurun_out/r02_sweep_low.txt 2>&1; cat gpurun_out/r02_sweep_low.txt
timeout 900 python tools/sweep.py --orders 5,6,7,8 --variants 0,30,31,32,33,34 --refine-high 1 > gpurun_out/r02_sweep_high_r1.txt 2>&1; cat gpurun_out/r02_sweep_high_r1.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_job1_bench.log 2>&1; tail -2 gpurun_out/r02_job1_bench.log
