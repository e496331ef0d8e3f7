#!/bin/bash
# round 2, call 18: order 6 with the compact aliased layout (three CTAs per SM); host-buffer apply with the size-dependent plan
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "alternative_variants and 6-" > gpurun_out/r02_job18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job18_pytest.log
tail -3 gpurun_out/r02_job18_pytest.log
timeout 600 python tools/sweep.py --verbose --orders 6 --variants 0,44,45,46,47 --refine-high 1 > gpurun_out/r02_sweep_p6_r1.txt 2>&1; cat gpurun_out/r02_sweep_p6_r1.txt
timeout 600 python tools/sweep.py --orders 6 --variants 0,44,45,46,47 --refine-high 2 > gpurun_out/r02_sweep_p6_r2.txt 2>&1; cat gpurun_out/r02_sweep_p6_r2.txt
timeout 600 python tools/e2e_host_ab.py --reps 5 2>&1 | head -3
