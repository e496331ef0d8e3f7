#!/bin/bash
# round 2, call 28: plain stores for element-interior nodes in the scatter (variant 42) against the default, orders 3-8
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "alternative_variants and 42" > gpurun_out/r02_job28_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job28_pytest.log
tail -3 gpurun_out/r02_job28_pytest.log
timeout 900 python tools/sweep.py --orders 3,4,5,6,7,8 --variants 0,42,0,42 --refine-low 2 --refine-high 2 > gpurun_out/r02_sweep_interior_store.txt 2>&1; cat gpurun_out/r02_sweep_interior_store.txt
