#!/bin/bash
# round 2, third GPU call: the two-warp-group kernel of orders 7 / 8 (apply_variant 35 / 36): parity, then the sweep
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "alternative_variants or all_orders" > gpurun_out/r02_job3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job3_pytest.log
tail -8 gpurun_out/r02_job3_pytest.log
timeout 900 python tools/sweep.py --orders 6,7,8 --variants 0,32,35,36 --refine-high 1 > gpurun_out/r02_sweep_split_r1.txt 2>&1; cat gpurun_out/r02_sweep_split_r1.txt
timeout 900 python tools/sweep.py --orders 9,10 --variants 0,35 --refine-high 0 > gpurun_out/r02_sweep_p9_10.txt 2>&1; cat gpurun_out/r02_sweep_p9_10.txt
