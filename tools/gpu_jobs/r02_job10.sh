#!/bin/bash
# round 2, tenth GPU call: the apply kernels without the end-of-batch barrier: parity of every variant, sweep of all orders,
# PCG leg; grid.sync vs graph-launch microbenchmark
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_parity.py -m gpu -q -k "variants or all_orders or ragged or c5 or c1 or affine" > gpurun_out/r02_job10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job10_pytest.log
tail -5 gpurun_out/r02_job10_pytest.log
timeout 600 python tools/sweep.py --orders 1,2,3,4 --variants 0 --refine-low 2 > gpurun_out/r02_sweep_nobar_low.txt 2>&1; cat gpurun_out/r02_sweep_nobar_low.txt
timeout 600 python tools/sweep.py --orders 5,6,7,8 --variants 0 --refine-high 1 > gpurun_out/r02_sweep_nobar_high.txt 2>&1; cat gpurun_out/r02_sweep_nobar_high.txt
python tools/pcg_small_ab.py . 1; python tools/pcg_small_ab.py . 0
tools/microbench/grid_sync > gpurun_out/r02_grid_sync.txt 2>&1; cat gpurun_out/r02_grid_sync.txt
