#!/bin/bash
# round 2, call 12: aliased stage buffers (A inside B) + early q-data release: parity of the new variants, sweep of orders 5-8
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "alternative_variants or all_orders" > gpurun_out/r02_job12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job12_pytest.log
tail -5 gpurun_out/r02_job12_pytest.log
timeout 900 python tools/sweep.py --verbose --orders 5,6,7,8 --variants 0,40,41,42,43,44 --refine-high 1 > gpurun_out/r02_sweep_alias_r1.txt 2>&1; cat gpurun_out/r02_sweep_alias_r1.txt
timeout 900 python tools/sweep.py --orders 6,7,8 --variants 0,41,42 --refine-high 2 > gpurun_out/r02_sweep_alias_r2.txt 2>&1; cat gpurun_out/r02_sweep_alias_r2.txt
