#!/bin/bash
# round 2, call 30: one coefficient-table copy per stage (TABS = 1) at orders 4 and 5 (variant 43): wider uniform loads, 6-8 % fewer
# instructions, 12-18 more registers
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "alternative_variants and 43" > gpurun_out/r02_job30_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job30_pytest.log
tail -3 gpurun_out/r02_job30_pytest.log
timeout 900 python tools/sweep.py --verbose --orders 4,5 --variants 0,43,0,43,0,43 --refine-low 2 --refine-high 2 > gpurun_out/r02_sweep_tabs_low.txt 2>&1; cat gpurun_out/r02_sweep_tabs_low.txt
python tools/pcg_small_ab.py . 1 apply_variant=0; python tools/pcg_small_ab.py . 1 apply_variant=43
