#!/bin/bash
# round 2, call 13: aliased stage buffers, early q-data release with a relaxed counter, refills without proxy fences
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "alternative_variants" > gpurun_out/r02_job13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job13_pytest.log
tail -5 gpurun_out/r02_job13_pytest.log
timeout 900 python tools/sweep.py --verbose --orders 4 --variants 0,40,41,42,43,44,45 --refine-low 2 > gpurun_out/r02_sweep_nofence_p4.txt 2>&1; grep -v "^lpf" gpurun_out/r02_sweep_nofence_p4.txt
timeout 900 python tools/sweep.py --verbose --orders 5,6,7,8,9 --variants 0,40,41,42,43,44,45 --refine-high 1 > gpurun_out/r02_sweep_nofence_r1.txt 2>&1; grep -v "^lpf" gpurun_out/r02_sweep_nofence_r1.txt
timeout 900 python tools/sweep.py --orders 7,8 --variants 0,40,41,42,43,45 --refine-high 2 > gpurun_out/r02_sweep_nofence_r2.txt 2>&1; cat gpurun_out/r02_sweep_nofence_r2.txt
