#!/bin/bash
# round 2, call 20: host-buffer apply, ramped chunk sizes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "host" > gpurun_out/r02_job19_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job19_pytest.log
tail -3 gpurun_out/r02_job19_pytest.log
timeout 600 python tools/e2e_host_ab.py > gpurun_out/r02_e2e_host_ab3.txt 2>&1; cat gpurun_out/r02_e2e_host_ab3.txt
timeout 600 python tools/e2e_host_ab.py --refine 1 > gpurun_out/r02_e2e_host_ab3_r1.txt 2>&1; cat gpurun_out/r02_e2e_host_ab3_r1.txt
