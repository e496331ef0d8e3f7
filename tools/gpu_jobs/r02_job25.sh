#!/bin/bash
# round 2, call 25: update + direction fused into one cooperative kernel (single GPU, vectors that fit the registers): PCG
# parity tests, A/B on big8 (299 520 dofs), big8 r=1 (2.25 M dofs: falls back to two kernels), cylinder-sized tank
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_parity.py -m gpu -q -x -k "pcg or solve or rk4 or c1 or c2 or c3 or c4 or stepper or smoke" > gpurun_out/r02_job25_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job25_pytest.log
tail -4 gpurun_out/r02_job25_pytest.log
for o in "fuse_update_dir=1" "fuse_update_dir=0"; do
  python tools/pcg_small_ab.py . 0 $o
  python tools/pcg_small_ab.py . 1 $o
done 2>&1 | grep "CG iteration" | tee gpurun_out/r02_pcg_fused_ab.txt
