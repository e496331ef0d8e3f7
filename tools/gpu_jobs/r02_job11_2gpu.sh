#!/bin/bash
# round 2, second 2-GPU call: the final library (no end-of-batch barrier in the single-GPU kernels, pruned variants, PDL off):
# parity for every exchange mode incl. prolong / restrict, multi-GPU pytest, bench at N = 2
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02_job11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job11_pytest.log
tail -5 gpurun_out/r02_job11_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 tools/check_multi_gpu.py --order 4 --mesh tank > gpurun_out/r02_parity2_final.txt 2>&1; echo "parity rc=$?"; cat gpurun_out/r02_parity2_final.txt | tail -22
timeout 600 $TR --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench2_final.log 2>&1; echo "bench2 rc=$?"; tail -1 gpurun_out/r02_bench2_final.log | cut -c1-400
