#!/bin/bash
# round 2, fourth 2-GPU call: duration (+ NVLink bytes) of the kernels that carry the exchange, single-pass ncu; NVLink counters of
# the driver (nvidia-smi nvlink -gt d) around a 2-GPU bench run; multi-GPU pytest with the final library
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29561 --no-python bash tools/ncu_rank.sh time gpurun_out/r02_xchg_time tools/check_multi_gpu.py --order 4 --mesh tank > gpurun_out/r02_xchg_time_run.log 2>&1; echo "ncu time rc=$?"; tail -2 gpurun_out/r02_xchg_time_run.log
timeout 400 $TR --master-port 29562 --no-python bash tools/ncu_rank.sh nvl gpurun_out/r02_xchg_nvl tools/check_multi_gpu.py --order 4 --mesh tank > gpurun_out/r02_xchg_nvl_run.log 2>&1; echo "ncu nvl rc=$?"; tail -2 gpurun_out/r02_xchg_nvl_run.log
nvidia-smi nvlink -gt d > gpurun_out/r02_nvlink_before.txt 2>&1
timeout 600 $TR --master-port 29563 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench2_c.log 2>&1; echo "bench2 rc=$?"
nvidia-smi nvlink -gt d > gpurun_out/r02_nvlink_after.txt 2>&1
tail -1 gpurun_out/r02_bench2_c.log | cut -c1-300
head -12 gpurun_out/r02_nvlink_after.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02_job21_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job21_pytest.log
tail -3 gpurun_out/r02_job21_pytest.log
