#!/bin/bash
# round 2, third 2-GPU call: aliased-layout kernels carrying the exchange hooks (orders 7, 8), multi-GPU pytest, ncu capture
# (duration + NVLink bytes, single pass) of the exchange kernels on both ranks, bench at N = 2
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02_job16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job16_pytest.log
tail -5 gpurun_out/r02_job16_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for p in 4 7 8; do
  timeout 400 $TR --master-port 2954$p tools/check_multi_gpu.py --order $p --mesh tank > gpurun_out/r02_parity2_p$p.txt 2>&1; echo "parity p=$p rc=$?"; tail -8 gpurun_out/r02_parity2_p$p.txt
done
ncu --query-metrics 2>/dev/null | grep -i "^nvl" | head -40 > gpurun_out/r02_ncu_nvl_metrics.txt
timeout 600 $TR --master-port 29551 --no-python bash tools/ncu_rank.sh gpurun_out/r02_nvl tools/check_multi_gpu.py --order 4 --mesh tank > gpurun_out/r02_nvl_run.log 2>&1; echo "ncu 2-rank rc=$?"; tail -3 gpurun_out/r02_nvl_run.log
timeout 600 $TR --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench2_b.log 2>&1; echo "bench2 rc=$?"; tail -1 gpurun_out/r02_bench2_b.log | cut -c1-600
