#!/bin/bash
# round 2, 2-GPU call: N-rank vs 1-rank parity for every exchange mode (overlapped inside the apply kernel = default,
# last-CTA tail, separate LL kernel, NCCL), tank and cylinder meshes; then the bench line at N = 2 for modes 2 and 0
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for fuse in 2 1 0; do
  timeout 300 $TR --master-port 2951$fuse tools/check_multi_gpu.py --order 4 --mesh tank --p2p-fuse $fuse > gpurun_out/r02_parity2_tank_fuse$fuse.txt 2>&1; echo "tank fuse=$fuse rc=$?"
  tail -3 gpurun_out/r02_parity2_tank_fuse$fuse.txt
done
timeout 300 $TR --master-port 29514 tools/check_multi_gpu.py --order 3 --mesh cylinder --p2p-fuse 2 > gpurun_out/r02_parity2_cyl_fuse2.txt 2>&1; echo "cyl fuse=2 rc=$?"; tail -3 gpurun_out/r02_parity2_cyl_fuse2.txt
timeout 300 $TR --master-port 29515 tools/check_multi_gpu.py --order 6 --mesh tank --p2p-fuse 2 > gpurun_out/r02_parity2_tank_p6_fuse2.txt 2>&1; echo "tank p6 fuse=2 rc=$?"; tail -3 gpurun_out/r02_parity2_tank_p6_fuse2.txt
timeout 300 $TR --master-port 29516 tools/check_multi_gpu.py --order 4 --mesh tank --comm nccl > gpurun_out/r02_parity2_tank_nccl.txt 2>&1; echo "tank nccl rc=$?"; tail -3 gpurun_out/r02_parity2_tank_nccl.txt
timeout 600 $TR --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench2_fuse2.log 2>&1; echo "bench2 rc=$?"; tail -1 gpurun_out/r02_bench2_fuse2.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench1.log 2>&1; echo "bench1 rc=$?"; tail -1 gpurun_out/r02_bench1.log
timeout 600 $TR --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --p2p-fuse 0 > gpurun_out/r02_bench2_fuse0.log 2>&1; echo "bench2 fuse0 rc=$?"; tail -1 gpurun_out/r02_bench2_fuse0.log
