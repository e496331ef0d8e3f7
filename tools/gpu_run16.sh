set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
O=gpurun_out/nccl2.log; : > $O
for comm in nccl p2p; do
  echo "=== check_multi_gpu 2 GPUs mesh=tank comm=$comm" >> $O
  timeout 300 $TR --master-port 29551 tools/check_multi_gpu.py --comm $comm --mesh tank 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -16 >> $O; echo "rc=$?" >> $O
done
echo "=== check_multi_gpu 2 GPUs mesh=cylinder comm=nccl" >> $O
timeout 300 $TR --master-port 29552 tools/check_multi_gpu.py --comm nccl --mesh cylinder 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -6 >> $O
timeout 400 $TR --master-port 29553 bench.py --gpus 2 --steps 20 --warmup 3 --comm nccl --no-cpu > gpurun_out/bench2_nccl.log 2> gpurun_out/bench2_nccl.err; echo "rc=$?" >> gpurun_out/bench2_nccl.err
cd master-thesis-lpf-in-mfem_b200/drivers/bin
./convergence-parallel-partial --mode h --rel-tol 1e-13 --gpus 2 --levels 1,2 > ../../../gpurun_out/drv_hconv_2gpu.log 2>&1
cd ../../..
grep "===\|PARITY\|rc=" $O; cat gpurun_out/drv_hconv_2gpu.log
