set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/check_multi_gpu.py --comm p2p > gpurun_out/multi2_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/multi2_p2p.log
timeout 400 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --comm p2p > gpurun_out/bench2_p2p.log 2> gpurun_out/bench2_p2p.err; echo "rc=$?" >> gpurun_out/bench2_p2p.err
timeout 400 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --comm nccl --no-cpu > gpurun_out/bench2_nccl.log 2> gpurun_out/bench2_nccl.err; echo "rc=$?" >> gpurun_out/bench2_nccl.err
cd master-thesis-lpf-in-mfem_b200/drivers/bin
for comm in p2p nccl; do
  echo "=== ss strong 2 GPUs $comm (2.2M dofs, rel 1e-12)" >> ../../../gpurun_out/ss2.log
  timeout 200 ./ss --gpus 2 --comm $comm --mode 0 --orders 4 --par-ref 1 --mesh wave-tank-big8.mesh --rel-tol 1e-12 --max-iter 2000 >> ../../../gpurun_out/ss2.log 2>&1
  echo "=== ss strong 2 GPUs $comm (big8 itself)" >> ../../../gpurun_out/ss2.log
  timeout 200 ./ss --gpus 2 --comm $comm --mode 0 --orders 4 --par-ref 0 --mesh wave-tank-big8.mesh --rel-tol 1e-12 --max-iter 2000 >> ../../../gpurun_out/ss2.log 2>&1
done
echo "=== cylinder 2 GPUs p2p" >> ../../../gpurun_out/ss2.log
timeout 200 ./cylinder-diffraction --gpus 2 --nsteps 35 --periods 1 --out /dev/null 2>&1 | grep "rank 0\|rms" >> ../../../gpurun_out/ss2.log
cd ../../..
tail -4 gpurun_out/multi2_p2p.log; cat gpurun_out/bench2_p2p.log; cat gpurun_out/ss2.log
