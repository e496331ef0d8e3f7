set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
O=gpurun_out/multi8.log; : > $O
for mesh in tank cylinder; do
  echo "=== check_multi_gpu 8 GPUs mesh=$mesh p2p" >> $O
  timeout 300 $TR --nproc-per-node 8 --master-port 29531 tools/check_multi_gpu.py --comm p2p --mesh $mesh 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -16 >> $O; echo "rc=$?" >> $O
done
timeout 500 $TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench8.log 2> gpurun_out/bench8.err; echo "rc=$?" >> gpurun_out/bench8.err
timeout 400 $TR --nproc-per-node 4 --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench4.log 2> gpurun_out/bench4.err; echo "rc=$?" >> gpurun_out/bench4.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench1_8box.log 2> gpurun_out/bench1_8box.err
cd master-thesis-lpf-in-mfem_b200/drivers/bin
S=../../../gpurun_out/ss8.log; : > $S
for n in 1 2 4 8; do
  echo "=== ss strong $n GPUs p2p (2.2M dofs, rel 1e-12)" >> $S
  timeout 200 ./ss --gpus $n --comm p2p --mode 0 --orders 4 --par-ref 1 --mesh wave-tank-big8.mesh --rel-tol 1e-12 --max-iter 2000 >> $S 2>&1
done
for n in 2 8; do
  echo "=== ss strong $n GPUs p2p (big8 itself, 299520 dofs)" >> $S
  timeout 200 ./ss --gpus $n --comm p2p --mode 0 --orders 4 --par-ref 0 --mesh wave-tank-big8.mesh --rel-tol 1e-12 --max-iter 2000 >> $S 2>&1
done
for n in 1 2 4 8; do
  echo "=== ss ws-mode $n GPUs p2p (reference ws.cpp protocol, par-ref 2)" >> $S
  timeout 200 ./ss --gpus $n --comm p2p --mode ws --orders 4 --par-ref 2 >> $S 2>&1
done
echo "=== cylinder 8 GPUs" >> $S
timeout 200 ./cylinder-diffraction --gpus 8 --nsteps 35 --periods 1 --out /dev/null 2>&1 | grep "rank 0\|rms" >> $S
cd ../../..
grep "===\|PARITY\|rc=\|timed" $O; grep -v "^---\|^procs\|^Strong\|^Weak" gpurun_out/ss8.log
