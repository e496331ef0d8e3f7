set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
O=gpurun_out/ll2.log; : > $O
for fuse in 1 0; do for mesh in tank cylinder; do
  echo "=== check_multi_gpu mesh=$mesh LPF_P2P_FUSE=$fuse" >> $O
  LPF_P2P_FUSE=$fuse timeout 300 $TR --master-port 2951$fuse tools/check_multi_gpu.py --comm p2p --mesh $mesh 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -14 >> $O; echo "rc=$?" >> $O
done; done
timeout 400 $TR --master-port 29520 bench.py --gpus 2 --steps 20 --warmup 3 --comm p2p --no-cpu > gpurun_out/bench2_p2p.log 2> gpurun_out/bench2_p2p.err; echo "rc=$?" >> gpurun_out/bench2_p2p.err
cd master-thesis-lpf-in-mfem_b200/drivers/bin
S=../../../gpurun_out/ss2.log; : > $S
for fuse in 1 0; do
  echo "=== ss strong 2 GPUs p2p FUSE=$fuse (big8 itself, 299520 dofs)" >> $S
  LPF_P2P_FUSE=$fuse timeout 200 ./ss --gpus 2 --comm p2p --mode 0 --orders 4 --par-ref 0 --mesh wave-tank-big8.mesh --rel-tol 1e-12 --max-iter 2000 >> $S 2>&1
done
echo "=== ss strong 2 GPUs p2p (2.2M dofs)" >> $S
timeout 200 ./ss --gpus 2 --comm p2p --mode 0 --orders 4 --par-ref 1 --mesh wave-tank-big8.mesh --rel-tol 1e-12 --max-iter 2000 >> $S 2>&1
echo "=== ss 1 GPU (299520 dofs)" >> $S
timeout 200 ./ss --gpus 1 --mode 0 --orders 4 --par-ref 0 --mesh wave-tank-big8.mesh --rel-tol 1e-12 --max-iter 2000 >> $S 2>&1
for fuse in 1 0; do
echo "=== cylinder 2 GPUs p2p FUSE=$fuse" >> $S
LPF_P2P_FUSE=$fuse timeout 200 ./cylinder-diffraction --gpus 2 --nsteps 35 --periods 1 --out /dev/null 2>&1 | grep "rank 0\|rms" >> $S
done
echo "=== cylinder 1 GPU" >> $S
timeout 200 ./cylinder-diffraction --gpus 1 --nsteps 35 --periods 1 --out /dev/null 2>&1 | grep "rank 0\|rms" >> $S
cd ../../..
cat $O | grep "===\|PARITY\|rc=\|timed"; cat gpurun_out/ss2.log | grep -v "^---\|^procs\|^Strong\|^Weak"
