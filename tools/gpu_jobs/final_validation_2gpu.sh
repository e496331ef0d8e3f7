set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
O=gpurun_out/final2.log; : > $O
for mesh in tank cylinder; do
  echo "=== check_multi_gpu 2 GPUs mesh=$mesh p2p (final kernels)" >> $O
  timeout 300 $TR --master-port 29571 tools/check_multi_gpu.py --comm p2p --mesh $mesh 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -17 >> $O; echo "rc=$?" >> $O
done
timeout 400 $TR --master-port 29572 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench2_p2p.log 2> gpurun_out/bench2_p2p.err; echo "rc=$?" >> gpurun_out/bench2_p2p.err
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench1_2box.log 2>&1
grep "===\|PARITY\|rc=\|host apply" $O
