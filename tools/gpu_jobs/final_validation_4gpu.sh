set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
O=gpurun_out/final4.log; : > $O
echo "=== check_multi_gpu 4 GPUs mesh=cylinder p2p (final kernels)" >> $O
timeout 200 $TR --master-port 29581 tools/check_multi_gpu.py --comm p2p --mesh cylinder 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -16 >> $O; echo "rc=$?" >> $O
timeout 400 $TR --master-port 29582 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench4.log 2> gpurun_out/bench4.err; echo "rc=$?" >> gpurun_out/bench4.err
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench1_4box.log 2>&1
grep "===\|PARITY\|rc=" $O
