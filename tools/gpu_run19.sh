set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
O=gpurun_out/hp2.log; : > $O
for comm in p2p nccl; do
  echo "=== check_multi_gpu 2 GPUs mesh=tank comm=$comm (+ pipelined host apply)" >> $O
  timeout 300 $TR --master-port 29561 tools/check_multi_gpu.py --comm $comm --mesh tank 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -16 >> $O; echo "rc=$?" >> $O
done
timeout 400 $TR --master-port 29562 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-rk4 > gpurun_out/bench2_hp.log 2> gpurun_out/bench2_hp.err; echo "rc=$?" >> gpurun_out/bench2_hp.err
grep "===\|PARITY\|rc=\|host apply" $O; grep -o '"e2e": {[^}]*}' gpurun_out/bench2_hp.log
