#!/bin/bash
# round 2, 4-GPU call with the final library: parity at order 7 (aliased stage buffers + exchange hooks, two neighbours per rank),
# bench line at N = 4 (carries the order-4 parity object, weak / strong CG legs, pipelined host apply)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29571 tools/check_multi_gpu.py --order 7 --mesh tank > gpurun_out/r02_parity4_p7.txt 2>&1; echo "parity p7 rc=$?"; tail -3 gpurun_out/r02_parity4_p7.txt
timeout 500 $TR --master-port 29572 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench4.log 2>&1; echo "bench4 rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench4.log") if l.startswith("{")][-1])
r, s = d["pcg_per_rk4_step"], d["strong_scaling"]
print("GDOF/s %.1f  ms/apply %.4f  kernel/rank %s  e2e %.2f  weak cg it %.4f ms  strong r1 %.4f  strong big8 %.4f  parity %s" % (
    d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_rank"], d["e2e"]["value"], r["ms_per_cg_iteration"], s["big8_r1"]["ms_per_cg_iteration"], s["big8"]["ms_per_cg_iteration"], d["parity"]["ok"]))
PY
