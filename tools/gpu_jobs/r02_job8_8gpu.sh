#!/bin/bash
# round 2, 8-GPU call: parity of the overlapped exchange at N = 8 (tank and cylinder), then the bench line with and without
# programmatic dependent launch, and with the separate LL kernel (p2p_fuse = 0) for the A/B of the overlap
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tools/check_multi_gpu.py --order 4 --mesh tank > gpurun_out/r02_parity8_tank.txt 2>&1; echo "tank rc=$?"; tail -2 gpurun_out/r02_parity8_tank.txt
timeout 300 $TR --master-port 29522 tools/check_multi_gpu.py --order 3 --mesh cylinder > gpurun_out/r02_parity8_cyl.txt 2>&1; echo "cyl rc=$?"; tail -2 gpurun_out/r02_parity8_cyl.txt
timeout 600 $TR --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench8.log 2>&1; echo "bench8 rc=$?"
timeout 600 $TR --master-port 29524 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --opt pdl=0 > gpurun_out/r02_bench8_pdl0.log 2>&1; echo "bench8 pdl0 rc=$?"
timeout 600 $TR --master-port 29525 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --p2p-fuse 0 > gpurun_out/r02_bench8_fuse0.log 2>&1; echo "bench8 fuse0 rc=$?"
python - <<'PY'
import json
for f in ("r02_bench8", "r02_bench8_pdl0", "r02_bench8_fuse0"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.log").read().strip().split("\n")[-1])
        r, s = d["pcg_per_rk4_step"], d["strong_scaling"]
        print(f, "GDOF/s %.1f  ms/apply %.4f  kernel/rank %s  weak cg it %.4f ms  strong r1 %.4f  strong big8 %.4f  parity %s" % (
            d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_rank"], r["ms_per_cg_iteration"], s["big8_r1"]["ms_per_cg_iteration"], s["big8"]["ms_per_cg_iteration"], d["parity"]["ok"]))
    except Exception as e:
        print(f, "failed:", e)
PY
