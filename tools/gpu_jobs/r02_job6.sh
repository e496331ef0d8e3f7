#!/bin/bash
# round 2, sixth GPU call: A/B of the PCG leg, round-1 library (worktree ab_r01) against the current one, same box;
# adapter tests after the stream fix
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for i in 1 2; do
  (cd ab_r01 && timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu) > gpurun_out/r02_ab_old_$i.log 2>&1; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_ab_old_$i.log').read().strip().split('\n')[-1]); r=d['pcg_per_rk4_step']
print('OLD $i: apply ms', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'rk4 ms', r['ms_per_rk4_step'], 'cg it ms', r['ms_per_cg_iteration'])
PY
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_ab_new_$i.log 2>&1; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_ab_new_$i.log').read().strip().split('\n')[-1]); r=d['pcg_per_rk4_step']
print('NEW $i: apply ms', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'rk4 ms', r['ms_per_rk4_step'], 'cg it ms', r['ms_per_cg_iteration'], 'strong big8', d['strong_scaling']['big8']['ms_per_cg_iteration'])
PY
done
for o in "max_ctas=592" "pdl=0" "use_graph=0"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --opt $o > gpurun_out/r02_ab_new_$o.log 2>&1; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_ab_new_$o.log').read().strip().split('\n')[-1]); r=d['pcg_per_rk4_step']
print('NEW $o: apply ms', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'rk4 ms', r['ms_per_rk4_step'], 'cg it ms', r['ms_per_cg_iteration'], 'strong big8', d['strong_scaling']['big8']['ms_per_cg_iteration'])
PY
done
timeout 900 python -m pytest tests/test_gpu_adapter.py tests/test_gpu_config_parity.py -m gpu -q -k "adapter or c3" > gpurun_out/r02_job6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job6_pytest.log
tail -6 gpurun_out/r02_job6_pytest.log
