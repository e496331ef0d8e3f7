#!/bin/bash
# round 2, call 23: e2e of bench.py against tools/e2e_host_ab.py on the SAME box
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/e2e_host_ab.py --reps 10 2>&1 | head -6
timeout 900 python bench.py --no-cpu --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['e2e'], d['value'])"
timeout 600 python tools/e2e_host_ab.py --reps 10 2>&1 | head -6
