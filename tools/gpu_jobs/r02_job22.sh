#!/bin/bash
# round 2, call 22: final library -- whole -m gpu suite, smoke(), bench line with its ncu launch list
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_job22_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job22_pytest.log
tail -4 gpurun_out/r02_job22_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_job22_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_job22_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu_c.json 2> gpurun_out/r02_bench_1gpu_c.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r02_bench_1gpu_c.json
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench_short.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r02_bench_launches_c.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench_ncu.log 2>&1; echo "ncu rc=$?"
