set -x
python -m pytest tests -m gpu -x -q -k "full_size or host_pipeline or golden or api_state" > gpurun_out/pytest_hp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_hp.log
python bench.py --steps 20 --warmup 3 --no-rk4 --no-cpu > gpurun_out/bench_hp.log 2> gpurun_out/bench_hp.err; echo "rc=$?" >> gpurun_out/bench_hp.err
tail -5 gpurun_out/pytest_hp.log; grep -o '"e2e": {[^}]*}' gpurun_out/bench_hp.log
