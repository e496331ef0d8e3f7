set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
python bench.py > gpurun_out/bench1.log 2> gpurun_out/bench1.err; echo "rc=$?" >> gpurun_out/bench1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1
tail -2 gpurun_out/pytest.log; tail -1 gpurun_out/smoke.log
