set -x
python bench.py > gpurun_out/bench1.log 2> gpurun_out/bench1.err; echo "rc=$?" >> gpurun_out/bench1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1
