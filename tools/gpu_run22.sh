set -x
for v in 0 33 34 0 33; do
  LPF_VERBOSE=1 python bench.py --steps 20 --warmup 3 --no-cpu --no-rk4 --order 3 --variant $v > gpurun_out/p3_v${v}_$RANDOM.log 2>&1
done
python -m pytest tests -m gpu -x -q -k "orders or variants or affine" > gpurun_out/pytest_p3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_p3.log
tail -3 gpurun_out/pytest_p3.log
