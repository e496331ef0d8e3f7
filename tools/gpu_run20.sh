set -x
for v in 0 33 0 33; do
  LPF_VERBOSE=1 python bench.py --steps 20 --warmup 3 --no-cpu --variant $v > gpurun_out/occ_v${v}_$RANDOM.log 2>&1
done
python -m pytest tests -m gpu -x -q -k "p4 or variants or orders or full_size or golden" > gpurun_out/pytest_occ.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_occ.log
grep -h "lpf: apply" gpurun_out/occ_v*.log | sort | uniq -c; tail -3 gpurun_out/pytest_occ.log
