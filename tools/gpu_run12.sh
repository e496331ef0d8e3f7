set -x
for i in 1 2; do
  python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/cs_$i.log 2>&1
done
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
for f in gpurun_out/cs_*.log; do echo $f; grep -o '"ms_per_cg_iteration": [0-9.]*' $f; done; tail -3 gpurun_out/pytest.log
