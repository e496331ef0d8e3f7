set -x
for h in 1 0 1 0; do
  LPF_L2_HINT=$h python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/l2hint_${h}_$RANDOM.log 2>&1
done
for f in gpurun_out/l2hint_*.log; do echo $f; grep -o '"kernel_ms": [0-9.]*' $f | head -1; grep -o '"ms_per_cg_iteration": [0-9.]*' $f; done
