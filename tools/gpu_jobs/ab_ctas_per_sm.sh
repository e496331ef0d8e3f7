set -x
for m in 444 0 444 0 518; do
  LPF_MAX_CTAS=$m python bench.py --steps 20 --warmup 3 --no-cpu --no-rk4 > gpurun_out/ctas_${m}_$RANDOM.log 2>&1
done
for f in gpurun_out/ctas_*.log; do echo $f; grep -o '"kernel_ms": [0-9.]*' $f | head -1; done
