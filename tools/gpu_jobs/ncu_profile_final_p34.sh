# final-kernel ncu evidence (orders 3 and 4 after the stride table / 4-CTA change); every ncu run follows a plain run
set -x
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
for P in 3 4; do
  A="--steps 3 --warmup 3 --order $P --refine 2 --no-rk4 --no-cpu"
  python bench.py $A > gpurun_out/plain_p$P.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:pa_apply_eo -s 3 -c 1 -f -o /tmp/prof_p$P python bench.py $A > gpurun_out/ncu_p$P.log 2>&1
  ncu -i /tmp/prof_p$P.ncu-rep --page raw --csv > gpurun_out/eo_p${P}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_p$P.ncu-rep --page source --csv > gpurun_out/eo_p${P}_source.csv 2>/dev/null
done
