set -x
cd master-thesis-lpf-in-mfem_b200/drivers/bin
O=../../../gpurun_out/smallvar.log; : > $O
for v in 0 32 31 30 20; do
  echo "=== LPF_APPLY_VARIANT=$v LPF_AFFINE=0 big8 r=0 (4096 hexes) / cylinder (3192 hexes)" >> $O
  LPF_APPLY_VARIANT=$v LPF_AFFINE=0 ./ss --mode 0 --orders 4 --par-ref 0 --mesh wave-tank-big8.mesh --rel-tol 1e-12 --max-iter 2000 2>&1 | tail -1 >> $O
  LPF_APPLY_VARIANT=$v ./cylinder-diffraction --nsteps 35 --periods 1 --out /dev/null 2>&1 | grep "rank 0" >> $O
done
cd ../../..
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_small.csv master-thesis-lpf-in-mfem_b200/drivers/bin/ss --mode 0 --orders 4 --par-ref 0 --mesh wave-tank-big8.mesh --rel-tol 1e-12 --max-iter 2000 --nsteps 1 > /dev/null 2>&1
cat gpurun_out/smallvar.log
