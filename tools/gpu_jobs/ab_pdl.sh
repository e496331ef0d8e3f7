set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench1.log 2> gpurun_out/bench1.err; echo "rc=$?" >> gpurun_out/bench1.err
cd master-thesis-lpf-in-mfem_b200/drivers/bin
SS="--mode 0 --orders 4 --par-ref 0 --mesh wave-tank-big8.mesh --rel-tol 1e-12 --max-iter 2000"
for pdl in 0 1; do for ch in 16 64; do
  echo "=== LPF_PDL=$pdl LPF_PCG_CHUNK=$ch ss $SS" >> ../../../gpurun_out/pdl_ab.log
  LPF_PDL=$pdl LPF_PCG_CHUNK=$ch ./ss $SS >> ../../../gpurun_out/pdl_ab.log 2>&1
done; done
for pdl in 0 1; do
  echo "=== LPF_PDL=$pdl ss par-ref 1" >> ../../../gpurun_out/pdl_ab.log
  LPF_PDL=$pdl ./ss --mode 0 --orders 4 --par-ref 1 --mesh wave-tank-big8.mesh --rel-tol 1e-12 --max-iter 2000 >> ../../../gpurun_out/pdl_ab.log 2>&1
  echo "=== LPF_PDL=$pdl cylinder" >> ../../../gpurun_out/pdl_ab.log
  LPF_PDL=$pdl ./cylinder-diffraction --nsteps 35 --periods 1 --out /dev/null 2>&1 | grep "rank 0" >> ../../../gpurun_out/pdl_ab.log
done
cd ../../..
tail -3 gpurun_out/pytest.log; cat gpurun_out/pdl_ab.log
