set -x
for p in 1 0 1 0; do
  LPF_L2_PERSIST=$p python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/persist_${p}_$RANDOM.log 2>&1
done
cd master-thesis-lpf-in-mfem_b200/drivers/bin
for p in 1 0; do
  echo "=== LPF_L2_PERSIST=$p ws 1 GPU par-ref 2 order 4 + strong par-ref 1 + big8" >> ../../../gpurun_out/persist_drv.log
  LPF_L2_PERSIST=$p ./ss --mode ws --par-ref 2 --orders 4 >> ../../../gpurun_out/persist_drv.log 2>&1
  LPF_L2_PERSIST=$p ./ss --mode 0 --orders 4 --par-ref 0 --mesh wave-tank-big8.mesh --rel-tol 1e-12 --max-iter 2000 >> ../../../gpurun_out/persist_drv.log 2>&1
  LPF_L2_PERSIST=$p ./cylinder-diffraction --nsteps 35 --periods 1 --out /dev/null 2>&1 | grep "rank 0" >> ../../../gpurun_out/persist_drv.log
done
cd ../../..
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
