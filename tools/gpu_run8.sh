set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench1.log 2> gpurun_out/bench1.err; echo "rc=$?" >> gpurun_out/bench1.err
python tools/sweep.py --help > /dev/null 2>&1
cd master-thesis-lpf-in-mfem_b200/drivers/bin
S=../../../gpurun_out/aff.log; : > $S
for aff in 0 1; do
  echo "=== LPF_AFFINE=$aff ss strong 1 GPU 2.2M dofs" >> $S
  LPF_AFFINE=$aff ./ss --mode 0 --orders 4 --par-ref 1 --mesh wave-tank-big8.mesh --rel-tol 1e-12 --max-iter 2000 >> $S 2>&1
  echo "=== LPF_AFFINE=$aff ws 1 GPU par-ref 2 orders 3,4" >> $S
  LPF_AFFINE=$aff ./ss --mode ws --par-ref 2 >> $S 2>&1
done
./PF_linear_par_partial --serial-params --nsteps 40 --periods 0.4 > ../../../gpurun_out/drv_serial.log 2>&1
cd ../../..
tail -3 gpurun_out/pytest.log; cat gpurun_out/aff.log | grep -v "^---\|^procs\|^Strong\|^Weak"; head -12 gpurun_out/drv_serial.log
