set -x
cd master-thesis-lpf-in-mfem_b200/drivers/bin
O=../../../gpurun_out/vecgrid.log; : > $O
for i in 1 2; do
  ./ss --mode 0 --orders 4 --par-ref 0 --mesh wave-tank-big8.mesh --rel-tol 1e-12 --max-iter 2000 2>&1 | tail -1 >> $O
  ./cylinder-diffraction --nsteps 35 --periods 1 --out /dev/null 2>&1 | grep "rank 0" >> $O
done
./ss --mode ws --par-ref 2 --orders 4 2>&1 | tail -1 >> $O
./PF_linear_par_partial --nsteps 40 --periods 1 2>&1 | grep "rank 0" >> $O
./convergence-parallel-partial --mode p --rel-tol 1e-13 --orders 4,6 >> $O 2>&1
cd ../../..
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_vg.log 2>&1
cat gpurun_out/vecgrid.log; tail -3 gpurun_out/pytest.log; grep -o '"ms_per_cg_iteration": [0-9.]*' gpurun_out/bench_vg.log
