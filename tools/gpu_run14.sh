set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
cd master-thesis-lpf-in-mfem_b200/drivers/bin
O=../../../gpurun_out/spec_drv.log; : > $O
for i in 1 2; do
./ss --mode 0 --orders 4 --par-ref 0 --mesh wave-tank-big8.mesh --rel-tol 1e-12 --max-iter 2000 >> $O 2>&1
./cylinder-diffraction --nsteps 35 --periods 1 --out /dev/null 2>&1 | grep "rank 0" >> $O
done
./ss --mode ws --par-ref 2 --orders 4 >> $O 2>&1
./laplace_solver >> $O 2>&1
cd ../../..
python bench.py --steps 20 --warmup 3 > gpurun_out/bench1.log 2> gpurun_out/bench1.err; echo "rc=$?" >> gpurun_out/bench1.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/pytest.log; grep -v "^---\|^procs\|^Strong\|^Weak" gpurun_out/spec_drv.log; tail -2 gpurun_out/smoke.log
