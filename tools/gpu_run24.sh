set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench1.log 2> gpurun_out/bench1.err; echo "rc=$?" >> gpurun_out/bench1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
cd master-thesis-lpf-in-mfem_b200/drivers/bin
(./laplace_solver; ./ss --mode ws --par-ref 1; ./ss --mode ws --par-ref 0 --mesh wave-tank-big8.mesh --orders 4; ./PF_linear_par_partial --nsteps 20 --periods 0.5; ./PF_linear_par_partial --serial-params --nsteps 40 --periods 0.4 | tail -4; ./cylinder-diffraction --out /dev/null | tail -4; ./cylinder-diffraction --order 3 --nsteps 70 --periods 2 --paraview cyl --pv-every 35 --out /dev/null | tail -2; ls ParaView/cyl; head -c 600 ParaView/cyl/cyl.pvd) > ../../../gpurun_out/drv_all.log 2>&1
cd ../../..
tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_ref.log | head -c 400; tail -12 gpurun_out/drv_all.log
