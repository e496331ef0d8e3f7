set -x
python -m pytest tests -m gpu -x -q -k "time_stepper" > gpurun_out/pytest_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_conv.log
cd master-thesis-lpf-in-mfem_b200/drivers/bin
(time ./convergence-parallel-partial --mode p --rel-tol 1e-13 --orders 1,2,3,4,5,6,7,8) > ../../../gpurun_out/drv_pconv.log 2>&1
(time ./convergence-parallel-partial --mode h --rel-tol 1e-13) > ../../../gpurun_out/drv_hconv.log 2>&1
(time ./convergence-parallel-partial --mode p --orders 4) > ../../../gpurun_out/drv_pconv_ref_settings.log 2>&1
cd ../../..
tail -4 gpurun_out/pytest_conv.log; cat gpurun_out/drv_pconv.log gpurun_out/drv_hconv.log gpurun_out/drv_pconv_ref_settings.log
