set -x
LPF_AFFINE=0 python tools/sweep.py --orders 1,2,3,5,6,7,8 --variants 0,30,31,32 --out gpurun_out/sweep_var2.json > gpurun_out/sweep_var2.log 2>&1
grep "^p=" gpurun_out/sweep_var2.log
