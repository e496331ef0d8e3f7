set -x
LPF_AFFINE=0 timeout 500 python tools/sweep.py --orders 5,6,7,8 --variants 0 --refine-high 2 --out gpurun_out/sweep_r2_high.json > gpurun_out/sweep_r2_high.log 2>&1
grep "^p=" gpurun_out/sweep_r2_high.log
