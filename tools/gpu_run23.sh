set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
LPF_AFFINE=0 LPF_VERBOSE=1 python tools/sweep.py --orders 1,2,3,4,5,6,7,8 --variants 0,33 --out gpurun_out/sweep_layout.json > gpurun_out/sweep_layout.log 2>&1
tail -3 gpurun_out/pytest.log; grep -v "^lpf:" gpurun_out/sweep_layout.log | tail -20
