set -x
for i in 1 2; do
LPF_VERBOSE=1 python bench.py --steps 20 --warmup 3 --no-cpu --no-rk4 > gpurun_out/pad_p4_$i.log 2>&1
LPF_VERBOSE=1 python bench.py --steps 20 --warmup 3 --no-cpu --no-rk4 --order 5 --refine 1 > gpurun_out/pad_p5_$i.log 2>&1
done
python -m pytest tests -m gpu -x -q -k "orders or variants or affine or full_size" > gpurun_out/pytest_pad.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pad.log
tail -2 gpurun_out/pytest_pad.log
for f in gpurun_out/pad_p*.log; do echo -n "$f "; grep -h "lpf: apply" $f | head -1 | sed 's/.*aff=0//'; grep -o '"kernel_ms": [0-9.]*' $f | head -1; done
