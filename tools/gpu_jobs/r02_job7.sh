#!/bin/bash
# round 2, seventh GPU call: PCG iteration A/B -- round-1 library vs current, grid size (CTAs/SM) x PDL, r=1 and big8
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
for r in 1 0; do
  python tools/pcg_small_ab.py ab_r01 $r
  python tools/pcg_small_ab.py ab_r01 $r pdl=0
  python tools/pcg_small_ab.py . $r
  python tools/pcg_small_ab.py . $r pdl=0
  python tools/pcg_small_ab.py . $r max_ctas=444
  python tools/pcg_small_ab.py . $r max_ctas=444 pdl=0
  python tools/pcg_small_ab.py . $r use_graph=0
  python tools/pcg_small_ab.py . $r pcg_chunk=32
done
} > gpurun_out/r02_pcg_ab.txt 2>&1
cat gpurun_out/r02_pcg_ab.txt
timeout 900 python -m pytest tests/test_gpu_adapter.py -m gpu -q > gpurun_out/r02_job7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_job7_pytest.log
tail -8 gpurun_out/r02_job7_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu --opt verbose=1 > gpurun_out/r02_job7_bench.log 2> gpurun_out/r02_job7_bench.err; tail -1 gpurun_out/r02_job7_bench.log | cut -c1-600; grep "lpf:" gpurun_out/r02_job7_bench.err | sort | uniq
