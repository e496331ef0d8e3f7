set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench1.log 2> gpurun_out/bench1.err; echo "rc=$?" >> gpurun_out/bench1.err
cd master-thesis-lpf-in-mfem_b200/drivers/bin && mkdir -p data && (time ./cylinder-diffraction --out data/cyl.txt) > ../../../gpurun_out/drv_cylfull.log 2>&1; cp data/cyl.txt ../../../gpurun_out/cyl_runup_p4.txt
(time ./cylinder-diffraction --order 6 --out data/cyl6.txt) > ../../../gpurun_out/drv_cylfull_p6.log 2>&1; cp data/cyl6.txt ../../../gpurun_out/cyl_runup_p6.txt
cd ../../..
tail -3 gpurun_out/pytest.log; cat gpurun_out/bench1.log; tail -8 gpurun_out/drv_cylfull.log
