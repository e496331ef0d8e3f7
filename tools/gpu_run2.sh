set -x
B="--steps 20 --warmup 3 --no-rk4 --no-cpu"
python bench.py $B > gpurun_out/ab_new1.log 2>&1
(cd _ab_old && python bench.py $B > ../gpurun_out/ab_old1.log 2>&1)
python bench.py $B > gpurun_out/ab_new2.log 2>&1
(cd _ab_old && python bench.py $B > ../gpurun_out/ab_old2.log 2>&1)
cd master-thesis-lpf-in-mfem_b200/drivers/bin && mkdir -p data && (time ./cylinder-diffraction --out data/cyl.txt) > ../../../gpurun_out/drv_cylfull.log 2>&1; cp data/cyl.txt ../../../gpurun_out/cyl_runup_p4.txt
(time ./cylinder-diffraction --order 6 --out data/cyl6.txt) > ../../../gpurun_out/drv_cylfull_p6.log 2>&1; cp data/cyl6.txt ../../../gpurun_out/cyl_runup_p6.txt
cd ../../..
grep -h -o '"kernel_ms": [0-9.]*' gpurun_out/ab_*.log; tail -4 gpurun_out/drv_cylfull.log
