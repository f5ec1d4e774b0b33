mkdir -p gpurun_out
for lg in 25 24 26 27; do python tools/midsize_sweep.py $lg > gpurun_out/r2j_midsize_$lg.csv 2>&1; grep "^#" gpurun_out/r2j_midsize_$lg.csv; done
F="--steps 10 --warmup 3 --no-cpu --no-f64 --no-parity --no-single --no-everyday --e2e-steps 3"
python bench.py $F > gpurun_out/r2j_ab_plain_1.json 2>/dev/null
LBK_LIBRARY=$PWD/build/liblbk_nc.so python bench.py $F > gpurun_out/r2j_ab_nc_1.json 2>/dev/null
python bench.py $F > gpurun_out/r2j_ab_plain_2.json 2>/dev/null
LBK_LIBRARY=$PWD/build/liblbk_nc.so python bench.py $F > gpurun_out/r2j_ab_nc_2.json 2>/dev/null
echo done
