mkdir -p gpurun_out
for lg in 28 30 22 20; do python tools/midsize_sweep.py $lg > gpurun_out/r2k_midsize_$lg.csv 2>&1; grep "^#" gpurun_out/r2k_midsize_$lg.csv; done
