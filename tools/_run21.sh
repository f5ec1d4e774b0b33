mkdir -p gpurun_out
python tools/profile_one.py 28 1 > gpurun_out/r2u_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:unit_kernel -c 10 -o /tmp/r2u_prof python tools/profile_one.py 28 1 > gpurun_out/r2u_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r2u_prof.ncu-rep --page raw --csv > gpurun_out/r2u_raw.csv 2>/dev/null
ncu -i /tmp/r2u_prof.ncu-rep --page source --csv --kernel-name regex:AxpyF > gpurun_out/r2u_source_saxpy.csv 2>/dev/null
ncu -i /tmp/r2u_prof.ncu-rep --page source --csv --kernel-name regex:DotOp > gpurun_out/r2u_source_sdot.csv 2>/dev/null
ls -la gpurun_out/r2u_*; du -sh gpurun_out
