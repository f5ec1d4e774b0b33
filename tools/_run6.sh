mkdir -p gpurun_out
ROUTINES=sdot,sasum,isamax,snrm2 python tools/strided_bench.py 26 > gpurun_out/r2f_strided.csv 2>&1; echo "strided rc=$?"
M="dram__bytes_read.sum,dram__bytes_write.sum,dram__cycles_active.sum,dram__cycles_active_read.sum,dram__cycles_active_write.sum,dram__cycles_elapsed.sum,gpu__time_duration.sum"
python tools/steady_counters.py 28 > gpurun_out/r2f_steady_manifest.csv 2> gpurun_out/r2f_steady_plain.err &&
ncu --replay-mode app-range --metrics $M --clock-control none --csv --log-file gpurun_out/r2f_steady.csv python tools/steady_counters.py 28 > gpurun_out/r2f_steady_ncu.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/r2f_steady_ncu.log
python -m pytest tests -m gpu -x -q -k "large or parity or golden" > gpurun_out/r2f_pytest.log 2>&1; tail -2 gpurun_out/r2f_pytest.log
