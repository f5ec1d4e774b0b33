mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -3 gpurun_out/r2e_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1; tail -2 gpurun_out/r2e_smoke.log
build/lb_example > gpurun_out/r2e_example.log 2>&1; tail -3 gpurun_out/r2e_example.log
build/phase_exp 28 > gpurun_out/r2e_phase.csv 2>&1; echo "phase rc=$?"
ROUTINES=isamax,sasum PAIRS=0/0 python tools/strided_bench.py 26 > gpurun_out/r2e_strided.csv 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__cycles_active.sum,dram__cycles_active_read.sum,dram__cycles_active_write.sum,dram__cycles_elapsed.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum"
GROUPS=streams python tools/ncu_more.py 28 > gpurun_out/r2e_streams_manifest.csv 2> gpurun_out/r2e_streams_plain.err &&
GROUPS=streams ncu --metrics $M --clock-control none -k regex:'unit_kernel|strided_kernel' --csv --log-file gpurun_out/r2e_streams.csv python tools/ncu_more.py 28 > gpurun_out/r2e_streams_ncu.log 2>&1
echo "ncu rc=$?"
