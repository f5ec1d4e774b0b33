mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -5 gpurun_out/r2c_bench.err
python tools/misaligned_bench.py 28 > gpurun_out/r2c_misaligned.csv 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__cycles_active.sum,dram__cycles_active_read.sum,dram__cycles_active_write.sum,dram__cycles_elapsed.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum"
python tools/ncu_more.py 26 > gpurun_out/r2c_more_manifest.csv 2> gpurun_out/r2c_more_plain.err &&
ncu --metrics $M --clock-control none -k regex:'unit_kernel|strided_kernel' --csv --log-file gpurun_out/r2c_more.csv python tools/ncu_more.py 26 > gpurun_out/r2c_more_ncu.log 2>&1
echo "ncu rc=$?"
