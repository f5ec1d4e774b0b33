mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log; tail -3 gpurun_out/r2g_pytest.log
ROUTINES=sdot,sasum,isamax,snrm2 python tools/strided_bench.py 26 > gpurun_out/r2g_strided.csv 2>&1; echo "strided rc=$?"
python tools/strided_bench.py 26 > gpurun_out/r2g_strided_config4.csv 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
