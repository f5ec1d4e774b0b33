mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -3 gpurun_out/r2h_pytest.log
ROUTINES=sdot,sasum,isamax,snrm2 python tools/strided_bench.py 26 > gpurun_out/r2h_strided.csv 2>&1; echo "strided rc=$?"
