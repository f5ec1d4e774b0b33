mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log; tail -3 gpurun_out/r2l_pytest.log
python tools/scan_n.py sasum,sdot,snrm2,isamax > gpurun_out/r2l_scan_n.csv 2>&1; grep "^#" gpurun_out/r2l_scan_n.csv
python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
