mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log; tail -3 gpurun_out/r2m_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2m_bench.err
python bench.py --criterion --quick > gpurun_out/r2m_criterion_quick.csv 2> gpurun_out/r2m_criterion.err; echo "criterion rc=$?"
