mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log; tail -3 gpurun_out/r2p_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2p_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p_ref.json 2>/dev/null; echo "ref rc=$?"
