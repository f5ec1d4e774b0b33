mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/exchange_bench.py > gpurun_out/r2q_exchange_n2.csv 2> gpurun_out/r2q_exchange_n2.err; echo "xb rc=$?"
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err ) 2>&1 | tail -3; echo "bench rc=$?"
python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log; tail -3 gpurun_out/r2o_pytest.log
