mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/exchange_bench.py > gpurun_out/r2r_exchange_n8.csv 2> gpurun_out/r2r_exchange_n8.err; echo "xb rc=$?"
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2r_bench_n8.json 2> gpurun_out/r2r_bench_n8.err ) 2>&1 | tail -3; echo "bench8 rc=$?"
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2r_bench_n4.json 2> gpurun_out/r2r_bench_n4.err ) 2>&1 | tail -3; echo "bench4 rc=$?"
