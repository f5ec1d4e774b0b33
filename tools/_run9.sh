mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2i_topo8.txt 2>&1; lscpu | grep -E "^CPU\(s\)|NUMA|Model name" >> gpurun_out/r2i_topo8.txt; free -g >> gpurun_out/r2i_topo8.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n8.json 2> gpurun_out/r2i_bench_n8.err ) 2>&1 | tail -3; echo "bench8 rc=$?"
tail -3 gpurun_out/r2i_bench_n8.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n4.json 2> gpurun_out/r2i_bench_n4.err ) 2>&1 | tail -3; echo "bench4 rc=$?"
python -m pytest tests/test_gpu_multi.py tests/test_gpu_multictx.py -m gpu -x -q > gpurun_out/r2i_pytest8.log 2>&1; tail -2 gpurun_out/r2i_pytest8.log
