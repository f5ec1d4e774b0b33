mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2d_topo.txt 2>&1; lscpu | head -25 >> gpurun_out/r2d_topo.txt; numactl -H >> gpurun_out/r2d_topo.txt 2>&1
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302" $d/class; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done >> gpurun_out/r2d_topo.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -3 gpurun_out/r2d_pytest.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -5 gpurun_out/r2d_bench_n2.err
