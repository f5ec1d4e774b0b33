mkdir -p gpurun_out
for w in 24 48 96; do
SIZES=27,28 LBK_TILES_PER_BLOCK=$w python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$((w%10)) tools/exchange_bench.py > gpurun_out/r2t_exchange28_n2_w$w.csv 2> gpurun_out/r2t_exchange28_n2_w$w.err; echo "w=$w rc=$?"
done
