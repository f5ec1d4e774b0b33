mkdir -p gpurun_out
for w in 12 24 48; do
LBK_TILES_PER_BLOCK=$w python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$((w%10)) tools/exchange_bench.py > gpurun_out/r2s_exchange_n2_w$w.csv 2> gpurun_out/r2s_exchange_n2_w$w.err; echo "w=$w rc=$?"
done
