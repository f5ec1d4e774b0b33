mkdir -p gpurun_out; nproc
for l in 1 2 4 6 8; do for c in 2 4 8 16; do LBK_STAGE_LANES=$l LBK_STAGE_CHUNK_MB=$c python tools/transfer_sweep.py; done; done > gpurun_out/r2n_transfer_sweep.csv 2>&1
cat gpurun_out/r2n_transfer_sweep.csv
