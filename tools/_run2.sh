mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -3 gpurun_out/r2b_pytest.log
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -5 gpurun_out/r2b_bench.err
export PAIRS=1/0,5/2
for v in 8 26 23 25 24; do for t in 256 512; do
  ROUTINES=saxpy,scopy,sswap,srot python tools/misaligned_bench.py 28 mapmisaligned:$v:$t:1048576 >> gpurun_out/r2b_mis_map.csv 2>&1
done; done
for v in 8 26 25; do for t in 256 512; do for b in 8 16; do
  ROUTINES=sdot python tools/misaligned_bench.py 28 redmisaligned:$v:$t:$b >> gpurun_out/r2b_mis_red.csv 2>&1
done; done; done
