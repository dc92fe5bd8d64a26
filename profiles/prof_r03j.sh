set -x
# r03j (1 GPU): the default bench line on the final tree
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r03j.json 2> gpurun_out/bench_r03j.err ) 2>&1 | tail -4; echo "bench rc $?"; tail -5 gpurun_out/bench_r03j.err
