set -x
# r02c: mccfr_static_kernel with lane-private nl-1 accumulators; bench.py full default run (N = 1); full capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_dropin.py tests/test_gpu_multideal.py -m gpu -q -x 2>&1 | tail -6
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err ) 2>&1 | tail -4; echo "bench rc $?"; tail -5 gpurun_out/bench_r02c.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mccfr_static_kernel -s 4 -c 1 -f -o gpurun_out/mccfr_r02c \
    python bench.py --steps 5 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_full_r02c.log 2>&1
ls -la gpurun_out/
