set -x
# r02j (1 GPU): md_blocked_kernel on the static walk: parity, bench, full capture
mkdir -p gpurun_out
python profiles/summarise_capture.py x --hash-only --sources scopa_b200/csrc/ms_multideal.cu scopa_b200/csrc/ms_static_walk.cuh scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_md_r02j.txt
timeout 900 python -m pytest tests/test_gpu_multideal.py tests/test_gpu_solver.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -5
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02j.json 2> gpurun_out/bench_r02j.err ) 2>&1 | tail -4; echo "bench rc $?"; tail -5 gpurun_out/bench_r02j.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:md_blocked_kernel -s 6 -c 1 -f -o gpurun_out/md_blocked_r02j \
    python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_md_r02j.log 2>&1
ls -la gpurun_out | tail -6
