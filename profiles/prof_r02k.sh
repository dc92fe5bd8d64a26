set -x
# r02k (1 GPU): re-capture of mccfr_static_kernel after the MODE template parameter (cb048f4) changed its sources' hash
mkdir -p gpurun_out
python profiles/summarise_capture.py x --hash-only --sources scopa_b200/csrc/ms_static_walk.cuh scopa_b200/csrc/ms_solver.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_static_r02k.txt
( time timeout 900 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/bench_r02k.json 2> gpurun_out/bench_r02k.err ) 2>&1 | tail -4; tail -5 gpurun_out/bench_r02k.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mccfr_static_kernel -s 4 -c 1 -f -o gpurun_out/mccfr_r02k \
    python bench.py --steps 5 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_full_r02k.log 2>&1
ls -la gpurun_out | tail -5
