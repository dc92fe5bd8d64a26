set -x
# r03b (1 GPU): zero-numerator-safe fp64 divisions in regret matching / sampling (ms_tree_walk.cuh, ms_multideal.cu, ms_solver.cu);
# full -m gpu suite, both bench arms, captures of the kernels whose sources changed
mkdir -p gpurun_out
H="python profiles/summarise_capture.py x --hash-only --sources"
$H scopa_b200/csrc/ms_static_walk.cuh scopa_b200/csrc/ms_solver.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_static_r03b.txt
$H scopa_b200/csrc/ms_multideal.cu scopa_b200/csrc/ms_static_walk.cuh scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_md_r03b.txt
$H scopa_b200/csrc/ms_env.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_env_r03b.txt
$H scopa_b200/csrc/ms_sdcfr.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_sd_r03b.txt
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r03b.json 2> gpurun_out/bench_r03b.err ) 2>&1 | tail -4; echo "bench rc $?"; tail -5 gpurun_out/bench_r03b.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_r03b.json 2> gpurun_out/bench_ref_r03b.err ) 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mccfr_static_kernel -s 4 -c 1 -f -o gpurun_out/mccfr_r03b \
    python bench.py --steps 5 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_full_r03b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:md_blocked_kernel -s 6 -c 1 -f -o gpurun_out/md_blocked_r03b \
    python bench.py --steps 5 --warmup 3 --no-cpu --only mccfr_multi_deal > gpurun_out/ncu_md_r03b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sd_level_mlp_kernel -s 19 -c 1 -f -o gpurun_out/sd_mlp_r03b \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_sd_mlp_r03b.log 2>&1
ls -la gpurun_out | tail -8
