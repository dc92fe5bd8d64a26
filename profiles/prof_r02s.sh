set -x
# r02s (1 GPU): full -m gpu suite, the driver's two arms at N = 1, full captures of md_blocked_kernel (sharded-table source),
# rollout_kernel (ms_env.cu changed: stage schedule of the host pipeline) and sd_level_mlp_kernel<1> with the box's source hashes
mkdir -p gpurun_out
H="python profiles/summarise_capture.py x --hash-only --sources"
$H scopa_b200/csrc/ms_static_walk.cuh scopa_b200/csrc/ms_solver.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_static_r02s.txt
$H scopa_b200/csrc/ms_multideal.cu scopa_b200/csrc/ms_static_walk.cuh scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_md_r02s.txt
$H scopa_b200/csrc/ms_env.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_env_r02s.txt
$H scopa_b200/csrc/ms_sdcfr.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_sd_r02s.txt
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02s.json 2> gpurun_out/bench_r02s.err ) 2>&1 | tail -4; echo "bench rc $?"; tail -5 gpurun_out/bench_r02s.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_r02s.json 2> gpurun_out/bench_ref_r02s.err ) 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:md_blocked_kernel -s 6 -c 1 -f -o gpurun_out/md_blocked_r02s \
    python bench.py --steps 5 --warmup 3 --no-cpu --only mccfr_multi_deal > gpurun_out/ncu_md_r02s.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 4 -c 1 -f -o gpurun_out/env_r02s \
    python bench.py --steps 5 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_env_r02s.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sd_level_mlp_kernel -s 19 -c 1 -f -o gpurun_out/sd_mlp_r02s \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_sd_mlp_r02s.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02s.csv \
    python bench.py --steps 5 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_launches_r02s.log 2>&1
ls -la gpurun_out | tail -12
