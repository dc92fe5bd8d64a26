set -x
# r02h (1 GPU): final headline kernel: parity, the driver's two arms, launch list, full captures with the source hash of the box
mkdir -p gpurun_out
python profiles/summarise_capture.py x --hash-only --sources scopa_b200/csrc/ms_static_walk.cuh scopa_b200/csrc/ms_solver.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_static_r02h.txt
python profiles/summarise_capture.py x --hash-only --sources scopa_b200/csrc/ms_env.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_env_r02h.txt
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err ) 2>&1 | tail -4; echo "bench rc $?"; tail -5 gpurun_out/bench_r02h.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_r02h.json 2> gpurun_out/bench_ref_r02h.err ) 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02h.csv \
    python bench.py --steps 5 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_launches_r02h.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mccfr_static_kernel -s 4 -c 1 -f -o gpurun_out/mccfr_r02h \
    python bench.py --steps 5 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_full_r02h.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 4 -c 1 -f -o gpurun_out/env_r02h \
    python bench.py --steps 5 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_env_r02h.log 2>&1
ls -la gpurun_out | tail -12
