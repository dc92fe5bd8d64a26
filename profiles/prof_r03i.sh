set -x
# r03i (1 GPU): captures of step_kernel (the HBM-bound kernel: DRAM traffic against its 42 B/step) and full_rollout_kernel
mkdir -p gpurun_out
python profiles/summarise_capture.py x --hash-only --sources scopa_b200/csrc/ms_env.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_env_r03i.txt
python profiles/summarise_capture.py x --hash-only --sources scopa_b200/csrc/ms_full.cu > gpurun_out/sha_full_r03i.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^step_kernel -s 6 -c 1 -f -o gpurun_out/step_r03i \
    python bench.py --steps 5 --warmup 3 --no-cpu --only env_step_api > gpurun_out/ncu_step_r03i.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:full_rollout_kernel -s 4 -c 1 -f -o gpurun_out/full_r03i \
    python bench.py --steps 5 --warmup 3 --no-cpu --only full_scopa > gpurun_out/ncu_full_r03i.log 2>&1
ls -la gpurun_out | tail -4
