set -x
# r02z (1 GPU): sd_backward_kernel at 4 CTAs per SM (64 registers), z loop of sd_policy_legal without the mask arithmetic
mkdir -p gpurun_out
python profiles/summarise_capture.py x --hash-only --sources scopa_b200/csrc/ms_sdcfr.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_sd_r02z.txt
timeout 900 python -m pytest tests/test_gpu_sdcfr.py -m gpu -q -x 2>&1 | tail -3
( time timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --only sdcfr > gpurun_out/bench_r02z.json 2> gpurun_out/bench_r02z.err ) 2>&1 | tail -4; tail -5 gpurun_out/bench_r02z.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sd_ -c 120 --csv --log-file gpurun_out/launches_sd_r02z.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_launches_sd_r02z.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sd_level_mlp_kernel -s 19 -c 1 -f -o gpurun_out/sd_mlp_r02z \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_sd_mlp_r02z.log 2>&1
ls -la gpurun_out | tail -3
