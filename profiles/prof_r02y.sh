set -x
# r02y (1 GPU): SDCFR after the zero-numerator-safe divisions with an OPAQUE stand-in (backward / expand / policy), final captures of the SDCFR kernels
mkdir -p gpurun_out
python profiles/summarise_capture.py x --hash-only --sources scopa_b200/csrc/ms_sdcfr.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_sd_r02y.txt
timeout 900 python -m pytest tests/test_gpu_sdcfr.py tests/test_gpu_sd_train.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -4
( time timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --only sdcfr > gpurun_out/bench_r02y.json 2> gpurun_out/bench_r02y.err ) 2>&1 | tail -4; tail -5 gpurun_out/bench_r02y.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sd_ -c 120 --csv --log-file gpurun_out/launches_sd_r02y.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_launches_sd_r02y.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sd_level_mlp_kernel -s 19 -c 1 -f -o gpurun_out/sd_mlp_r02y \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_sd_mlp_r02y.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sd_backward_kernel -s 17 -c 1 -f -o gpurun_out/sd_bwd_r02y \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_sd_bwd_r02y.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sd_expand_kernel -s 19 -c 1 -f -o gpurun_out/sd_expand_r02y \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_sd_expand_r02y.log 2>&1
ls -la gpurun_out | tail -4
