set -x
# r03a (1 GPU): zero-numerator-safe divisions in the optimiser / average-policy kernels too; SDCFR GPU tests, SDCFR bench section
mkdir -p gpurun_out
python profiles/summarise_capture.py x --hash-only --sources scopa_b200/csrc/ms_sdcfr.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_sd_r03a.txt
timeout 900 python -m pytest tests/test_gpu_sdcfr.py tests/test_gpu_sd_train.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -4
( time timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --only sdcfr > gpurun_out/bench_r03a.json 2> gpurun_out/bench_r03a.err ) 2>&1 | tail -4; tail -5 gpurun_out/bench_r03a.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sd_ -c 120 --csv --log-file gpurun_out/launches_sd_r03a.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_launches_sd_r03a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sd_level_mlp_kernel -s 19 -c 1 -f -o gpurun_out/sd_mlp_r03a \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_sd_mlp_r03a.log 2>&1
ls -la gpurun_out | tail -3
