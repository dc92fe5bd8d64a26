set -x
# r02m (1 GPU): SDCFR forward split into sd_level_mlp_kernel (tcgen05, biases through the tensor cores) + sd_expand_kernel
mkdir -p gpurun_out
python profiles/summarise_capture.py x --hash-only --sources scopa_b200/csrc/ms_sdcfr.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_sd_r02m.txt
timeout 900 python -m pytest tests/test_gpu_sdcfr.py tests/test_gpu_sd_train.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -8
( time timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r02m.json 2> gpurun_out/bench_r02m.err ) 2>&1 | tail -4; tail -5 gpurun_out/bench_r02m.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sd_ -c 120 --csv --log-file gpurun_out/launches_sd_r02m.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_launches_sd_r02m.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:sd_level_mlp_kernel<1>" -s 19 -c 1 -f -o gpurun_out/sd_mlp_r02m \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_sd_mlp_r02m.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:sd_expand_kernel" -s 19 -c 1 -f -o gpurun_out/sd_expand_r02m \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_sd_expand_r02m.log 2>&1
ls -la gpurun_out | tail -8
