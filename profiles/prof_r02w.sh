set -x
# r02w (1 GPU): the tensor-core level kernel with two threads per row (sd_level_mlp_team_kernel) against one thread per row
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sdcfr.py tests/test_gpu_sd_train.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -4
for F in 0 1; do
  MS_SD_LEVEL_FORM=$F timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --only sdcfr > gpurun_out/bench_r02w_f$F.json 2> gpurun_out/bench_r02w_f$F.err
  MS_SD_LEVEL_FORM=$F timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sd_level_mlp -s 14 -c 14 --csv --log-file gpurun_out/lat_r02w_f$F.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > /dev/null 2>&1
done
MS_SD_LEVEL_FORM=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sd_level_mlp_team_kernel -s 19 -c 1 -f -o gpurun_out/sd_mlp_team_r02w \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_sd_mlp_team_r02w.log 2>&1
ls -la gpurun_out | tail -4
