set -x
# r02q (1 GPU): the eight variants of sd_level_mlp_kernel<1, VAR> (MS_SD_VARIANT), whole-traversal bench of each
mkdir -p gpurun_out
for V in 0 1 2 3 4 5 6 7; do
  MS_SD_VARIANT=$V timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --only sdcfr > gpurun_out/bench_r02q_v$V.json 2> gpurun_out/bench_r02q_v$V.err
done
MS_SD_VARIANT=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sd_level_mlp -c 60 --csv --log-file gpurun_out/launches_sd_r02q_v0.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > /dev/null 2>&1
MS_SD_VARIANT=6 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sd_level_mlp -c 60 --csv --log-file gpurun_out/launches_sd_r02q_v6.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > /dev/null 2>&1
ls gpurun_out | tail -3
