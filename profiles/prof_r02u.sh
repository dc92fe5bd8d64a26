set -x
mkdir -p gpurun_out
for V in 2 6 10 14; do
  MS_SD_VARIANT=$V timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sd_level_mlp -s 14 -c 14 --csv --log-file gpurun_out/lat_v$V.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > /dev/null 2>&1
done
ls gpurun_out | grep lat_
