set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --trav 151552 --sd-trav 8192 --step-states 4000000 --md-deals 16384 --md-log2-capacity 24 --md-trav 113664 --full-games 200000"
$CMD > gpurun_out/plain_g.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01g.csv $CMD > gpurun_out/ncu_g.log 2>&1
$CMD > gpurun_out/plain_g2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mccfr_tree -s 3 -c 1 -f -o gpurun_out/prof_mccfr_r01g $CMD > gpurun_out/ncu_g2.log 2>&1
tail -c 400 gpurun_out/plain_g.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r01g.csv
