set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --trav 113664 --sd-trav 8192 --step-states 4000000"
$CMD > gpurun_out/plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01d.csv $CMD > gpurun_out/ncu_a.log 2>&1
$CMD > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mccfr_batch -s 3 -c 1 -f -o gpurun_out/prof_mccfr_r01d $CMD > gpurun_out/ncu_b.log 2>&1
$CMD > gpurun_out/plain_c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'deal_kernel|step_kernel|mccfr_es_kernel' -s 1 -c 4 -f -o gpurun_out/prof_env_r01d $CMD > gpurun_out/ncu_c.log 2>&1
$CMD > gpurun_out/plain_d.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'sd_forward_kernel' -s 20 -c 4 -f -o gpurun_out/prof_sd_r01d $CMD > gpurun_out/ncu_d.log 2>&1
cat gpurun_out/plain_a.log | tail -c 1500
ls -la gpurun_out/*.ncu-rep
