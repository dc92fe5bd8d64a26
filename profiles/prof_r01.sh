set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --trav 65536"
$CMD > gpurun_out/plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_a.log 2>&1
$CMD > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mccfr_batch -s 3 -c 1 -f -o gpurun_out/prof_mccfr_r01 $CMD > gpurun_out/ncu_b.log 2>&1
$CMD > gpurun_out/plain_c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'rollout_kernel|deal_kernel|cfr_kernel' -c 3 -f -o gpurun_out/prof_env_r01 $CMD > gpurun_out/ncu_c.log 2>&1
tail -3 gpurun_out/ncu_b.log gpurun_out/ncu_c.log
ls -la gpurun_out/
