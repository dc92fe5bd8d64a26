set -x
# r01i: launch list of the bench (reduced sizes so that the run under ncu stays short; --curve-max keeps the 10 M-traversal
# curves from eating the launch budget) + full captures of the kernels this stage changed
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --trav 151552 --sd-trav 8192 --step-states 4000000 --md-deals 16384 --md-log2-capacity 24 --md-trav 113664 --full-games 200000 --curve-max 10000"
$CMD > gpurun_out/plain_i.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01i.csv $CMD > gpurun_out/ncu_i.log 2>&1
python profiles/e2e_probe.py > gpurun_out/e2e_probe_i.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:deal_kernel -s 4 -c 1 -f -o gpurun_out/prof_deal_r01i python profiles/e2e_probe.py > gpurun_out/ncu_i2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:full_rollout -s 2 -c 1 -f -o gpurun_out/prof_full_r01j python profiles/e2e_probe.py > gpurun_out/ncu_i3.log 2>&1
tail -c 300 gpurun_out/plain_i.log
cat gpurun_out/e2e_probe_i.log
ls -la gpurun_out/*r01i* gpurun_out/*r01j*
