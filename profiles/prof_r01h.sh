set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --trav 151552 --sd-trav 4096 --step-states 1000000 --md-deals 0 --full-games 0 --games 100000"
$CMD > gpurun_out/plain_h.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mccfr_tree -s 3 -c 1 -f -o gpurun_out/prof_mccfr_r01h $CMD > gpurun_out/ncu_h.log 2>&1
tail -c 300 gpurun_out/plain_h.log
ls -la gpurun_out/prof_mccfr_r01h.ncu-rep
