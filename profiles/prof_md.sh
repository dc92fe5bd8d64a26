set -x
CMD="python profiles/md_probe.py 65536:27 warm=12"
$CMD > gpurun_out/md_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:md_mccfr -s 13 -c 1 -f -o gpurun_out/prof_md_r01f $CMD > gpurun_out/md_ncu.log 2>&1
cat gpurun_out/md_plain.log | tail -3
ls -la gpurun_out/*.ncu-rep
