set -x
# r03h (1 GPU): capture of mccfr_inplace_tree_kernel (reference semantics, one dependent chain)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mccfr_inplace_tree_kernel -s 1 -c 1 -f -o gpurun_out/inplace_r03h \
    python bench.py --steps 3 --warmup 3 --no-cpu --only mccfr_in_place > gpurun_out/ncu_inplace_r03h.log 2>&1
ls -la gpurun_out | tail -3
