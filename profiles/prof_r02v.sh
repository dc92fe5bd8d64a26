set -x
# r02v (1 GPU): smoke() with the traversal / multi-deal checks, the SDCFR section with the cluster-optimiser drop-in variant,
# first ncu captures of the optimiser kernels (sd_train_kernel, sd_train_cluster_kernel)
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12
( time timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --only sdcfr > gpurun_out/bench_r02v.json 2> gpurun_out/bench_r02v.err ) 2>&1 | tail -4; tail -5 gpurun_out/bench_r02v.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sd_train_kernel -s 3 -c 1 -f -o gpurun_out/sd_train_r02v \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_sd_train_r02v.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sd_train_cluster_kernel -s 3 -c 1 -f -o gpurun_out/sd_train_cluster_r02v \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_sd_train_cluster_r02v.log 2>&1
ls -la gpurun_out | tail -5
