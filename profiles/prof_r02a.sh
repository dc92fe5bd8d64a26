set -x
# r02a (prepared at the end of round 1, when the GPU budget was spent): first things to run in round 2.
#   gpurun --timeout 1500 -- 'bash profiles/prof_r02a.sh > gpurun_out/prof_r02a.log 2>&1'
# 1. the new GPU tests (fused optimiser, average-policy kernels) and the SDCFR drop-in tests that now go through them
python -m pytest tests/test_gpu_sd_train.py tests/test_gpu_dropin.py tests/test_gpu_sdcfr.py -x -q 2>&1 | tail -3
# 2. torch-free parity + timing of both kernels (CUDA events inside the checker)
C=./tests/emu/_build/sd_train_check
for a in "128 10 100000 100" "128 1 100000 200" "32 10 100 100"; do $C $a; done
for a in "avgpol 20 200 100"; do $C $a; done
# 2b. the cluster form of the optimiser: first run on a device (bit parity with its 8-CTA emulation, then timing)
for a in "cluster 17 3 40 0" "cluster 128 6 4096 100" "cluster 128 1 4096 200" "cluster 32 10 100 100"; do timeout 120 $C $a; done
for a in "sample 128 10 100000 200" "sample 128 3 130 0" "sample 32 3 32 0"; do timeout 60 $C $a; done
SCOPA_B200_UNVERIFIED=1 python -m pytest tests/test_gpu_sd_train.py -x -q -m gpu -k "cluster or sampler" 2>&1 | tail -3
SD_CHECK_TIMING_ONLY=1 $C avgpol 100 1 500
SD_CHECK_TIMING_ONLY=1 $C avgpol 100 738 100
# 3. full captures (each after its plain run above exited 0): optimiser kernel, average-policy kernel
ncu --set full --clock-control none --import-source on -k regex:sd_train_kernel -s 3 -c 1 -f -o gpurun_out/prof_sd_train_r02a $C 128 10 100000 5 > gpurun_out/ncu_r02a_1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sd_train_cluster -s 3 -c 1 -f -o gpurun_out/prof_sd_train_cluster_r02a $C cluster 128 10 100000 5 > gpurun_out/ncu_r02a_1b.log 2>&1
SD_CHECK_TIMING_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:sd_avgpol_kernel -s 3 -c 1 -f -o gpurun_out/prof_sd_avgpol_r02a $C avgpol 100 738 5 > gpurun_out/ncu_r02a_2.log 2>&1
# 4. bench line (the sdcfr.train section compares the fused optimiser with the PyTorch step) + launch list
SCOPA_B200_BENCH_CLUSTER=1 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r02a.json').readline()); print(json.dumps(d['sdcfr'].get('train'), indent=1)); print(d['sdcfr'].get('roofline'))"
ls -la gpurun_out/*r02a*
