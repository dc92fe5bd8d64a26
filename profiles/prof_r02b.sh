set -x
# r02b: first device run of round 2: mccfr_static_kernel (new headline), touched-through-delta, in-place many, the
# restructured bench.py, and the kernels round 1 left unverified (prof_r02a.sh).
#   gpurun --timeout 2400 -- 'bash profiles/prof_r02b.sh > gpurun_out/prof_r02b.log 2>&1'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv
# 1. parity
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multigpu.py 2>&1 | tail -15
# 2. the driver's two arms, N = 1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench rc $?"; tail -5 gpurun_out/bench_r02b.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_r02b.json 2> gpurun_out/bench_ref_r02b.err; echo "ref rc $?"
# 3. launch list of the headline sections + full capture of the headline kernel (each after the plain run above exited 0)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02b.csv \
    python bench.py --steps 5 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_launches_r02b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mccfr_static_kernel -s 4 -c 1 -f -o gpurun_out/mccfr_r02b \
    python bench.py --steps 5 --warmup 3 --no-extras --no-cpu > gpurun_out/ncu_full_r02b.log 2>&1
# 4. the kernels round 1 left unverified
SCOPA_B200_UNVERIFIED=1 timeout 900 python -m pytest tests/test_gpu_sd_train.py tests/test_gpu_sdcfr.py -q -m gpu 2>&1 | tail -8
C=./tests/emu/_build/sd_train_check
for a in "cluster 17 3 40 0" "cluster 128 6 4096 100" "cluster 128 1 4096 200" "cluster 32 10 100 100"; do timeout 120 $C $a; done
for a in "sample 128 10 100000 200" "sample 128 3 130 0" "sample 32 3 32 0"; do timeout 60 $C $a; done
for a in "128 10 100000 100" "avgpol 20 200 100"; do timeout 120 $C $a; done
SD_CHECK_TIMING_ONLY=1 timeout 120 $C avgpol 100 738 100
ls -la gpurun_out/
