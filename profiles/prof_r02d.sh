set -x
# r02d (2 GPUs): the multi-GPU tests, the driver's exact N = 2 command line, peer-exchange timings, the reference arm
mkdir -p gpurun_out
nvidia-smi -L
ls oracle/_ref/src/algorithms/ | head
timeout 900 python -m pytest tests/test_gpu_multigpu.py -m gpu -q -x 2>&1 | tail -15
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/scale2_r02d.json 2> gpurun_out/scale2_r02d.err ) 2>&1 | tail -4; echo "scale2 rc $?"; tail -3 gpurun_out/scale2_r02d.err
PEERS_CHECK_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tests/multigpu_peers_check.py 2>&1 | tail -3
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_r02d.json 2> gpurun_out/bench_ref_r02d.err ) 2>&1 | tail -4; tail -3 gpurun_out/bench_ref_r02d.err
( time timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err ) 2>&1 | tail -4; tail -3 gpurun_out/bench_r02d.err
