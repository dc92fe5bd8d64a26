set -x
# r03d (2 GPUs): the collected multi-GPU tests and the driver's N = 2 line on the final sources
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multigpu.py -m gpu -q -x 2>&1 | tail -4
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/scale2_r03d.json 2> gpurun_out/scale2_r03d.err ) 2>&1 | tail -3
echo "rc=$?"
