set -x
# r02i (8 GPUs): final defaults: the driver's scaling command lines N = 8, 4, 2, 1 back to back
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multigpu.py -m gpu -q -x 2>&1 | tail -3
for N in 8 4 2; do
  ( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale${N}_r02i.json 2> gpurun_out/scale${N}_r02i.err ) 2>&1 | tail -4; echo "scale$N rc $?"; tail -2 gpurun_out/scale${N}_r02i.err
done
( time timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu > gpurun_out/scale1_r02i.json 2> gpurun_out/scale1_r02i.err ) 2>&1 | tail -4
