set -x
# r02e (8 GPUs): the driver's exact scaling command lines, N = 2, 4, 8, back to back, default flags
mkdir -p gpurun_out
nvidia-smi -L | head -8
for N in 8 4 2; do
  ( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale${N}_r02e.json 2> gpurun_out/scale${N}_r02e.err ) 2>&1 | tail -4; echo "scale$N rc $?"; tail -2 gpurun_out/scale${N}_r02e.err
done
( time timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu > gpurun_out/scale1_r02e.json 2> gpurun_out/scale1_r02e.err ) 2>&1 | tail -4
