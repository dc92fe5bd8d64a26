set -x
# r02g (8 GPUs): push-form peer exchange: the 2-rank tests, then the driver's scaling command lines N = 8, 4, 2, 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multigpu.py -m gpu -q -x 2>&1 | tail -5
for N in 8 4 2; do
  ( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale${N}_r02g.json 2> gpurun_out/scale${N}_r02g.err ) 2>&1 | tail -4; echo "scale$N rc $?"; tail -2 gpurun_out/scale${N}_r02g.err
done
( time timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu > gpurun_out/scale1_r02g.json 2> gpurun_out/scale1_r02g.err ) 2>&1 | tail -4
PEERS_CHECK_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29719 tests/multigpu_peers_check.py 2>&1 | tail -2
