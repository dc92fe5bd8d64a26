set -x
# r03f (8 GPUs, final sources): the driver's scaling command lines (default flags) at N = 8 / 4 / 2, then the sharded multi-deal check with timing at 8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4 2; do
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale${N}_r03f.json 2> gpurun_out/scale${N}_r03f.err ) 2>&1 | tail -3
  echo "N=$N rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/scale${N}_r03f.err | tail -3
done
MD_CHECK_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_md_check.py 2>&1 | grep "MD_CHECK" 
