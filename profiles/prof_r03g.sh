set -x
# r03g (4 GPUs): the driver's N = 4 line with and without binding each rank to its GPU's NUMA-local CPUs
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14
for A in 0 1; do
  SCOPA_B200_BENCH_AFFINITY=$A timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29510 + A)) \
      bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/scale4_r03g_aff$A.json 2> gpurun_out/scale4_r03g_aff$A.err
  echo "aff=$A rc=$?"
done
