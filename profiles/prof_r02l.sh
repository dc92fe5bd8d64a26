set -x
# r02l (2 GPUs): the sharded multi-deal table: collected 2-rank tests, then the timing form of the check
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_multideal.py -m gpu -q -x 2>&1 | tail -8
MD_CHECK_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_md_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -12
