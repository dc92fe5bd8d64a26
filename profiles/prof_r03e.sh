set -x
# r03e (1 GPU): the multi-deal tests incl. the sharding entry points in a world of one; smoke()
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multideal.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
