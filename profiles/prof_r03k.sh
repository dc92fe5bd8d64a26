set -x
# r03k (1 GPU): the full -m gpu suite on the final tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
