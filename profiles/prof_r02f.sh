set -x
# r02f (1 GPU): kernel variants (Philox inlined vs called), rollout / static captures with the source hash recorded on the box
mkdir -p gpurun_out
python profiles/summarise_capture.py x --hash-only --sources scopa_b200/csrc/ms_static_walk.cuh scopa_b200/csrc/ms_solver.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_static_r02f.txt
python profiles/summarise_capture.py x --hash-only --sources scopa_b200/csrc/ms_env.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_env_r02f.txt
for lib in "" scopa_b200/variants/lib_inline.so; do
  echo "=== lib [$lib]"
  SCOPA_B200_LIB=$lib timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('updates/s %.1f G  ms/step %.4f  kernel_ms %.4f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['kernel_ms']))"
done
SCOPA_B200_LIB=scopa_b200/variants/lib_inline.so timeout 300 python -m pytest tests/test_gpu_solver.py -m gpu -q -x -k "frozen_sigma or shards" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -4
