set -x
# r02n (1 GPU): full capture of sd_level_mlp_kernel<1> (level 5 of a 65 536-traversal batch: 12 288 tiles)
mkdir -p gpurun_out
python profiles/summarise_capture.py x --hash-only --sources scopa_b200/csrc/ms_sdcfr.cu scopa_b200/csrc/ms_state.cuh > gpurun_out/sha_sd_r02n.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sd_level_mlp_kernel -s 19 -c 1 -f -o gpurun_out/sd_mlp_r02n \
    python bench.py --steps 3 --warmup 3 --no-cpu --only sdcfr > gpurun_out/ncu_sd_mlp_r02n.log 2>&1
ls -la gpurun_out | tail -3
