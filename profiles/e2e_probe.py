"""Where the end-to-end rollout call (seeds in host memory -> actions + rewards in host memory) spends its time:
PCIe copies alone, the two kernels alone, and the pipelined C-ABI call.  Run on the GPU box:
    python profiles/e2e_probe.py
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scopa_b200 import _lib  # noqa: E402
from scopa_b200.batch import BatchedMiniScopa, rollout_random_host  # noqa: E402

G = 1_000_000
dev = torch.device("cuda:0")
seeds_np = np.arange(1, G + 1, dtype=np.int64)
h_seeds = torch.from_numpy(seeds_np).pin_memory()
h_act = torch.empty((G, 8), dtype=torch.uint8).pin_memory()
h_rew = torch.empty((G, 2), dtype=torch.float32).pin_memory()
d_seeds = torch.empty(G, dtype=torch.int64, device=dev)
d_act = torch.empty((G, 8), dtype=torch.uint8, device=dev)
d_rew = torch.empty((G, 2), dtype=torch.float32, device=dev)


def ev_time(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def wall_time(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / reps


t_h2d = ev_time(lambda: d_seeds.copy_(h_seeds, non_blocking=True))
t_d2h_a = ev_time(lambda: h_act.copy_(d_act, non_blocking=True))
t_d2h_r = ev_time(lambda: h_rew.copy_(d_rew, non_blocking=True))
print(f"H2D  8 MB seeds   : {t_h2d:.3f} ms = {8e-3 * G / t_h2d / 1e3:.1f} GB/s")
print(f"D2H  8 MB actions : {t_d2h_a:.3f} ms = {8e-3 * G / t_d2h_a / 1e3:.1f} GB/s")
print(f"D2H  8 MB rewards : {t_d2h_r:.3f} ms = {8e-3 * G / t_d2h_r / 1e3:.1f} GB/s")

s2 = torch.cuda.Stream()


def both():
    d_seeds.copy_(h_seeds, non_blocking=True)
    with torch.cuda.stream(s2):
        h_act.copy_(d_act, non_blocking=True)
        h_rew.copy_(d_rew, non_blocking=True)


t_both = wall_time(both)
print(f"H2D 8 MB || D2H 16 MB (two streams, wall): {t_both:.3f} ms")

b = BatchedMiniScopa(dev)
d_seeds.copy_(h_seeds)
t_deal = ev_time(lambda: b.reset(d_seeds))
t_roll = ev_time(lambda: b.rollout_random(philox_seed=1, actions=d_act, rewards=d_rew))
print(f"deal_kernel 1 M seeds    : {t_deal:.3f} ms")
print(f"rollout_kernel 1 M games : {t_roll:.3f} ms")

lib = _lib.load()
has_chunk = hasattr(lib, "ms_debug_set_host_chunk")
for chunk in ([0] if not has_chunk else [32768, 65536, 131072, 262144, 524288, 1048576]):
    if has_chunk:
        lib.ms_debug_set_host_chunk(chunk)
    t = wall_time(lambda: rollout_random_host(h_seeds, 1, 0, h_act, h_rew))
    print(f"ms_rollout_random_host chunk {chunk:>8}: {t:.3f} ms = {8e-9 * G / (t * 1e-3):.2f} G env steps/s")

# ---- 40-card game: FullDeck(seed) + deal + 36-ply rollout
from scopa_b200 import full as fs  # noqa: E402

p_act = torch.empty((G, fs.PLIES), dtype=torch.uint8).pin_memory()
fb = fs.BatchedFullScopa(dev)
t_fdeal = ev_time(lambda: fb.reset(d_seeds), reps=10)
t_froll = ev_time(lambda: fb.rollout_random(philox_seed=1), reps=10)
print(f"full deck + init kernels 1 M seeds : {t_fdeal:.3f} ms")
print(f"full_rollout_kernel 1 M games      : {t_froll:.3f} ms")
for chunk in ([0] if not has_chunk else [65536, 131072, 262144, 524288]):
    if has_chunk:
        lib.ms_debug_set_host_chunk(chunk)
    t = wall_time(lambda: _lib.check(lib.ms_full_rollout_random_host(h_seeds.data_ptr(), G, 1, 0, p_act.data_ptr(), h_rew.data_ptr())), reps=10)
    print(f"ms_full_rollout_random_host chunk {chunk:>8}: {t:.3f} ms = {36e-9 * G / (t * 1e-3):.2f} G env steps/s")
if has_chunk:
    lib.ms_debug_set_host_chunk(0)
