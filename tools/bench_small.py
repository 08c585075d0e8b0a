#!/usr/bin/env python
"""Step latency at small batch sizes (BASELINE configs[1] = 4096 envs): CUDA events, L2 flushed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_treasure_game_b200 import VectorTreasureGame  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [4096]:
    env = VectorTreasureGame(n, seed=0, max_episode_steps=100, auto_reset=True, render=False)
    g = torch.Generator(device="cuda").manual_seed(7)
    pool = [torch.randint(0, 9, (n,), generator=g, dtype=torch.int32, device="cuda") for _ in range(400)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for k in range(100):
        env.step_raw(pool[k])
    tot = 0.0
    ts = []
    for k in range(300):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); env.step_raw(pool[100 + k]); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    mean = sum(ts) / len(ts)
    print("n=%d  mean %.1f us  median %.1f us  p10 %.1f  p90 %.1f   %.1f M env-steps/s" % (
        n, mean * 1e3, ts[len(ts) // 2] * 1e3, ts[30] * 1e3, ts[270] * 1e3, n / mean / 1e3))
    env.close()
