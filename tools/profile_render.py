#!/usr/bin/env python
"""Small render workload for ncu captures (BASELINE configs[3] shape at 4096 envs = 5.2 GB per launch)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_treasure_game_b200 import VectorTreasureGame  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = VectorTreasureGame(n, seed=0, max_episode_steps=100, auto_reset=True)
g = torch.Generator(device="cuda").manual_seed(0)
frames = torch.empty((n, 624, 672, 3), dtype=torch.uint8, device="cuda")
for k in range(30):
    env.step_raw(torch.randint(0, 9, (n,), generator=g, dtype=torch.int32, device="cuda"))
for k in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); env.render(out=frames); e.record(); e.synchronize()
    print("render %d frames: %.3f ms  %.1f GB/s" % (n, s.elapsed_time(e), n * 1257984 / s.elapsed_time(e) / 1e6))
print("checksum", int(frames[::97].to(torch.int64).sum()))
