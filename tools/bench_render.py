#!/usr/bin/env python
"""Render throughput: single-layout and mixed-layout batches (frames/s, TB/s of frame bytes written)."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import py_oracle as po
from gym_treasure_game_b200 import VectorTreasureGame, Level
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
def timed(env, frames, iters=8):
    for _ in range(3): env.render(out=frames)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); env.render(out=frames); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2] * 1e-3
frames = torch.empty((n, 624, 672, 3), dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda").manual_seed(3)
for name, kw in (("single layout", {}),
                 ("two layouts interleaved", dict(levels=[Level.from_strings(*po.level_to_strings(l)) for l in (po.default_level(), po.mirrored_level(po.default_level()))],
                                                  level_ids=(np.arange(n) % 2).astype(np.uint8)))):
    env = VectorTreasureGame(n, seed=1, max_episode_steps=100, auto_reset=True, **kw)
    for _ in range(60):
        env.step_raw(torch.randint(0, 9, (n,), generator=g, dtype=torch.int32, device="cuda"))
    t = timed(env, frames)
    print("%-26s n=%d: %.3f ms  %.2f M frames/s  %.2f TB/s" % (name, n, t * 1e3, n / t / 1e6, n * 1258024 / t / 1e12))
    env.close()
