#!/usr/bin/env python
"""Per-tick latency of one option class: every env runs the same option from the same kind of state."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_treasure_game_b200 import VectorTreasureGame
for n in (4096, 1 << 20):
    env = VectorTreasureGame(n, seed=0, render=False, auto_reset=False)
    def act(k): return torch.full((n,), k, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    for rep in range(12):
        env.reset(); env.clear_stats()
        # up_ladder at the start cell is not runnable: the cost of a step in which no option runs
        seq = [("not_runnable", 2), ("down_ladder", 3), ("go_left", 0), ("interact", 4), ("go_right", 1)]
        for name, k in seq:
            flush.zero_()
            t0 = env.stats()["primitive_ticks"]
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); env.step_raw(act(k)); e.record(); e.synchronize()
            ticks = (env.stats()["primitive_ticks"] - t0) / n
            res.setdefault(name, []).append((s.elapsed_time(e) * 1e3, ticks))
    for name, v in res.items():
        v.sort()
        us, ticks = v[len(v) // 2]
        print("n=%8d %-12s median %7.1f us  mean ticks/env %5.1f" % (n, name, us, ticks))
    env.close()
