#!/usr/bin/env python
"""Fixed cost of one tg_step launch: every option not runnable (go_left at the start cell)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_treasure_game_b200 import VectorTreasureGame
for n in (4096, 65536, 1 << 20):
    env = VectorTreasureGame(n, seed=0, render=False, auto_reset=False)
    a0 = torch.zeros(n, dtype=torch.int32, device="cuda")          # go_left: never runnable at the start
    a3 = torch.full((n,), 3, dtype=torch.int32, device="cuda")     # down_ladder: runnable for everyone, ~18 ticks
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, a, prep in (("none runnable", a0, False), ("all down_ladder", a3, True)):
        ts = []
        for k in range(40):
            if prep:
                env.reset()
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); env.step_raw(a); e.record(); e.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        ts.sort()
        print("n=%8d %-16s median %.1f us  min %.1f us" % (n, name, ts[len(ts) // 2], ts[0]))
    env.close()
