#!/usr/bin/env python
"""Target for ncu: a 1,048,576-env batch in the steady state of the bench workload (desynchronised episode
phases), then a few tg_step launches with fresh i.i.d. actions.  usage: profile_step.py [n] [steps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gym_treasure_game_b200 import VectorTreasureGame

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
env = VectorTreasureGame(n, seed=0, max_episode_steps=100, auto_reset=True, render=False)
g = torch.Generator(device="cuda").manual_seed(1234)
acts = torch.empty((n,), dtype=torch.int32, device="cuda")
new_actions = lambda: torch.randint(0, 9, (n,), generator=g, dtype=torch.int32, device="cuda", out=acts)
bench.desynchronise(env, torch, new_actions)
env.clear_stats()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
ts = []
for k in range(steps):
    new_actions(); flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); env.step_raw(acts); e.record(); e.synchronize()
    ts.append(s.elapsed_time(e) * 1e3)
st = env.stats()
print("n=%d steps=%d  us/step %s  episodes %d  ticks/step %.2f  runnable %.3f" % (
    n, steps, " ".join("%.1f" % t for t in ts), st["episodes"], st["primitive_ticks"] / st["gym_steps"], st["runnable_steps"] / st["gym_steps"]))
