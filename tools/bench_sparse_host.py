#!/usr/bin/env python
"""End-to-end step through tg_step_host_sparse / tg_step_host at steady state: ms per call.  usage: [n] [chunks...]"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gym_treasure_game_b200 import VectorTreasureGame
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
env = VectorTreasureGame(n, seed=0, max_episode_steps=100, auto_reset=True, render=False)
g = torch.Generator(device="cuda").manual_seed(1)
a = torch.empty((n,), dtype=torch.int32, device="cuda")
new_actions = lambda: torch.randint(0, 9, (n,), generator=g, dtype=torch.int32, device="cuda", out=a)
bench.desynchronise(env, torch, new_actions)
host = env.make_host_buffers()
pool = [new_actions().cpu().pin_memory() for _ in range(16)]
for name, fn in (("sparse", env.step_host_sparse), ("dense", env.step_host)):
    for k in range(12):
        host["actions"] = pool[k % 16]; fn(host)
    ts = []
    for k in range(40):
        host["actions"] = pool[k % 16]
        t0 = time.perf_counter(); fn(host); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    print("n=%d %s TG_SPARSE_CHUNKS=%s threads=%s: median %.3f ms  p10 %.3f  p90 %.3f  -> %.2f G env-steps/s" % (
        n, name, os.environ.get("TG_SPARSE_CHUNKS", "auto"), os.environ.get("TG_HOST_THREADS", "auto"), ts[20], ts[4], ts[36], n / ts[20] / 1e6))
