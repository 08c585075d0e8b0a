#!/usr/bin/env python
"""Timing of the drawer helpers on the device: blend() of K states onto one picture, and onto K pictures."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_treasure_game_b200 import VectorTreasureGame

n = 4096
env = VectorTreasureGame(n, seed=0, max_episode_steps=100, auto_reset=True)
g = torch.Generator(device="cuda").manual_seed(1)
for _ in range(60):
    env.step_raw(torch.randint(0, 9, (n,), generator=g, dtype=torch.int32, device="cuda"))


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / reps


for k in (64, 1024, 4096):
    pic = env.draw_background_to_surface()
    t = timed(lambda: env.blend(pic, 0.05, 0.05, first=0, count=k, accumulate=True))
    print("accumulate %5d states on one picture: %8.3f ms  (%.2f us per state)" % (k, t, t * 1e3 / k))
k = 1024
pics = env.draw_background_to_surface().unsqueeze(0).repeat(k, 1, 1, 1).contiguous()
t = timed(lambda: env.blend(pics, 0.5, 0.5, first=0, count=k))
print("blend %d states onto %d pictures: %8.3f ms  (%.1f GB/s read+write)" % (k, k, t, 2 * pics.numel() / t / 1e6))
