#!/usr/bin/env python
"""Small workload touching every kernel once or twice (for compute-sanitizer runs)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_treasure_game_b200 import Level, VectorTreasureGame  # noqa: E402

lv = Level.default()
for levels, ids in (([lv], None), ([lv, lv.mirrored()], (np.arange(700) % 2).astype(np.uint8))):
    n = 700
    env = VectorTreasureGame(n, seed=1, max_episode_steps=10, auto_reset=True, levels=levels, level_ids=ids)
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in range(25):
        a = torch.randint(0, 9, (n,), generator=g, dtype=torch.int32, device="cuda")
        obs, rew, done, info = env.step(a, want_available=True)
    env.primitive_step(torch.randint(0, 7, (n,), generator=g, dtype=torch.int32, device="cuda"))
    m = env.available_mask
    st = env.get_state()
    env.set_state(st)
    env.init_with_state(obs.double().cpu().numpy())
    fr = env.render()
    host = env.make_host_buffers()
    host["actions"].copy_(torch.randint(0, 9, (n,), dtype=torch.int32))
    env.step_host(host)
    env.reset(mask=(torch.arange(n) % 3 == 0))
    tape = [np.random.default_rng(i).random(4000) for i in range(n)]
    env.set_draw_tape(tape)
    env.reset()
    for t in range(5):
        env.step_raw(torch.randint(0, 9, (n,), generator=g, dtype=torch.int32, device="cuda"))
    torch.cuda.synchronize()
    print("ok", int(fr.sum()), env.stats())
    env.close()
