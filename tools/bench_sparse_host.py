import os, sys, time, torch
sys.path.insert(0, '/root/repo')
from gym_treasure_game_b200 import VectorTreasureGame
n = 1 << 20
mode = sys.argv[1]
env = VectorTreasureGame(n, seed=0, max_episode_steps=100, auto_reset=True, render=False)
if mode in ("dev250", "dev250_1t"):
    a = torch.empty((n,), dtype=torch.int32, device="cuda")
    for k in range(250):
        env.step_raw(torch.randint(0, 9, (n,), dtype=torch.int32, device="cuda", out=a))
if mode == "dev250_1t":
    torch.set_num_threads(1)
host = env.make_host_buffers()
pool = [torch.randint(0, 9, (n,), dtype=torch.int32).pin_memory() for _ in range(8)]
for k in range(30):
    host["actions"] = pool[k % 8]; env.step_host_sparse(host)
t0 = time.perf_counter()
for k in range(20):
    host["actions"] = pool[k % 8]; env.step_host_sparse(host)
print(mode, "ms/step", (time.perf_counter() - t0) / 20 * 1e3)
