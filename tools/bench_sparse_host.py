#!/usr/bin/env python
"""End-to-end step through host buffers at steady state: one blocking tg_step_host_sparse call per step, the pipelined
form (tg_step_host_sparse_begin / _end over sub-batches in flight), and the dense tg_step_host; with the library's own
split of the host time (enqueue / wait for records / patch).  usage: [n] ; env TG_HOST_THREADS, TG_SPARSE_CHUNKS"""
import ctypes as C
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gym_treasure_game_b200 import PipelinedHostEnv, VectorTreasureGame

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
STEPS = 40
tag = "chunks=%s threads=%s" % (os.environ.get("TG_SPARSE_CHUNKS", "auto"), os.environ.get("TG_HOST_THREADS", "auto"))


def steady(env, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.empty((env.num_envs,), dtype=torch.int32, device="cuda")
    bench.desynchronise(env, torch, lambda: torch.randint(0, 9, (env.num_envs,), generator=g, dtype=torch.int32, device="cuda", out=a))


def host_times(envs):
    tot = [0.0, 0.0, 0.0]
    for e in envs:
        out = (C.c_double * 3)()
        e._L.tg_debug_host_times(e._h, out)
        tot = [t + v for t, v in zip(tot, out)]
    return tot


g = torch.Generator().manual_seed(7)
pool = [torch.randint(0, 9, (n,), generator=g, dtype=torch.int32).pin_memory() for _ in range(16)]

env = VectorTreasureGame(n, seed=0, max_episode_steps=100, auto_reset=True, render=False)
steady(env, 1)
host = env.make_host_buffers()
for name, fn in (("sparse, one call per step", env.step_host_sparse), ("dense", env.step_host)):
    for k in range(12):
        host["actions"] = pool[k % 16]; fn(host)
    host_times([env])
    ts = []
    for k in range(STEPS):
        host["actions"] = pool[k % 16]
        t0 = time.perf_counter(); fn(host); ts.append((time.perf_counter() - t0) * 1e3)
    ht = host_times([env])
    ts.sort()
    print("n=%d %-28s %s: median %.3f ms  p10 %.3f  p90 %.3f  -> %.2f G env-steps/s   host ms/step: enqueue %.3f wait %.3f patch %.3f" % (
        n, name, tag, ts[STEPS // 2], ts[STEPS // 10], ts[STEPS * 9 // 10], n / ts[STEPS // 2] / 1e6, *(1e3 * t / STEPS for t in ht)))
env.close()

for parts in (2, 3, 4):
    p = PipelinedHostEnv(n, parts=parts, seed=0, max_episode_steps=100, auto_reset=True, render=False)
    for k, e in enumerate(p.envs):
        steady(e, 10 + k)

    def run(count):
        for k, (lo, hi) in enumerate(p.ranges):
            p.hosts[k]["actions"] = pool[0][lo:hi]; p.begin(k)
        for it in range(1, count):
            for k, (lo, hi) in enumerate(p.ranges):
                p.end(k); p.hosts[k]["actions"] = pool[it % 16][lo:hi]; p.begin(k)
        for k in range(parts):
            p.end(k)
    run(12)
    host_times(p.envs)
    t0 = time.perf_counter(); run(STEPS); dt = (time.perf_counter() - t0) * 1e3 / STEPS
    ht = host_times(p.envs)
    print("n=%d pipelined, %d parts in flight %s: %.3f ms per step -> %.2f G env-steps/s   host ms/step: enqueue %.3f wait %.3f patch %.3f" % (
        n, parts, tag, dt, n / dt / 1e6, *(1e3 * t / STEPS for t in ht)))
    p.close()
