#!/usr/bin/env python
"""Where the time of one step kernel goes: per-CTA %globaltimer stamps at the phase boundaries
(tg_debug_phase_buffer).  Prints, per scenario, the kernel's event time and the median / max over CTAs of
each phase's end (ns after the first CTA started)."""
import ctypes as C
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_treasure_game_b200 import VectorTreasureGame

NAMES = ["start", "levels", "phase-A", "sort", "t0-left-B", "phase-B", "phase-C", "stats"]


def run(n, scenario, flush_mb=512, reps=7):
    env = VectorTreasureGame(n, seed=0, render=False, auto_reset=True, max_episode_steps=100)
    stamps_all = torch.zeros(8192 * 8 + 8192 * 512, dtype=torch.int64, device="cuda")   # phase stamps, then the chunk stamps (bench_chunks.py)
    stamps = stamps_all[: 8192 * 8].view(8192, 8)
    flush = torch.empty(max(flush_mb, 1) << 20, dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.empty((n,), dtype=torch.int32, device="cuda")
    if scenario == "steady":      # the bench workload: desynchronised episodes, 1 % resets per step
        import bench
        bench.desynchronise(env, torch, lambda: torch.randint(0, 9, (n,), generator=g, dtype=torch.int32, device="cuda", out=acts))
    else:
        for _ in range(30):
            env.step_raw(torch.randint(0, 9, (n,), generator=g, dtype=torch.int32, device="cuda", out=acts))
    env._L.tg_debug_phase_buffer(env._h, C.c_void_p(stamps_all.data_ptr()))
    rows = []
    for rep in range(reps):
        if scenario in ("random", "steady"):
            torch.randint(0, 9, (n,), generator=g, dtype=torch.int32, device="cuda", out=acts)
        else:
            env.reset(); acts.fill_(2)          # up_ladder at the start cell: nothing is runnable
        stamps.zero_()
        if flush_mb:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); env.step_raw(acts); e.record(); e.synchronize()
        t = stamps[stamps[:, 0] > 0].cpu()
        t0 = int(t[:, 0].min())
        rel = (t - t0).float()
        rows.append((s.elapsed_time(e) * 1e3, rel.median(0).values.tolist(), rel.max(0).values.tolist(), t.shape[0]))
    rows.sort(key=lambda r: r[0])
    us, med, mx, ctas = rows[len(rows) // 2]
    print("n=%8d %-12s flush=%4d MiB  event %7.1f us  CTAs %d" % (n, scenario, flush_mb, us, ctas))
    print("    phase end (us after first CTA start)  " + "  ".join("%8s" % k for k in NAMES))
    print("    median over CTAs                      " + "  ".join("%8.1f" % (v / 1e3) for v in med[:8]))
    print("    max over CTAs                         " + "  ".join("%8.1f" % (v / 1e3) for v in mx[:8]))
    env._L.tg_debug_phase_buffer(env._h, None)
    env.close()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        for n in (4096, 131072, 1 << 20):
            run(n, sys.argv[1], 512)
        sys.exit(0)
    for n in (4096, 1 << 20):
        for sc in ("not_runnable", "random", "steady"):
            for fl in (512, 0):
                run(n, sc, fl)
