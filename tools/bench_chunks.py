#!/usr/bin/env python
"""Per-chunk timing of the step kernel's phase B (debug stamps, lane 0 of the warp that ran the chunk): duration by
sort class, and the critical path of a CTA.  Steady-state workload of bench.py.  usage: bench_chunks.py [n]"""
import ctypes as C
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gym_treasure_game_b200 import VectorTreasureGame

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
env = VectorTreasureGame(n, seed=0, max_episode_steps=100, auto_reset=True, render=False)
g = torch.Generator(device="cuda").manual_seed(1234)
acts = torch.empty((n,), dtype=torch.int32, device="cuda")
new_actions = lambda: torch.randint(0, 9, (n,), generator=g, dtype=torch.int32, device="cuda", out=acts)
bench.desynchronise(env, torch, new_actions)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
grid_max = 8192
stamps = torch.zeros(grid_max * 8 + grid_max * 512, dtype=torch.int64, device="cuda")
env._L.tg_debug_phase_buffer(env._h, C.c_void_p(stamps.data_ptr()))
NAMES = ["walk", "ladder", "drop", "jump", "interact"]
for rep in range(3):
    new_actions(); stamps.zero_(); flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); env.step_raw(acts); e.record(); e.synchronize()
    ph = stamps[: grid_max * 8].view(grid_max, 8)
    grid = int((ph[:, 0] > 0).sum())
    ph = ph[:grid].cpu()
    ch = stamps[grid_max * 8: grid_max * 8 + grid * 512].view(grid, 128, 4).cpu()
    t0 = int(ph[:, 0].min())
    print("rep %d: event %.1f us, %d CTAs; phase B start (median) %.1f us, end median %.1f / max %.1f us" % (
        rep, s.elapsed_time(e) * 1e3, grid, float((ph[:, 3] - t0).float().median()) / 1e3,
        float((ph[:, 5] - t0).float().median()) / 1e3, float((ph[:, 5] - t0).float().max()) / 1e3))
    valid = ch[:, :, 0] > 0
    dur = (ch[:, :, 1] & 0xFFFFFFFF).float() / 1e3
    cls = (ch[:, :, 1] >> 32) & 0xFF
    nt = (ch[:, :, 1] >> 40) & 0xFFFF
    for c in range(5):
        m = valid & (cls == c)
        if int(m.sum()) == 0: continue
        d = dur[m]
        seg = [((ch[:, :, 2] & 0xFFFFFFFF).float() / 1e3)[m], ((ch[:, :, 2] >> 32).float() / 1e3)[m],
               ((ch[:, :, 3] & 0xFFFFFFFF).float() / 1e3)[m], ((ch[:, :, 3] >> 32).float() / 1e3)[m]]
        print("   %-9s chunks/CTA %5.2f  duration us: mean %6.2f  p50 %6.2f  p90 %6.2f  max %6.2f   lane-0 ticks mean %5.1f   load+setup %5.2f  run %5.2f  store+plan %5.2f  obs+out %5.2f" % (
            NAMES[c], float(m.sum()) / grid, float(d.mean()), float(d.median()), float(d.quantile(0.9)), float(d.max()), float(nt[m].float().mean()),
            float(seg[0].median()), float(seg[1].median()), float(seg[2].median()), float(seg[3].median())))
    tot = (dur * valid).sum(1)
    print("   sum of chunk durations per CTA: mean %.1f us (/12 warps = %.1f), longest chunk per CTA: mean %.1f us" % (
        float(tot.mean()), float(tot.mean()) / 12, float((dur * valid).max(1).values.mean())))
    # slowest CTA
    k = int((ph[:, 5]).argmax())
    rel0 = (ch[k, :, 0] - int(ph[k, 3])).float() / 1e3
    rows = [(float(rel0[q]), float(dur[k, q]), NAMES[int(cls[k, q])] if int(cls[k, q]) < 5 else "?", int(nt[k, q])) for q in range(128) if bool(valid[k, q])]
    print("   slowest CTA %d chunks (start after sort, duration, class, lane-0 ticks): %s" % (k, " ".join("%.1f+%.1f:%s:%d" % r for r in rows)))
env._L.tg_debug_phase_buffer(env._h, None)
