#!/usr/bin/env python
"""Benchmark of the Treasure Game hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU algorithm

Workload (BASELINE.json configs[2], the configuration the headline metric is quoted on):
treasure_game-v0, vector-state observations, 1,048,576 environments in total, sharded contiguously
over the N GPUs (N = 1: all of them on one GPU; "scaling": "strong"), uniform-random option ids,
100-step time limit with auto-reset, in the steady state of that process: episode phases are
desynchronised before anything is timed, so about 1 % of the envs hit the time limit in every
step.  One "step" = one TreasureGame.step for every env of the batch (runnable or not).  N > 1: one
process per GPU (torchrun), no data-path collective; the episode-statistics vector is all-reduced
with NCCL on a side stream every 100 steps and once at the end.  The weak-scaling reading
(1,048,576 envs per GPU) is timed next to it and reported under aux.

Timing: CUDA events on the launching stream around every step kernel; between timed steps a
512 MiB buffer is overwritten to flush the 126 MB L2 (outside the event pairs); the K steps are
bracketed by barrier + synchronize; the per-rank sum of event times is MAX-reduced over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_ENV_STEP = 125          # SURVEY.md 8(d); this layout moves 126 (52 read + 74 written)
ALGO_BYTES_PER_FRAME = 1258024         # 624*672*3 written + 40 state bytes read
TOTAL_ENVS = 1 << 20
MAX_EPISODE_STEPS = 100
METRIC, UNIT = "env_steps_per_s", "env-steps/s"
WORKLOAD = ("treasure_game-v0 vector-state obs, 1,048,576 envs in total sharded over the GPUs (BASELINE configs[2]), "
            "uniform-random option ids, 100-step time limit + auto-reset, steady state (desynchronised episodes)")


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:       # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------- steady state
def desynchronise(env, torch, new_actions, max_episode_steps=MAX_EPISODE_STEPS, seed=99):
    """Brings a freshly created batch to the steady state of the workload, outside any timed region: one episode
    length of steps, then every env's episode age is redrawn uniformly in [0, max_episode_steps) (tg_set_state),
    then one more episode length so that every env has been reset at its own time.  From then on about 1 % of
    the envs hit the time limit in every step (instead of all of them together every 100th step)."""
    n = env.num_envs
    for _ in range(max_episode_steps):
        env.step_raw(new_actions())
    st = env.get_state()
    g = torch.Generator(device=env.device).manual_seed(seed)
    acct = st["acct"]
    acct[:, 1] = torch.randint(0, max_episode_steps, (n,), generator=g, device=env.device, dtype=torch.int64)
    env.set_state({"acct": acct})
    for _ in range(max_episode_steps + 10):
        env.step_raw(new_actions())


# ----------------------------------------------------------------------------- CPU baselines
def _py_worker(args):
    seed, seconds = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import py_oracle as po
    t0 = time.perf_counter()
    steps = ticks = 0
    ep = 0
    while time.perf_counter() - t0 < seconds:
        s, t = po.readme_loop(seed * 100003 + ep, episodes=5)        # README loop: 5 episodes x 100 steps
        steps += s
        ticks += t
        ep += 1
    return steps, ticks, time.perf_counter() - t0


def cpu_baseline_python(seconds, procs=None, pool=None):
    """The reference's algorithm (pure-Python port, oracle/py_oracle.py, pinned bit-exact against the
    reference) in P independent processes -- BASELINE.md section 3."""
    import multiprocessing as mp
    procs = procs or os.cpu_count() or 1
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(procs)
    try:
        res = pool.map(_py_worker, [(i + 1, seconds) for i in range(procs)])
    finally:
        if own:
            pool.close()
            pool.join()
    steps = sum(r[0] for r in res)
    ticks = sum(r[1] for r in res)
    wall = max(r[2] for r in res)
    return {"value": steps / wall, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": "README loop (5 episodes x 100 uniform-random steps, reset per episode) repeated for "
                      "%.2f s in %d processes; pure-Python port of the reference (oracle/py_oracle.py)" % (seconds, procs),
            "primitive_ticks_per_s": ticks / wall, "per_core": steps / wall / procs}


def cpu_baseline_c(seconds, threads=None, envs_per_thread=4096):
    """Secondary, much stronger baseline: the plain-C port (oracle/tg_oracle.c), one batch per thread."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import c_oracle
    import py_oracle as po
    threads = threads or os.cpu_count() or 1
    lv = c_oracle.CLevel(po.default_level())
    batches = [c_oracle.CBatch(lv, envs_per_thread, first_env_id=i * envs_per_thread, seed=0,
                               max_episode_steps=MAX_EPISODE_STEPS, auto_reset=True) for i in range(threads)]
    for b in batches:
        b.reset()
    counts = [0] * threads

    def work(i):
        rng = np.random.default_rng(i)
        rew = np.zeros(envs_per_thread, np.float32)
        done = np.zeros(envs_per_thread, np.uint8)
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            a = rng.integers(0, 9, envs_per_thread, dtype=np.int32)
            batches[i].step_fast(a, rew, done)          # ctypes releases the GIL
            counts[i] += envs_per_thread

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    wall = time.perf_counter() - t0
    return {"value": sum(counts) / wall, "unit": UNIT, "cores": threads, "kind": "port-c",
            "sample": "%d envs per thread, random actions, auto-reset, %.0f s" % (envs_per_thread, seconds)}


# ----------------------------------------------------------------------------- reference arm
def run_reference(args, rank):
    """CPU arm: every "step" is a bounded sample (README loops for `per_step` seconds on all host
    cores); the whole run is sized to about two minutes whatever K and W are."""
    if rank != 0:
        return
    import multiprocessing as mp
    procs = os.cpu_count() or 1
    total = args.warmup + args.steps
    per_step = max(0.25, min(2.0, 120.0 / total))
    pool = mp.get_context("spawn").Pool(procs)
    t_all = time.perf_counter()
    vals = []
    try:
        for k in range(total):
            r = cpu_baseline_python(per_step, procs, pool)
            if k >= args.warmup:
                vals.append(r)
    finally:
        pool.close()
        pool.join()
    wall = time.perf_counter() - t_all
    value = sum(v["value"] for v in vals) / len(vals)
    res = dict(vals[-1])
    res["value"] = value
    res["primitive_ticks_per_s"] = sum(v["primitive_ticks_per_s"] for v in vals) / len(vals)
    res["per_core"] = value / procs
    res["sample"] = "each step = " + res["sample"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU arm: reference algorithm (pure-Python port; the Python "
                       "reference itself cannot travel to the GPU box), all host cores, README loop"},
            "cpu_baseline": res,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": wall}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- main arm
def timed_steps(torch, dist, env, dev, world, n, K, W, seed, stats_every=100):
    """W untimed + K timed steps of one batch in steady state.  Returns per-step event times (ms), the max-over-ranks
    total (s), launches, statistics (all ranks), clocks, wall time and draws per step."""
    g = torch.Generator(device=dev).manual_seed(seed)
    actions = torch.empty((n,), dtype=torch.int32, device=dev)

    def new_actions():
        # i.i.d. uniform option ids, regenerated on the device for every step (BASELINE configs[1]/[2]);
        # a short cycled pool would make every env's action sequence periodic and shrink the work
        return torch.randint(0, 9, (n,), generator=g, dtype=torch.int32, device=dev, out=actions)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream(device=dev)
    stats_buf = torch.zeros(8, dtype=torch.int64, device=dev)

    def stats_allreduce():
        # the path's only collective: 64-byte int64[8] sum, off the critical path
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            stats_buf.copy_(env.stats_tensor())
            if world > 1:
                dist.all_reduce(stats_buf)

    desynchronise(env, torch, new_actions)            # outside the timed region: steady state of the workload
    for k in range(W):
        env.step_raw(new_actions())
    env.clear_stats()
    draws0 = int(env.get_state()["misc"][:, 3].to(torch.int64).sum().item())     # work accounting, outside the timed region
    barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    launches0 = env.launch_count
    t_wall0 = time.perf_counter()
    for k in range(K):
        new_actions()                                   # not timed: inputs are resident before the event pair
        flush.zero_()                                   # L2 flush, outside the event pair
        starts[k].record()
        env.step_raw(actions)
        ends[k].record()
        if (k + 1) % stats_every == 0:
            stats_allreduce()
    stats_allreduce()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = env.launch_count - launches0
    draws1 = int(env.get_state()["misc"][:, 3].to(torch.int64).sum().item())    # after the launch count was taken
    clocks = sampler.finish()
    per_step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = torch.tensor([sum(per_step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    stats = dict(zip(["episodes", "successes", "return_sum", "episode_steps_sum", "primitive_ticks",
                      "runnable_steps", "gym_steps", "errors"], (int(v) for v in stats_buf.cpu())))
    del flush
    return dict(per_step_ms=per_step_ms, total_s=float(total_ms.item()) / 1000.0, launches=launches, stats=stats,
                clocks=clocks, wall=t_wall, draws_per_step=(draws1 - draws0) / max(n * K, 1), new_actions=new_actions,
                barrier=barrier)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--total-envs", type=int, default=TOTAL_ENVS)
    ap.add_argument("--no-aux", action="store_true", help="skip the auxiliary configs (4096 envs, shards, RGB render, weak scaling, CPU baselines)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from gym_treasure_game_b200 import VectorTreasureGame, shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    total = args.total_envs
    lo, hi = shard_range(total, rank, world)            # strong scaling: BASELINE configs[2] shards 1,048,576 envs over the GPUs
    n = hi - lo
    K, W = args.steps, args.warmup

    env = VectorTreasureGame(n, device=dev, seed=0, max_episode_steps=MAX_EPISODE_STEPS, auto_reset=True,
                             first_env_id=lo, render=False)
    r = timed_steps(torch, dist, env, dev, world, n, K, W, seed=1234 + rank)
    per_step_ms, total_s, stats, barrier, new_actions = r["per_step_ms"], r["total_s"], r["stats"], r["barrier"], r["new_actions"]
    value = total * K / total_s
    if stats["episodes"] <= 0:
        raise SystemExit("bench.py: no episode ended inside the timed region -- the workload is not in its steady state")

    # ---- end-to-end through the C ABI with HOST buffers: H2D actions, step, D2H results, host arrays complete on return.
    # Headline: tg_step_host_sparse (only the rows of envs that ran or were reset cross the bus; the host arrays of the
    # previous call are patched in place).  tg_step_host (everything crosses) is timed next to it under its own key.
    Ke = min(K, 100)                                 # host-side time varies with the box's other tenants: more steps than a kernel timing needs
    host = env.make_host_buffers()
    hpool = [new_actions().cpu().pin_memory() for _ in range(min(Ke, 20))]      # cycled

    def time_host(fn):
        for k in range(10):                           # untimed: first-touch of the pinned / record buffers, host threads up
            host["actions"] = hpool[k % len(hpool)]
            fn(host)
        barrier()
        h0, d0 = env.host_traffic()
        ts = []
        t0 = time.perf_counter()
        for k in range(Ke):
            host["actions"] = hpool[k % len(hpool)]
            t1 = time.perf_counter()
            fn(host)
            ts.append(time.perf_counter() - t1)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h1, d1 = env.host_traffic()
        ts.sort()
        return total * Ke / float(dt.item()), (h1 - h0) // Ke, (d1 - d0) // Ke, 1e3 * ts[len(ts) // 2], 1e3 * ts[-1]

    dense = time_host(env.step_host)
    single = time_host(env.step_host_sparse)

    # Headline: the same entry point in its two halves (tg_step_host_sparse_begin / _end) over sub-batches kept in
    # flight (PipelinedHostEnv; three parts at this size, two for small shards): while the host patches one part's arrays the
    # others' kernels run and their records cross.  The loop stays closed per part: the actions of a part's step are handed
    # over after its previous step ended.
    from gym_treasure_game_b200 import PipelinedHostEnv
    penv = PipelinedHostEnv(n, first_env_id=lo, device=dev, seed=0, max_episode_steps=MAX_EPISODE_STEPS,
                            auto_reset=True, render=False)
    for k, sub in enumerate(penv.envs):
        gk = torch.Generator(device=dev).manual_seed(4321 + 2 * rank + k)
        ak = torch.empty((sub.num_envs,), dtype=torch.int32, device=dev)
        desynchronise(sub, torch, lambda: torch.randint(0, 9, (sub.num_envs,), generator=gk, dtype=torch.int32, device=dev, out=ak))
    torch.cuda.synchronize()

    round_ms = []

    def pipelined_steps(count):
        for k, (plo, phi) in enumerate(penv.ranges):
            penv.hosts[k]["actions"] = hpool[0][plo:phi]
            penv.begin(k)
        for it in range(1, count):
            t1 = time.perf_counter()
            for k, (plo, phi) in enumerate(penv.ranges):
                penv.end(k)                                     # host arrays of part k complete: obs / reward / done / ran
                penv.hosts[k]["actions"] = hpool[it % len(hpool)][plo:phi]
                penv.begin(k)
            round_ms.append(1e3 * (time.perf_counter() - t1))
        for k in range(penv.parts):
            penv.end(k)

    pipelined_steps(10)                                         # untimed: dense first step, first touch, host threads up
    barrier()
    ph0, pd0 = penv.host_traffic()
    del round_ms[:]
    t0 = time.perf_counter()
    pipelined_steps(Ke)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    ph1, pd1 = penv.host_traffic()
    pst = penv.stats()
    sparse = (total * Ke / float(dt.item()), (ph1 - ph0) // Ke, (pd1 - pd0) // Ke, 1e3 * float(dt.item()) / Ke)
    if pst["errors"]:
        raise SystemExit("bench.py: pipelined e2e run flagged %d envs" % pst["errors"])
    parts_in_flight = penv.parts
    penv.close()
    del penv

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1000.0 * total_s / K, "ms_per_step_median": sorted(per_step_ms)[K // 2], "ms_per_step_max": max(per_step_ms),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32+f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "total_envs": total, "envs_per_gpu": n, "max_episode_steps": MAX_EPISODE_STEPS,
                   "steady_state": "100 steps, episode ages redrawn uniformly in [0, 100), 110 more steps; then the W warm-up steps",
                   "l2": "512 MiB buffer overwritten between timed steps (outside the event pairs)",
                   "actions": "torch.randint on the device before every step, outside the timed event pair",
                   "collective": "NCCL all-reduce of int64[8] stats every 100 steps, side stream"},
        "clocks": r["clocks"],
        "e2e": {"value": sparse[0], "unit": UNIT, "h2d_bytes_per_step": sparse[1], "d2h_bytes_per_step": sparse[2], "steps": Ke,
                "ms_per_step": sparse[3], "ms_per_step_median_rank0": sorted(round_ms)[len(round_ms) // 2] if round_ms else None,
                "ms_per_step_max_rank0": max(round_ms) if round_ms else None,
                "parts_in_flight": parts_in_flight,
                "api": "PipelinedHostEnv = tg_step_host_sparse_begin / _end over sub-batches in flight (pinned host buffers; per "
                       "part: H2D actions, step kernels that compact the envs whose outputs changed into records, one D2H copy per "
                       "chunk of header + tile table + records, host threads patch obs/reward/done/ran in place while the other parts "
                       "run; a part's next actions are handed over after its previous step ended; bytes counted by the library per copy)",
                "single_call": {"name": "tg_step_host_sparse", "value": single[0], "h2d_bytes_per_step": single[1],
                                "d2h_bytes_per_step": single[2], "ms_per_step_median_rank0": single[3], "ms_per_step_max_rank0": single[4],
                                "api": "one batch, one blocking call per step (begin + end back to back)"},
                "dense_path": {"name": "tg_step_host", "value": dense[0], "h2d_bytes_per_step": dense[1], "d2h_bytes_per_step": dense[2],
                               "ms_per_step_median_rank0": dense[3],
                               "api": "tg_step_host (pinned host buffers; H2D actions, chunked step kernels overlapping the D2H of every "
                                      "env's obs/reward/done/ran; stream sync inside the call)"}},
        "gpu_launches": r["launches"],
        "roofline": {"bound": "hbm", "achieved": (n * ALGO_BYTES_PER_ENV_STEP) / (total_s / K) / 1e9,
                     "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                     "kernel": "tg_step_kernel<false,2> (%d envs per launch, automatic tile size)" % n,
                     "algorithmic_bytes_per_launch": n * ALGO_BYTES_PER_ENV_STEP,
                     "note": "per GPU; serial-chain / instruction-fetch bound, not HBM bound (DESIGN.md 3.1); see work.primitive_ticks_per_s"},
        "work": {"runnable_fraction": stats["runnable_steps"] / max(stats["gym_steps"], 1),
                 "primitive_ticks_per_step": stats["primitive_ticks"] / max(stats["gym_steps"], 1),
                 "primitive_ticks_per_s": stats["primitive_ticks"] / total_s if stats["gym_steps"] else None,
                 "episodes_per_step_fraction": stats["episodes"] / max(stats["gym_steps"], 1),
                 "rng_draws_per_step_rank0": r["draws_per_step"],
                 "stats_all_ranks": stats},
        "wall_s_timed_region_incl_flush": r["wall"],
    }
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = json.load(open(peaks_path))["hbm_gbs"]
        line["roofline"]["peak_source"] = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak = 6650.0
        line["roofline"]["peak_source"] = "B200_PROFILING.md fallback (of fallback)"
    line["roofline"]["peak"] = peak
    line["roofline"]["frac"] = line["roofline"]["achieved"] / peak
    traffic_path = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    if os.path.exists(traffic_path) and n == TOTAL_ENVS:
        line["roofline"]["traffic"] = json.load(open(traffic_path)).get("dram_bytes_per_launch")

    env.close()
    del env, host, hpool
    torch.cuda.empty_cache()

    if not args.no_aux:
        aux = {}
        if world > 1:
            # the weak-scaling reading of configs[2]: 1,048,576 envs per GPU, same steady state, 20 timed steps
            wn = TOTAL_ENVS
            wenv = VectorTreasureGame(wn, device=dev, seed=0, max_episode_steps=MAX_EPISODE_STEPS, auto_reset=True,
                                      first_env_id=rank * wn, render=False)
            wr = timed_steps(torch, dist, wenv, dev, world, wn, 20, 5, seed=99 + rank)
            aux["weak_1048576_envs_per_gpu"] = {"env_steps_per_s": world * wn * 20 / wr["total_s"], "ms_per_step": 1000.0 * wr["total_s"] / 20,
                                                "scaling": "weak", "total_envs": world * wn}
            wenv.close()
            del wenv
            torch.cuda.empty_cache()
        if rank == 0 and world == 1:
            aux.update(aux_configs(torch, dev, peak))
            line["cpu_baseline"] = cpu_baseline_python(args.cpu_seconds)
            line["cpu_baseline_c"] = cpu_baseline_c(min(args.cpu_seconds, 6.0))
        if aux:
            line["aux"] = aux
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def aux_configs(torch, dev, peak):
    """One GPU: BASELINE configs[1] (4096 envs), the per-GPU shards of configs[2] at 2 / 4 / 8 GPUs, and configs[3]
    (16384 envs, RGB render).  Every batch is brought to the same steady state as the main line first."""
    from gym_treasure_game_b200 import VectorTreasureGame
    out = {}
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(7)

    def timed(fn, iters, warm=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            e.synchronize()
            tot += s.elapsed_time(e)
        return tot / iters / 1000.0

    def steady(n, render=False):
        env = VectorTreasureGame(n, device=dev, seed=0, max_episode_steps=MAX_EPISODE_STEPS, auto_reset=True, render=render)
        acts = torch.empty((n,), dtype=torch.int32, device=dev)
        new_actions = lambda: torch.randint(0, 9, (n,), generator=g, dtype=torch.int32, device=dev, out=acts)   # fresh i.i.d. ids every step
        desynchronise(env, torch, new_actions)
        return env, acts, new_actions

    for name, n, iters in (("cfg2_4096_envs", 4096, 300), ("cfg3_shard_131072_envs", 131072, 200),
                           ("cfg3_shard_262144_envs", 262144, 100), ("cfg3_shard_524288_envs", 524288, 100)):
        env, acts, new_actions = steady(n)

        def step():
            env.step_raw(acts)
        # the action refresh stays outside the event pair, as in the main line
        for _ in range(5):
            new_actions(); step()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            new_actions(); flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); step(); e.record()
            e.synchronize()
            tot += s.elapsed_time(e)
        t = tot / iters / 1000.0
        out[name] = {"env_steps_per_s": n / t, "ms_per_step": t * 1e3, "roofline_frac": n * ALGO_BYTES_PER_ENV_STEP / t / 1e9 / peak}
        if n > 4096:
            out[name]["note"] = "one GPU's shard when 1,048,576 envs are split over %d GPUs" % (TOTAL_ENVS // n)
        env.close()

    # configs[3]: 16384 envs, RGB observations
    n = 16384
    env, acts, new_actions = steady(n, render=True)
    frames = torch.empty((n, 624, 672, 3), dtype=torch.uint8, device=dev)
    t_r = timed(lambda: env.render(out=frames), 10, warm=3)

    def step_render():
        new_actions(); env.step_raw(acts)
        env.render(out=frames)
    t_sr = timed(step_render, 10, warm=3)
    out["cfg4_render_16384_envs"] = {
        "frames_per_s_render_only": n / t_r, "frames_per_s_step_plus_render": n / t_sr, "ms_render": t_r * 1e3,
        "roofline": {"bound": "hbm", "achieved": n * ALGO_BYTES_PER_FRAME / t_r / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": n * ALGO_BYTES_PER_FRAME / t_r / 1e9 / peak, "kernel": "tg_render_stream_kernel",
                     "note": "20.6 GB written per launch (> L2)"}}
    env.close()
    return out


if __name__ == "__main__":
    main()
