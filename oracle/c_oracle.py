"""TEST INFRASTRUCTURE ONLY -- ctypes binding of ``oracle/libtg_oracle.so``
(the plain-C restatement, ``oracle/tg_oracle.c``).  Not importable from the
product package; see the header of ``tg_oracle.c``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libtg_oracle.so")

K_DOOR, K_HANDLE, K_KEY, K_BOLT, K_GOLD = range(5)


def build(force=False):
    src = os.path.join(HERE, "tg_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B", "libtg_oracle.so"])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        L.tgo_level_new.restype = C.c_void_p
        L.tgo_level_new.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.tgo_level_free.argtypes = [C.c_void_p]
        L.tgo_obs_dim.argtypes = [C.c_void_p]
        L.tgo_count.argtypes = [C.c_void_p, C.c_int]
        L.tgo_batch_new.restype = C.c_void_p
        L.tgo_batch_new.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_int, C.c_int]
        L.tgo_batch_free.argtypes = [C.c_void_p]
        L.tgo_batch_set_tape.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.tgo_batch_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.tgo_batch_step.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.tgo_batch_mask.argtypes = [C.c_void_p, C.c_void_p]
        L.tgo_batch_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.tgo_batch_get.argtypes = [C.c_void_p] + [C.c_void_p] * 9
        L.tgo_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.tgo_batch_init_with_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.tgo_batch_get_pt.argtypes = [C.c_void_p, C.c_void_p]
        L.tgo_batch_set_draws.argtypes = [C.c_void_p, C.c_void_p]
        L.tgo_batch_prim_step.argtypes = [C.c_void_p] * 5
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().tgo_philox(_p(c), _p(k), _p(out))
    return out


class CLevel:
    """``level`` is an ``oracle.py_oracle.LevelText`` (tiles/objects/triggers)."""

    def __init__(self, level):
        L = lib()
        self.text = level
        self.ch, self.cw = len(level.tiles), len(level.tiles[0])
        tiles = "".join(r.ljust(self.cw)[: self.cw] for r in level.tiles).encode()
        objs = np.array([[k, cx, cy, int(f)] for k, cx, cy, f in level.objects], dtype=np.int32).reshape(-1, 4)
        trg = np.array([[a, b, int(c), d, e, int(f)] for a, b, c, d, e, f in level.triggers],
                       dtype=np.int32).reshape(-1, 6)
        self.h = L.tgo_level_new(tiles, self.cw, self.ch, _p(objs), len(objs), _p(trg), len(trg))
        if not self.h:
            raise ValueError("level too large for the C oracle")
        self.obs_dim = L.tgo_obs_dim(self.h)
        self.nd, self.nh, self.nb = (L.tgo_count(self.h, k) for k in (K_DOOR, K_HANDLE, K_BOLT))
        self.ni = L.tgo_count(self.h, K_KEY) + L.tgo_count(self.h, K_GOLD)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.tgo_level_free(self.h)
            self.h = None


class CBatch:
    def __init__(self, level: CLevel, n, first_env_id=0, seed=0, max_episode_steps=0, auto_reset=False):
        self.level, self.n = level, int(n)
        self.h = lib().tgo_batch_new(level.h, self.n, int(first_env_id), int(seed),
                                     int(max_episode_steps), int(bool(auto_reset)))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.tgo_batch_free(self.h)
            self.h = None

    def set_tape(self, tapes):
        """tapes: list (len n) of 1-D float64 sequences, or None for Philox mode."""
        if tapes is None:
            lib().tgo_batch_set_tape(self.h, None, None)
            return
        off = np.zeros(self.n + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(t) for t in tapes])
        flat = np.concatenate([np.asarray(t, dtype=np.float64) for t in tapes]) if off[-1] else np.zeros(1)
        lib().tgo_batch_set_tape(self.h, _p(flat), _p(off))

    def reset(self, mask=None):
        obs = np.zeros((self.n, self.level.obs_dim), dtype=np.float64)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().tgo_batch_reset(self.h, _p(m), _p(obs))
        return obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.zeros((self.n, self.level.obs_dim), dtype=np.float64)
        rew = np.zeros(self.n, dtype=np.float32)
        done = np.zeros(self.n, dtype=np.uint8)
        ran = np.zeros(self.n, dtype=np.uint8)
        ticks = np.zeros(self.n, dtype=np.int32)
        lib().tgo_batch_step(self.h, _p(a), _p(obs), _p(rew), _p(done), _p(ran), _p(ticks))
        return obs, rew, done, ran, ticks

    def step_fast(self, actions, rew, done):
        """No per-call allocations (for timing)."""
        lib().tgo_batch_step(self.h, _p(actions), None, _p(rew), _p(done), None, None)

    def prim_step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.zeros((self.n, self.level.obs_dim), dtype=np.float64)
        rew = np.zeros(self.n, dtype=np.float32)
        done = np.zeros(self.n, dtype=np.uint8)
        lib().tgo_batch_prim_step(self.h, _p(a), _p(obs), _p(rew), _p(done))
        return obs, rew, done

    def init_with_state(self, states, mask=None):
        st = np.ascontiguousarray(states, dtype=np.float64).reshape(self.n, self.level.obs_dim)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().tgo_batch_init_with_state(self.h, _p(st), _p(m))

    def set_draws(self, draws):
        """Test hook: every env's 32-bit draw index (the counter its Philox blocks are numbered from)."""
        d = np.ascontiguousarray(draws, dtype=np.uint32).reshape(self.n)
        lib().tgo_batch_set_draws(self.h, _p(d))

    def handles_pt(self):
        pt = np.zeros((self.n, max(self.level.nh, 1)), dtype=np.uint8)
        lib().tgo_batch_get_pt(self.h, _p(pt))
        return pt[:, : self.level.nh]

    def mask(self):
        m = np.zeros((self.n, 9), dtype=np.uint8)
        lib().tgo_batch_mask(self.h, _p(m))
        return m

    def stats(self):
        s = np.zeros(8, dtype=np.int64)
        lib().tgo_batch_stats(self.h, _p(s))
        return s

    def state(self):
        lv, n = self.level, self.n
        out = dict(
            pos=np.zeros((n, 2), np.int32), misc=np.zeros((n, 4), np.int32),
            doors=np.zeros((n, max(lv.nd, 1)), np.uint8), handles=np.zeros((n, max(lv.nh, 1)), np.uint8),
            bolts=np.zeros((n, max(lv.nb, 1)), np.uint8), angles=np.zeros((n, max(lv.nh, 1)), np.float64),
            items=np.zeros((n, max(lv.ni, 1), 4), np.int32), bag=np.full((n, 4), -1, np.int32),
            acct=np.zeros((n, 3), np.int64))
        lib().tgo_batch_get(self.h, _p(out["pos"]), _p(out["misc"]), _p(out["doors"]), _p(out["handles"]),
                            _p(out["bolts"]), _p(out["angles"]), _p(out["items"]), _p(out["bag"]), _p(out["acct"]))
        out["doors"] = out["doors"][:, : lv.nd]
        out["handles"] = out["handles"][:, : lv.nh]
        out["bolts"] = out["bolts"][:, : lv.nb]
        out["angles"] = out["angles"][:, : lv.nh]
        out["items"] = out["items"][:, : lv.ni]
        return out

    def snapshot(self, i=0):
        """Same dict layout as ``py_oracle.OracleEnv.snapshot`` for env ``i``."""
        s = self.state()
        return dict(
            px=int(s["pos"][i, 0]), py=int(s["pos"][i, 1]), facing=int(s["misc"][i, 0]),
            ticker=int(s["misc"][i, 1]), doors=[int(v) for v in s["doors"][i]],
            handles_up=[int(v) for v in s["handles"][i]], angles=[float(v) for v in s["angles"][i]],
            bolts=[int(v) for v in s["bolts"][i]], items=[tuple(int(v) for v in it) for it in s["items"][i]],
            bag=[int(v) for v in s["bag"][i] if v >= 0], total_actions=int(s["misc"][i, 2]),
            handles_pt=[int(v) for v in self.handles_pt()[i]])


class ShardedCBatch:
    """``n`` envs split into contiguous sub-batches stepped by one Python thread each (ctypes releases the GIL;
    every env owns its Philox stream keyed by the global env id, so the split is invisible).  Same interface
    as ``CBatch`` for what the full-size parity tests use: 1,048,576 envs x 130 steps take seconds."""

    def __init__(self, level: CLevel, n, first_env_id=0, seed=0, max_episode_steps=0, auto_reset=False, threads=None):
        import os
        from concurrent.futures import ThreadPoolExecutor
        self.level, self.n = level, int(n)
        threads = max(1, min(threads or os.cpu_count() or 1, (self.n + 4095) // 4096))
        cuts = [self.n * k // threads for k in range(threads + 1)]
        self.ranges = [(cuts[k], cuts[k + 1]) for k in range(threads) if cuts[k + 1] > cuts[k]]
        self.parts = [CBatch(level, hi - lo, first_env_id + lo, seed, max_episode_steps, auto_reset) for lo, hi in self.ranges]
        self.pool = ThreadPoolExecutor(len(self.parts))

    def _map(self, fn):
        return list(self.pool.map(fn, range(len(self.parts))))

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        return np.concatenate(self._map(lambda k: self.parts[k].reset(None if m is None else m[self.ranges[k][0]:self.ranges[k][1]])))

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        outs = self._map(lambda k: self.parts[k].step(a[self.ranges[k][0]:self.ranges[k][1]]))
        return tuple(np.concatenate([o[j] for o in outs]) for j in range(5))

    def mask(self):
        return np.concatenate(self._map(lambda k: self.parts[k].mask()))

    def stats(self):
        return np.sum(self._map(lambda k: self.parts[k].stats()), axis=0)

    def state(self):
        sts = self._map(lambda k: self.parts[k].state())
        return {key: np.concatenate([s[key] for s in sts]) for key in sts[0]}

    def handles_pt(self):
        return np.concatenate(self._map(lambda k: self.parts[k].handles_pt()))

    def snapshot(self, i=0):
        for (lo, hi), p in zip(self.ranges, self.parts):
            if lo <= i < hi:
                return p.snapshot(i - lo)
        raise IndexError(i)
