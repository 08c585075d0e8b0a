#!/usr/bin/env python
"""Offline fuzz campaign for the kernels' device functions WITHOUT a GPU: random levels (tests/level_fuzz.py) through the
host build of csrc/tg_device.cuh (tests/hostdev/hostdev.cpp) against the C oracle, bit-exact after every step -- the test
tests/test_device_code_on_host.py::test_device_functions_match_c_oracle on many more levels, with drops and jumps weighted
up.  usage: fuzz_device_on_host.py FIRST_SEED LAST_SEED [envs] [steps] [closure]     (one process per seed range; ~0.4 s per
level; `closure`: INTERACT through the level's closure table, levels with triggers only, ~0.6 s per level)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("tests", "oracle", ""):
    sys.path.insert(0, os.path.join(ROOT, d))
import ctypes as C

import numpy as np

import test_device_code_on_host as T
from level_fuzz import random_level, usable


def main():
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 160
    closure = len(sys.argv) > 5 and sys.argv[5] == "closure"
    L = C.CDLL(T._build())
    L.hostdev_blob_size.restype = C.c_size_t
    L.hostdev_create.restype = C.c_void_p
    L.hostdev_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int]
    L.hostdev_destroy.argtypes = [C.c_void_p]
    L.hostdev_reset.argtypes = [C.c_void_p, C.c_void_p]
    L.hostdev_step.argtypes = [C.c_void_p] * 7
    L.hostdev_mask.argtypes = [C.c_void_p, C.c_void_p]
    L.hostdev_flags.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.hostdev_state.argtypes = [C.c_void_p] * 10
    T.RARE_WEIGHT = np.array([1, 1, 2, 2, 1, 12, 12, 12, 12], dtype=np.float64)
    ok = bad = 0
    t0 = time.time()
    for s in range(lo, hi):
        lv = random_level(s)
        if not usable(lv) or (closure and not lv.triggers):
            continue
        try:
            T._run(L, lv, n=n, steps=steps, seed=s * 7 + 1, max_steps=60 if s % 2 else 0, with_closure=closure, first_env_id=s)
            ok += 1
        except AssertionError as e:
            bad += 1
            print("MISMATCH level seed", s, str(e)[:300])
            sys.stdout.flush()
    print("levels ok %d, mismatching %d, in %.0f s" % (ok, bad, time.time() - t0))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
