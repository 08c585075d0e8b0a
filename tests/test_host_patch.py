"""CPU: the host half of tg_step_host_sparse (csrc/tg_host_patch.h) against a numpy restatement of its contract:
reported envs take their record; envs that are not reported but still show the previous step's outputs (ran or done
set) go back to reward 0 / done 0 / ran 0; every other byte of the caller's arrays stays as it was."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gym_treasure_game_b200", "csrc")
SHIM = os.path.join(ROOT, "tests", "host_patch", "patch_shim.cpp")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("patch") / "libpatch_shim.so")
    subprocess.check_call(["g++", "-O2", "-fopenmp", "-shared", "-fPIC", "-I", CSRC, "-I", os.path.join(ROOT, "include"), SHIM, "-o", so])
    L = C.CDLL(so)
    L.patch_apply.restype = C.c_int
    L.patch_apply.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4
    return L


def make_step(rng, n, lo, cnt, tile, od, words, density):
    """Records of one launch over envs [lo, lo + cnt): per tile a contiguous block in arbitrary order; blocks in arbitrary order."""
    grid = (cnt + tile - 1) // tile
    blocks = []
    for t in range(grid):
        m = min(tile, cnt - t * tile)
        els = np.nonzero(rng.random(m) < density)[0]
        rng.shuffle(els)
        blocks.append((t, lo + t * tile + els))
    order = rng.permutation(grid)
    table = np.zeros((grid, 2), np.uint32)
    recs = []
    pos = 0
    for t in order:
        idx = blocks[t][1]
        table[t] = (pos, len(idx))
        r = np.zeros((len(idx), words), np.uint32)
        r[:, 0] = idx
        r[:, 1] = rng.integers(-4000, 0, len(idx)).astype(np.float32).view(np.uint32)
        r[:, 2] = rng.integers(0, 4, len(idx)).astype(np.uint32) | (rng.integers(0, 2, len(idx)).astype(np.uint32) << 8)
        r[:, 3:3 + od] = rng.random((len(idx), od)).astype(np.float32).view(np.uint32)
        recs.append(r)
        pos += len(idx)
    recs = np.concatenate(recs) if recs else np.zeros((0, words), np.uint32)
    return table, np.ascontiguousarray(recs), grid


def expected(recs, lo, cnt, od, obs, reward, done, ran):
    obs, reward, done, ran = obs.copy(), reward.copy(), done.copy(), ran.copy()
    sl = slice(lo, lo + cnt)
    stale = (ran[sl] | done[sl]) != 0
    reward[sl][stale] = 0.0; done[sl][stale] = 0; ran[sl][stale] = 0
    idx = recs[:, 0].astype(np.int64)
    reward[idx] = recs[:, 1].view(np.float32)
    done[idx] = (recs[:, 2] & 255).astype(np.uint8)
    ran[idx] = (recs[:, 2] >> 8).astype(np.uint8)
    obs[idx] = recs[:, 3:3 + od].view(np.float32)
    return obs, reward, done, ran


@pytest.mark.parametrize("n,lo,cnt,tile,od,threads", [
    (5000, 0, 5000, 256, 9, "1"), (100001, 0, 100001, 3544, 9, "4"), (70000, 2048, 40003, 1000, 9, "3"),
    (30000, 0, 30000, 32, 12, "2"), (9000, 0, 9000, 4096, 17, "4"), (64, 0, 64, 64, 9, "2")])
def test_patch_matches_contract(lib, n, lo, cnt, tile, od, threads):
    os.environ["TG_HOST_THREADS"] = threads          # read once per process: the first case decides; all cases are thread-count agnostic
    rng = np.random.default_rng(n + tile)
    words = (3 + od + 3) // 4 * 4
    obs = rng.random((n, od)).astype(np.float32)
    reward = np.zeros(n, np.float32); done = np.zeros(n, np.uint8); ran = np.zeros(n, np.uint8)
    for step in range(4):                            # consecutive steps: what one step reports is stale in the next
        density = (0.0, 0.2, 1.0, 0.05)[step]
        table, recs, grid = make_step(rng, n, lo, cnt, tile, od, words, density)
        want = expected(recs, lo, cnt, od, obs, reward, done, ran)
        rc = lib.patch_apply(table.ctypes.data, grid, tile, lo, cnt, recs.ctypes.data if len(recs) else None, words, od,
                             obs.ctypes.data, reward.ctypes.data, done.ctypes.data, ran.ctypes.data)
        assert rc == 0
        for got, exp, name in zip((obs, reward, done, ran), want, ("obs", "reward", "done", "ran")):
            assert np.array_equal(got, exp), (step, name)


def test_patch_rejects_foreign_records(lib):
    n, tile, od, words = 1000, 250, 9, 12
    rng = np.random.default_rng(0)
    table, recs, grid = make_step(rng, n, 0, n, tile, od, words, 0.3)
    recs[5, 0] = (recs[5, 0] + 500) % n              # an env of another tile
    obs = np.zeros((n, od), np.float32); reward = np.zeros(n, np.float32); done = np.zeros(n, np.uint8); ran = np.zeros(n, np.uint8)
    rc = lib.patch_apply(table.ctypes.data, grid, tile, 0, n, recs.ctypes.data, words, od,
                         obs.ctypes.data, reward.ctypes.data, done.ctypes.data, ran.ctypes.data)
    assert rc == 1
