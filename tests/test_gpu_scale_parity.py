"""GPU, BASELINE.json full sizes against the ORACLE (not against invariants, not CUDA against CUDA):
the step kernel at the tile sizes and instantiations the benchmark runs (1,048,576 envs; the 131,072-env
strong-scaling shard; a ragged 300,001-env batch; forced tile sizes), the two host-buffer entry points at
1,048,576 envs, and sampled frames of the 16,384-env render -- all compared with the threaded C oracle
(oracle/tg_oracle.c via c_oracle.ShardedCBatch, Philox mode, bit-exact) or the CPU render restatement.
Reference behaviour: treasure_game.py:91-96, _option.py:20-36, _treasure_game_impl.py:290-359."""
import numpy as np
import pytest
import torch

import c_oracle
import py_oracle as po
import render_oracle as ro
from gpu_util import assert_state_equal, assert_step_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def VTG():
    from gym_treasure_game_b200 import VectorTreasureGame
    return VectorTreasureGame


def _pair(VTG, n, seed, max_steps, **kw):
    lvt = po.default_level()
    env = VTG(n, seed=seed, max_episode_steps=max_steps, auto_reset=True, render=False, **kw)
    cb = c_oracle.ShardedCBatch(c_oracle.CLevel(lvt), n, first_env_id=kw.get("first_env_id", 0), seed=seed,
                                max_episode_steps=max_steps, auto_reset=True)
    cb.reset()
    return env, cb


def _actions(g, n, cb, t):
    """Uniform-random option ids (the benchmark's law); every fourth step draws from the runnable options so that
    the episodes leave the start corridor (keys, doors, drops and jumps get exercised)."""
    if t % 4 != 3:
        return torch.randint(0, 9, (n,), generator=g, dtype=torch.int32)
    m = torch.from_numpy(cb.mask().astype(np.float32)) + 1e-6
    return torch.multinomial(m, 1, generator=g).squeeze(1).to(torch.int32)


def test_million_envs_match_oracle(VTG):
    """BASELINE configs[2] on one GPU: the benchmarked launch (1,048,576 envs, automatic tile size), 130 steps so
    that the step-100 mass reset is inside; outputs every step, full state every 25 steps, statistics at the end."""
    n, steps = 1 << 20, 130
    env, cb = _pair(VTG, n, seed=2025, max_steps=100)
    assert_state_equal(env, cb, "after construction")
    g = torch.Generator().manual_seed(11)
    for t in range(steps):
        a = _actions(g, n, cb, t)
        out = env.step_raw(a.cuda())
        ref = cb.step(a.numpy())
        assert_step_equal(out, ref, "step %d" % t)
        if t % 25 == 24 or t in (99, 100):
            assert_state_equal(env, cb, "step %d" % t)
    np.testing.assert_array_equal(env.available_mask.cpu().numpy(), cb.mask())
    st = env.stats()
    assert list(st.values()) == cb.stats().tolist()
    assert st["errors"] == 0 and st["gym_steps"] == n * steps and st["episodes"] >= n
    env.close()


@pytest.mark.parametrize("n,max_steps", [(131072, 100), (300001, 37)])
def test_shard_and_ragged_sizes_match_oracle(VTG, n, max_steps):
    """The 131,072-env strong-scaling shard of configs[2] and a ragged batch (last tile partial, not a multiple of 4)."""
    env, cb = _pair(VTG, n, seed=77 + n, max_steps=max_steps, first_env_id=5 * n)
    g = torch.Generator().manual_seed(n)
    for t in range(120):
        a = _actions(g, n, cb, t)
        assert_step_equal(env.step_raw(a.cuda()), cb.step(a.numpy()), "step %d" % t)
        if t % 40 == 39:
            assert_state_equal(env, cb, "step %d" % t)
    assert list(env.stats().values()) == cb.stats().tolist()
    env.close()


@pytest.mark.parametrize("tile", [32, 36, 256, 1000, 2364, 4096])
def test_forced_tile_sizes_match_oracle(VTG, tile):
    """Every tile-size-dependent code path (chunk ranking, record blocks, row alignment of the vector stores)
    at n = 20,000 with the tile forced through tg_debug_set_step_tile."""
    n = 20000
    env, cb = _pair(VTG, n, seed=900 + tile, max_steps=23)
    env.set_step_tile(tile)
    g = torch.Generator().manual_seed(tile)
    for t in range(80):
        a = _actions(g, n, cb, t)
        assert_step_equal(env.step_raw(a.cuda()), cb.step(a.numpy()), "tile %d step %d" % (tile, t))
    assert_state_equal(env, cb, "final")
    assert list(env.stats().values()) == cb.stats().tolist()
    env.close()


@pytest.mark.parametrize("path", ["step_host", "step_host_sparse"])
def test_host_paths_million_envs_match_oracle(VTG, path):
    """tg_step_host / tg_step_host_sparse at 1,048,576 envs against the oracle (not against each other):
    host arrays after every call, incl. the step at which every env is reset (max_episode_steps = 12)."""
    n = 1 << 20
    env, cb = _pair(VTG, n, seed=31337, max_steps=12)
    host = env.make_host_buffers()
    fn = getattr(env, path)
    g = torch.Generator().manual_seed(5)
    for t in range(30):
        a = _actions(g, n, cb, t)
        host["actions"].copy_(a)
        fn(host)
        o2, r2, d2, ran2, _ = cb.step(a.numpy())
        np.testing.assert_array_equal(host["reward"].numpy(), r2, err_msg="reward step %d" % t)
        np.testing.assert_array_equal(host["done"].numpy(), d2, err_msg="done step %d" % t)
        np.testing.assert_array_equal(host["ran"].numpy(), ran2, err_msg="ran step %d" % t)
        np.testing.assert_array_equal(host["obs"].numpy(), o2.astype(np.float32), err_msg="obs step %d" % t)
    assert_state_equal(env, cb, "final")
    assert list(env.stats().values()) == cb.stats().tolist()
    env.close()


def test_render_16384_frames_sampled_against_restatement(VTG):
    """BASELINE configs[3]: one 16,384-frame render launch; 72 sampled frames (first and last job, every region of
    the batch) pixel-exact against oracle/render_oracle.render_frame on the oracle's state of the same env."""
    n, seed = 16384, 4242
    lvt = po.default_level()
    env = VTG(n, seed=seed, max_episode_steps=0, auto_reset=False)
    cb = c_oracle.ShardedCBatch(c_oracle.CLevel(lvt), n, first_env_id=0, seed=seed)
    cb.reset()
    g = torch.Generator().manual_seed(9)
    for t in range(36):
        m = torch.from_numpy(cb.mask().astype(np.float32)) + 1e-6
        a = torch.multinomial(m, 1, generator=g).squeeze(1).to(torch.int32)
        env.step_raw(a.cuda())
        cb.step(a.numpy())
    assert_state_equal(env, cb, "before render")
    frames = env.render()
    assert frames.shape == (n, 624, 672, 3)
    bg = ro.background(lvt.tiles)
    rng = np.random.default_rng(0)
    idx = sorted(set([0, 1, 2, 255, 256, 257, n - 257, n - 256, n - 2, n - 1] + rng.integers(0, n, 62).tolist()))
    for i in idx:
        want = ro.render_frame(lvt, cb.snapshot(i), bg)
        got = frames[i].cpu().numpy()
        assert np.array_equal(got, want), (i, int((got != want).sum()))
    assert len(idx) >= 64
    env.close()
