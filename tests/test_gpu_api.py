"""GPU: the reference-facing single-env Gym surface (treasure_game.py:38-114) served by the CUDA path."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_readme_loop_and_none_reward():
    import gym_treasure_game_b200 as tgb
    random.seed(0)
    env = tgb.make("treasure_game-v0")
    assert env.action_space.n == 9 and env.observation_space.shape == (9,)
    s = env.reset()
    assert isinstance(s, list) and len(s) == 9 and all(isinstance(v, float) for v in s)
    assert env.available_mask.tolist() == [0, 0, 0, 1, 0, 0, 0, 0, 0]
    s2, r, done, info = env.step(0)                                        # go_left is not runnable at the start
    assert r is None and done is False and info == {} and s2 == s
    s3, r, done, _ = env.step(3)                                           # down the first ladder
    assert isinstance(r, int) and r < 0 and s3 != s
    for episode in range(2):                                               # README.md:40-48
        env.reset()
        for _ in range(100):
            _, reward, done, _ = env.step(env.action_space.sample())
            assert reward is None or reward < 0
            if done:
                break
    with pytest.raises(IndexError):
        env.step(9)
    env.close()


def test_observation_wrapper():
    from gym_treasure_game_b200.envs import ObservationWrapper, TreasureGame
    env = ObservationWrapper(TreasureGame(seed=5))
    obs = env.reset()
    assert obs.shape == (624, 672, 3) and obs.dtype == np.uint8
    frame, r, done, info = env.step(3)
    assert frame.shape == (624, 672, 3) and len(info["world_state"]) == 9
    assert (frame != obs).any()
    env.close()


def test_seed_reproducibility():
    from gym_treasure_game_b200.envs import TreasureGame
    outs = []
    for _ in range(2):
        random.seed(123)
        e = TreasureGame()
        tr = [e.reset()]
        for a in [3, 0, 4, 4, 1, 1, 3]:
            tr.append(e.step(a))
        outs.append(tr)
        e.close()
    assert outs[0] == outs[1]


@pytest.mark.parametrize("n", [1000, 300000])
def test_step_host_sparse_equals_dense(n):
    """tg_step_host_sparse patches the host arrays of the previous call; after every step they must equal what the
    dense tg_step_host of a twin batch delivers (obs, reward, done, ran), including resets and the fallbacks."""
    import torch
    from gym_treasure_game_b200 import VectorTreasureGame
    a = VectorTreasureGame(n, seed=5, max_episode_steps=7, auto_reset=True, render=False)
    b = VectorTreasureGame(n, seed=5, max_episode_steps=7, auto_reset=True, render=False)
    ha, hb = a.make_host_buffers(), b.make_host_buffers()
    g = torch.Generator().manual_seed(3)
    d2h0 = a.host_traffic()[1]
    for t in range(24):
        acts = torch.randint(0, 9, (n,), generator=g, dtype=torch.int32)
        ha["actions"].copy_(acts); hb["actions"].copy_(acts)
        if t == 12:
            a.reset(); b.reset()                      # a device-side call in between: the next sparse call must fall back
        a.step_host_sparse(ha)
        b.step_host(hb)
        for k in ("obs", "reward", "done", "ran"):
            assert torch.equal(ha[k], hb[k]), (t, k)
    sparse_bytes = a.host_traffic()[1] - d2h0
    dense_bytes = b.host_traffic()[1]
    assert sparse_bytes < 0.6 * dense_bytes           # 2 dense fallbacks + 22 sparse steps
    assert a.stats() == b.stats()


@pytest.mark.parametrize("n,parts", [(5000, 2), (300001, 2), (40000, 3)])
def test_pipelined_host_env_equals_one_batch(n, parts):
    """PipelinedHostEnv (tg_step_host_sparse_begin / _end over sub-batches in flight on their own streams) delivers, part
    by part, what one batch over the whole id range delivers through tg_step_host -- every step, resets included."""
    import torch
    from gym_treasure_game_b200 import PipelinedHostEnv, VectorTreasureGame
    p = PipelinedHostEnv(n, parts=parts, first_env_id=1000, seed=5, max_episode_steps=7, auto_reset=True, render=False)
    b = VectorTreasureGame(n, first_env_id=1000, seed=5, max_episode_steps=7, auto_reset=True, render=False)
    hb = b.make_host_buffers()
    g = torch.Generator().manual_seed(3)
    acts = [torch.randint(0, 9, (n,), generator=g, dtype=torch.int32) for _ in range(20)]
    pinned = [a.pin_memory() for a in acts]
    for k, (lo, hi) in enumerate(p.ranges):
        p.hosts[k]["actions"] = pinned[0][lo:hi]
        p.begin(k)
    for t in range(20):
        hb["actions"].copy_(acts[t])
        b.step_host(hb)
        for k, (lo, hi) in enumerate(p.ranges):
            h = p.end(k)
            for key in ("obs", "reward", "done", "ran"):
                assert torch.equal(h[key], hb[key][lo:hi]), (t, k, key)
            if t + 1 < 20:
                p.hosts[k]["actions"] = pinned[t + 1][lo:hi]
                p.begin(k)
    assert p.stats() == b.stats()
    h2d, d2h = p.host_traffic()
    assert h2d == 20 * n * 4 and d2h < 0.6 * b.host_traffic()[1]
    p.close(); b.close()


def test_sparse_begin_end_contract():
    """Between _begin and _end every other state-changing call on the env is refused; _end needs a _begin."""
    import torch
    from gym_treasure_game_b200 import VectorTreasureGame
    from gym_treasure_game_b200._lib import TreasureError
    e = VectorTreasureGame(3000, seed=1, max_episode_steps=9, auto_reset=True, render=False)
    h = e.make_host_buffers()
    h["actions"].zero_()
    with pytest.raises(TreasureError):
        e.step_host_sparse_end()
    e.step_host_sparse(h)                             # primes the arrays (dense first call)
    e.step_host_sparse_begin(h)
    with pytest.raises(TreasureError):
        e.step_host_sparse_begin(h)
    with pytest.raises(TreasureError):
        e.step(torch.zeros(3000, dtype=torch.int32, device="cuda"))
    with pytest.raises(TreasureError):
        e.reset()
    e.step_host_sparse_end()
    e.step(torch.zeros(3000, dtype=torch.int32, device="cuda"))    # allowed again
    e.close()


def test_step_host_sparse_without_auto_reset():
    """Without auto-reset a finished env reports done in every step without being touched; the sparse entry point must
    deliver that too (it takes the dense path)."""
    import torch
    from gym_treasure_game_b200 import VectorTreasureGame
    n = 6000
    a = VectorTreasureGame(n, seed=2, max_episode_steps=5, auto_reset=False, render=False)
    b = VectorTreasureGame(n, seed=2, max_episode_steps=5, auto_reset=False, render=False)
    ha, hb = a.make_host_buffers(), b.make_host_buffers()
    g = torch.Generator().manual_seed(9)
    for t in range(12):
        acts = torch.randint(0, 9, (n,), generator=g, dtype=torch.int32)
        ha["actions"].copy_(acts); hb["actions"].copy_(acts)
        a.step_host_sparse(ha); b.step_host(hb)
        for k in ("obs", "reward", "done", "ran"):
            assert torch.equal(ha[k], hb[k]), (t, k)
    assert int(ha["done"].ne(0).sum()) == n            # past the time limit every env stays done


def test_step_host_sparse_mixed_layouts():
    """Sparse records with several layouts in one batch: observation rows of different lengths (zero padded),
    the generic observation program and the 4-item kernels."""
    import os, sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import gen_golden
    from gpu_util import product_level
    from gym_treasure_game_b200 import VectorTreasureGame
    lv = gen_golden.variant_levels()
    names = ["default", "mirror", "twin", "altinit"]
    n = 4099                                                    # ragged: the last tile is not a multiple of 32
    ids = np.random.default_rng(2).integers(0, len(names), n).astype(np.uint8)
    mk = lambda: VectorTreasureGame(n, seed=8, max_episode_steps=9, auto_reset=True, render=False,
                                    levels=[product_level(lv[k]) for k in names], level_ids=ids)
    a, b = mk(), mk()
    ha, hb = a.make_host_buffers(), b.make_host_buffers()
    g = torch.Generator().manual_seed(6)
    for t in range(30):
        acts = torch.randint(0, 9, (n,), generator=g, dtype=torch.int32)
        ha["actions"].copy_(acts); hb["actions"].copy_(acts)
        a.step_host_sparse(ha); b.step_host(hb)
        for k in ("obs", "reward", "done", "ran"):
            assert torch.equal(ha[k], hb[k]), (t, k)
    assert a.obs_dim > 9 and a.stats() == b.stats()
