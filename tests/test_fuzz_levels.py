"""Differential fuzzing on random layouts: the level is data, not code.

CPU: Python oracle vs C oracle on a shared MT draw tape (and, where /root/reference exists, the
unmodified reference under a watchdog).  GPU: the CUDA path vs the C oracle in Philox mode on the
same layouts.  Environments that hit a reference failure mode (option that never terminates /
target None) are flagged by both implementations and excluded from further comparison."""
import os
import random
import signal
import tempfile

import numpy as np
import pytest

import c_oracle
import py_oracle as po
import ref_harness as rh
from level_fuzz import random_level, usable

SEEDS = [s for s in range(60) if usable(random_level(s))][:28]


def test_generator_yields_enough_levels():
    assert len(SEEDS) >= 20


@pytest.mark.parametrize("seed", SEEDS)
def test_py_vs_c_oracle_random_level(seed):
    lv = random_level(seed)
    tape_rng = random.Random(seed)
    tape = [tape_rng.random() for _ in range(60000)]
    pe = po.OracleEnv(lv, po.TapeUniform(tape))
    cb = c_oracle.CBatch(c_oracle.CLevel(lv), 1)
    cb.set_tape([tape])
    cb.reset()
    arng = random.Random(seed + 1)
    for t in range(250):
        m = pe.mask()
        assert cb.mask()[0].tolist() == m, t
        run = [i for i in range(9) if m[i]]
        a = arng.choice(run) if run and arng.random() < 0.8 else arng.randrange(9)
        try:
            obs, r, done, _ = pe.gym_step(a)
        except po.ReferenceWouldFail:
            _, _, _, _, _ = cb.step([a])
            assert cb.state()["acct"][0, 2] == 1          # the C oracle flags the same situation
            break
        o2, rew, d2, ran, _ = cb.step([a])
        assert obs == o2[0].tolist() and (r is None) == (not ran[0]) and int(rew[0]) == (r or 0), t
        assert cb.snapshot() == pe.snapshot(), t
        if t % 40 == 39:
            pe.reset()
            cb.reset()


class _Timeout(Exception):
    pass


@pytest.mark.skipif(not rh.reference_available(), reason="/root/reference not present")
@pytest.mark.parametrize("seed", SEEDS[:12])
def test_reference_vs_py_oracle_random_level(seed):
    lv = random_level(seed)

    def on_alarm(sig, frm):
        raise _Timeout()

    old = signal.signal(signal.SIGALRM, on_alarm)
    try:
        with tempfile.TemporaryDirectory() as td:
            paths = []
            for fn, txt in zip(("d", "o", "i"), po.level_to_strings(lv)):
                p = os.path.join(td, fn)
                open(p, "w").write(txt)
                paths.append(p)
            random.seed(seed)
            arng = random.Random(seed + 1)
            with rh.DrawTap() as tap:
                g = rh.RefGame(*paths)
                g.reset()
                pe = po.OracleEnv(lv, po.TapeUniform(tap.tape))
                pe.reset()
                assert pe.snapshot() == rh.impl_snapshot(g.env)
                for t in range(150):
                    m = g.mask()
                    assert m == pe.mask(), t
                    run = [i for i in range(9) if m[i]]
                    a = arng.choice(run) if run and arng.random() < 0.8 else arng.randrange(9)
                    signal.alarm(2)
                    try:
                        s, r, d, _ = g.step(a)
                    except (_Timeout, TypeError):
                        with pytest.raises(po.ReferenceWouldFail):       # hang or close_enough_x(None)
                            pe.gym_step(a)
                        break
                    finally:
                        signal.alarm(0)
                    s2, r2, d2, _ = pe.gym_step(a)
                    assert (s, r, d) == (s2, r2, d2), t
                    assert pe.snapshot() == rh.impl_snapshot(g.env), t
    finally:
        signal.signal(signal.SIGALRM, old)


@pytest.mark.gpu
def test_cuda_vs_c_oracle_random_levels():
    import torch
    from gpu_util import assert_state_equal, product_level
    from gym_treasure_game_b200 import VectorTreasureGame
    g = torch.Generator().manual_seed(77)
    for seed in SEEDS:
        lv = random_level(seed)
        n = 256
        env = VectorTreasureGame(n, seed=seed, max_episode_steps=30, auto_reset=True, levels=[product_level(lv)], render=False)
        cb = c_oracle.CBatch(c_oracle.CLevel(lv), n, first_env_id=0, seed=seed, max_episode_steps=30, auto_reset=True)
        cb.reset()
        assert_state_equal(env, cb, "level %d after construction" % seed)
        for t in range(90):
            m = torch.from_numpy(cb.mask().astype(np.float32)) + 0.05
            a = torch.multinomial(m, 1, generator=g).squeeze(1).to(torch.int32)
            obs, rew, done, ran = env.step_raw(a.cuda())
            o2, r2, d2, ran2, _ = cb.step(a.numpy())
            np.testing.assert_array_equal(env.available_mask.cpu().numpy(), cb.mask(), err_msg="mask level %d step %d" % (seed, t))
            np.testing.assert_array_equal(rew.cpu().numpy(), r2, err_msg="reward level %d step %d" % (seed, t))
            np.testing.assert_array_equal(done.cpu().numpy(), d2, err_msg="done level %d step %d" % (seed, t))
            np.testing.assert_array_equal(obs.cpu().numpy(), o2.astype(np.float32), err_msg="obs level %d step %d" % (seed, t))
        assert_state_equal(env, cb, "level %d final" % seed)
        assert list(env.stats().values()) == cb.stats().tolist()
        env.close()


@pytest.mark.gpu
def test_cuda_graph_walk_equals_closure_table(monkeypatch):
    """The INTERACT tick reads the trigger closure table; batches created with TG_NO_CLOSURE walk the trigger graph on the
    device instead (the path kept for sticky handles and untabulated cascades).  Both must give the oracle's results."""
    import torch
    from gpu_util import assert_state_equal, product_level
    from gym_treasure_game_b200 import VectorTreasureGame
    g = torch.Generator().manual_seed(5)
    for seed in [s for s in SEEDS if random_level(s).triggers][:6]:
        lv = random_level(seed)
        n = 512
        monkeypatch.setenv("TG_NO_CLOSURE", "1")
        walk = VectorTreasureGame(n, seed=seed, max_episode_steps=40, auto_reset=True, levels=[product_level(lv)], render=False)
        monkeypatch.delenv("TG_NO_CLOSURE")
        table = VectorTreasureGame(n, seed=seed, max_episode_steps=40, auto_reset=True, levels=[product_level(lv)], render=False)
        cb = c_oracle.CBatch(c_oracle.CLevel(lv), n, first_env_id=0, seed=seed, max_episode_steps=40, auto_reset=True)
        cb.reset()
        for t in range(80):
            m = torch.from_numpy(cb.mask().astype(np.float32))
            m[:, 4] *= 6.0                               # interact whenever it can run, mostly
            a = torch.multinomial(m + 0.02, 1, generator=g).squeeze(1).to(torch.int32)
            outs = [e.step_raw(a.cuda()) for e in (walk, table)]
            o2, r2, d2, ran2, _ = cb.step(a.numpy())
            for e, (obs, rew, done, ran) in zip((walk, table), outs):
                np.testing.assert_array_equal(obs.cpu().numpy(), o2.astype(np.float32), err_msg="obs level %d step %d" % (seed, t))
                np.testing.assert_array_equal(rew.cpu().numpy(), r2); np.testing.assert_array_equal(done.cpu().numpy(), d2)
        assert_state_equal(walk, cb, "graph walk, level %d" % seed)
        assert_state_equal(table, cb, "closure table, level %d" % seed)
        walk.close(); table.close()
