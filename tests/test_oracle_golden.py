"""CPU: both oracle restatements (pure Python, plain C) against the
reference-generated golden trajectories (tests/golden, tools/gen_golden.py).
Bit-exact on every field, including float64 handle angles and state vectors."""
import numpy as np
import pytest

import c_oracle
import py_oracle as po
from conftest import golden_files, golden_level, load_golden, norm_snap


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-8])
def test_py_oracle_replays_golden(path):
    rec = load_golden(path)
    env = po.OracleEnv(golden_level(rec), po.TapeUniform(rec["tape"]))   # ctor draws
    obs = env.reset()                                                      # reset draws
    assert env.draws == rec["draws_ctor_and_reset"]
    assert env.snapshot() == norm_snap(rec["init"]["snap"])
    assert env.obs() == rec["init"]["obs"]
    for t, st in enumerate(rec["steps"]):
        assert env.mask() == st["mask"], t
        if "restore" in st:                        # impl:447-481 init_with_state
            env.init_with_state(list(st["restore"]))
            assert env.obs() == st["obs"] and env.draws == st["draws"], t
            assert env.snapshot() == norm_snap(st["snap"]), t
            continue
        before = env.total_actions
        obs, r, done, _ = env.gym_step(st["a"])
        assert r == st["r"], t
        assert done == st["done"], t
        assert obs == st["obs"], t
        assert env.total_actions - before == st["ticks"], t
        assert env.draws == st["draws"], t
        assert env.snapshot() == norm_snap(st["snap"]), t


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-8])
def test_c_oracle_replays_golden(path):
    rec = load_golden(path)
    lv = c_oracle.CLevel(golden_level(rec))
    b = c_oracle.CBatch(lv, 1)
    b.set_tape([rec["tape"]])
    b.reset()                      # constructor-equivalent draws
    obs = b.reset()                # TreasureGame.reset()
    assert b.snapshot() == norm_snap(rec["init"]["snap"])
    assert obs[0].tolist() == rec["init"]["obs"]
    for t, st in enumerate(rec["steps"]):
        assert b.mask()[0].tolist() == st["mask"], t
        if "restore" in st:
            b.init_with_state([st["restore"]])
            assert b.snapshot() == norm_snap(st["snap"]), t
            assert int(b.state()["misc"][0, 3]) == st["draws"], t
            continue
        obs, rew, done, ran, ticks = b.step([st["a"]])
        assert bool(ran[0]) == (st["r"] is not None), t
        assert int(rew[0]) == (st["r"] or 0), t
        assert bool(done[0] & 1) == st["done"], t
        assert obs[0].tolist() == st["obs"], t
        assert int(ticks[0]) == st["ticks"], t
        snap = b.snapshot()
        assert snap == norm_snap(st["snap"]), t
        assert int(b.state()["misc"][0, 3]) == st["draws"], t


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10."""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert c_oracle.philox(ctr, key).tolist() == want


def test_c_oracle_philox_mode_matches_py_oracle_with_same_uniforms():
    """C oracle in Philox mode == Python oracle fed the same Philox uniforms."""
    lvt = po.default_level()
    lv = c_oracle.CLevel(lvt)
    n, seed, first = 16, 0x1234567890ABCDEF, 1000
    b = c_oracle.CBatch(lv, n, first_env_id=first, seed=seed)
    b.reset()
    rng = np.random.default_rng(0)

    class Stream:
        """draw j of an API call that began at draw index d0 = word j&3 of Philox(ctr=(d0, j>>2, env id), key=seed)"""
        def __init__(self, env_id):
            self.env_id, self.d, self.d0 = env_id, 0, 0

        def begin_call(self):
            self.d0 = self.d

        def __call__(self):
            j = self.d - self.d0
            self.d += 1
            w = c_oracle.philox([self.d0, j >> 2, self.env_id & 0xffffffff, self.env_id >> 32],
                                [seed & 0xffffffff, seed >> 32])
            return int(w[j & 3]) / 4294967296.0

    envs, streams = [], []
    for i in range(n):
        g = Stream(first + i)
        e = po.OracleEnv.__new__(po.OracleEnv)
        e.level, e.u = lvt, g
        e.ch, e.cw = len(lvt.tiles), len(lvt.tiles[0])
        e.width, e.height, e.draws = e.cw * po.S, e.ch * po.S, 0
        e.reset()
        envs.append(e)
        streams.append(g)
    for t in range(60):
        m = b.mask()
        acts = [int(rng.choice(np.flatnonzero(m[i]))) if t % 2 else int(rng.integers(9)) for i in range(n)]
        obs, rew, done, ran, ticks = b.step(acts)
        for i, e in enumerate(envs):
            streams[i].begin_call()
            o2, r2, d2, _ = e.gym_step(acts[i])
            assert obs[i].tolist() == o2
            assert (r2 is None) == (not ran[i]) and int(rew[i]) == (r2 or 0)
            assert b.snapshot(i) == e.snapshot()
