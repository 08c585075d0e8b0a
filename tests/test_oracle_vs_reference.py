"""CPU, development container only: live differential test of the oracle
restatements against the UNMODIFIED reference imported from /root/reference
(skipped where that tree does not exist, e.g. on the GPU box -- the committed
golden trajectories cover that case)."""
import os
import random
import tempfile

import pytest

import c_oracle
import py_oracle as po
import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="/root/reference not present")


def _ref_files():
    d = rh.load_reference().default_dir
    return [os.path.join(d, f) for f in ("domain.txt", "domain-objects.txt", "domain-interactions.txt")]


def test_embedded_default_level_equals_reference_files():
    a = po.LevelText.from_files(*_ref_files())
    b = po.default_level()
    assert [r for r in a.tiles if r] == b.tiles
    assert a.objects == b.objects and a.triggers == b.triggers


@pytest.mark.parametrize("mirrored", [False, True])
@pytest.mark.parametrize("runnable_only", [False, True])
def test_live_differential(mirrored, runnable_only):
    lvt = po.default_level()
    if mirrored:
        lvt = po.mirrored_level(lvt)
    clv = c_oracle.CLevel(lvt)
    with tempfile.TemporaryDirectory() as td:
        paths = []
        for fn, txt in zip(("d", "o", "i"), po.level_to_strings(lvt)):
            p = os.path.join(td, fn)
            open(p, "w").write(txt)
            paths.append(p)
        for seed in range(6):
            random.seed(seed + 1000 * mirrored)
            arng = random.Random(seed)
            with rh.DrawTap() as tap:
                g = rh.RefGame(*paths)
                g.reset()
                pe = po.OracleEnv(lvt, po.TapeUniform(tap.tape))
                pe.reset()
                acts, snaps = [], []
                for t in range(200):
                    m = g.mask()
                    assert m == pe.mask()
                    a = arng.choice([i for i in range(9) if m[i]]) if runnable_only else arng.randrange(9)
                    s, r, d, _ = g.step(a)
                    s2, r2, d2, _ = pe.gym_step(a)
                    assert (s, r, d) == (s2, r2, d2)
                    assert pe.snapshot() == rh.impl_snapshot(g.env)
                    assert pe.draws == tap.pos
                    acts.append(a)
                    snaps.append((rh.impl_snapshot(g.env), s, r, d))
                tape = list(tap.tape)
            cb = c_oracle.CBatch(clv, 1)
            cb.set_tape([tape])
            cb.reset()
            cb.reset()
            for a, (snap, s, r, d) in zip(acts, snaps):
                obs, rew, done, ran, _ = cb.step([a])
                assert cb.snapshot() == snap
                assert obs[0].tolist() == s and bool(done[0] & 1) == d
                assert (r is None) == (not ran[0]) and int(rew[0]) == (r or 0)


def test_appendix_e_known_answer():
    """seed 7 solved trajectory, regenerated here."""
    acts = [3, 0, 4, 4, 1, 1, 3, 1, 4, 4, 0, 0, 5, 7, 7, 1, 8, 7, 0, 1, 6, 0, 3, 0, 4, 1, 1, 8, 8, 8, 1, 0, 5, 5,
            5, 0, 2, 1, 8, 8, 0, 7, 8, 1, 1, 4, 0, 2, 0, 2]
    ref = rh.load_reference()
    random.seed(7)
    with rh.DrawTap() as tap:
        g = ref.TreasureGame()
        g.reset()
        tot, d = 0, False
        for a in acts:
            _, r, d, _ = g.step(a)
            tot += r
        assert (tot, d, g._env.playerx, g._env.playery, g._env.total_actions, tap.pos) == (-1836, True, 216, 1, 1792, 1570)
        tape = list(tap.tape)
    pe = po.OracleEnv(po.default_level(), po.TapeUniform(tape))
    pe.reset()
    tot2 = sum(pe.gym_step(a)[1] for a in acts)
    assert (tot2, pe.is_done(), pe.px, pe.py, pe.total_actions, pe.draws) == (-1836, True, 216, 1, 1792, 1570)


def test_primitive_step_live():
    """_TreasureGameImpl.step(act) (impl:290-359) with raw actions, reference vs both oracles."""
    lvt = po.default_level()
    clv = c_oracle.CLevel(lvt)
    for seed in range(4):
        random.seed(seed)
        arng = random.Random(seed + 50)
        with rh.DrawTap() as tap:
            g = rh.RefGame()
            g.reset()
            pe = po.OracleEnv(lvt, po.TapeUniform(tap.tape))
            pe.reset()
            acts, snaps = [], []
            for t in range(1500):
                a = arng.choice([0, 1, 2, 3, 3, 3, 4, 4, 4, 5, 6, 7])      # 7: not an action id (falls through like NOP)
                r = g.env.step(a)
                assert pe.tick(a) == r
                assert pe.snapshot() == rh.impl_snapshot(g.env)
                acts.append(a)
                snaps.append((rh.impl_snapshot(g.env), g.env.get_state(), r))
            tape = list(tap.tape)
        cb = c_oracle.CBatch(clv, 1)
        cb.set_tape([tape])
        cb.reset()
        cb.reset()
        for a, (snap, st, r) in zip(acts, snaps):
            obs, rew, _ = cb.prim_step([a])
            assert cb.snapshot() == snap and obs[0].tolist() == st and int(rew[0]) == r
