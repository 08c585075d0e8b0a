"""GPU: per-tick frame streaming (tg_step_frames) -- the option layer with a drawer (_option.py:20-36, :33-34:
drawer.draw_domain() after every primitive tick).  The tick-by-tick states come from the reference-pinned Python
oracle replaying a reference-generated trajectory (golden fixture, reference draws injected), the frames from the CPU
restatement of the drawer; the CUDA path replays the same trajectory with the same draws."""
import numpy as np
import pytest
import torch

import py_oracle as po
import render_oracle as ro
from conftest import golden_files, golden_level, load_golden, norm_snap
from gpu_util import product_level

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", [p for p in golden_files() if "default_solve_s7" in p or "twin_solve" in p],
                         ids=lambda p: p.split("/")[-1][:-8])
def test_frames_after_every_tick_match_oracle(path):
    from gym_treasure_game_b200 import VectorTreasureGame
    rec = load_golden(path)
    lvt = golden_level(rec)
    T = 112
    env = VectorTreasureGame(1, seed=1, auto_reset=False, levels=[product_level(lvt)])
    env.set_draw_tape([rec["tape"]])
    env.reset(); env.reset()
    oe = po.OracleEnv(lvt, po.TapeUniform(rec["tape"]))      # constructor draws
    oe.reset()
    bg = ro.background(lvt.tiles)
    checked = 0
    for t, st in enumerate(rec["steps"]):
        snaps = []
        obs2, r2, done2, _ = oe.gym_step(st["a"], on_tick=lambda: snaps.append(oe.snapshot()))
        frames, n_ticks, obs, rew, done, ran = env.step_frames([0], [st["a"]], max_ticks=T)
        n = int(n_ticks[0])
        assert n == len(snaps), (t, n, len(snaps))
        assert bool(ran[0]) == (r2 is not None) and int(rew[0]) == (r2 or 0), t
        assert bool(done[0] & 1) == done2, t
        np.testing.assert_array_equal(obs[0].cpu().numpy(), np.asarray(obs2, dtype=np.float32), err_msg=str(t))
        assert env.snapshot(0) == norm_snap(st["snap"]), t          # the step itself is the reference's
        if n == 0:
            continue
        # every tick of a few options, first / middle / last tick of the others (one oracle frame takes ~0.5 s)
        ticks = range(n) if t % 7 == 0 and n <= 40 else sorted({0, n // 2, n - 1})
        fr = frames[0].cpu().numpy()
        for k in ticks:
            want = ro.render_frame(lvt, snaps[k], bg)
            assert np.array_equal(fr[k], want), (t, k, int((fr[k] != want).sum()))
            checked += 1
        assert np.array_equal(fr[min(n, T) - 1], fr[T - 1])         # frames after the last tick repeat it
    assert checked >= 60
    env.close()


def test_frames_batch_subset_leaves_other_envs_alone():
    """A subset of a batch steps with frames; the other envs keep their state and their episode clocks."""
    import c_oracle
    from gpu_util import assert_state_equal
    from gym_treasure_game_b200 import VectorTreasureGame
    n, seed = 64, 77
    lvt = po.default_level()
    env = VectorTreasureGame(n, seed=seed, max_episode_steps=30, auto_reset=True)
    ref = VectorTreasureGame(n, seed=seed, max_episode_steps=30, auto_reset=True, render=False)
    g = torch.Generator().manual_seed(2)
    for _ in range(12):
        a = torch.randint(0, 9, (n,), generator=g, dtype=torch.int32).cuda()
        env.step_raw(a); ref.step_raw(a)
    ids = torch.tensor([3, 17, 40, 63])
    acts = torch.tensor([3, 1, 0, 3], dtype=torch.int32)
    before = {k: v.clone() for k, v in env.get_state().items()}
    frames, n_ticks, obs, rew, done, ran = env.step_frames(ids, acts, max_ticks=40)
    assert int(n_ticks.max()) > 0
    after = env.get_state()
    others = torch.ones(n, dtype=torch.bool); others[ids] = False
    for k in before:
        assert torch.equal(before[k][others.cuda()], after[k][others.cuda()]), k
    # the selected envs: same result as a whole-batch step of the twin with those actions (Philox streams are per env)
    full = torch.full((n,), 2, dtype=torch.int32); full[ids] = acts
    o2, r2, d2, ran2 = ref.step_raw(full.cuda())
    sel = ids.cuda()
    assert torch.equal(obs, o2[sel]) and torch.equal(rew, r2[sel]) and torch.equal(ran, ran2[sel])
    st2 = ref.get_state()
    for k in ("pos", "misc", "doors", "handles", "bolts", "angles", "items", "bag"):
        assert torch.equal(after[k][sel], st2[k][sel]), k
    assert torch.equal(after["acct"][sel][:, 1], st2["acct"][sel][:, 1])        # episode steps advanced by one
    # frame of the last tick == a plain render of the env afterwards (no reset happened: 13 < 30 steps)
    plain = env.render()
    for j, i in enumerate(ids.tolist()):
        nt = int(n_ticks[j])
        assert nt >= 0
        if nt <= 40:                                   # ticks beyond max_ticks are run but not drawn
            assert torch.equal(frames[j, max(nt, 1) - 1], plain[i])
        assert torch.equal(frames[j, 39], plain[i]) or nt > 40
    env.close(); ref.close()
