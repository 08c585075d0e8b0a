"""GPU parity tests proper: the CUDA path, called through the C ABI, against
(a) the reference-generated golden trajectories with the reference's uniform
draws injected, and (b) the C oracle in Philox mode on seeded batches.
Integer/boolean/index state is compared bit-exactly; float64 handle angles
bit-exactly; float32 observations against the float64 oracle value rounded to
float32 (tolerance stated by north_star: 1e-6 relative -- we get 0)."""
import numpy as np
import pytest
import torch

import c_oracle
import py_oracle as po
from conftest import golden_files, golden_level, load_golden, norm_snap
from gpu_util import assert_state_equal, assert_step_equal, product_level

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def VTG():
    from gym_treasure_game_b200 import VectorTreasureGame
    return VectorTreasureGame


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-8])
def test_cuda_replays_reference_golden(VTG, path):
    rec = load_golden(path)
    lvt = golden_level(rec)
    env = VTG(1, seed=1, auto_reset=False, levels=[product_level(lvt)], render=False)
    env.set_draw_tape([rec["tape"]])
    env.reset()                                   # the reference constructor's draws
    obs = env.reset()                             # TreasureGame.reset()
    assert env.snapshot(0) == norm_snap(rec["init"]["snap"])
    np.testing.assert_array_equal(obs[0].cpu().numpy(), np.asarray(rec["init"]["obs"], dtype=np.float32))
    a = torch.zeros(1, dtype=torch.int32, device="cuda")
    for t, st in enumerate(rec["steps"]):
        assert env.available_mask[0].tolist() == st["mask"], t
        if "restore" in st:                         # impl:447-481 init_with_state, quirks included
            env.init_with_state([st["restore"]])
            state = env.get_state()
            assert env.snapshot(0, state) == norm_snap(st["snap"]), t
            assert int(state["misc"][0, 3]) == st["draws"], t
            continue
        a[0] = st["a"]
        obs, rew, done, info = env.step(a, want_available=True)
        state = env.get_state()
        assert bool(info["ran"][0]) == (st["r"] is not None), t
        assert int(rew[0]) == (st["r"] or 0), t
        assert bool(info["terminated"][0]) == st["done"], t
        np.testing.assert_array_equal(obs[0].cpu().numpy(), np.asarray(st["obs"], dtype=np.float32), err_msg=str(t))
        assert env.snapshot(0, state) == norm_snap(st["snap"]), t          # incl. float64 angles, bag order
        assert int(state["misc"][0, 3]) == st["draws"], t                   # draws consumed
        if t + 1 < len(rec["steps"]) and "restore" not in rec["steps"][t + 1]:  # avail-after-step == next mask
            nxt = rec["steps"][t + 1]["mask"]
            assert [(int(info["available"][0]) >> k) & 1 for k in range(9)] == nxt, t
    env.close()


@pytest.mark.parametrize("name,n,steps,max_steps", [("default", 4096, 250, 100), ("twin", 2048, 250, 60),
                                                    ("mirror", 1024, 150, 0), ("wide", 1536, 200, 40)])
def test_cuda_matches_c_oracle_philox(VTG, name, n, steps, max_steps):
    """BASELINE config 2 shape: batched envs, uniform-random actions, time-limit auto-reset."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import gen_golden
    lvt = gen_golden.variant_levels()[name]
    seed = 0xC0FFEE + n
    env = VTG(n, seed=seed, max_episode_steps=max_steps, auto_reset=True, levels=[product_level(lvt)], render=False)
    cb = c_oracle.CBatch(c_oracle.CLevel(lvt), n, first_env_id=0, seed=seed, max_episode_steps=max_steps, auto_reset=True)
    cb.reset()
    assert_state_equal(env, cb, "after construction")
    g = torch.Generator().manual_seed(n)
    for t in range(steps):
        if t % 3 == 2:       # runnable actions two steps out of three keep the episodes moving
            a = torch.randint(0, 9, (n,), generator=g, dtype=torch.int32)
        else:
            m = torch.from_numpy(cb.mask().astype(np.float32)) + 1e-6
            a = torch.multinomial(m, 1, generator=g).squeeze(1).to(torch.int32)
        out = env.step_raw(a.cuda())
        ref = cb.step(a.numpy())
        assert_step_equal(out, ref, "step %d" % t)
        if t % 25 == 24:
            assert_state_equal(env, cb, "step %d" % t)
            np.testing.assert_array_equal(env.available_mask.cpu().numpy(), cb.mask())
    assert_state_equal(env, cb, "final")
    st = env.stats()
    assert list(st.values()) == cb.stats().tolist()
    assert st["errors"] == 0 and st["gym_steps"] == n * steps
    env.close()


def test_mixed_level_batch(VTG):
    """BASELINE config 5: one batch mixing several layouts."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import gen_golden
    lv = gen_golden.variant_levels()
    names = ["default", "mirror", "twin", "altinit"]
    n, seed = 1536, 99
    rng = np.random.default_rng(5)
    ids = rng.integers(0, len(names), n).astype(np.uint8)
    env = VTG(n, seed=seed, max_episode_steps=50, auto_reset=True, levels=[product_level(lv[k]) for k in names],
              level_ids=ids, render=False)
    refs = [c_oracle.CBatch(c_oracle.CLevel(lv[k]), n, first_env_id=0, seed=seed, max_episode_steps=50, auto_reset=True)
            for k in names]
    for r in refs:
        r.reset()
    g = torch.Generator().manual_seed(1)
    for t in range(120):
        a = torch.randint(0, 9, (n,), generator=g, dtype=torch.int32)
        if t % 2:
            masks = np.stack([r.mask() for r in refs])[ids, np.arange(n)]
            a = torch.multinomial(torch.from_numpy(masks.astype(np.float32)) + 1e-6, 1, generator=g).squeeze(1).to(torch.int32)
        obs, rew, done, ran = env.step_raw(a.cuda())
        outs = [r.step(a.numpy()) for r in refs]
        for l in range(len(names)):
            sel = ids == l
            o2, r2, d2, ran2, _ = outs[l]
            np.testing.assert_array_equal(rew.cpu().numpy()[sel], r2[sel])
            np.testing.assert_array_equal(done.cpu().numpy()[sel], d2[sel])
            np.testing.assert_array_equal(ran.cpu().numpy()[sel], ran2[sel])
            od = o2.shape[1]
            np.testing.assert_array_equal(obs.cpu().numpy()[sel][:, :od], o2[sel].astype(np.float32))
            assert not obs.cpu().numpy()[sel][:, od:].any()      # padding columns are zero
    st = {k: v.cpu().numpy() for k, v in env.get_state().items()}
    for l in range(len(names)):
        sel = ids == l
        cs = refs[l].state()
        np.testing.assert_array_equal(st["pos"][sel], cs["pos"][sel])
        np.testing.assert_array_equal(st["misc"][sel], cs["misc"][sel])
        np.testing.assert_array_equal(st["angles"][sel][:, : refs[l].level.nh], cs["angles"][sel])
    env.close()


def test_sharding_is_invisible(VTG):
    """N envs on one device == the same env ids split over shards, bit-exact."""
    from gym_treasure_game_b200 import shard_range
    n, seed = 3000, 7
    whole = VTG(n, seed=seed, max_episode_steps=40, render=False)
    parts = []
    for r in range(3):
        lo, hi = shard_range(n, r, 3)
        parts.append((lo, hi, VTG(hi - lo, seed=seed, max_episode_steps=40, first_env_id=lo, render=False)))
    g = torch.Generator().manual_seed(3)
    for t in range(80):
        a = torch.randint(0, 9, (n,), generator=g, dtype=torch.int32).cuda()
        o, r, d, _ = whole.step_raw(a)
        for lo, hi, p in parts:
            o2, r2, d2, _ = p.step_raw(a[lo:hi].contiguous())
            assert torch.equal(o[lo:hi], o2) and torch.equal(r[lo:hi], r2) and torch.equal(d[lo:hi], d2)
    tot = sum(np.array(list(p.stats().values())) for _, _, p in parts)
    assert tot.tolist() == list(whole.stats().values())


def test_state_roundtrip_and_injection(VTG):
    n = 512
    env = VTG(n, seed=5, render=False)
    g = torch.Generator().manual_seed(0)
    for _ in range(30):
        env.step_raw(torch.randint(0, 9, (n,), generator=g, dtype=torch.int32).cuda())
    s0 = {k: v.clone() for k, v in env.get_state().items()}
    env2 = VTG(n, seed=5, render=False)
    env2.set_state(s0)
    s1 = env2.get_state()
    for k in s0:
        assert torch.equal(s0[k], s1[k]), k
    a = torch.randint(0, 9, (n,), generator=g, dtype=torch.int32).cuda()
    o0, r0, d0, _ = env.step_raw(a)
    o1, r1, d1, _ = env2.step_raw(a)
    assert torch.equal(o0, o1) and torch.equal(r0, r1) and torch.equal(d0, d1)


def test_host_step_matches_device_step(VTG):
    n = 2048
    e1, e2 = VTG(n, seed=11, max_episode_steps=30, render=False), VTG(n, seed=11, max_episode_steps=30, render=False)
    host = e2.make_host_buffers()
    g = torch.Generator().manual_seed(2)
    for _ in range(40):
        a = torch.randint(0, 9, (n,), generator=g, dtype=torch.int32)
        o, r, d, ran = e1.step_raw(a.cuda())
        host["actions"].copy_(a)
        e2.step_host(host)
        assert torch.equal(o.cpu(), host["obs"]) and torch.equal(r.cpu(), host["reward"])
        assert torch.equal(d.cpu(), host["done"]) and torch.equal(ran.cpu(), host["ran"])


def test_init_with_state_batch_matches_c_oracle(VTG):
    """Batched save/restore: random -99 wildcards over earlier state vectors, then keep playing (Philox mode)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import gen_golden
    lvt = gen_golden.variant_levels()["twin"]
    n, seed = 1024, 404
    env = VTG(n, seed=seed, max_episode_steps=0, auto_reset=False, levels=[product_level(lvt)], render=False)
    cb = c_oracle.CBatch(c_oracle.CLevel(lvt), n, first_env_id=0, seed=seed)
    cb.reset()
    g = torch.Generator().manual_seed(5)
    rng = np.random.default_rng(6)
    saved = None
    for t in range(90):
        if t % 15 == 14 and saved is not None:
            vec = saved.copy()
            keep_pos = rng.random(n) < 0.5
            wild = rng.random(vec.shape) < 0.4
            wild[:, :2] = keep_pos[:, None]
            vec[wild] = -99.0
            mask = (rng.random(n) < 0.7).astype(np.uint8)
            env.init_with_state(vec, mask=torch.from_numpy(mask))
            cb.init_with_state(vec, mask=mask)
            assert_state_equal(env, cb, "after init_with_state at %d" % t)
            continue
        m = torch.from_numpy(cb.mask().astype(np.float32)) + 1e-6
        a = torch.multinomial(m, 1, generator=g).squeeze(1).to(torch.int32)
        out = env.step_raw(a.cuda())
        ref = cb.step(a.numpy())
        assert_step_equal(out, ref, "step %d" % t)
        if t % 15 == 7:
            saved = ref[0].copy()                    # float64 state vectors of an earlier step
    assert_state_equal(env, cb, "final")


def test_primitive_step_matches_c_oracle(VTG):
    """tg_primitive_step: one raw action per env (impl:290-359), no option layer."""
    lvt = po.default_level()
    n, seed = 2048, 909
    env = VTG(n, seed=seed, max_episode_steps=400, auto_reset=True, render=False)
    cb = c_oracle.CBatch(c_oracle.CLevel(lvt), n, first_env_id=0, seed=seed, max_episode_steps=400, auto_reset=True)
    cb.reset()
    g = torch.Generator().manual_seed(12)
    weights = torch.tensor([1.0, 2, 3, 4, 4, 1, 1, 0.2])          # id 7 is not an action: falls through like NOP
    for t in range(900):
        a = torch.multinomial(weights.expand(n, -1), 1, generator=g).squeeze(1).to(torch.int32)
        obs, rew, done = env.primitive_step(a.cuda())
        o2, r2, d2 = cb.prim_step(a.numpy())
        np.testing.assert_array_equal(rew.cpu().numpy(), r2, err_msg=str(t))
        np.testing.assert_array_equal(done.cpu().numpy(), d2, err_msg=str(t))
        np.testing.assert_array_equal(obs.cpu().numpy(), o2.astype(np.float32), err_msg=str(t))
        if t % 150 == 149:
            assert_state_equal(env, cb, "tick %d" % t)
    assert list(env.stats().values()) == cb.stats().tolist()


def test_philox_mode_statistics(VTG):
    """Production RNG (Philox4x32-10, one 32-bit word per draw) drives the reference's transforms with
    the right probabilities: noisy() (impl:361-366), the JUMP coin (impl:318), flip() (objs:117-122) and
    the reset's handle angles (impl:99-104).  Exact law of noisy(-4) from the oracle's own code on a dense
    grid of u; tolerances are 6 sigma of the binomial / sample-mean error at n = 262 144."""
    n = 1 << 18
    env = VTG(n, seed=2024, render=False, auto_reset=False)
    full = lambda k: torch.full((n,), k, dtype=torch.int32, device="cuda")
    env.reset()
    st0 = env.get_state()
    ang = st0["angles"].cpu().numpy()
    up = st0["handles"].cpu().numpy().astype(bool)
    nh = sum(1 for ob in po.default_level().objects if ob[0] == po.K_HANDLE)
    for h in range(nh):
        a = ang[:, h]
        lo, hi = (0.85, 1.0) if up[0, h] else (0.0, 0.15)
        assert a.min() >= lo and a.max() <= hi
        assert abs(a.mean() - (lo + hi) / 2) < 6 * 0.15 / np.sqrt(12 * n)
        hist = np.histogram(a, bins=16, range=(lo, hi))[0]
        assert np.abs(hist - n / 16).max() < 6 * np.sqrt(n / 16)

    # noisy(-INCR): law from the oracle's formula on 2^16 midpoints of u
    o = po.OracleEnv(po.default_level(), lambda: 0.5)
    law = {}
    for i in range(1 << 16):
        o._draw = lambda u=(i + 0.5) / 65536.0: u
        d = o.noisy(-po.INCR)
        law[d] = law.get(d, 0) + 1.0 / 65536
    env.step_raw(full(3))                                   # down the first ladder: everyone is in a corridor
    x0 = env.get_state()["pos"][:, 0].clone()
    env.primitive_step(full(po.LEFT))
    dx = (env.get_state()["pos"][:, 0] - x0).cpu().numpy()
    assert set(np.unique(dx)) == set(law), (np.unique(dx), law)
    for d, p in law.items():
        assert abs((dx == d).mean() - p) < 6 * np.sqrt(p * (1 - p) / n), (d, p)

    # JUMP: ticker 23 with probability 3/4 (u > 0.25), else 22; the same tick then consumes one.  The
    # reference only jumps with the feet on the ground and head room: find such a state with the Python
    # oracle (a random walk until a jump option is runnable) and restore it into every env.
    import random as _random
    r = _random.Random(0)
    walker = po.OracleEnv(po.default_level(), r.random)
    for _ in range(5000):
        m = walker.mask()
        if m[7] or m[8]:
            break
        walker.gym_step(r.choice([k for k in range(9) if m[k]]))
    assert m[7] or m[8]
    env.init_with_state(np.tile(np.asarray(walker.obs(), dtype=np.float64), (n, 1)))
    env.primitive_step(full(po.JUMP))
    tk = env.get_state()["misc"][:, 1].cpu().numpy()
    assert set(np.unique(tk)) == {21, 22}
    assert abs((tk == 22).mean() - 0.75) < 6 * np.sqrt(0.75 * 0.25 / n)

    # flip(): walk to the first handle (closed-loop solver prefix: go_left reaches it), interact
    env.reset()
    env.step_raw(full(3)); env.step_raw(full(0))
    m = env.available_mask.cpu().numpy()[:, 4].astype(bool)
    assert m.mean() > 0.9, "go_left from the first ladder ends next to the first handle"
    h0 = env.get_state()["handles"].cpu().numpy()
    env.primitive_step(full(po.INTERACT))
    h1 = env.get_state()["handles"].cpu().numpy()
    flipped = (h0 != h1).any(axis=1)[m]
    assert abs(flipped.mean() - 0.8) < 6 * np.sqrt(0.8 * 0.2 / m.sum())
    env.close()


def test_draw_index_wraps_like_the_oracle():
    """The per-env draw index is 32 bits wide (RNG): envs started a few draws below 2^32 wrap
    inside the run, and CUDA and the C oracle must keep agreeing draw for draw through the wrap (outputs every
    step, full state at the end; Philox mode, auto-reset on)."""
    from gym_treasure_game_b200 import VectorTreasureGame
    n = 2048
    env = VectorTreasureGame(n, seed=99, max_episode_steps=40, auto_reset=True, render=False)
    cb = c_oracle.CBatch(c_oracle.CLevel(po.default_level()), n, seed=99, max_episode_steps=40, auto_reset=True)
    cb.reset()
    start = (0xFFFFFFFF - np.arange(n, dtype=np.int64) % 257).astype(np.uint32)      # 0 .. 256 draws before the wrap
    st = env.get_state()
    st["misc"][:, 3] = torch.from_numpy(start.view(np.int32)).to(st["misc"].device)
    env.set_state(st)
    cb.set_draws(start)
    assert_state_equal(env, cb, "after moving the draw index")
    g = torch.Generator().manual_seed(4)
    for t in range(90):
        if t % 3 == 2:                                                              # runnable options: draws get consumed
            m = torch.from_numpy(cb.mask().astype(np.float32)) + 1e-6
            a = torch.multinomial(m, 1, generator=g).squeeze(1).to(torch.int32)
        else:
            a = torch.randint(0, 9, (n,), generator=g, dtype=torch.int32)
        assert_step_equal(env.step_raw(a.cuda()), cb.step(a.numpy()), "step %d" % t)
    assert_state_equal(env, cb, "after the wrap")
    wrapped = env.get_state()["misc"][:, 3].cpu().numpy().view(np.uint32)
    assert (wrapped < 0x80000000).all(), "every env should have passed 2^32 draws"
    env.close()
