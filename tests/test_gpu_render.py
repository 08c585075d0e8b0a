"""GPU render parity: frames from the CUDA renderer (through tg_render) must be
pixel-exact against the CPU restatement of the reference drawer
(oracle/render_oracle.py).  Parity versus real pygame is unpinned (see that file)."""
import numpy as np
import pytest
import torch

import c_oracle
import py_oracle as po
import render_oracle as ro
from conftest import golden_files, golden_level, load_golden
from gpu_util import product_level

pytestmark = pytest.mark.gpu


def test_frames_match_cpu_restatement_random_states():
    from gym_treasure_game_b200 import VectorTreasureGame
    n, seed = 256, 31
    lvt = po.default_level()
    env = VectorTreasureGame(n, seed=seed, max_episode_steps=0, auto_reset=False)
    cb = c_oracle.CBatch(c_oracle.CLevel(lvt), n, first_env_id=0, seed=seed)
    cb.reset()
    bg = ro.background(lvt.tiles)
    g = torch.Generator().manual_seed(4)
    checked = 0
    for t in range(60):
        m = torch.from_numpy(cb.mask().astype(np.float32)) + 1e-6
        a = torch.multinomial(m, 1, generator=g).squeeze(1).to(torch.int32)
        env.step_raw(a.cuda())
        cb.step(a.numpy())
        if t % 10 == 9:
            frames = env.render().cpu().numpy()
            assert frames.shape == (n, 624, 672, 3) and frames.dtype == np.uint8
            for i in range(t % 7, n, 37):
                want = ro.render_frame(lvt, cb.snapshot(i), bg)
                assert np.array_equal(frames[i], want), (t, i, int((frames[i] != want).sum()))
                checked += 1
    assert checked >= 30


@pytest.mark.parametrize("path", [p for p in golden_files() if "solve" in p], ids=lambda p: p.split("/")[-1][:-8])
def test_frames_along_solved_trajectories(path):
    """States with open doors, key/gold in the bag, dropped key, hero facing left, on ladders."""
    from gym_treasure_game_b200 import VectorTreasureGame
    rec = load_golden(path)
    lvt = golden_level(rec)
    env = VectorTreasureGame(1, seed=1, auto_reset=False, levels=[product_level(lvt)])
    env.set_draw_tape([rec["tape"]])
    env.reset(); env.reset()
    bg = ro.background(lvt.tiles)
    a = torch.zeros(1, dtype=torch.int32, device="cuda")
    for t, st in enumerate(rec["steps"]):
        a[0] = st["a"]
        env.step_raw(a)
        if t % 4 == 0 or t == len(rec["steps"]) - 1:
            frame = env.render()[0].cpu().numpy()
            snap = dict(st["snap"]); snap["items"] = [tuple(i) for i in snap["items"]]
            want = ro.render_frame(lvt, snap, bg)
            assert np.array_equal(frame, want), (t, int((frame != want).sum()))


def test_partial_range_and_out_buffer():
    from gym_treasure_game_b200 import VectorTreasureGame
    env = VectorTreasureGame(64, seed=3)
    full = env.render()
    out = torch.zeros((10, 624, 672, 3), dtype=torch.uint8, device="cuda")
    part = env.render(first=20, count=10, out=out)
    assert part.data_ptr() == out.data_ptr() and torch.equal(part, full[20:30])


def test_mixed_level_batch_frames():
    """Two layouts in one batch go through the per-band kernel (one background per level)."""
    from gym_treasure_game_b200 import VectorTreasureGame
    lvts = [po.default_level(), po.mirrored_level(po.default_level())]
    n, seed = 96, 17
    ids = (np.arange(n) % 2).astype(np.uint8)
    env = VectorTreasureGame(n, seed=seed, auto_reset=False, levels=[product_level(l) for l in lvts], level_ids=ids)
    refs = [c_oracle.CBatch(c_oracle.CLevel(l), n, first_env_id=0, seed=seed) for l in lvts]
    for r in refs:
        r.reset()
    bgs = [ro.background(l.tiles) for l in lvts]
    g = torch.Generator().manual_seed(9)
    for t in range(25):
        masks = np.stack([r.mask() for r in refs])[ids, np.arange(n)]
        a = torch.multinomial(torch.from_numpy(masks.astype(np.float32)) + 1e-6, 1, generator=g).squeeze(1).to(torch.int32)
        env.step_raw(a.cuda())
        for r in refs:
            r.step(a.numpy())
    frames = env.render().cpu().numpy()
    for i in range(0, n, 7):
        want = ro.render_frame(lvts[ids[i]], refs[ids[i]].snapshot(i), bgs[ids[i]])
        assert np.array_equal(frames[i], want), i


def test_other_grid_size_frames():
    """An 18 x 7 layout (864 x 336 frames): unit rows, job grid and tile layer all follow the level."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import gen_golden
    from gym_treasure_game_b200 import VectorTreasureGame
    lvt = gen_golden.variant_levels()["wide"]
    n, seed = 200, 23
    env = VectorTreasureGame(n, seed=seed, auto_reset=False, levels=[product_level(lvt)])
    cb = c_oracle.CBatch(c_oracle.CLevel(lvt), n, first_env_id=0, seed=seed)
    cb.reset()
    bg = ro.background(lvt.tiles)
    g = torch.Generator().manual_seed(2)
    for t in range(45):
        m = torch.from_numpy(cb.mask().astype(np.float32)) + 1e-6
        a = torch.multinomial(m, 1, generator=g).squeeze(1).to(torch.int32)
        env.step_raw(a.cuda())
        cb.step(a.numpy())
    frames = env.render().cpu().numpy()
    assert frames.shape == (n, 336, 864, 3)
    for i in range(0, n, 9):
        want = ro.render_frame(lvt, cb.snapshot(i), bg)
        assert np.array_equal(frames[i], want), (i, int((frames[i] != want).sum()))
