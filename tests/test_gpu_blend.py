"""GPU parity of the drawer's analysis helpers: blend / blit_alpha /
draw_background_to_surface / draw_to_surface / draw_to_file through the C ABI, pixel-exact against the CPU
restatement in oracle/render_oracle.py (unpinned versus real pygame, see that file)."""
import os

import numpy as np
import pytest
import torch

import c_oracle
import py_oracle as po
import render_oracle as ro
from conftest import golden_files, golden_level, load_golden
from gpu_util import product_level

pytestmark = pytest.mark.gpu


def _advance(env, cb, steps, seed):
    g = torch.Generator().manual_seed(seed)
    for _ in range(steps):
        m = torch.from_numpy(cb.mask().astype(np.float32)) + 1e-6
        a = torch.multinomial(m, 1, generator=g).squeeze(1).to(torch.int32)
        env.step_raw(a.cuda())
        cb.step(a.numpy())


def test_background_and_surface():
    from gym_treasure_game_b200 import VectorTreasureGame
    lvt = po.default_level()
    env = VectorTreasureGame(4, seed=3, auto_reset=False)
    bg = ro.draw_background_to_surface(lvt.tiles)
    assert np.array_equal(env.draw_background_to_surface().cpu().numpy(), bg)
    cb = c_oracle.CBatch(c_oracle.CLevel(lvt), 4, first_env_id=0, seed=3)
    cb.reset()
    surf = env.draw_to_surface(first=1, count=2).cpu().numpy()
    for k in range(2):
        assert np.array_equal(surf[k], ro.draw_to_surface(lvt, cb.snapshot(1 + k)))


@pytest.mark.parametrize("alphas", [(0.25, 0.5), (1.0, 1.0), (128 / 255 + 1e-9, 0.1), (0.0, 0.7)])
def test_blend_per_env_matches_cpu_restatement(alphas):
    from gym_treasure_game_b200 import VectorTreasureGame
    n, seed = 64, 11
    lvt = po.default_level()
    env = VectorTreasureGame(n, seed=seed, auto_reset=False)
    cb = c_oracle.CBatch(c_oracle.CLevel(lvt), n, first_env_id=0, seed=seed)
    cb.reset()
    _advance(env, cb, 40, 5)
    bg = ro.background(lvt.tiles)
    pick = [3, 17, 42]
    for i in pick:
        surf = torch.from_numpy(bg.copy()).cuda().unsqueeze(0).contiguous()
        env.blend(surf, alphas[0], alphas[1], first=i, count=1)
        want = ro.blend(lvt, cb.snapshot(i), bg.copy(), alphas[0], alphas[1])
        got = surf[0].cpu().numpy()
        assert np.array_equal(got, want), (i, int((got != want).sum()))
    assert int(255 * alphas[0]) in (63, 255, 128, 0)         # the four SDL per-surface alpha code paths asked for


def test_blend_accumulates_a_set_of_states_in_order():
    """The use the research code makes of blend(): many states on one picture; order matters (integer blends)."""
    from gym_treasure_game_b200 import VectorTreasureGame
    n, seed = 32, 19
    lvt = po.default_level()
    env = VectorTreasureGame(n, seed=seed, auto_reset=False)
    cb = c_oracle.CBatch(c_oracle.CLevel(lvt), n, first_env_id=0, seed=seed)
    cb.reset()
    _advance(env, cb, 60, 8)
    bg = ro.background(lvt.tiles)
    surf = torch.from_numpy(bg.copy()).cuda().contiguous()
    env.blend(surf, 0.3, 0.2, first=4, count=6, accumulate=True)
    want = bg.copy()
    for i in range(4, 10):
        ro.blend(lvt, cb.snapshot(i), want, 0.3, 0.2)
    got = surf.cpu().numpy()
    assert np.array_equal(got, want), int((got != want).sum())
    # the batched per-env form equals six separate calls
    many = torch.from_numpy(np.stack([bg] * 6)).cuda().contiguous()
    env.blend(many, 0.3, 0.2, first=4, count=6)
    one = torch.from_numpy(bg.copy()).cuda().unsqueeze(0).contiguous()
    env.blend(one, 0.3, 0.2, first=6, count=1)
    assert torch.equal(many[2], one[0])


@pytest.mark.parametrize("path", [p for p in golden_files() if "twin_solve" in p or "wide_runnable" in p],
                         ids=lambda p: p.split("/")[-1][:-8])
def test_blend_other_layouts(path):
    """Layouts with several keys / golds / bolts / handles, open doors, items in the bag, a dropped key."""
    from gym_treasure_game_b200 import VectorTreasureGame
    rec = load_golden(path)
    lvt = golden_level(rec)
    env = VectorTreasureGame(1, seed=1, auto_reset=False, levels=[product_level(lvt)])
    env.set_draw_tape([rec["tape"]])
    env.reset(); env.reset()
    bg = ro.background(lvt.tiles)
    a = torch.zeros(1, dtype=torch.int32, device="cuda")
    last = len(rec["steps"]) - 1
    for t, st in enumerate(rec["steps"]):
        a[0] = st["a"]
        env.step_raw(a)
        if t in (last // 2, last):
            snap = dict(st["snap"]); snap["items"] = [tuple(i) for i in snap["items"]]
            surf = torch.from_numpy(bg.copy()).cuda().contiguous()
            env.blend(surf, 0.6, 0.9, accumulate=True)
            want = ro.blend(lvt, snap, bg.copy(), 0.6, 0.9)
            got = surf.cpu().numpy()
            assert np.array_equal(got, want), (t, int((got != want).sum()))


def test_blit_alpha_generic_sources():
    from gym_treasure_game_b200 import VectorTreasureGame
    env = VectorTreasureGame(1, seed=2, auto_reset=False)
    rng = np.random.default_rng(5)
    target = rng.integers(0, 256, (60, 90, 3), dtype=np.uint8)
    for channels, loc, op in [(4, (10, 7), 77), (3, (-5, 50), 200), (4, (70, -9), 128), (4, (0, 0), 255), (3, (3, 3), 0)]:
        src = rng.integers(0, 256, (23, 31, channels), dtype=np.uint8)
        if channels == 4:
            src[::3, ::2, 3] = 0
            src[1::3, 1::2, 3] = 255
        want = target.copy()
        ro.blit_alpha(want, src, loc, op)
        got = env.blit_alpha(torch.from_numpy(target.copy()).cuda(), torch.from_numpy(src).cuda(), loc, op).cpu().numpy()
        assert np.array_equal(got, want), (channels, loc, op, int((got != want).sum()))


def test_drawer_object_and_draw_to_file(tmp_path):
    import zlib
    from gym_treasure_game_b200.envs import TreasureGame
    env = TreasureGame(seed=4)
    frame = env.render(mode="rgb_array")
    assert env.drawer is not None and np.array_equal(env.drawer.screen.cpu().numpy(), frame)
    assert np.array_equal(env.drawer.draw_to_surface().cpu().numpy(), frame)
    f = str(tmp_path / "frame.png")
    env.drawer.draw_to_file(f)
    data = open(f, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    # decode the single IDAT chunk back (filter 0 scanlines)
    i = data.index(b"IDAT")
    n = int.from_bytes(data[i - 4:i], "big")
    raw = zlib.decompress(data[i + 4:i + 4 + n])
    h, w, _ = frame.shape
    rows = np.frombuffer(raw, dtype=np.uint8).reshape(h, 1 + 3 * w)
    assert not rows[:, 0].any() and np.array_equal(rows[:, 1:].reshape(h, w, 3), frame)
    surf = env.drawer.draw_background_to_surface()
    env.drawer.blend(surf, 0.5, 0.5)
    assert not torch.equal(surf, env.drawer.draw_background_to_surface())
    env.close()


def test_blend_and_frames_on_random_layouts():
    """Messy generated layouts (objects next to ladders and doors, several handles / bolts / keys): one rendered frame
    and one accumulated blend per layout against the CPU restatement."""
    from level_fuzz import random_level
    from test_fuzz_levels import SEEDS
    from gym_treasure_game_b200 import VectorTreasureGame
    for seed in SEEDS[3:6]:
        lvt = random_level(seed)
        n = 8
        env = VectorTreasureGame(n, seed=seed, auto_reset=False, levels=[product_level(lvt)])
        cb = c_oracle.CBatch(c_oracle.CLevel(lvt), n, first_env_id=0, seed=seed)
        cb.reset()
        _advance(env, cb, 25, seed)
        bg = ro.background(lvt.tiles)
        frame = env.render(first=5, count=1)[0].cpu().numpy()
        want = ro.render_frame(lvt, cb.snapshot(5), bg)
        assert np.array_equal(frame, want), (seed, int((frame != want).sum()))
        surf = torch.from_numpy(bg.copy()).cuda().contiguous()
        env.blend(surf, 0.45, 0.8, first=2, count=3, accumulate=True)
        want = bg.copy()
        for i in range(2, 5):
            ro.blend(lvt, cb.snapshot(i), want, 0.45, 0.8)
        got = surf.cpu().numpy()
        assert np.array_equal(got, want), (seed, int((got != want).sum()))
        env.close()
