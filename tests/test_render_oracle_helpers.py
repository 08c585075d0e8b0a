"""CPU checks of the drawer-helper restatement (oracle/render_oracle.py blend / blit_alpha): internal consistency
properties that follow from the reference code (drawer.py:198-231); no GPU."""
import numpy as np

import py_oracle as po
import render_oracle as ro


def _snap():
    return po.OracleEnv(po.default_level(), po.TapeUniform([0.5] * 64)).snapshot()


def test_blit_alpha_opaque_source_full_opacity_is_a_copy():
    rng = np.random.default_rng(0)
    target = rng.integers(0, 256, (20, 30, 3), dtype=np.uint8)
    src = rng.integers(0, 256, (8, 9, 3), dtype=np.uint8)
    out = target.copy()
    ro.blit_alpha(out, src, (5, 6), 255)
    want = target.copy(); want[6:14, 5:14] = src
    assert np.array_equal(out, want)
    out = target.copy()
    ro.blit_alpha(out, src, (5, 6), 0)                          # opacity 0: d + ((s - d) * 0 >> 8) = d
    assert np.array_equal(out, target)
    out = target.copy()
    ro.blit_alpha(out, src, (-4, 17), 128)                      # clipped; 128 = SDL's averaging special case
    reg_t, reg_s = target[17:20, 0:5].astype(int), src[0:3, 4:9].astype(int)
    assert np.array_equal(out[17:20, 0:5], ((reg_s & 0xFE) + (reg_t & 0xFE) >> 1) + (reg_s & reg_t & 1))


def test_blend_zero_opacity_only_draws_handle_bases():
    lvt = po.default_level()
    snap = _snap()
    bg = ro.background(lvt.tiles)
    out = ro.blend(lvt, snap, bg.copy(), 0.0, 0.0)
    diff = np.argwhere((out != bg).any(axis=2))
    cells = {(int(y) // 48, int(x) // 48) for y, x in diff}
    handles = {(cy, cx) for kind, cx, cy, _ in lvt.objects if kind == ro.K_HANDLE}
    assert cells and cells <= handles                            # handle bases are blitted straight onto surf (:220)


def test_blend_full_opacity_matches_draw_domain_outside_handle_cells():
    """alpha 1.0: blit_alpha copies (opacity 255), so blend() over the tile layer shows every object as draw_domain
    does -- except around the handles, whose base is drawn *before* the lever there (:212-220) and after it in
    draw_object (:257-266)."""
    lvt = po.default_level()
    snap = _snap()
    bg = ro.background(lvt.tiles)
    out = ro.blend(lvt, snap, bg.copy(), 1.0, 1.0)
    frame = ro.render_frame(lvt, snap, bg)
    mask = np.ones(frame.shape[:2], dtype=bool)
    for kind, cx, cy, _ in lvt.objects:
        if kind == ro.K_HANDLE:
            mask[max(cy * 48 - 8, 0):cy * 48 + 52, max(cx * 48 - 8, 0):cx * 48 + 56] = False
    assert np.array_equal(out[mask], frame[mask])
    assert (out != bg).any(axis=2).sum() > 3000                  # door, key, bolt, gold, hero are there
