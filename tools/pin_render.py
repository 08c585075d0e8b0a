#!/usr/bin/env python
"""Pins the render restatement (oracle/render_oracle.py) against the REAL drawer of the reference.

Needs what this build image does not have: an importable ``pygame`` (any version; the reference README names 1.9.6)
and the reference tree (``TG_REFERENCE_ROOT``, default /root/reference).  Where both exist it

  1. replays every committed golden trajectory through the unmodified reference with the recorded draws injected,
  2. every fourth step draws the state with the reference's own ``_TreasureGameDrawer.draw_domain``
     (``_treasure_game_drawer.py:136-163``) and reads the frame back exactly as ``TreasureGame.render('rgb_array')``
     does (``treasure_game.py:98-104``: ``surfarray.array3d(...).swapaxes(0, 1)``),
  3. compares it with ``render_oracle.render_frame`` on the same state, reports the differing pixels per frame,
  4. with ``--write`` stores the reference frames as ``tests/golden/frames_<trajectory>.npz`` (uint8, (k, H, W, 3), plus
     the step indices), which ``tests/test_render_pinned.py`` then checks on every machine -- from then on the renderer
     is pinned and the note in ``render_oracle.py`` / DESIGN.md 3.5 can be flipped.

Exit code: 0 all frames equal, 1 differences (the summary says where: sprite blits, lever line, lever disc), 2 pygame
or the reference missing (nothing done).
"""
from __future__ import annotations

import argparse
import glob
import gzip
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--write", action="store_true", help="store the reference frames under tests/golden/")
    ap.add_argument("--every", type=int, default=4, help="draw every k-th step of a trajectory")
    args = ap.parse_args()
    os.environ.setdefault("SDL_VIDEODRIVER", "dummy")           # headless: the drawer opens a display (drawer.py:44-48)
    try:
        import pygame                                            # the REAL one, before ref_harness installs its stub
    except ImportError:
        print("pin_render: pygame is not importable here -- the renderer stays unpinned (see DESIGN.md 3.5)")
        return 2
    import numpy as np
    import ref_harness as rh
    import render_oracle as ro
    import py_oracle as po
    if not rh.reference_available():
        print("pin_render: no reference tree at %s" % rh.REFERENCE_ROOT)
        return 2
    import importlib
    ref = rh.load_reference()
    drawer_mod = importlib.import_module("gym_treasure_game.envs._treasure_game_impl._treasure_game_drawer")
    print("pin_render: pygame %s, SDL %s" % (pygame.version.ver, ".".join(str(v) for v in pygame.get_sdl_version())))
    bad = total = 0
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.json.gz"))):
        with gzip.open(path, "rt") as f:
            rec = json.load(f)
        if any("restore" in st for st in rec["steps"]):
            continue
        lv = rec["level"]
        lvt = po.LevelText.from_strings(lv["domain"], lv["objects"], lv["interactions"])
        name = os.path.basename(path)[:-8]
        with tempfile.TemporaryDirectory() as td:
            paths = []
            for fn, text in (("domain.txt", lv["domain"]), ("domain-objects.txt", lv["objects"]), ("domain-interactions.txt", lv["interactions"])):
                p = os.path.join(td, fn)
                with open(p, "w") as f:
                    f.write(text)
                paths.append(p)
            frames, steps = [], []
            with rh.DrawTap(rec["tape"]):
                g = rh.RefGame(*paths)
                g.reset()
                drawer = drawer_mod._TreasureGameDrawer(g.env)                    # treasure_game.py:99-100
                bg = ro.background(lvt.tiles)
                for t, st in enumerate(rec["steps"]):
                    g.step(st["a"])
                    if t % args.every and t != len(rec["steps"]) - 1:
                        continue
                    drawer.draw_domain()                                          # treasure_game.py:102
                    rgb = pygame.surfarray.array3d(drawer.screen).swapaxes(0, 1)  # treasure_game.py:103
                    want = np.ascontiguousarray(rgb, dtype=np.uint8)
                    got = ro.render_frame(lvt, rh.impl_snapshot(g.env), bg)
                    total += 1
                    nd = int((got != want).any(axis=2).sum())
                    if nd:
                        bad += 1
                        ys, xs = np.nonzero((got != want).any(axis=2))
                        print("  %s step %d: %d pixels differ, rows %d-%d cols %d-%d, max |d| %d" % (
                            name, t, nd, ys.min(), ys.max(), xs.min(), xs.max(), int(np.abs(got.astype(int) - want.astype(int)).max())))
                    frames.append(want); steps.append(t)
            if args.write and frames:
                out = os.path.join(ROOT, "tests", "golden", "frames_%s.npz" % name)
                np.savez_compressed(out, frames=np.stack(frames), steps=np.asarray(steps, dtype=np.int32),
                                    pygame=np.asarray(pygame.version.ver), sdl=np.asarray(pygame.get_sdl_version()))
                print("  wrote", out)
        print("%s: %d frames compared" % (name, len(steps)))
    print("pin_render: %d of %d frames differ from oracle/render_oracle.py" % (bad, total))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
