#!/usr/bin/env python
"""Build ``gym_treasure_game_b200/assets/sprites32.npz`` from the reference's PNG
sprites (development container only; the reference tree is read-only and does
not travel to the GPU box, so the decoded pixels are committed as an asset).

Only the sprites that ``_treasure_game_drawer.py`` actually blits are kept
(SURVEY.md Appendix D.2): 5 background / 5 wall / 5 floor variants
(drawer.py:59-76, ``range(5)``), ladder, open/closed door, key, gold, open/locked
bolt, handle base, hero (drawer.py:83-134).  Pixels are stored at their native
32x32 size as straight-alpha RGBA (what ``pygame.image.load(...).convert_alpha()``
yields; palette PNGs without tRNS become opaque); the 32->48 nearest-neighbour
``pygame.transform.scale`` is applied at load time by ``sprites.py``.

Art credits: see assets/SPRITES_ATTRIBUTION.txt (copied from the reference's
sprites/attribution.txt).
"""
import os
import shutil
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/gym_treasure_game/envs/_treasure_game_impl/sprites"
DST = os.path.join(ROOT, "gym_treasure_game_b200", "assets")

FILES = {
    "ladder": "ladder.png", "gold": "gold.png", "key": "key.png", "hero": "hero.png",
    "door_open": "open-door.png", "door_closed": "closeddoor.png",
    "bolt_open": "bolt-open.png", "bolt_locked": "bolt-locked.png", "handle_base": "handle-base.png",
}
for i in range(5):
    FILES["background_%d" % i] = "background/background_%d.png" % i
    FILES["wall_%d" % i] = "wall/wall_%d.png" % i
    FILES["floor_%d" % i] = "floor/floor-%d.png" % i


def main():
    os.makedirs(DST, exist_ok=True)
    out = {}
    for name, rel in sorted(FILES.items()):
        im = Image.open(os.path.join(SRC, rel))
        a = np.asarray(im.convert("RGBA"), dtype=np.uint8)
        assert a.shape == (32, 32, 4), (name, a.shape, im.mode)
        out[name] = a
        print("%-14s %-5s alpha: %4d transparent %4d partial" % (
            name, im.mode, int((a[..., 3] == 0).sum()), int(((a[..., 3] > 0) & (a[..., 3] < 255)).sum())))
    np.savez_compressed(os.path.join(DST, "sprites32.npz"), **out)
    shutil.copyfile(os.path.join(SRC, "attribution.txt"), os.path.join(DST, "SPRITES_ATTRIBUTION.txt"))
    print("wrote", os.path.join(DST, "sprites32.npz"), os.path.getsize(os.path.join(DST, "sprites32.npz")), "bytes")


if __name__ == "__main__":
    sys.exit(main())
