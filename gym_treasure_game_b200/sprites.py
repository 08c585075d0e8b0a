"""Render assets (host side): sprite atlas and the precomposed static tile layer.

Replaces ``_TreasureGameDrawer.load_sprites`` / ``load_random_images``
(``_treasure_game_drawer.py:59-134``) and the tile loop of ``draw_domain``
(``:140-152``).  pygame is not used: the PNGs were decoded once into
``assets/sprites32.npz`` (``tools/build_sprite_atlas.py``) and the documented
pygame/SDL semantics (DESIGN.md "Render semantics") are applied here with numpy.
"""
from __future__ import annotations

import functools
import os

import numpy as np

from .level import ASSET_DIR, CELL, Level

# order of the dynamic sprites == enum tg_sprite_id in include/treasure_b200.h
DYNAMIC_SPRITES = ["door_closed", "door_open", "key", "gold", "bolt_open", "bolt_locked",
                   "handle_base", "hero_right", "hero_left"]


@functools.lru_cache(maxsize=None)
def _atlas32():
    with np.load(os.path.join(ASSET_DIR, "sprites32.npz")) as z:
        return {k: z[k].copy() for k in z.files}


def scale_nearest(img: np.ndarray, size: int = CELL) -> np.ndarray:
    """``pygame.transform.scale`` (1.9.x ``stretch``): nearest neighbour with an integer error
    accumulator; for 32 -> 48 the source index is ``floor(2 * dst / 3)`` on both axes."""
    n = img.shape[0]
    idx = (np.arange(size) * n) // size
    return img[idx][:, idx]


@functools.lru_cache(maxsize=None)
def sprite48(name: str) -> np.ndarray:
    if name == "hero_right":
        return sprite48("hero")
    if name == "hero_left":                                    # pygame.transform.flip(img, True, False), drawer.py:160
        return np.ascontiguousarray(sprite48("hero")[:, ::-1])
    return np.ascontiguousarray(scale_nearest(_atlas32()[name]))


def dynamic_atlas() -> np.ndarray:
    """(TG_NUM_SPRITES, 48, 48, 4) uint8 RGBA, in ``tg_sprite_id`` order."""
    return np.stack([sprite48(n) for n in DYNAMIC_SPRITES])


def blend_over(dst_rgb: np.ndarray, src_rgba: np.ndarray) -> None:
    """In-place blit of a per-pixel-alpha sprite onto an opaque RGB region (same shape):
    alpha 0 keeps dst, 255 copies src, else ``d + (((s - d) * a) >> 8)`` per channel."""
    a = src_rgba[..., 3:4].astype(np.int32)
    s = src_rgba[..., :3].astype(np.int32)
    d = dst_rgb.astype(np.int32)
    out = d + (((s - d) * a) >> 8)
    out = np.where(a == 255, s, out)
    out = np.where(a == 0, d, out)
    dst_rgb[...] = out.astype(np.uint8)


def compose_background(level: Level) -> np.ndarray:
    """The constant tile layer of a level: (H, W, 3) uint8 (drawer.py:137-152)."""
    H, W = level.frame_size
    img = np.zeros((H, W, 3), dtype=np.uint8)                  # screen.fill((0, 0, 0)), drawer.py:138
    for i, row in enumerate(level.tile_variants()):
        for j, (kind, var) in enumerate(row):
            if kind == "none":
                continue
            name = "ladder" if kind == "ladder" else "%s_%d" % (kind, var)
            blend_over(img[i * CELL:(i + 1) * CELL, j * CELL:(j + 1) * CELL], sprite48(name))
    return img
