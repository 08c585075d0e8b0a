"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference renderer
(``_treasure_game_drawer.py:136-163`` ``draw_domain``, ``:238-269`` ``draw_object``,
``:59-134`` sprite loading, ``treasure_game.py:98-104`` ``render('rgb_array')``).

Parity status: UNPINNED against real pygame.  pygame/SDL are not installed in
the build image and the reference ships no golden frames, so this restatement
follows the pygame 1.9.6 / SDL 1.2 semantics named in the reference README
(README.md:15-19) as documented in DESIGN.md "Render semantics"; the tile-variant
table (the only RNG-dependent part) IS pinned: it is produced by CPython's own
``random.Random(12).choice`` and matches SURVEY.md Appendix D.3.

Deliberately written with per-pixel Python loops and none of the product's
numpy helpers (``gym_treasure_game_b200/sprites.py``) so that agreement between
the two -- and with the CUDA kernel -- is meaningful.  Sprite *pixels* come from
the committed asset ``gym_treasure_game_b200/assets/sprites32.npz`` (decoded
reference PNGs; data, not code).
"""
from __future__ import annotations

import math
import os
import random

import numpy as np

S = 48
_ASSET = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                      "gym_treasure_game_b200", "assets", "sprites32.npz")
K_DOOR, K_HANDLE, K_KEY, K_BOLT, K_GOLD = range(5)
_cache = {}


def _load32(name):
    if "z" not in _cache:
        with np.load(_ASSET) as z:
            _cache["z"] = {k: z[k].copy() for k in z.files}
    return _cache["z"][name]


def _scaled(name):
    """pygame.transform.scale 32->48 (1.9.x stretch(): integer error stepping), drawer.py:62-75, :86-132."""
    key = ("s", name)
    if key not in _cache:
        src = _load32(name)
        n = src.shape[0]
        dst = np.zeros((S, S, 4), dtype=np.uint8)
        # restated literally from pygame's stretch(): w_err = 2*srcw - 2*dstw, step while w_err >= 0
        rows = []
        sy, h_err = 0, 2 * n - 2 * S
        for _ in range(S):
            rows.append(sy)
            while h_err >= 0:
                sy += 1
                h_err -= 2 * S
            h_err += 2 * n
        for dy, sy in enumerate(rows):
            sx, w_err = 0, 2 * n - 2 * S
            for dx in range(S):
                dst[dy, dx] = src[sy, sx]
                while w_err >= 0:
                    sx += 1
                    w_err -= 2 * S
                w_err += 2 * n
        _cache[key] = dst
    return _cache[key]


def _blit(screen, spr, ox, oy):
    """Surface.blit of a per-pixel-alpha sprite on the opaque screen (SDL 1.2 ARGB->RGB blit)."""
    H, W, _ = screen.shape
    ox, oy = int(ox), int(oy)
    for sy in range(spr.shape[0]):
        y = oy + sy
        if y < 0 or y >= H:
            continue
        for sx in range(spr.shape[1]):
            x = ox + sx
            if x < 0 or x >= W:
                continue
            r, g, b, a = (int(v) for v in spr[sy, sx])
            if a == 0:
                continue
            if a == 255:
                screen[y, x] = (r, g, b)
            else:
                d = [int(v) for v in screen[y, x]]
                screen[y, x] = (d[0] + (((r - d[0]) * a) >> 8), d[1] + (((g - d[1]) * a) >> 8),
                                d[2] + (((b - d[2]) * a) >> 8))


def _set(screen, x, y, col):
    H, W, _ = screen.shape
    if 0 <= x < W and 0 <= y < H:
        screen[y, x] = col


def _drawline(screen, x1, y1, x2, y2, col):
    """pygame 1.9.x draw.c drawline()."""
    dx, dy = x2 - x1, y2 - y1
    sx = -1 if dx < 0 else 1
    sy = -1 if dy < 0 else 1
    dx, dy = sx * dx + 1, sy * dy + 1
    x, y, e = x1, y1, 0
    if dx >= dy:
        for _ in range(dx):
            _set(screen, x, y, col)
            x += sx
            e += dy
            if e >= dx:
                e -= dx
                y += sy
    else:
        for _ in range(dy):
            _set(screen, x, y, col)
            y += sy
            e += dx
            if e >= dy:
                e -= dy
                x += sx


def _line_width(screen, p1, p2, width, col):
    """pygame 1.9.x clip_and_draw_line_width()."""
    xinc = yinc = 0
    if abs(p1[0] - p2[0]) > abs(p1[1] - p2[1]):
        yinc = 1
    else:
        xinc = 1
    _drawline(screen, p1[0], p1[1], p2[0], p2[1], col)
    loop = 1
    while loop < width:
        k = loop // 2 + 1
        _drawline(screen, p1[0] + xinc * k, p1[1] + yinc * k, p2[0] + xinc * k, p2[1] + yinc * k, col)
        if loop + 1 < width:
            _drawline(screen, p1[0] - xinc * k, p1[1] - yinc * k, p2[0] - xinc * k, p2[1] - yinc * k, col)
        loop += 2


def _hline(screen, x1, y, x2, col):
    if x1 > x2:
        x1, x2 = x2, x1
    for x in range(x1, x2 + 1):
        _set(screen, x, y, col)


def _fill_circle(screen, x, y, rad, col):
    """pygame 1.9.x draw.circle(width=0) -> draw_fillellipse(rx = ry = radius)."""
    rx = ry = rad
    if rx == 0:
        _set(screen, x, y, col)
        return
    oj = ok = 0xFFFF
    ix, iy = 0, rx * 64
    while True:
        h, i = (ix + 8) >> 6, (iy + 8) >> 6
        j, k = (h * ry) // rx, (i * ry) // rx
        if ok != k and oj != k and k < ry:
            _hline(screen, x - h, y - k - 1, x + h - 1, col)
            _hline(screen, x - h, y + k, x + h - 1, col)
            ok = k
        if oj != j and ok != j and k != j:
            _hline(screen, x - i, y + j, x + i - 1, col)
            _hline(screen, x - i, y - j - 1, x + i - 1, col)
            oj = j
        ix = ix + iy // rx
        iy = iy - ix // rx
        if not i > h:
            break


def background(tiles):
    """Tile layer of draw_domain (drawer.py:137-152)."""
    ch, cw = len(tiles), len(tiles[0])
    screen = np.zeros((ch * S, cw * S, 3), dtype=np.uint8)       # fill((0,0,0)), :138
    rng = random.Random()
    rng.seed(12)                                                 # :55, :137
    five = [0, 1, 2, 3, 4]
    for i in range(ch):
        for j in range(cw):
            c = tiles[i][j]
            if c == "/":
                key = "floor" if (i > 0 and tiles[i - 1][j] != "/") else "wall"    # :144-146
                _blit(screen, _scaled("%s_%d" % (key, rng.choice(five))), j * S, i * S)
            elif c == "L":
                _blit(screen, _scaled("ladder"), j * S, i * S)
            elif c == " ":
                _blit(screen, _scaled("background_%d" % rng.choice(five)), j * S, i * S)
    return screen


def render_frame(level_text, snap, bg=None):
    """level_text: oracle.py_oracle.LevelText; snap: oracle snapshot dict.  Returns (H, W, 3) uint8."""
    screen = (background(level_text.tiles) if bg is None else bg).copy()
    d = h = b = it = 0
    for kind, cx, cy, _ in level_text.objects:                   # drawer.py:154-155, file order
        if kind == K_DOOR:
            _blit(screen, _scaled("door_closed" if snap["doors"][d] else "door_open"), cx * S, cy * S)
            d += 1
        elif kind in (K_KEY, K_GOLD):
            x, y = snap["items"][it][0], snap["items"][it][1]
            it += 1
            if x < 0:                                            # :240-241
                continue
            _blit(screen, _scaled("key" if kind == K_KEY else "gold"), x, y)
        elif kind == K_BOLT:
            _blit(screen, _scaled("bolt_locked" if snap["bolts"][b] else "bolt_open"), cx * S, cy * S)
            b += 1
        elif kind == K_HANDLE:                                   # :257-266
            angle = ((math.pi / 2.0) * snap["angles"][h]) + math.pi / 4.0
            h += 1
            r = S * 0.75
            start = (cx * S + S / 2, cy * S + S)
            end = (int(start[0] + (r * math.cos(angle))), int(start[1] - (r * math.sin(angle))))
            _line_width(screen, (int(start[0]), int(start[1])), end, 5, (47, 79, 79))
            _fill_circle(screen, end[0], end[1], int(S / 10), (255, 0, 0))
            _blit(screen, _scaled("handle_base"), cx * S, cy * S)
    hero = _scaled("hero")
    if not snap["facing"]:
        hero = hero[:, ::-1]                                     # transform.flip(img, True, False), :160
    _blit(screen, hero, snap["px"] - S / 2, snap["py"])          # :157-161
    return screen


# ---------------------------------------------------------------------------------------------------
# Analysis helpers of the drawer: draw_background_to_surface (drawer.py:165-182),
# draw_to_surface (:184-196), blit_alpha (:198-205), blend (:207-231).  UNPINNED like the rest of this
# file; the pygame 1.9.6 / SDL 1.2 semantics restated here are:
#   * Surface((w, h), SRCALPHA, 32) starts as transparent black; draw.line / draw.circle store the mapped
#     colour (alpha 255) without blending;
#   * blit of a per-pixel-alpha sprite onto a SRCALPHA surface = pygame's own alphablit_alpha (ALPHA_BLEND):
#     dA == 0 copies the source pixel, else d = ((d << 8) + (s - d) * sA + s) >> 8 per colour channel and
#     dA = sA + dA - sA * dA / 255;
#   * blit of a SRCALPHA surface onto an opaque one = the SDL per-pixel-alpha blit of _blit above;
#   * set_alpha(a) on an opaque surface + blit onto an opaque surface = SDL per-surface alpha:
#     a == 255 copies, a == 128 averages ((s & 0xfe) + (d & 0xfe) >> 1) + (s & d & 1), else
#     d + (((s - d) * a) >> 8) per channel.
# ---------------------------------------------------------------------------------------------------
def draw_background_to_surface(tiles):
    """drawer.py:165-182 -- the tile layer on a fresh surface."""
    return background(tiles)


def draw_to_surface(level_text, snap):
    """drawer.py:184-196 -- same pixels as draw_domain, on a fresh surface."""
    return render_frame(level_text, snap)


def _surface_alpha_px(s, d, a):
    """SDL 1.2 per-surface alpha, one channel."""
    if a == 255:
        return s
    if a == 128:
        return (((s & 0xFE) + (d & 0xFE)) >> 1) + (s & d & 1)
    return d + (((s - d) * a) >> 8)


def _pixel_alpha_px(s, d, a):
    if a == 0:
        return d
    if a == 255:
        return s
    return d + (((s - d) * a) >> 8)


def blit_alpha(target, source, location, opacity):
    """drawer.py:198-205.  target (H, W, 3) uint8 modified in place; source (h, w, 3 or 4) uint8."""
    H, W, _ = target.shape
    x0, y0 = int(location[0]), int(location[1])
    h, w = source.shape[0], source.shape[1]
    temp = np.zeros((h, w, 3), dtype=np.uint8)                   # Surface((w, h)).convert(): opaque black
    for ty in range(h):                                          # temp.blit(target, (-x, -y))
        for tx in range(w):
            sx, sy = tx + x0, ty + y0
            if 0 <= sx < W and 0 <= sy < H:
                temp[ty, tx] = target[sy, sx]
    for ty in range(h):                                          # temp.blit(source, (0, 0))
        for tx in range(w):
            px = source[ty, tx]
            a = int(px[3]) if source.shape[2] == 4 else 255
            temp[ty, tx] = [_pixel_alpha_px(int(px[c]), int(temp[ty, tx, c]), a) for c in range(3)]
    opacity = int(opacity)                                       # temp.set_alpha(opacity); target.blit(temp, location)
    for ty in range(h):
        for tx in range(w):
            sx, sy = tx + x0, ty + y0
            if 0 <= sx < W and 0 <= sy < H:
                target[sy, sx] = [_surface_alpha_px(int(temp[ty, tx, c]), int(target[sy, sx, c]), opacity) for c in range(3)]


def _alphablit(dst, spr, ox, oy):
    """pygame alphablit_alpha (ALPHA_BLEND) of an RGBA sprite onto an RGBA surface."""
    H, W, _ = dst.shape
    ox, oy = int(ox), int(oy)
    for sy in range(spr.shape[0]):
        y = oy + sy
        if y < 0 or y >= H:
            continue
        for sx in range(spr.shape[1]):
            x = ox + sx
            if x < 0 or x >= W:
                continue
            sR, sG, sB, sA = (int(v) for v in spr[sy, sx])
            dR, dG, dB, dA = (int(v) for v in dst[y, x])
            if dA:
                dR = ((dR << 8) + (sR - dR) * sA + sR) >> 8
                dG = ((dG << 8) + (sG - dG) * sA + sG) >> 8
                dB = ((dB << 8) + (sB - dB) * sA + sB) >> 8
                dA = sA + dA - ((sA * dA) // 255)
            else:
                dR, dG, dB, dA = sR, sG, sB, sA
            dst[y, x] = (dR, dG, dB, dA)


class _Opaque:
    """Adapter: the line / circle helpers above write 3-tuples; on a SRCALPHA surface the colour carries alpha 255."""

    def __init__(self, rgba):
        self.a = rgba
        self.shape = rgba.shape

    def __setitem__(self, idx, col):
        self.a[idx] = (col[0], col[1], col[2], 255)


def blend(level_text, snap, surf, alpha_objs, alpha_player):
    """drawer.py:207-231: objects (but the handle bases) at opacity alpha_objs, the hero at alpha_player, blended
    onto ``surf`` (H, W, 3) uint8 in place."""
    H, W, _ = surf.shape
    new_surf = np.zeros((H, W, 4), dtype=np.uint8)               # Surface(..., SRCALPHA, 32), :209
    d = h = b = it = 0
    for kind, cx, cy, _ in level_text.objects:                   # :211
        if kind == K_HANDLE:                                     # :212-220
            angle = ((math.pi / 2.0) * snap["angles"][h]) + math.pi / 4.0
            h += 1
            r = S * 0.75
            start = (cx * S + S / 2, cy * S + S)
            end = (int(start[0] + (r * math.cos(angle))), int(start[1] - (r * math.sin(angle))))
            canvas = _Opaque(new_surf)
            _line_width(canvas, (int(start[0]), int(start[1])), end, 5, (47, 79, 79))
            _fill_circle(canvas, end[0], end[1], int(S / 10), (255, 0, 0))
            _blit(surf, _scaled("handle_base"), cx * S, cy * S)  # straight onto surf, full opacity (:220)
        elif kind == K_DOOR:                                     # draw_object(obj, new_surf), :222
            _alphablit(new_surf, _scaled("door_closed" if snap["doors"][d] else "door_open"), cx * S, cy * S)
            d += 1
        elif kind in (K_KEY, K_GOLD):
            x, y = snap["items"][it][0], snap["items"][it][1]
            it += 1
            if x < 0:
                continue
            _alphablit(new_surf, _scaled("key" if kind == K_KEY else "gold"), x, y)
        elif kind == K_BOLT:
            _alphablit(new_surf, _scaled("bolt_locked" if snap["bolts"][b] else "bolt_open"), cx * S, cy * S)
            b += 1
    blit_alpha(surf, new_surf, (0, 0), int(255 * alpha_objs))    # :223
    hero = _scaled("hero")
    if not snap["facing"]:
        hero = hero[:, ::-1]
    blit_alpha(surf, hero, (snap["px"] - S / 2, snap["py"]), int(255 * alpha_player))   # :225-231
    return surf
