"""Random Treasure Game layouts for differential fuzzing (tests only).

The generator produces levels the reference's constructor accepts and the product's level compiler
accepts (``tg_level_create``): walled border, all-wall bottom row (so the bag cells are walls), object
counts inside the ``TG_MAX_*`` limits, at most one door per cell.  Geometry is deliberately messy --
corridors of random length, ladders through one or more floors, gaps, pedestals, objects next to
ladders and doors -- because the interesting bugs live in the option target / can_run logic.
"""
from __future__ import annotations

import random

import py_oracle as po


def random_level(seed: int) -> "po.LevelText":
    rng = random.Random(seed)
    cw, ch = rng.randint(8, 20), rng.randint(6, 13)
    g = [["/"] * cw for _ in range(ch)]
    # corridor rows: every second row starting at 1, sometimes skipping one
    rows = []
    y = 1
    while y < ch - 1:
        rows.append(y)
        y += rng.choice([2, 2, 2, 3])
    for y in rows:
        x = 1
        while x < cw - 1:
            span = rng.randint(2, cw)
            for xx in range(x, min(x + span, cw - 1)):
                g[y][xx] = " "
            x += span + rng.choice([0, 0, 1, 2])       # walls of width 0-2 between corridor pieces
    # gaps in the floors (lets the player fall / jump down) and pedestals
    for y in rows[:-1]:
        for _ in range(rng.randint(0, 2)):
            x = rng.randint(1, cw - 2)
            if g[y][x] == " " and y + 1 < ch - 1:
                g[y + 1][x] = " "
                if rng.random() < 0.4 and x + 1 < cw - 1:
                    g[y + 1][x + 1] = " "
    # ladders: a run of 'L' from the floor row under corridor a down through to corridor b
    for a, b in zip(rows[:-1], rows[1:]):
        for _ in range(rng.randint(1, 2)):
            x = rng.randint(1, cw - 2)
            if g[a][x] == " " and g[b][x] == " ":
                for yy in range(a + 1, b):
                    g[yy][x] = "L"
                if rng.random() < 0.3:
                    g[b][x] = "L"                        # ladder continues into the lower corridor
    # the start: a ladder cell in row 0 above the first corridor (first non-wall cell, row-major)
    xs = [x for x in range(1, cw - 1) if g[1][x] == " "]
    if xs:
        g[0][rng.choice(xs)] = "L"
    open_cells = [(x, y) for y in rows for x in range(1, cw - 1) if g[y][x] == " " and g[y + 1][x] != " "]
    rng.shuffle(open_cells)
    objs = []
    kinds = ([po.K_DOOR] * rng.randint(0, 4) + [po.K_HANDLE] * rng.randint(0, 3) + [po.K_BOLT] * rng.randint(0, 2)
             + [po.K_KEY] * rng.randint(0, 2) + [po.K_GOLD] * rng.randint(0, 2))
    rng.shuffle(kinds)
    for k in kinds:
        if not open_cells:
            break
        x, y = open_cells.pop()
        objs.append((k, x, y, rng.random() < 0.5))
    count = {k: sum(1 for o in objs if o[0] == k) for k in range(5)}
    trig = []
    srcs = [(k, i) for k in (po.K_DOOR, po.K_HANDLE, po.K_BOLT) for i in range(count[k])]
    for _ in range(rng.randint(0, 10)):
        if len(srcs) < 2:
            break
        (k1, i1), (k2, i2) = rng.sample(srcs, 2)
        trig.append((k1, i1, rng.random() < 0.5, k2, i2, rng.random() < 0.5))
    return po.LevelText(["".join(r) for r in g], objs, trig)


def usable(lv: "po.LevelText") -> bool:
    """A level is worth fuzzing if the start cell exists and at least one option is runnable there."""
    try:
        env = po.OracleEnv(lv, random.Random(1).random)
    except Exception:
        return False
    return any(env.mask())
