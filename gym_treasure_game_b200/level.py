"""Level description and text front end (host side).

Replaces the three parsers of the reference's ``_TreasureGameImpl``:
``get_file_description`` (``_treasure_game_impl.py:180-202``), ``read_objects``
(``:119-166``) and ``extract_interactives`` (``:75-117``).  The reference re-reads
the three files on *every* reset (``:57-60``); here a level is parsed once into an
immutable record that ``tg_level_create`` compiles into the device blob.
"""
from __future__ import annotations

import os
import random
from dataclasses import dataclass, field
from typing import List, Tuple

DOOR, HANDLE, KEY, BOLT, GOLD = range(5)          # == TG_DOOR .. TG_GOLD in treasure_b200.h
KIND_WORDS = {"door": DOOR, "handle": HANDLE, "key": KEY, "bolt": BOLT, "gold": GOLD}
WORD_OF_KIND = {v: k for k, v in KIND_WORDS.items()}
OPEN_SPACE, WALL, LADDER = " ", "/", "L"          # _cell_types.py:7-13
CELL = 48                                         # _scale.py:8-9

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


@dataclass(frozen=True)
class Level:
    tiles: Tuple[str, ...]                                   # ch rows of cw characters
    objects: Tuple[Tuple[int, int, int, bool], ...]          # (kind, cx, cy, flag), file order
    triggers: Tuple[Tuple[int, int, bool, int, int, bool], ...]
    name: str = field(default="level", compare=False)

    # ---- geometry -------------------------------------------------------
    @property
    def cw(self) -> int:
        return len(self.tiles[0])

    @property
    def ch(self) -> int:
        return len(self.tiles)

    @property
    def frame_size(self) -> Tuple[int, int]:
        return self.ch * CELL, self.cw * CELL               # (H, W)

    @property
    def obs_dim(self) -> int:                               # impl:368-378
        return 2 + sum(1 if k in (HANDLE, BOLT) else 2 if k in (KEY, GOLD) else 0 for k, _, _, _ in self.objects)

    def state_descriptors(self) -> List[str]:               # impl:380-400
        names = ["playerx", "playery"]
        handle_no = 1
        for k, _, _, _ in self.objects:
            if k == HANDLE:
                names.append("handle%d.angle" % handle_no)
                handle_no += 1
            elif k == BOLT:
                names.append("bolt.locked")
            elif k == KEY:
                names += ["key.x", "key.y"]
            elif k == GOLD:
                names += ["goldcoin.x", "goldcoin.y"]
        return names

    # ---- parsing ----------------------------------------------------------
    @staticmethod
    def from_strings(domain: str, objects: str, interactions: str, name: str = "level") -> "Level":
        rows = [ln.strip() for ln in domain.splitlines()]    # impl:188 strips every line
        while rows and rows[-1] == "":                       # tolerate trailing blank lines
            rows.pop()
        if not rows or any(len(r) != len(rows[0]) for r in rows):
            raise ValueError("tile rows must be non-empty and of equal length after strip()")
        objs = []
        for ln in objects.splitlines():                      # impl:127-163: dispatch on line.startswith
            w = ln.split()
            for word, kind in KIND_WORDS.items():
                if ln.startswith(word):
                    objs.append((kind, int(w[1]), int(w[2]), len(w) > 3 and w[3] == "True"))
                    break
        trigs = []
        for ln in interactions.splitlines():                 # impl:90-115
            if ln.strip():
                t1, i1, b1, t2, i2, b2 = ln.split()
                for t in (t1, t2):
                    if t not in ("door", "handle", "bolt"):
                        raise ValueError("interaction on unsupported object type %r" % t)
                trigs.append((KIND_WORDS[t1], int(i1), b1 == "True", KIND_WORDS[t2], int(i2), b2 == "True"))
        return Level(tuple(rows), tuple(objs), tuple(trigs), name)

    @staticmethod
    def from_reference_files(domain_file, object_file, interaction_file, name=None) -> "Level":
        """The reference's own three-file layout (constructor args of impl:31)."""
        with open(domain_file) as a, open(object_file) as b, open(interaction_file) as c:
            return Level.from_strings(a.read(), b.read(), c.read(), name or os.path.basename(domain_file))

    @staticmethod
    def from_file(path) -> "Level":
        """Single-file ``.tglevel``: ``[tiles]`` / ``[objects]`` / ``[interactions]`` sections."""
        sec = {"tiles": [], "objects": [], "interactions": []}
        cur = None
        with open(path) as f:
            for ln in f.read().splitlines():
                if cur is None and (ln.startswith("#") or not ln.strip()):
                    continue
                if ln.strip() in ("[tiles]", "[objects]", "[interactions]"):
                    cur = ln.strip()[1:-1]
                    continue
                if cur is None:
                    raise ValueError("text before the first section in %s" % path)
                sec[cur].append(ln)
        return Level.from_strings("\n".join(sec["tiles"]), "\n".join(sec["objects"]),
                                  "\n".join(sec["interactions"]), os.path.splitext(os.path.basename(path))[0])

    @staticmethod
    def default() -> "Level":
        """The single layout the reference ships (treasure_game.py:67-70)."""
        return Level.from_file(os.path.join(ASSET_DIR, "default.tglevel"))

    def mirrored(self) -> "Level":
        cw = self.cw
        return Level(tuple(r[::-1] for r in self.tiles),
                     tuple((k, cw - 1 - cx, cy, f) for k, cx, cy, f in self.objects),
                     self.triggers, self.name + "-mirrored")

    def to_strings(self) -> Tuple[str, str, str]:
        dom = "\n".join(self.tiles) + "\n"
        ob = ""
        for k, cx, cy, f in self.objects:
            ob += "%s %d %d" % (WORD_OF_KIND[k], cx, cy)
            ob += " %s\n" % f if k in (DOOR, HANDLE, BOLT) else "\n"
        tr = "".join("%s %d %s %s %d %s\n" % (WORD_OF_KIND[a], b, c, WORD_OF_KIND[d], e, f)
                     for a, b, c, d, e, f in self.triggers)
        return dom, ob, tr

    # ---- static tile layer of the renderer ------------------------------------
    def tile_variants(self, seed: int = 12) -> List[List[Tuple[str, int]]]:
        """Which sprite each cell gets in ``draw_domain`` (``_treasure_game_drawer.py:137-152``).

        The drawer reseeds a private ``random.Random`` with 12 on every frame and calls
        ``choice`` on a 5-element list once per WALL cell (floor set when the cell above is
        not WALL, ``:145-146``) and once per OPEN cell, row-major; LADDER cells draw nothing.
        CPython's own generator is used here, so the sequence is the reference's.
        Returns rows of (set, variant) with set in {'wall','floor','background','ladder','none'}.
        """
        rng = random.Random()
        rng.seed(seed)
        five = list(range(5))
        out = []
        for i, row in enumerate(self.tiles):
            line = []
            for j, c in enumerate(row):
                if c == WALL:
                    key = "floor" if i > 0 and self.tiles[i - 1][j] != WALL else "wall"
                    line.append((key, rng.choice(five)))
                elif c == LADDER:
                    line.append(("ladder", 0))
                elif c == OPEN_SPACE:
                    line.append(("background", rng.choice(five)))
                else:
                    line.append(("none", 0))
            out.append(line)
        return out
