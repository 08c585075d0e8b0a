"""TEST INFRASTRUCTURE ONLY -- CPU restatement (pure Python) of the reference
Treasure Game dynamics.  Never imported by the product package; used by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs only.

Parity status: PINNED.  ``tests/test_oracle_vs_reference.py`` replays seeded
trajectories through the *unmodified* reference (``oracle/ref_harness.py``) and
through this file with the same uniform draws and compares every state field
after every gym step; ``tests/golden/*.json`` hold reference-generated
trajectories (``tools/gen_golden.py``) for machines without ``/root/reference``.

Every function cites the reference lines it restates.  Paths are relative to
``/root/reference/gym_treasure_game/envs/`` :
  impl = _treasure_game_impl/_treasure_game_impl.py
  opts = _treasure_game_impl/_move_options.py
  opt  = _treasure_game_impl/_option.py
  objs = _treasure_game_impl/_objects.py
  tg   = treasure_game.py

The structure is deliberately flat (one state record, table-driven options)
but the arithmetic is kept at the reference's granularity -- pixel probes,
104-probe ``can_go_down``, float ``near_enough`` -- so that timing it is a fair
stand-in for timing the reference itself ("CPU baseline").
"""
from __future__ import annotations

import math
import random as _random

S = 48                      # _scale.py:8-9 (xscale == yscale)
INCR = S // 10              # impl:46-47  x_incr == y_incr == 4
PW = S // 2                 # impl:49     player_width
OPEN, WALL, LADDER, DOOR = " ", "/", "L", "D"     # _cell_types.py:7-13
NOP, UP, DOWN, LEFT, RIGHT, JUMP, INTERACT = range(7)   # _actions.py:7-13
JUMP_REWARD, STEP_REWARD = -5, -1                   # impl:15-16
OPTION_NAMES = ["go_left", "go_right", "up_ladder", "down_ladder", "interact",
                "down_left", "down_right", "jump_left", "jump_right"]  # impl:495
TICK_CAP = 4096             # per option, like TG_TICK_CAP: the reference would hang; we raise instead

K_DOOR, K_HANDLE, K_KEY, K_BOLT, K_GOLD = range(5)
_KIND_BY_WORD = {"door": K_DOOR, "handle": K_HANDLE, "key": K_KEY,
                 "bolt": K_BOLT, "gold": K_GOLD}


class ReferenceWouldFail(RuntimeError):
    """The reference would raise TypeError or loop forever here."""


# --------------------------------------------------------------------------
# level text formats (impl:180-202, impl:119-166, impl:75-117)
# --------------------------------------------------------------------------
class LevelText:
    def __init__(self, tile_rows, object_rows, trigger_rows):
        self.tiles = tile_rows          # list[str]
        self.objects = object_rows      # list[(kind, cx, cy, flag)]
        self.triggers = trigger_rows    # list[(k1, i1, v1, k2, i2, v2)]

    @staticmethod
    def from_strings(domain: str, objects: str, interactions: str) -> "LevelText":
        # impl:183-192 -- readlines + str.strip on every line (blank lines kept)
        tiles = [ln.strip() for ln in domain.splitlines()]
        objs = []
        for ln in objects.splitlines():          # impl:127-163 (startswith tests)
            w = ln.split()
            for word, kind in _KIND_BY_WORD.items():
                if ln.startswith(word):
                    flag = len(w) > 3 and w[3] == "True"
                    objs.append((kind, int(w[1]), int(w[2]), flag))
                    break
        trigs = []
        for ln in interactions.splitlines():     # impl:90-115
            if ln.strip():
                t1, i1, b1, t2, i2, b2 = ln.split()
                trigs.append((_KIND_BY_WORD[t1], int(i1), b1 == "True",
                              _KIND_BY_WORD[t2], int(i2), b2 == "True"))
        return LevelText(tiles, objs, trigs)

    @staticmethod
    def from_files(domain, objects, interactions) -> "LevelText":
        with open(domain) as a, open(objects) as b, open(interactions) as c:
            return LevelText.from_strings(a.read(), b.read(), c.read())


class _Obj:
    __slots__ = ("kind", "cx", "cy", "x", "y", "val", "angle", "radius", "pt",
                 "trig")

    def __init__(self, kind, cx, cy, val):
        self.kind, self.cx, self.cy = kind, cx, cy
        self.x, self.y = cx * S, cy * S                 # objs:21-22
        self.val = val            # door.closed / handle.up / bolt.locked
        self.angle = 0.0
        self.radius = S * 0.75 if kind == K_HANDLE else S / 2   # objs:23,115
        self.pt = False           # previously_triggered, objs:32
        self.trig = ([], [])      # [False-list, True-list] of (obj, val); objs:65-71


class OracleEnv:
    """One Treasure Game instance (restates impl._TreasureGameImpl + the option
    layer + the gym wrapper's step/done)."""

    def __init__(self, level: LevelText, uniform=None):
        self.level = level
        self.u = uniform if uniform is not None else _random.Random().random
        self.ch = len(level.tiles)                       # impl:196-200
        self.cw = len(level.tiles[0])
        self.width, self.height = self.cw * S, self.ch * S
        self.draws = 0
        self.reset()

    # ---- RNG transforms ------------------------------------------------
    def _draw(self):
        self.draws += 1
        return self.u()

    def _uniform(self, a, b):            # CPython random.uniform: a + (b-a)*random()
        return a + (b - a) * self._draw()

    def _gauss_pair(self):               # CPython random.gauss, fresh pair
        x2pi = self._draw() * (2.0 * math.pi)
        g2rad = math.sqrt(-2.0 * math.log(1.0 - self._draw()))
        return math.cos(x2pi) * g2rad, math.sin(x2pi) * g2rad

    # ---- construction / reset (impl:31-73) -----------------------------
    def reset(self):
        lv = self.level
        # impl:204-216 + objs:246-253: the pixel map is cell-granular because
        # the 48 pixel rows of a cell row alias one list; keep one char row per
        # cell row and patch door cells in place.
        self.cells = [list(r) for r in lv.tiles]
        self.objects = []
        for kind, cx, cy, flag in lv.objects:            # impl:127-163, file order
            o = _Obj(kind, cx, cy, flag)
            if kind == K_HANDLE:                         # objs:111-114 (one draw)
                o.angle = (self._uniform(0.85, 1.0) if flag
                           else self._uniform(0, 0.15))
            elif kind == K_DOOR:                         # objs:226
                self._door_patch(o)
            self.objects.append(o)
        by_kind = {k: [o for o in self.objects if o.kind == k]
                   for k in (K_DOOR, K_HANDLE, K_BOLT)}  # impl:80-86
        for k1, i1, v1, k2, i2, v2 in lv.triggers:       # impl:115, objs:65-71
            by_kind[k1][i1].trig[int(v1)].append((by_kind[k2][i2], v2))
        # impl:168-178
        z0, z1 = self._gauss_pair()
        nx = int(0 + z0 * (S / 24))
        ny = int(abs(0 + z1 * (S / 36)))
        self.px, self.py = 0, 0
        found = False
        for y in range(self.ch):
            for x in range(self.cw):
                if lv.tiles[y][x] != WALL:
                    self.px, self.py = x * S + S // 2 + nx, y * S + ny
                    found = True
                    break
            if found:
                break
        self.bag = []
        self.ticker = 0
        self.facing = True
        self.total_actions = 0

    def _door_patch(self, d):            # objs:246-253
        self.cells[d.cy][d.cx] = DOOR if d.val else OPEN

    # ---- tile + collision predicates -----------------------------------
    def tile(self, x, y):                # impl:218-225
        if x >= self.width or x < 0 or y >= self.height or y < 0:
            return WALL
        return self.cells[y // S][x // S]

    def tile_cell(self, xc, yc):         # impl:227-230
        return self.tile(xc * S + S // 2, yc * S + S // 2)

    def up_clear(self):                  # impl:232-238
        for dx in (-INCR, 0, INCR):
            for dy in range(-INCR, 0):
                if self.tile(self.px + dx, self.py + dy) != OPEN:
                    return False
        return True

    def can_go_up(self):                 # impl:240-250
        if self.py <= 1:
            return False
        for dy in (-INCR, 0, S - INCR):
            for dx in (-PW // 2, PW // 2):
                if self.tile(self.px + dx, self.py + dy) == LADDER:
                    return True
        return False

    def can_go_down(self):               # impl:252-257 (0 .. 51 inclusive)
        for dy in range(0, S + INCR):
            for dx in (-PW // 2, PW // 2):
                if self.tile(self.px + dx, self.py + dy) == LADDER:
                    return True
        return False

    def _side_free(self, x):             # impl:259-281
        for dy in (INCR, S - INCR):
            if self.tile(x, self.py + dy) in (WALL, DOOR):
                return False
        return True

    def can_go_left(self):
        return self._side_free(self.px - PW // 2 - INCR)

    def can_go_right(self):
        return self._side_free(self.px + PW // 2 + INCR)

    def can_fall(self):                  # impl:283-288
        for dx in (-PW // 2 + 2, -2 + PW // 2):
            for dy in (0, S + 2):
                if self.tile(self.px + dx, self.py + dy) != OPEN:
                    return False
        return True

    def near(self, o):                   # objs:46-53, called with (px, py + 24.0)
        cx, cy = o.x + S / 2, o.y + S / 2
        d = math.pow(self.px - cx, 2) + math.pow(self.py + S / 2 - cy, 2)
        return math.sqrt(d) < o.radius

    def player_cell(self):               # impl:441-445
        return self.px // S, (self.py + S // 2) // S

    def has_kind_in_bag(self, kind):     # impl:418-428
        return any(o.kind == kind for o in self.bag)

    # ---- trigger graph (objs:73-94,145-149,175-178,231-235) ------------
    def _set_val(self, o, v):
        if o.val == v:
            return
        o.val = v
        if o.kind == K_HANDLE:
            self._wiggle(o)
        elif o.kind == K_DOOR:
            self._door_patch(o)
        self._fire(o, v)

    def _fire(self, o, v):               # objs:76-94
        o.pt = True
        for tgt, tv in o.trig[int(bool(v))]:
            if not tgt.pt:
                self._set_val(tgt, tv)
        o.pt = False

    def _wiggle(self, h):                # objs:127-131
        h.angle = self._uniform(0.85, 1.0) if h.val else self._uniform(0, 0.15)

    def _flip(self, h):                  # objs:117-122
        if self._uniform(0, 1) <= 0.8:
            self._set_val(h, not h.val)
        else:
            self._wiggle(h)

    def _drop_key(self):                 # impl:434-439, objs:34-38
        for o in self.bag:
            if o.kind == K_KEY:
                self.bag.remove(o)
                o.cx = o.cy = -1
                o.x = o.y = -S
                return

    # ---- primitive tick (impl:290-359) ----------------------------------
    def noisy(self, val):                # impl:361-366
        mid = val / 2.0
        if val < mid:
            return int(round(self._uniform(val, mid)))
        return int(round(self._uniform(mid, val)))

    def tick(self, act):
        xd = yd = 0
        self.total_actions += 1
        if act == UP:
            if self.can_go_up():
                yd = self.noisy(-INCR)
        elif act == DOWN:
            if self.can_go_down():
                yd = self.noisy(INCR)
        elif act == LEFT:
            if self.can_go_left():
                xd = self.noisy(-INCR)
                self.facing = False
        elif act == RIGHT:
            if self.can_go_right():
                xd = self.noisy(INCR)
                self.facing = True
        elif act == JUMP:
            if (not self.can_go_down()) and self.up_clear():
                self.ticker = 22
                if self._draw() > 0.25:
                    self.ticker = 23
        elif act == INTERACT:
            for o in self.objects:                       # impl:322-329
                if self.near(o):
                    if o.kind == K_HANDLE:
                        self._flip(o)
                    elif o.kind == K_BOLT:
                        if self.has_kind_in_bag(K_KEY):
                            self._set_val(o, False)      # impl:430-432, objs:172-173
                            self._drop_key()
        if self.ticker > 0:                              # impl:331-337
            if self.up_clear():
                yd = -INCR
            self.ticker -= 1
        elif self.can_fall():
            self.ticker = 0
            yd = INCR
        self.px += xd                                    # impl:339
        if self.can_fall() and yd > 0:                   # impl:341-348
            while yd > 0:
                self.py += 1
                yd -= 1
                if not self.can_fall():
                    yd = 0
        else:
            self.py += yd
        for o in self.objects:                           # impl:350-354
            if o.kind in (K_KEY, K_GOLD) and self.near(o):
                o.cx, o.cy = self.cw - 1 - len(self.bag), self.ch - 1
                o.x, o.y = o.cx * S, o.cy * S
                self.bag.append(o)
        return JUMP_REWARD if act == JUMP else STEP_REWARD

    # ---- option layer (opts + opt) --------------------------------------
    def is_object_at(self, xc, yc):      # impl:402-409
        for o in self.objects:
            if o.cx == xc and o.cy == yc and (o.kind != K_DOOR or o.val):
                return True
        return False

    def is_closed_door_at(self, xc, yc):  # impl:411-416
        return any(o.kind == K_DOOR and o.val and o.cx == xc and o.cy == yc
                   for o in self.objects)

    def _walk_target(self, pc, s):       # opts:43-67 (s=-1) / opts:115-139 (s=+1)
        xc, yc = pc[0] + s, pc[1]
        T = self.tile_cell
        while not (T(xc, yc - 1) == LADDER or T(xc, yc + 1) == LADDER
                   or T(xc + s, yc) == WALL
                   or self.is_object_at(xc, yc)
                   or self.is_closed_door_at(xc + s, yc)
                   or T(xc + s, yc + 1) == OPEN):
            xc += s
            if xc < 0:
                return None
        return xc, yc

    def _walk_can_run(self, s):          # opts:23-41 / opts:95-113
        pc = self.player_cell()
        t = self._walk_target(pc, s)
        if t is None:
            return False
        xc, yc = pc
        while (xc >= t[0]) if s < 0 else (xc <= t[0]):
            if self.tile_cell(xc, yc) != OPEN or self.tile_cell(xc, yc + 1) == OPEN:
                return False
            xc += s
        return True

    def _drop_target(self, pc, s):       # opts:211-221 / opts:406-416
        xc, yc = pc[0] + s, pc[1] + 1
        while self.tile_cell(xc, yc) == OPEN:
            yc += 1
            if yc >= self.ch:
                return None
        return xc, yc

    def _landing(self, xc, yc):          # opts:281-287 / opts:351-357
        return self.tile_cell(xc, yc) == OPEN and self.tile_cell(xc, yc + 1) == WALL

    def _jump_target(self, pc, s):       # opts:269-279 / opts:339-349
        xc, yc = pc
        if self._landing(xc + s, yc - 1):
            return xc + s, yc - 1
        if self._landing(xc + 2 * s, yc - 1):
            return xc + 2 * s, yc - 1
        return None

    def _aligned(self, t):               # opts:69-72 etc. (close_enough_*)
        if t is None:
            raise ReferenceWouldFail("target cell is None")
        return abs(t[0] * S + S / 2 - self.px) < INCR

    def can_run(self, k):
        T = self.tile_cell
        if k == 0:
            return self._walk_can_run(-1)
        if k == 1:
            return self._walk_can_run(+1)
        if k == 2:
            return self.can_go_up()                      # opts:165-166
        if k == 3:
            return self.can_go_down()                    # opts:181-182
        if k == 4:                                       # opts:446-455
            for o in self.objects:
                if self.near(o):
                    if o.kind == K_HANDLE:
                        return True
                    if o.kind == K_BOLT and self.has_kind_in_bag(K_KEY):
                        return True
            return False
        xc, yc = self.player_cell()
        if k in (5, 6):                                  # opts:199-209 / 394-404
            s = -1 if k == 5 else 1
            return T(xc + s, yc) == OPEN and T(xc + s, yc + 1) == OPEN
        s = -1 if k == 7 else 1                          # opts:254-267 / 324-337
        if T(xc, yc - 1) != OPEN or T(xc + s, yc - 1) != OPEN:
            return False
        return self._landing(xc + s, yc - 1) or self._landing(xc + 2 * s, yc - 1)

    def mask(self):                      # tg:83-89
        return [int(self.can_run(k)) for k in range(9)]

    def run_option(self, k, on_tick=None):   # opt:20-36 with the policies of opts; on_tick = drawer.draw_domain (opt:33-34)
        if not self.can_run(k):
            return None
        tot = 0
        target = None
        first = True
        done = False
        n = 0
        while not done:
            if k in (0, 1):                              # opts:74-85 / 146-157
                if first:
                    target = self._walk_target(self.player_cell(), -1 if k == 0 else 1)
                if self._aligned(target):
                    done = True
                act = LEFT if k == 0 else RIGHT
            elif k == 2:                                 # opts:168-173
                if not self.can_go_up():
                    done, act = True, NOP
                else:
                    act = UP
            elif k == 3:                                 # opts:184-189
                if not self.can_go_down():
                    done, act = True, NOP
                else:
                    act = DOWN
            elif k == 4:                                 # opts:457-460
                done, act = True, INTERACT
            elif k in (5, 6):                            # opts:231-244 / 426-439
                if first:
                    target = self._drop_target(self.player_cell(), -1 if k == 5 else 1)
                if self._aligned(target):
                    if not self.can_fall():
                        done = True
                    act = NOP
                else:
                    act = LEFT if k == 5 else RIGHT
            else:                                        # opts:297-314 / 367-384
                s = -1 if k == 7 else 1
                if first:
                    target = self._jump_target(self.player_cell(), s)
                    act = JUMP
                elif self._aligned(target):
                    if not self.can_fall():
                        done = True
                    act = NOP
                else:
                    blocked = not (self.can_go_left() if s < 0 else self.can_go_right())
                    if (not self.can_fall()) and blocked:
                        act = RIGHT if s < 0 else LEFT
                    else:
                        act = LEFT if s < 0 else RIGHT
            first = False
            tot += self.tick(act)
            n += 1
            if on_tick is not None:                      # opt:33-34: a frame after every primitive tick
                on_tick()
            if n >= TICK_CAP and not done:
                raise ReferenceWouldFail("option does not terminate")
        return tot

    # ---- gym surface (tg:78-96) ------------------------------------------
    def obs(self):                       # impl:368-378 + objs get_state
        v = [float(self.px) / self.width, float(self.py) / self.height]
        for o in self.objects:
            if o.kind == K_HANDLE:
                v.append(o.angle)
            elif o.kind == K_BOLT:
                v.append(1.0 if o.val else 0.0)
            elif o.kind in (K_KEY, K_GOLD):
                v += [float(o.x) / self.width, float(o.y) / self.height]
        return v

    def is_done(self):                   # tg:95
        return self.has_kind_in_bag(K_GOLD) and self.player_cell()[1] == 0

    def gym_step(self, a, on_tick=None):     # tg:91-96
        r = self.run_option(a, on_tick)
        return self.obs(), r, self.is_done(), {}

    # ---- save / restore (impl:380-400, impl:447-481) ----------------------
    def state_descriptors(self):
        d = ["playerx", "playery"]
        hn = 1
        for o in self.objects:
            if o.kind == K_HANDLE:
                d.append("handle%d.angle" % hn)
                hn += 1
            elif o.kind == K_BOLT:
                d.append("bolt.locked")
            elif o.kind == K_KEY:
                d += ["key.x", "key.y"]
            elif o.kind == K_GOLD:
                d += ["goldcoin.x", "goldcoin.y"]
        return d

    def init_with_state(self, state):
        """impl:447-481, quirks included: -99 keeps the current value (through a float round trip
        that can lose a pixel), every key / gold / bolt reads the FIRST slot of its name
        (``desc.index``), facing is forced right, the bag and the jump ticker are untouched,
        ``handle.set_angle`` propagates triggers (targets redraw their angle), and the handles keep
        ``previously_triggered = True`` afterwards (impl:473; the reset at :479-481 hits the env)."""
        desc = self.state_descriptors()
        state = list(state)
        self.facing = True
        st = self.obs()
        for v in range(len(st)):
            if state[v] == -99:
                state[v] = st[v]
        self.px = int(state[desc.index("playerx")] * self.width)
        self.py = int(state[desc.index("playery")] * self.height)
        hn = 1
        for o in self.objects:
            if o.kind in (K_KEY, K_GOLD):
                name = "key" if o.kind == K_KEY else "goldcoin"
                o.x = int(state[desc.index(name + ".x")] * self.width)        # objs:40-44 move_to_xy
                o.y = int(state[desc.index(name + ".y")] * self.height)
                o.cx, o.cy = int(o.x / S), int(o.y / S)
            elif o.kind == K_HANDLE:
                angle = state[desc.index("handle%d.angle" % hn)]
                old = o.val                                                   # objs:133-143 set_angle
                o.angle = angle
                o.val = not (angle <= 0.15)
                if o.val != old:
                    self._fire(o, o.val)
                o.pt = True
                hn += 1
            elif o.kind == K_BOLT:
                self._set_val(o, state[desc.index("bolt.locked")] > 0.5)

    # ---- full snapshot for differential tests -----------------------------
    def snapshot(self):
        items = [o for o in self.objects if o.kind in (K_KEY, K_GOLD)]
        return dict(
            px=self.px, py=self.py, facing=int(self.facing), ticker=self.ticker,
            doors=[int(o.val) for o in self.objects if o.kind == K_DOOR],
            handles_up=[int(o.val) for o in self.objects if o.kind == K_HANDLE],
            angles=[o.angle for o in self.objects if o.kind == K_HANDLE],
            bolts=[int(o.val) for o in self.objects if o.kind == K_BOLT],
            items=[(o.x, o.y, o.cx, o.cy) for o in items],
            bag=[items.index(o) for o in self.bag],
            total_actions=self.total_actions,
            handles_pt=[int(o.pt) for o in self.objects if o.kind == K_HANDLE])


# --------------------------------------------------------------------------
# the shipped level (reference impl/domain*.txt), restated as data so that the
# oracle can run where /root/reference does not exist (GPU box).  Checked
# against the reference files in tests/test_oracle_vs_reference.py.
# --------------------------------------------------------------------------
DEFAULT_DOMAIN = "\n".join([
    "////L/////////",
    "/          ///",
    "//////////L///",
    "/    /////L///",
    "/            /",
    "/////   //////",
    "/     /      /",
    "///L//////////",
    "/  L         /",
    "/  L      ////",
    "/  L     /////",
    "/       //////",
    "//////////////",
]) + "\n"
DEFAULT_OBJECTS = ("door 9 1 True\ndoor 9 4 False\ndoor 10 8 True\n"
                   "handle 1 1 True\nhandle 12 4 False\nkey 1 4\n"
                   "bolt 1 11 True\ngold 12 8\n")
DEFAULT_INTERACTIONS = "".join(
    "%s %d %s %s %d %s\n" % t for t in [
        ("handle", 0, True, "door", 0, True), ("handle", 0, False, "door", 0, False),
        ("handle", 0, True, "door", 1, False), ("handle", 0, False, "door", 1, True),
        ("handle", 0, True, "handle", 1, False), ("handle", 0, False, "handle", 1, True),
        ("handle", 1, False, "door", 0, True), ("handle", 1, True, "door", 0, False),
        ("handle", 1, False, "door", 1, False), ("handle", 1, True, "door", 1, True),
        ("handle", 1, False, "handle", 0, True), ("handle", 1, True, "handle", 0, False),
        ("bolt", 0, True, "door", 2, True), ("bolt", 0, False, "door", 2, False),
    ])


def default_level() -> LevelText:
    return LevelText.from_strings(DEFAULT_DOMAIN, DEFAULT_OBJECTS, DEFAULT_INTERACTIONS)


def mirrored_level(lv: LevelText) -> LevelText:
    """Horizontally mirrored layout (a cheap synthetic
    second layout that the reference constructor also accepts)."""
    cw = len(lv.tiles[0])
    return LevelText([r[::-1] for r in lv.tiles],
                     [(k, cw - 1 - cx, cy, f) for k, cx, cy, f in lv.objects],
                     list(lv.triggers))


def level_to_strings(lv: LevelText):
    words = {v: k for k, v in _KIND_BY_WORD.items()}
    dom = "\n".join(lv.tiles) + "\n"
    ob = ""
    for k, cx, cy, f in lv.objects:
        ob += "%s %d %d" % (words[k], cx, cy)
        ob += (" %s\n" % f) if k in (K_DOOR, K_HANDLE, K_BOLT) else "\n"
    tr = "".join("%s %d %s %s %d %s\n" % (words[a], b, c, words[d], e, f)
                 for a, b, c, d, e, f in lv.triggers)
    return dom, ob, tr


class TapeUniform:
    """Uniform source that replays a recorded draw tape (parity mode)."""

    def __init__(self, tape):
        self.tape, self.pos = tape, 0

    def __call__(self):
        v = self.tape[self.pos]
        self.pos += 1
        return v


def readme_loop(seed, episodes, steps_per_episode=100, level=None):
    """BASELINE config 1: reference README loop (README.md:40-48, tg:119-127)
    without render('human').  Actions come from a private generator (the
    reference uses gym's ``action_space.sample()`` which does not touch
    Python's ``random``).  Returns (gym_steps, primitive_ticks)."""
    rng = _random.Random(seed)
    env = OracleEnv(level or default_level(), rng.random)   # TreasureGame()
    act = _random.Random(seed ^ 0x5EED)
    gym_steps = ticks = 0
    for _ in range(episodes):
        env.reset()                                          # tg:78-81
        for _ in range(steps_per_episode):
            _, _, done, _ = env.gym_step(act.randrange(9))
            gym_steps += 1
            if done:
                break
        ticks += env.total_actions
    return gym_steps, ticks
