"""TEST INFRASTRUCTURE ONLY -- closed-loop scripted solver for the shipped
layout (and its horizontal mirror): key -> bolt -> gold -> back to the top row.

It is the generator of BASELINE config-5 action traces (SURVEY.md Appendix C
"Solution route" / Appendix E).  It looks only at a snapshot dict (player
cell, door bits, bolt bits, bag), so the same controller drives the reference
(``ref_harness.RefGame``), the Python oracle and the CUDA path.
"""
from __future__ import annotations

S = 48
GL, GR, UL, DL, IN, DNL, DNR, JL, JR = range(9)


def _table():
    # phase A: no key yet, bolt locked  (door0 = doors[0] at (9,1); door1 at (9,4))
    A = {
        (4, 0): DL,
        (4, 1): lambda s: GL if s["doors"][0] else GR,
        (1, 1): lambda s: IN if s["doors"][0] else GR,
        (10, 1): DL,
        (10, 4): lambda s: GR if s["doors"][1] else GL,
        (12, 4): lambda s: IN if s["doors"][1] else GL,
        (8, 4): DNL, (7, 6): JL, (6, 5): JL,
        (4, 6): GR, (5, 6): JR, (8, 6): GL,
        (4, 4): GL,
    }
    # phase B: key in bag, bolt locked
    B = {
        (1, 4): GR, (4, 4): DNR, (5, 6): GL, (4, 6): GL, (3, 6): DL,
        (3, 11): GL, (1, 11): IN, (7, 11): GL, (6, 5): JL, (7, 6): JL, (8, 6): GL,
    }
    # phase C: bolt open, no gold
    C = {
        (1, 11): GR, (3, 11): GR, (7, 11): JR, (8, 10): JR, (9, 9): JR,
        (10, 8): GR,
    }
    # phase D: gold in bag -> return to row 0
    D = {
        (12, 8): GL, (10, 8): DNL, (9, 9): DNL, (8, 10): DNL, (7, 11): GL,
        (3, 11): UL, (3, 6): GR, (5, 6): JR, (6, 5): JR, (4, 6): GR,
        (7, 6): JL, (8, 6): GL, (4, 4): DNR,
        (8, 4): GR,
        (10, 4): lambda s: GR if s["doors"][0] else UL,
        (12, 4): lambda s: IN if s["doors"][0] else GL,
        (10, 1): GL, (4, 1): UL,
    }
    return A, B, C, D


_A, _B, _C, _D = _table()
_MIRROR_ACT = {GL: GR, GR: GL, DNL: DNR, DNR: DNL, JL: JR, JR: JL}


def choose_action(snap, mask, kinds=(0, 1), cw=14, mirrored=False, fallback_rng=None):
    """snap: oracle/reference snapshot dict; kinds[i] = 0 key / 1 gold of item i;
    mask: 9 ints.  Returns an option id that is runnable."""
    px, py = snap["px"], snap["py"]
    cell = (px // S, (py + S // 2) // S)
    if mirrored:
        cell = (cw - 1 - cell[0], cell[1])
    has_key = any(kinds[i] == 0 for i in snap["bag"])
    has_gold = any(kinds[i] == 1 for i in snap["bag"])
    bolt_locked = bool(snap["bolts"][0])
    if has_gold:
        tab = _D
    elif not bolt_locked:
        tab = _C
    elif has_key:
        tab = _B
    else:
        tab = _A
    a = tab.get(cell)
    if callable(a):
        a = a(snap)
    if a is not None and mirrored:
        a = _MIRROR_ACT.get(a, a)
    if a is None or not mask[a]:
        runnable = [i for i in range(9) if mask[i]]
        a = fallback_rng.choice(runnable) if fallback_rng else runnable[0]
    return a
