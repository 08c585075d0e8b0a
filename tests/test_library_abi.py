"""CPU: the C-ABI shared library loads, exports every symbol include/treasure_b200.h declares,
validates levels on the host, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from gym_treasure_game_b200 import Level, _build, _lib
from gym_treasure_game_b200._lib import TgLevelInfo, TgObject, TgTrigger


def header_functions():
    src = open(os.path.join(ROOT, "include", "treasure_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tg_[a-z_0-9]+)\s*\(", src)))


def test_library_exists_and_exports_every_declared_symbol():
    assert os.path.exists(_build.SO), "build with __graft_entry__.build()"
    L = C.CDLL(_build.SO)
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "missing export %s" % n
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)      # the binding covers the whole header
    assert L.tg_abi_version() == 1


def _create(level, L):
    tiles = "".join(level.tiles).encode()
    objs = (TgObject * max(len(level.objects), 1))(*[TgObject(k, x, y, int(f)) for k, x, y, f in level.objects])
    trg = (TgTrigger * max(len(level.triggers), 1))(*[TgTrigger(a, b, int(c), d, e, int(f)) for a, b, c, d, e, f in level.triggers])
    h = C.c_void_p()
    rc = L.tg_level_create(tiles, level.cw, level.ch, objs, len(level.objects), trg, len(level.triggers), C.byref(h))
    return rc, h


def test_level_compiler_info_and_validation():
    L = _lib.lib()
    lv = Level.default()
    rc, h = _create(lv, L)
    assert rc == 0
    info = TgLevelInfo()
    assert L.tg_level_get_info(h, C.byref(info)) == 0
    assert (info.cw, info.ch, info.n_doors, info.n_handles, info.n_bolts, info.n_items) == (14, 13, 3, 2, 1, 2)
    assert (info.obs_dim, info.start_cx, info.start_cy, info.frame_w, info.frame_h) == (9, 4, 0, 672, 624)
    L.tg_level_destroy(h)
    # unsupported tile character
    bad = Level(tuple(r.replace("L", "X", 1) if i == 0 else r for i, r in enumerate(lv.tiles)), lv.objects, lv.triggers)
    rc, _ = _create(bad, L)
    assert rc == -1 and b"unsupported character" in L.tg_last_error()
    # too many doors
    many = Level(lv.tiles, lv.objects + tuple((0, 2 + i, 1, True) for i in range(6)), lv.triggers)
    rc, _ = _create(many, L)
    assert rc == -1 and b"doors" in L.tg_last_error()
    # a bag cell that is not a wall (impl:353 would let an item be picked up twice)
    rows = list(lv.tiles)
    rows[-1] = rows[-1][:-1] + " "
    rc, _ = _create(Level(tuple(rows), lv.objects, lv.triggers), L)
    assert rc == -1 and b"bag cell" in L.tg_last_error()
    # trigger on a missing object
    rc, _ = _create(Level(lv.tiles, lv.objects, lv.triggers + ((1, 5, True, 0, 0, True),)), L)
    assert rc == -1 and b"trigger" in L.tg_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    L = _lib.lib()
    assert L.tg_device_count() == 0
    rc, h = _create(Level.default(), L)
    arr = (C.c_void_p * 1)(h)
    env = C.c_void_p()
    rc = L.tg_create(arr, 1, None, 16, 0, 0, 0, 0, 1, C.byref(env))
    assert rc == -2 and b"no CPU fallback" in L.tg_last_error()
    from gym_treasure_game_b200 import VectorTreasureGame
    from gym_treasure_game_b200._lib import TreasureError
    with pytest.raises(TreasureError):
        VectorTreasureGame(4)
