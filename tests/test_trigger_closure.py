"""CPU: the per-level trigger closure table (tg_capi.cu build_closure -> LevelBlob::closure, what the INTERACT tick reads
on the device) against the reference-pinned Python oracle's set_val + process_trigger (py_oracle._set_val / _fire =
objects.py:76-94, :145-149): for every (object, value, door / handle / bolt bits) the bits afterwards and the handles
whose angle is redrawn, in order.  No device needed (tg_level_create is host code)."""
import ctypes as C
import itertools

import pytest

import py_oracle as po
from level_fuzz import random_level, usable

KINDS = (po.K_DOOR, po.K_HANDLE, po.K_BOLT)
BASE = {po.K_DOOR: 0, po.K_HANDLE: 6, po.K_BOLT: 10}


def oracle_entry(env, objs, o_idx, v, bits):
    """Run the oracle's set_val on a fresh value assignment; returns (bits afterwards, [(handle index, value), ...])."""
    by_kind = {k: [o for o in env.objects if o.kind == k] for k in KINDS}
    for k in KINDS:
        for i, o in enumerate(by_kind[k]):
            o.val = bool((bits >> (BASE[k] + i)) & 1)
            o.pt = False
    events = []
    env._wiggle = lambda h: events.append((by_kind[po.K_HANDLE].index(h), bool(h.val)))
    env._door_patch = lambda d: None
    env._set_val(env.objects[o_idx], bool(v))
    out = 0
    for k in KINDS:
        for i, o in enumerate(by_kind[k]):
            out |= int(bool(o.val)) << (BASE[k] + i)
    return out, events


def levels():
    yield "default", po.default_level()
    n = 0
    for s in range(60):
        lv = random_level(s)
        if usable(lv) and lv.triggers:
            yield "fuzz%d" % s, lv
            n += 1
            if n >= 12:
                break


@pytest.mark.parametrize("name,lv", list(levels()), ids=lambda x: x if isinstance(x, str) else "")
def test_closure_table_matches_oracle(name, lv):
    from gpu_util import product_level
    from gym_treasure_game_b200 import _lib
    from gym_treasure_game_b200.vector_env import CompiledLevel
    L = _lib.lib()
    cl = CompiledLevel(product_level(lv), with_sprites=False)
    env = po.OracleEnv(lv, po.TapeUniform([0.5] * 64))
    counts = {k: sum(1 for o in env.objects if o.kind == k) for k in KINDS}
    valid_bits = [BASE[k] + i for k in KINDS for i in range(counts[k])]
    checked = 0
    for combo in itertools.product((0, 1), repeat=len(valid_bits)):
        bits = sum(b << p for b, p in zip(combo, valid_bits))
        for o_idx, o in enumerate(env.objects):
            if o.kind not in KINDS:
                continue
            for v in (0, 1):
                want_bits, want_ev = oracle_entry(env, env.objects, o_idx, v, bits)
                ent = C.c_uint32()
                _lib.check(L.tg_debug_level_closure(cl.handle, o_idx, v, bits, C.byref(ent)))
                ne = (ent.value >> 13) & 7
                if ne == 7:
                    assert len(want_ev) > 5, (name, o_idx, v, bits)          # only cascades with more than five redraws are left out
                    continue
                got_ev = [((ent.value >> (16 + 3 * k)) & 3, bool((ent.value >> (18 + 3 * k)) & 1)) for k in range(ne)]
                assert (ent.value & 0x1FFF, got_ev) == (want_bits, want_ev), (name, o_idx, v, bits)
                checked += 1
    assert checked > 0
