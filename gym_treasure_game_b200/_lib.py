"""ctypes binding of ``libtreasure_b200.so`` (the C ABI in ``include/treasure_b200.h``).

There is no CPU implementation behind this module: if the CUDA library is
missing it raises, and every compute call fails loudly without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

from ._build import SO as _DEFAULT_SO

# TG_B200_LIB selects another build of the same library (kernel tuning experiments)
SO = os.environ.get("TG_B200_LIB") or _DEFAULT_SO

MAX_DOORS, MAX_HANDLES, MAX_BOLTS, MAX_ITEMS = 6, 4, 3, 4
MAX_OBJECTS, MAX_TRIGGERS, MAX_GRID, MAX_LEVELS = 16, 32, 26, 8
NUM_OPTIONS, NUM_SPRITES = 9, 9
DONE_TERMINATED, DONE_TRUNCATED = 1, 2


class TgObject(C.Structure):
    _fields_ = [("kind", C.c_int32), ("cx", C.c_int32), ("cy", C.c_int32), ("flag", C.c_int32)]


class TgTrigger(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("src_kind", "src_index", "src_value", "dst_kind", "dst_index", "dst_value")]


class TgLevelInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("cw", "ch", "n_doors", "n_handles", "n_bolts", "n_items", "n_objects",
                                         "n_triggers", "obs_dim", "start_cx", "start_cy", "frame_w", "frame_h",
                                         "has_sprites")]


class TgStateView(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("pos", "misc", "doors", "handles", "bolts", "angles", "items", "bag", "acct")]


class TreasureError(RuntimeError):
    pass


_lib = None

_SIGNATURES = {
    "tg_last_error": (C.c_char_p, []),
    "tg_abi_version": (C.c_int, []),
    "tg_device_count": (C.c_int, []),
    "tg_level_create": (C.c_int, [C.c_char_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                  C.POINTER(C.c_void_p)]),
    "tg_level_get_info": (C.c_int, [C.c_void_p, C.POINTER(TgLevelInfo)]),
    "tg_level_set_sprites": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "tg_level_destroy": (None, [C.c_void_p]),
    "tg_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                            C.c_uint64, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "tg_destroy": (None, [C.c_void_p]),
    "tg_num_envs": (C.c_int64, [C.c_void_p]),
    "tg_obs_dim": (C.c_int32, [C.c_void_p]),
    "tg_reset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tg_step": (C.c_int, [C.c_void_p] * 8),
    "tg_bind_obs": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tg_bind_flags": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tg_step_host": (C.c_int, [C.c_void_p] * 7),
    "tg_step_host_sparse": (C.c_int, [C.c_void_p] * 7),
    "tg_step_host_sparse_begin": (C.c_int, [C.c_void_p] * 7),
    "tg_step_host_sparse_end": (C.c_int, [C.c_void_p]),
    "tg_available_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "tg_render": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "tg_step_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32] + [C.c_void_p] * 7),
    "tg_blend": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "tg_background": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "tg_blit_alpha": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                C.c_int32, C.c_int32, C.c_void_p]),
    "tg_get_state": (C.c_int, [C.c_void_p, C.POINTER(TgStateView), C.c_void_p]),
    "tg_set_state": (C.c_int, [C.c_void_p, C.POINTER(TgStateView), C.c_void_p]),
    "tg_primitive_step": (C.c_int, [C.c_void_p] * 6),
    "tg_init_with_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tg_set_draw_tape": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tg_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "tg_stats_clear": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tg_launch_count": (C.c_int64, [C.c_void_p]),
    "tg_host_traffic": (None, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "tg_debug_level_closure": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_uint32, C.POINTER(C.c_uint32)]),
    "tg_debug_host_times": (None, [C.c_void_p, C.POINTER(C.c_double)]),
    "tg_debug_phase_buffer": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tg_debug_set_step_tile": (C.c_int, [C.c_void_p, C.c_int32]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """Load the CUDA library (never a fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise TreasureError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback." % SO)
        L = C.CDLL(SO)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.tg_abi_version() != 1:
            raise TreasureError("ABI version mismatch: library %d, binding 1" % L.tg_abi_version())
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise TreasureError("treasure_b200 error %d: %s" % (rc, lib().tg_last_error().decode()))
