"""CPU: the CUDA device functions themselves -- csrc/tg_device.cuh: run_option_to_end with its straight-line walker /
ladder loops, the register-window drop / jump ticks and their blocked-mid-air loop, the closed-form fall, the
INTERACT tick (closure table and graph walk), compute_plan, reset_env, write_obs with its two-slot rows -- compiled for
the host by g++ (tests/hostdev/hostdev.cpp: intrinsic shims + a serial driver that mirrors what the step kernel does
for one env) and compared step by step (a) with the C oracle in Philox mode: observation, reward, done, ran, primitive
ticks, available mask, full state, error flag and draw index of every env (option steps and primitive ticks), and (b) in
parity mode with all reference-generated golden trajectories, the reference's own draws injected (init_with_state included).  The level blob is the one the product library compiles
(tg_level_create is host code), so no GPU is needed.  This is test infrastructure: nothing here is a CPU path of the
product (the library still refuses to run without a device, tests/test_library_abi.py::test_no_cpu_fallback).
Reference behaviour: treasure_game.py:91-96, _option.py:20-36, _move_options.py, _treasure_game_impl.py:290-359."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import c_oracle
import py_oracle as po
from level_fuzz import random_level, usable

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostdev", "hostdev.cpp")
OUT = os.path.join(HERE, "hostdev", "_build", "libhostdev.so")
CSRC = os.path.join(os.path.dirname(HERE), "gym_treasure_game_b200", "csrc")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
RARE_WEIGHT = np.array([1, 1, 1, 1, 1, 4, 4, 4, 4], dtype=np.float64)


def _build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("tg_device.cuh", "tg_types.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                           "-I", CUDA_INC, "-o", OUT, SRC])
    return OUT


@pytest.fixture(scope="module")
def hostdev():
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    L = C.CDLL(_build())
    L.hostdev_blob_size.restype = C.c_size_t
    L.hostdev_create.restype = C.c_void_p
    L.hostdev_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int]
    L.hostdev_destroy.argtypes = [C.c_void_p]
    L.hostdev_reset.argtypes = [C.c_void_p, C.c_void_p]
    L.hostdev_step.argtypes = [C.c_void_p] * 7
    L.hostdev_mask.argtypes = [C.c_void_p, C.c_void_p]
    L.hostdev_flags.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.hostdev_set_tape.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.hostdev_init_with_state.argtypes = [C.c_void_p, C.c_void_p]
    L.hostdev_primitive_step.argtypes = [C.c_void_p] * 5
    L.hostdev_state.argtypes = [C.c_void_p] * 10
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _state(hostdev, h, n):
    """Unpacked state of every env (the layout of tg_get_state)."""
    st = dict(pos=np.zeros((n, 2), np.int32), misc=np.zeros((n, 4), np.int32), doors=np.zeros((n, 6), np.uint8),
              handles=np.zeros((n, 4), np.uint8), bolts=np.zeros((n, 3), np.uint8), angles=np.zeros((n, 4), np.float64),
              items=np.zeros((n, 4, 2), np.int32), bag=np.zeros((n, 4), np.int32), sticky=np.zeros(n, np.uint8))
    hostdev.hostdev_state(h, *[_ptr(st[k]) for k in ("pos", "misc", "doors", "handles", "bolts", "angles", "items", "bag", "sticky")])
    return st


def _assert_state_equal(hostdev, h, cb, msg):
    """Full state against the C oracle batch, bit-exact (what tests/gpu_util.py::assert_state_equal checks on the GPU)."""
    st, cs, lv = _state(hostdev, h, cb.n), cb.state(), cb.level
    np.testing.assert_array_equal(st["pos"], cs["pos"], err_msg="pos " + msg)
    np.testing.assert_array_equal(st["misc"], cs["misc"], err_msg="facing/ticker/total_actions/draws " + msg)
    np.testing.assert_array_equal(st["doors"][:, : lv.nd], cs["doors"], err_msg="doors " + msg)
    np.testing.assert_array_equal(st["handles"][:, : lv.nh], cs["handles"], err_msg="handles " + msg)
    np.testing.assert_array_equal(st["bolts"][:, : lv.nb], cs["bolts"], err_msg="bolts " + msg)
    np.testing.assert_array_equal(st["angles"][:, : lv.nh], cs["angles"], err_msg="angles (f64 bit-exact) " + msg)
    np.testing.assert_array_equal(st["items"][:, : lv.ni], cs["items"][:, :, :2], err_msg="items " + msg)
    np.testing.assert_array_equal(st["bag"], cs["bag"], err_msg="bag " + msg)
    pt = cb.handles_pt()
    if lv.nh:
        np.testing.assert_array_equal(np.stack([(st["sticky"] >> k) & 1 for k in range(lv.nh)], axis=1), pt, err_msg="sticky " + msg)


def _snapshot(hostdev, h, info):
    """Env 0 in the dict layout of the golden records (VectorTreasureGame.snapshot)."""
    c = {k: v[0] for k, v in _state(hostdev, h, 1).items()}
    return dict(
        px=int(c["pos"][0]), py=int(c["pos"][1]), facing=int(c["misc"][0]), ticker=int(c["misc"][1]),
        doors=[int(v) for v in c["doors"][: info.n_doors]], handles_up=[int(v) for v in c["handles"][: info.n_handles]],
        angles=[float(v) for v in c["angles"][: info.n_handles]], bolts=[int(v) for v in c["bolts"][: info.n_bolts]],
        items=[(int(x), int(y), int(np.trunc(x / 48)), int(np.trunc(y / 48))) for x, y in c["items"][: info.n_items]],
        bag=[int(v) for v in c["bag"] if v >= 0], total_actions=int(c["misc"][2]),
        handles_pt=[(int(c["sticky"]) >> k) & 1 for k in range(info.n_handles)])


def _compiled(lvt):
    from gpu_util import product_level
    from gym_treasure_game_b200.vector_env import CompiledLevel
    return CompiledLevel(product_level(lvt), with_sprites=False)


def _blob(hostdev, cl):
    """The level's LevelBlob bytes: tg_level begins with its LevelBlob (csrc/tg_capi.cu).  Checked, not assumed: the two
    int16 behind the 32 x 32 tile table are the grid size."""
    blob = C.string_at(cl.handle, hostdev.hostdev_blob_size())
    cw, ch = np.frombuffer(blob, dtype=np.int16, count=2, offset=32 * 32)
    assert (int(cw), int(ch)) == (cl.info.cw, cl.info.ch), "tg_level no longer starts with its LevelBlob"
    return blob


def _closure_table(cl, n_objs):
    """The level's closure table, entry by entry through the library's debug accessor."""
    from gym_treasure_game_b200 import _lib
    L = _lib.lib()
    tab = np.zeros((n_objs * 2) << 13, dtype=np.uint32)
    ent = C.c_uint32()
    for o in range(n_objs):
        for v in (0, 1):
            base = ((o * 2 + v) << 13)
            for bits in range(1 << 13):
                _lib.check(L.tg_debug_level_closure(cl.handle, o, v, bits, C.byref(ent)))
                tab[base | bits] = ent.value
    return tab


def _levels():
    yield "default", po.default_level()
    yield "mirrored", po.mirrored_level(po.default_level())
    k = 0
    for s in range(200):
        lv = random_level(s)
        if usable(lv):
            yield "fuzz%d" % s, lv
            k += 1
            if k >= 24:
                break


def _run(hostdev, lvt, n, steps, seed, max_steps, with_closure, first_env_id=0):
    cl = _compiled(lvt)
    blob = _blob(hostdev, cl)
    info = cl.info
    tab = _closure_table(cl, info.n_objects) if with_closure else None
    h = hostdev.hostdev_create(blob, _ptr(tab) if tab is not None else None, 0 if tab is None else tab.size, n,
                               first_env_id, seed, max_steps, 1, info.frame_w, info.frame_h)
    cb = c_oracle.CBatch(c_oracle.CLevel(lvt), n, first_env_id=first_env_id, seed=seed, max_episode_steps=max_steps,
                         auto_reset=True)
    try:
        od = info.obs_dim
        obs = np.zeros((n, od), np.float32)
        rew = np.zeros(n, np.float32); done = np.zeros(n, np.uint8); ran = np.zeros(n, np.uint8); ticks = np.zeros(n, np.int32)
        mask = np.zeros((n, 9), np.uint8); err = np.zeros(n, np.uint8); draws = np.zeros(n, np.uint32)
        hostdev.hostdev_reset(h, _ptr(obs))
        o0 = cb.reset()
        np.testing.assert_array_equal(obs, o0.astype(np.float32), err_msg="obs after reset")
        rng = np.random.default_rng(seed)
        ran_total = 0
        for t in range(steps):
            m = cb.mask()
            hostdev.hostdev_mask(h, _ptr(mask))
            np.testing.assert_array_equal(mask, m, err_msg="available mask before step %d" % t)
            if t % 3 == 0:                              # the benchmark's law ...
                a = rng.integers(0, 9, n).astype(np.int32)
            else:                                       # ... and runnable options (drops and jumps four times as likely:
                p = m.astype(np.float64) * RARE_WEIGHT + 1e-9     # they are 0.04 % of the steps under the benchmark's law)
                p /= p.sum(1, keepdims=True)
                a = (p.cumsum(1) > rng.random((n, 1))).argmax(1).astype(np.int32)
            if t % 17 == 5:
                a[::7] = 11                             # invalid ids: treasure_game.py:92 raises; here: not run
            hostdev.hostdev_step(h, _ptr(a), _ptr(obs), _ptr(rew), _ptr(done), _ptr(ran), _ptr(ticks))
            o2, r2, d2, ran2, tk2 = cb.step(a)
            msg = "step %d" % t
            np.testing.assert_array_equal(ran, ran2, err_msg="ran " + msg)
            np.testing.assert_array_equal(ticks, tk2, err_msg="primitive ticks " + msg)
            np.testing.assert_array_equal(rew, r2, err_msg="reward " + msg)
            np.testing.assert_array_equal(done, d2, err_msg="done " + msg)
            np.testing.assert_array_equal(obs, o2.astype(np.float32), err_msg="obs " + msg)
            ran_total += int(ran.sum())
            if t % 50 == 49:
                _assert_state_equal(hostdev, h, cb, msg)
        _assert_state_equal(hostdev, h, cb, "final")
        st = cb.state()
        hostdev.hostdev_flags(h, _ptr(err), _ptr(draws))
        np.testing.assert_array_equal(draws, st["misc"][:, 3].view(np.uint32), err_msg="draw index")
        np.testing.assert_array_equal(err.astype(np.int64), (st["acct"][:, 2] & 1), err_msg="error flag")
        return ran_total
    finally:
        hostdev.hostdev_destroy(h)


@pytest.mark.parametrize("name,lvt", list(_levels()), ids=lambda x: x if isinstance(x, str) else "")
def test_device_functions_match_c_oracle(hostdev, name, lvt):
    """Graph-walk INTERACT (no closure table), 384 envs x 150 steps per level, time limit 40 with auto-reset."""
    ran = _run(hostdev, lvt, n=384, steps=150, seed=1000 + len(name), max_steps=40, with_closure=False, first_env_id=12345)
    assert ran > 384 * 20


def test_device_functions_with_closure_table(hostdev):
    """The shipped level with the per-level closure table behind INTERACT (what the step kernel reads), 1024 envs."""
    ran = _run(hostdev, po.default_level(), n=1024, steps=200, seed=7, max_steps=100, with_closure=True)
    assert ran > 1024 * 40


def _golden_paths():
    from conftest import golden_files
    return golden_files()


@pytest.mark.parametrize("path", _golden_paths(), ids=lambda p: os.path.basename(p)[:-8])
def test_device_functions_replay_reference_golden(hostdev, path):
    """The reference-generated golden trajectories (tools/gen_golden.py ran the unmodified reference and recorded its
    uniform draws) through the host-compiled device code in parity mode: available mask before every step, whether the
    option ran (the reference's None), reward, primitive ticks, done (treasure_game.py:95), the float32 observation, the
    full state snapshot (float64 angles, ordered bag, previously_triggered flags) and the number of draws consumed, every
    step, incl. the init_with_state calls of the "restore" trajectories (impl:447-481, quirks and all)."""
    from conftest import golden_level, load_golden, norm_snap
    rec = load_golden(path)
    lvt = golden_level(rec)
    cl = _compiled(lvt)
    info = cl.info
    blob = _blob(hostdev, cl)
    h = hostdev.hostdev_create(blob, None, 0, 1, 0, 1, 0, 0, info.frame_w, info.frame_h)
    try:
        tape = np.asarray(rec["tape"], dtype=np.float64)
        off = np.array([0, tape.size], dtype=np.int64)
        hostdev.hostdev_set_tape(h, _ptr(tape), _ptr(off))
        od = info.obs_dim
        obs = np.zeros((1, od), np.float32)
        rew = np.zeros(1, np.float32); done = np.zeros(1, np.uint8); ran = np.zeros(1, np.uint8); ticks = np.zeros(1, np.int32)
        mask = np.zeros((1, 9), np.uint8); err = np.zeros(1, np.uint8); draws = np.zeros(1, np.uint32)
        hostdev.hostdev_reset(h, None)                    # the reference constructor's draws
        hostdev.hostdev_reset(h, _ptr(obs))               # TreasureGame.reset()
        np.testing.assert_array_equal(obs[0], np.asarray(rec["init"]["obs"], dtype=np.float32))
        assert _snapshot(hostdev, h, info) == norm_snap(rec["init"]["snap"])
        a = np.zeros(1, np.int32)
        for t, st in enumerate(rec["steps"]):
            hostdev.hostdev_mask(h, _ptr(mask))
            assert mask[0].tolist() == st["mask"], t
            if "restore" in st:
                states = np.asarray([st["restore"]], dtype=np.float64)
                hostdev.hostdev_init_with_state(h, _ptr(states))
                assert _snapshot(hostdev, h, info) == norm_snap(st["snap"]), t
                hostdev.hostdev_flags(h, _ptr(err), _ptr(draws))
                assert int(draws[0]) == st["draws"], t
                continue
            a[0] = st["a"]
            hostdev.hostdev_step(h, _ptr(a), _ptr(obs), _ptr(rew), _ptr(done), _ptr(ran), _ptr(ticks))
            assert bool(ran[0]) == (st["r"] is not None), t
            assert int(rew[0]) == (st["r"] or 0), t
            if st["r"] is not None and st.get("ticks") is not None:
                assert int(ticks[0]) == st["ticks"], t
            assert bool(done[0] & 1) == st["done"], t
            np.testing.assert_array_equal(obs[0], np.asarray(st["obs"], dtype=np.float32), err_msg=str(t))
            assert _snapshot(hostdev, h, info) == norm_snap(st["snap"]), t
            hostdev.hostdev_flags(h, _ptr(err), _ptr(draws))
            assert int(draws[0]) == st["draws"] and not err[0], t
    finally:
        hostdev.hostdev_destroy(h)


def test_primitive_tick_matches_c_oracle(hostdev):
    """tick() in its general form (impl:290-359, INTERACT included) as tg_primitive_step drives it: one raw action per
    env per call, 768 envs x 900 calls, against the C oracle's prim_step; full state every 150 calls."""
    lvt = po.default_level()
    n, seed = 768, 909
    cl = _compiled(lvt)
    info = cl.info
    blob = _blob(hostdev, cl)
    h = hostdev.hostdev_create(blob, None, 0, n, 0, seed, 400, 1, info.frame_w, info.frame_h)
    cb = c_oracle.CBatch(c_oracle.CLevel(lvt), n, first_env_id=0, seed=seed, max_episode_steps=400, auto_reset=True)
    try:
        obs = np.zeros((n, info.obs_dim), np.float32)
        rew = np.zeros(n, np.float32); done = np.zeros(n, np.uint8)
        hostdev.hostdev_reset(h, _ptr(obs))
        cb.reset()
        rng = np.random.default_rng(12)
        w = np.array([1.0, 2, 3, 4, 4, 1, 1, 0.2]); w /= w.sum()         # id 7 is not an action: falls through like NOP
        for t in range(900):
            a = rng.choice(8, size=n, p=w).astype(np.int32)
            hostdev.hostdev_primitive_step(h, _ptr(a), _ptr(obs), _ptr(rew), _ptr(done))
            o2, r2, d2 = cb.prim_step(a)
            np.testing.assert_array_equal(rew, r2, err_msg=str(t))
            np.testing.assert_array_equal(done, d2, err_msg=str(t))
            np.testing.assert_array_equal(obs, o2.astype(np.float32), err_msg=str(t))
            if t % 150 == 149:
                _assert_state_equal(hostdev, h, cb, "tick %d" % t)
    finally:
        hostdev.hostdev_destroy(h)
