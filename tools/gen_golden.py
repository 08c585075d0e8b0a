#!/usr/bin/env python
"""Generate ``tests/golden/*.json`` by running the UNMODIFIED reference
(``/root/reference``, dev container only) through ``oracle/ref_harness.py``.

Each golden file is one trajectory: the level (three text blobs), the seed, the
full uniform-draw tape recorded from CPython's MT through both RNG entry points
of the reference, the option ids, and after every gym step the complete
reference state (positions, facing, jump ticker, door/handle/bolt bits, handle
angles, item locations, bag order, total_actions, draws consumed), the state
vector, the reward (None when not runnable), done, and the 9-bit available mask
evaluated before the step.

Usage:  python tools/gen_golden.py        (re-creates every file deterministically)
"""
import gzip
import json
import os
import random
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import py_oracle as po          # noqa: E402  (level text helpers + solver input only)
import ref_harness as rh        # noqa: E402
import solver                   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def variant_levels():
    base = po.default_level()
    lv = {"default": base, "mirror": po.mirrored_level(base)}
    # initial object states swapped (door0 open / door1 closed / handles swapped)
    alt = po.LevelText(list(base.tiles), list(base.objects), list(base.triggers))
    alt.objects = [(k, cx, cy, (not f) if k in (po.K_DOOR, po.K_HANDLE) and (cx, cy) != (10, 8) else f)
                   for k, cx, cy, f in base.objects]
    lv["altinit"] = alt
    # more objects than the shipped level: two keys, two golds, two bolts, a
    # third handle chained to door 2, door->door trigger
    twin = po.LevelText(list(base.tiles), list(base.objects), list(base.triggers))
    twin.objects = list(base.objects) + [
        (po.K_KEY, 3, 4, False), (po.K_GOLD, 6, 11, False),
        (po.K_BOLT, 5, 11, True), (po.K_HANDLE, 6, 1, False),
        (po.K_DOOR, 6, 8, False),
    ]
    twin.triggers = list(base.triggers) + [
        (po.K_HANDLE, 2, True, po.K_DOOR, 3, True),
        (po.K_HANDLE, 2, False, po.K_DOOR, 3, False),
        (po.K_DOOR, 3, True, po.K_HANDLE, 0, False),
        (po.K_BOLT, 1, False, po.K_DOOR, 3, True),
        (po.K_BOLT, 1, False, po.K_HANDLE, 2, True),
    ]
    lv["twin"] = twin
    # a different grid size (18 x 7): exercises the data-driven level path end to end
    wide = po.LevelText.from_strings(
        "\n".join([
            "//////////////L///",
            "/                /",
            "////L/////////////",
            "/                /",
            "//////L//  ///////",
            "/                /",
            "//////////////////"]) + "\n",
        "handle 15 1 True\ndoor 8 1 True\nkey 12 3\nbolt 2 5 True\ndoor 12 5 True\ngold 15 5\n",
        "handle 0 True door 0 True\nhandle 0 False door 0 False\nbolt 0 True door 1 True\nbolt 0 False door 1 False\n")
    lv["wide"] = wide
    return lv


def record(name, level, seed, mode, nsteps):
    dom, obj, trg = po.level_to_strings(level)
    kinds = [0 if k == po.K_KEY else 1 for k, _, _, _ in level.objects
             if k in (po.K_KEY, po.K_GOLD)]
    with tempfile.TemporaryDirectory() as td:
        paths = []
        for fn, txt in (("d.txt", dom), ("o.txt", obj), ("i.txt", trg)):
            p = os.path.join(td, fn)
            with open(p, "w") as f:
                f.write(txt)
            paths.append(p)
        random.seed(seed)
        arng = random.Random(seed * 7919 + 13)
        with rh.DrawTap() as tap:
            g = rh.RefGame(*paths)          # constructor draws (like TreasureGame())
            obs0 = g.reset()                # reset draws
            rec = dict(name=name, level=dict(domain=dom, objects=obj, interactions=trg),
                       seed=seed, mode=mode, item_kinds=kinds,
                       draws_ctor_and_reset=tap.pos,
                       init=dict(snap=rh.impl_snapshot(g.env), obs=obs0),
                       steps=[])
            saved = []
            for t in range(nsteps):
                m = g.mask()
                snap = rh.impl_snapshot(g.env)
                if mode == "restore" and t % 8 == 7 and saved:
                    # impl:447-481 init_with_state from an earlier state vector with random -99 wildcards;
                    # the player position is kept or restored as a pair (a mixed pair can land inside a
                    # wall, where the reference's option loops never terminate)
                    vec = list(arng.choice(saved))
                    keep_pos = arng.random() < 0.5
                    for i in range(len(vec)):
                        if (i < 2 and keep_pos) or (i >= 2 and arng.random() < 0.4):
                            vec[i] = -99
                    g.env.init_with_state(list(vec))
                    rec["steps"].append(dict(restore=vec, mask=m, obs=g.env.get_state(), draws=tap.pos,
                                             snap=rh.impl_snapshot(g.env)))
                    continue
                if mode == "solve":
                    a = solver.choose_action(snap, m, kinds=kinds, mirrored=(name.startswith("mirror")),
                                             fallback_rng=arng)
                elif mode in ("runnable", "restore"):
                    run = [i for i in range(9) if m[i]]
                    a = arng.choice(run) if run else arng.randrange(9)
                else:
                    a = arng.randrange(9)
                st, r, d, _ = g.step(a)
                rec["steps"].append(dict(a=a, mask=m, r=r, done=d, obs=st, ticks=g.ticks_last,
                                         draws=tap.pos, snap=rh.impl_snapshot(g.env)))
                saved.append(st)
                if d and mode == "solve":
                    break
            rec["tape"] = list(tap.tape)
    fn = os.path.join(OUT, "%s_%s_s%d.json.gz" % (name, mode, seed))
    with open(fn, "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as f:
        f.write(json.dumps(rec, separators=(",", ":")).encode())
    return fn, len(rec["steps"]), len(rec["tape"])


def main():
    os.makedirs(OUT, exist_ok=True)
    lv = variant_levels()
    plan = [
        ("default", 7, "solve", 400), ("default", 11, "solve", 400), ("default", 23, "solve", 400),
        ("default", 1, "random", 150), ("default", 2, "random", 150),
        ("default", 3, "runnable", 150), ("default", 4, "runnable", 150),
        ("mirror", 5, "solve", 400), ("mirror", 6, "runnable", 150), ("mirror", 8, "random", 150),
        ("altinit", 9, "runnable", 150),
        ("twin", 10, "runnable", 200), ("twin", 12, "runnable", 200), ("twin", 14, "solve", 400),
        ("default", 31, "restore", 240), ("default", 32, "restore", 240), ("twin", 33, "restore", 240),
        ("mirror", 34, "restore", 160),
        ("wide", 41, "runnable", 300), ("wide", 42, "random", 200), ("wide", 43, "restore", 200),
    ]
    for name, seed, mode, n in plan:
        print(record(name, lv[name], seed, mode, n))


if __name__ == "__main__":
    main()
