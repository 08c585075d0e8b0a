"""TEST INFRASTRUCTURE ONLY -- live binding to the *unmodified* reference.

This module imports the reference's own Python dynamics from ``/root/reference``
(read-only, only present in the development container -- never on the GPU box)
so that the CPU restatements in this directory can be pinned against it and so
that ``tools/gen_golden.py`` can record golden trajectories.

Nothing under ``gym_treasure_game_b200/`` may import this file.

Recipe: the reference imports ``pygame`` and ``gym`` at
module import time (reference ``_treasure_game_impl.py:4``, ``treasure_game.py:5-10``,
``gym_treasure_game/__init__.py:1``), neither of which is installed.  Empty stub
modules are registered in ``sys.modules`` first; the dynamics only use stdlib
``random``/``math`` and one ``np.arange``, so they then run unmodified.

RNG tap: the reference draws through two *different* bound callables --
``random.uniform``/``random.gauss`` reach ``random._inst.random`` while
``_treasure_game_impl.py:318`` calls the module attribute ``random.random``.
Both are replaced by the same wrapper, so a recording contains every draw in
consumption order and an injected tape is consumed in the same order.
"""
from __future__ import annotations

import contextlib
import os
import random
import sys
import types

REFERENCE_ROOT = os.environ.get("TG_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_treasure_game"))


def _install_stubs() -> None:
    if "pygame" not in sys.modules:
        pg = types.ModuleType("pygame")
        pg.__path__ = []  # behave like a package so `pygame.locals` resolves
        sys.modules["pygame"] = pg
        sys.modules["pygame.locals"] = types.ModuleType("pygame.locals")
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        gym.__path__ = []

        class Env:  # noqa: D401 - stand-in for gym.Env
            pass

        class Wrapper:
            def __init__(self, env):
                self.env = env

            def reset(self, **kw):
                return self.env.reset(**kw)

            def step(self, a):
                return self.env.step(a)

            def render(self, mode="human"):
                return self.env.render(mode=mode)

        gym.Env, gym.Wrapper = Env, Wrapper
        envs = types.ModuleType("gym.envs")
        envs.__path__ = []
        reg = types.ModuleType("gym.envs.registration")
        reg.register = lambda **kw: None
        cc = types.ModuleType("gym.envs.classic_control")
        cc.__path__ = []
        rendering = types.ModuleType("gym.envs.classic_control.rendering")
        cc.rendering = rendering
        spaces = types.ModuleType("gym.spaces")

        class Discrete:
            def __init__(self, n):
                self.n = n

        class Box:
            def __init__(self, low, high, shape=None, dtype=None):
                self.low, self.high, self.shape = low, high, shape

        spaces.Discrete, spaces.Box = Discrete, Box
        gym.envs, gym.spaces = envs, spaces
        envs.registration, envs.classic_control = reg, cc
        sys.modules.update({
            "gym": gym, "gym.envs": envs, "gym.envs.registration": reg,
            "gym.envs.classic_control": cc,
            "gym.envs.classic_control.rendering": rendering,
            "gym.spaces": spaces,
        })


_REF = None


def load_reference():
    """Return a namespace with the reference's classes (imported unmodified)."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    impl = importlib.import_module(
        "gym_treasure_game.envs._treasure_game_impl._treasure_game_impl")
    tg = importlib.import_module("gym_treasure_game.envs.treasure_game")
    objs = importlib.import_module(
        "gym_treasure_game.envs._treasure_game_impl._objects")
    ns = types.SimpleNamespace(
        impl_mod=impl, tg_mod=tg, objs_mod=objs,
        Impl=impl._TreasureGameImpl, create_options=impl.create_options,
        TreasureGame=tg.TreasureGame,
        default_dir=os.path.dirname(impl.__file__),
    )
    _REF = ns
    return ns


class DrawTap:
    """Replaces both RNG entry points of the reference by one callable.

    mode 'record': draws come from CPython's global MT and are appended to
    ``self.tape``.  mode 'inject': draws are popped from a supplied tape.
    """

    def __init__(self, tape=None):
        self.inject = tape is not None
        self.tape = list(tape) if tape is not None else []
        self.pos = 0
        self._orig_inst = None
        self._orig_mod = None

    def __call__(self):
        if self.inject:
            v = self.tape[self.pos]
        else:
            v = self._orig_inst()
            self.tape.append(v)
        self.pos += 1
        return v

    def __enter__(self):
        self._orig_inst = random._inst.random
        self._orig_mod = random.random
        random._inst.random = self  # instance attribute shadows the C method
        random.random = self
        random._inst.gauss_next = None
        return self

    def __exit__(self, *exc):
        del random._inst.random
        random.random = self._orig_mod
        return False


def impl_snapshot(env) -> dict:
    """Every piece of dynamic state of a reference ``_TreasureGameImpl``."""
    ref = load_reference()
    o = ref.objs_mod
    doors = [int(d.closed) for d in env.doors]
    handles_up = [int(h.up) for h in env.handles]
    angles = [float(h.angle) for h in env.handles]
    handles_pt = [int(h.previously_triggered) for h in env.handles]
    bolts = [int(b.locked) for b in env.bolts]
    items = [(int(x.x), int(x.y), int(x.cx), int(x.cy)) for x in env.objects
             if isinstance(x, (o.key, o.goldcoin))]
    item_objs = [x for x in env.objects if isinstance(x, (o.key, o.goldcoin))]
    bag = [item_objs.index(b) for b in env.player_bag]
    return dict(px=int(env.playerx), py=int(env.playery),
                facing=int(env.facing_right), ticker=int(env.jump_ticker),
                doors=doors, handles_up=handles_up, angles=angles, bolts=bolts,
                items=items, bag=bag, total_actions=int(env.total_actions), handles_pt=handles_pt)


class RefGame:
    """The four lines of reference ``TreasureGame.step`` (treasure_game.py:91-96)
    around an arbitrary level trio, so non-default layouts can be driven too
    (the gym class hard-codes the default files, treasure_game.py:67-70)."""

    def __init__(self, domain=None, objects=None, interactions=None):
        ref = load_reference()
        d = ref.default_dir
        self.files = (domain or os.path.join(d, "domain.txt"),
                      objects or os.path.join(d, "domain-objects.txt"),
                      interactions or os.path.join(d, "domain-interactions.txt"))
        self.env = ref.Impl(*self.files)
        self.options, self.names = ref.create_options(self.env)
        self.ticks_last = 0

    def reset(self):
        ref = load_reference()
        self.env.reset_game()
        self.options, self.names = ref.create_options(self.env, None)
        return self.env.get_state()

    def mask(self):
        return [int(o.can_run()) for o in self.options]

    def step(self, action):
        before = self.env.total_actions
        r = self.options[action].run()
        self.ticks_last = self.env.total_actions - before
        st = self.env.get_state()
        done = self.env.player_got_goldcoin() and self.env.get_player_cell()[1] == 0
        return st, r, bool(done), {}


@contextlib.contextmanager
def seeded_tap(seed):
    random.seed(seed)
    with DrawTap() as tap:
        yield tap
