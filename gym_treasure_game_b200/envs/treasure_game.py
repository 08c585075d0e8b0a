"""Single-environment Gym surface, drop-in for the reference's
``gym_treasure_game/envs/treasure_game.py``: ``TreasureGame`` (``:54-114``) and
``ObservationWrapper`` (``:38-51``) keep their names, argument meaning, return
types and error behaviour; the work is done by the CUDA library through a
one-env ``VectorTreasureGame`` (there is no CPU implementation).
"""
from __future__ import annotations

import ctypes as C
import random

import numpy as np
import torch

from .. import _lib
from ..level import BOLT, GOLD, HANDLE, KEY, Level
from ..spaces import Box, Discrete
from ..vector_env import OPTION_NAMES, VectorTreasureGame

try:                                    # gym is optional (it is not installed in the build image)
    import gym as _gym
    _EnvBase, _WrapperBase = _gym.Env, _gym.Wrapper
except Exception:                       # pragma: no cover - exercised where gym is absent
    _gym = None

    class _EnvBase:
        pass

    class _WrapperBase:
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            return getattr(self.env, name)

        def reset(self, **kwargs):
            return self.env.reset(**kwargs)

        def step(self, action):
            return self.env.step(action)

        def render(self, mode="human"):
            return self.env.render(mode=mode)

        def close(self):
            return self.env.close()


class TreasureGame(_EnvBase):
    metadata = {"render.modes": ["human", "rgb_array"]}          # treasure_game.py:55

    def __init__(self, seed=None, device="cuda:0", level: Level = None):
        """``seed`` selects the Philox stream; by default it is drawn from Python's global
        ``random`` so that ``random.seed(k)`` before construction makes a run reproducible, as
        with the reference (which has no seeding API of its own)."""
        if seed is None:
            seed = random.getrandbits(63)
        self.level = level or Level.default()
        # constructor = first reset (4 draws in the shipped level), treasure_game.py:67-70 -> impl:31-53
        self._vec = VectorTreasureGame(1, device=device, seed=seed, auto_reset=False, levels=[self.level])
        self.option_names = list(OPTION_NAMES)                    # impl:496
        self.action_space = Discrete(len(OPTION_NAMES))          # treasure_game.py:73
        self.observation_space = Box(np.float32(0.0), np.float32(1.0), shape=(self.level.obs_dim,))   # :75
        self.viewer = None
        self.drawer = None                                        # created by the first render(), treasure_game.py:99-100
        # One env: every input and output of a step lives in pinned host memory that the kernels read and write
        # directly (unified addressing), so a step is two launches (tg_step, tg_get_state) and one stream
        # synchronisation -- no torch kernels, no copies.
        pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory()
        self._act = pin((4,), torch.int32)
        self._obs = pin((self._vec.obs_dim,), torch.float32)
        self._rew, self._done, self._ran = pin((4,), torch.float32), pin((4,), torch.uint8), pin((4,), torch.uint8)
        self._st = dict(pos=pin((1, 2), torch.int32), misc=pin((1, 4), torch.int32), doors=pin((1, _lib.MAX_DOORS), torch.uint8),
                        handles=pin((1, _lib.MAX_HANDLES), torch.uint8), bolts=pin((1, _lib.MAX_BOLTS), torch.uint8),
                        angles=pin((1, _lib.MAX_HANDLES), torch.float64), items=pin((1, _lib.MAX_ITEMS, 2), torch.int32),
                        bag=pin((1, _lib.MAX_ITEMS), torch.int32), acct=pin((1, 3), torch.int64))
        self._st_view = _lib.TgStateView(**{k: v.data_ptr() for k, v in self._st.items()})
        self._L = self._vec._L
        _lib.check(self._L.tg_bind_obs(self._vec._h, C.c_void_p(self._obs.data_ptr())))

    # -- helpers ---------------------------------------------------------------
    def _sync_state(self):
        """tg_get_state into the pinned buffers, then wait for the stream: the host sees the step's results."""
        _lib.check(self._L.tg_get_state(self._vec._h, C.byref(self._st_view), self._vec._stream()))
        torch.cuda.current_stream(self._vec.device).synchronize()

    def _state_vector(self):
        """float64 state vector exactly as impl:368-378 builds it (python list), from the pinned state buffers."""
        s = self._st
        H, W = self.level.frame_size
        pos = s["pos"][0].tolist()
        angles = s["angles"][0].tolist()
        bolts = s["bolts"][0].tolist()
        items = s["items"][0].tolist()
        v = [float(pos[0]) / W, float(pos[1]) / H]
        h = b = it = 0
        for kind, _, _, _ in self.level.objects:
            if kind == HANDLE:
                v.append(angles[h]); h += 1
            elif kind == BOLT:
                v.append(1.0 if bolts[b] else 0.0); b += 1
            elif kind in (KEY, GOLD):
                v += [float(items[it][0]) / W, float(items[it][1]) / H]; it += 1
        return v

    # -- save / restore (the reference keeps these on ``env._env``: impl:368-400, 447-481) -------
    def get_state(self):
        self._sync_state()
        return self._state_vector()

    def get_state_descriptors(self):
        return self.level.state_descriptors()

    def init_with_state(self, state):
        self._vec.init_with_state([list(state)])

    # -- gym API ------------------------------------------------------------------
    def reset(self):
        _lib.check(self._L.tg_reset(self._vec._h, None, C.c_void_p(self._obs.data_ptr()), self._vec._stream()))
        self._sync_state()
        return self._state_vector()

    @property
    def available_mask(self):
        return self._vec.available_mask[0].cpu().numpy().astype(np.int64)   # treasure_game.py:89

    def step(self, action):
        action = range(len(OPTION_NAMES))[action]     # IndexError on a bad id, like option_list[action] (:92)
        self._act[0] = action                         # pinned host memory: read by the kernel in place
        _lib.check(self._L.tg_step(self._vec._h, C.c_void_p(self._act.data_ptr()), C.c_void_p(self._obs.data_ptr()),
                                   C.c_void_p(self._rew.data_ptr()), C.c_void_p(self._done.data_ptr()),
                                   C.c_void_p(self._ran.data_ptr()), None, self._vec._stream()))
        self._sync_state()
        ran = bool(self._ran[0])
        r = int(self._rew[0]) if ran else None        # option.run() returns None when it cannot run (_option.py:22-23)
        return self._state_vector(), r, bool(int(self._done[0]) & _lib.DONE_TERMINATED), {}

    def render(self, mode="human"):
        if self.drawer is None:
            self.drawer = _TreasureGameDrawer(self)               # treasure_game.py:99-100
        rgb = self.drawer.draw_domain().cpu().numpy()
        if mode == "rgb_array":
            return rgb
        raise NotImplementedError("render('human') needs gym's SimpleImageViewer and a display; use mode='rgb_array'")

    def close(self):
        self._vec.close()


class _TreasureGameDrawer:
    """The drawer object the reference keeps on ``env.drawer`` (``_treasure_game_drawer.py:37-269``), for one
    environment: same method names; surfaces are ``(H, W, 3)`` uint8 CUDA tensors instead of pygame Surfaces."""

    def __init__(self, env: "TreasureGame"):
        self._vec = env._vec
        self.screen = None

    def draw_domain(self, show_screen=True):                     # drawer.py:136-163
        self.screen = self._vec.render("rgb_array")[0]
        return self.screen

    def draw_background_to_surface(self):                        # :165-182
        return self._vec.draw_background_to_surface()

    def draw_to_surface(self):                                   # :184-196
        return self._vec.draw_to_surface()[0]

    def blit_alpha(self, target, source, location, opacity):     # :198-205
        return self._vec.blit_alpha(target, source, location, opacity)

    def blend(self, surf, alpha_objs, alpha_player):             # :207-231
        return self._vec.blend(surf, alpha_objs, alpha_player, accumulate=True)

    def draw_to_file(self, fname):                               # :233-236
        self._vec.draw_to_file(fname)


class ObservationWrapper(_WrapperBase):
    """RGB observations (treasure_game.py:38-51): the vector state moves to info['world_state']."""

    def reset(self, **kwargs):
        self.env.reset(**kwargs)
        return self.env.render(mode="rgb_array")

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        info["world_state"] = obs
        screen = self.env.render(mode="rgb_array")
        return screen, reward, done, info
