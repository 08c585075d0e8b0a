"""``VectorTreasureGame``: N Treasure Game environments stepped by one CUDA kernel.

Host-side mirror of the reference's Gym surface (``treasure_game.py:54-114``)
for a batch: ``reset`` / ``step`` / ``available_mask`` / ``render`` keep the
reference's names and argument meaning; results are PyTorch CUDA tensors that
view buffers owned by this object (they are overwritten by the next call --
clone them if they must survive).  All compute happens in
``libtreasure_b200.so``; torch only provides device memory and the stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import TgLevelInfo, TgObject, TgStateView, TgTrigger, TreasureError, check
from .level import Level
from .spaces import Box, Discrete
from . import sprites as _sprites

OPTION_NAMES = ["go_left_option", "go_right_option", "up_ladder_option", "down_ladder_option",
                "interact_option", "down_left_option", "down_right_option", "jump_left_option",
                "jump_right_option"]                      # _treasure_game_impl.py:495-496
STAT_NAMES = ["episodes", "successes", "return_sum", "episode_steps_sum", "primitive_ticks",
              "runnable_steps", "gym_steps", "errors"]


class CompiledLevel:
    """A ``Level`` compiled by ``tg_level_create`` (+ render assets)."""

    def __init__(self, level: Level, with_sprites: bool = True):
        L = _lib.lib()
        self.level = level
        tiles = "".join(level.tiles).encode("latin-1")
        objs = (TgObject * max(len(level.objects), 1))(*[TgObject(k, cx, cy, int(f)) for k, cx, cy, f in level.objects])
        trg = (TgTrigger * max(len(level.triggers), 1))(*[TgTrigger(a, b, int(c), d, e, int(f))
                                                         for a, b, c, d, e, f in level.triggers])
        h = C.c_void_p()
        check(L.tg_level_create(tiles, level.cw, level.ch, objs, len(level.objects), trg, len(level.triggers), C.byref(h)))
        self.handle = h
        self.info = TgLevelInfo()
        check(L.tg_level_get_info(self.handle, C.byref(self.info)))
        if with_sprites:
            atlas = np.ascontiguousarray(_sprites.dynamic_atlas())
            bg = np.ascontiguousarray(_sprites.compose_background(level))
            check(L.tg_level_set_sprites(self.handle, atlas.ctypes.data_as(C.c_void_p), bg.ctypes.data_as(C.c_void_p)))
            check(L.tg_level_get_info(self.handle, C.byref(self.info)))
            self.background = bg

    def __del__(self):
        try:
            h = getattr(self, "handle", None)
            if h and _lib._lib is not None:
                _lib._lib.tg_level_destroy(h)
                self.handle = None
        except Exception:      # interpreter shutdown
            pass


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class VectorTreasureGame:
    metadata = {"render.modes": ["rgb_array"]}

    def __init__(self, num_envs: int, device="cuda:0", seed: int = 0, max_episode_steps: int = 0,
                 auto_reset: bool = True, levels: Optional[Sequence[Level]] = None,
                 level_ids: Optional[Sequence[int]] = None, first_env_id: int = 0, render: bool = True):
        if not torch.cuda.is_available():
            raise TreasureError("VectorTreasureGame needs a CUDA device; there is no CPU fallback")
        self._L = _lib.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise TreasureError("device must be a CUDA device, got %s" % device)
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        self.num_envs = int(num_envs)
        self.levels = list(levels) if levels else [Level.default()]
        self._compiled = [CompiledLevel(lv, with_sprites=render) for lv in self.levels]
        arr = (C.c_void_p * len(self._compiled))(*[c.handle for c in self._compiled])
        ids = None
        if level_ids is not None:
            ids = np.ascontiguousarray(level_ids, dtype=np.uint8)
            if ids.shape != (self.num_envs,):
                raise ValueError("level_ids must have one entry per env")
        h = C.c_void_p()
        torch.cuda.init()
        with torch.cuda.device(self.device):
            check(self._L.tg_create(arr, len(self._compiled), None if ids is None else ids.ctypes.data_as(C.c_void_p),
                                    self.num_envs, int(first_env_id), dev_index, int(seed) & (2 ** 64 - 1),
                                    int(max_episode_steps), int(bool(auto_reset)), C.byref(h)))
        self._h = h
        self.level_ids = ids
        self.obs_dim = int(self._L.tg_obs_dim(self._h))
        self.frame_shape = (self.levels[0].frame_size[0], self.levels[0].frame_size[1], 3)
        self.max_episode_steps, self.auto_reset = int(max_episode_steps), bool(auto_reset)
        self.option_names = list(OPTION_NAMES)
        self.single_action_space = Discrete(len(OPTION_NAMES))                       # treasure_game.py:73
        self.single_observation_space = Box(np.float32(0.0), np.float32(1.0), shape=(self.obs_dim,))   # :75
        self.action_space, self.observation_space = self.single_action_space, self.single_observation_space
        n, d = self.num_envs, self.device
        self._obs = torch.empty((n, self.obs_dim), dtype=torch.float32, device=d)
        self._reward = torch.empty((n,), dtype=torch.float32, device=d)
        self._done = torch.empty((n,), dtype=torch.uint8, device=d)
        self._ran = torch.empty((n,), dtype=torch.uint8, device=d)
        self._avail = None
        self._flags = None
        self._mask = None
        self._stats = torch.zeros((8,), dtype=torch.int64, device=d)
        # the observation buffer lives as long as this object and only the library writes it: steps update just the
        # rows of the envs whose state changed (tg_bind_obs).  Results are views: clone before modifying them.
        check(self._L.tg_bind_obs(self._h, _ptr(self._obs)))
        self._tape = None
        self._host = None

    # ------------------------------------------------------------------ utils
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "_h", None):
            self._L.tg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ gym surface
    def reset(self, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``TreasureGame.reset`` (treasure_game.py:78-81) for every env, or for ``mask != 0``."""
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        check(self._L.tg_reset(self._h, _ptr(mask), _ptr(self._obs), self._stream()))
        return self._obs

    def step(self, actions: torch.Tensor, want_available: bool = False):
        """``TreasureGame.step`` (treasure_game.py:91-96).  Returns ``(obs, reward, done, info)``:
        obs (N, obs_dim) f32, reward (N,) f32 -- 0 where the option was not runnable (the
        reference returns ``None``; see ``info['ran']``), done (N,) bool, info with
        ``ran``, ``terminated``, ``truncated`` (and ``available`` (N,) int16 bit masks)."""
        if actions.dtype != torch.int32 or actions.device != self.device or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
        if actions.shape != (self.num_envs,):
            raise ValueError("actions must have shape (%d,)" % self.num_envs)
        if want_available and self._avail is None:
            self._avail = torch.empty((self.num_envs,), dtype=torch.int16, device=self.device)
        if self._flags is None:
            # 0 / 1 byte arrays the step kernel fills itself (tg_bind_flags), handed out as bool views: one launch per step()
            self._flags = tuple(torch.zeros((self.num_envs,), dtype=torch.uint8, device=self.device) for _ in range(3))
            check(self._L.tg_bind_flags(self._h, _ptr(self._flags[0]), _ptr(self._flags[1]), _ptr(self._flags[2])))
            self._flag_views = tuple(f.view(torch.bool) for f in self._flags)
            self._ran_bool = self._ran.view(torch.bool)
        check(self._L.tg_step(self._h, _ptr(actions), _ptr(self._obs), _ptr(self._reward), _ptr(self._done),
                              _ptr(self._ran), _ptr(self._avail) if want_available else None, self._stream()))
        done, terminated, truncated = self._flag_views
        info = {"ran": self._ran_bool, "terminated": terminated, "truncated": truncated}
        if want_available:
            info["available"] = self._avail
        return self._obs, self._reward, done, info

    def step_raw(self, actions: torch.Tensor):
        """Hot-loop variant: no tensor post-processing; returns the raw output buffers
        ``(obs, reward, done_bits, ran)``.  ``actions`` must already be a contiguous int32 tensor of shape
        (N,) on this env's device (``step`` converts; this one only checks)."""
        if actions.dtype is not torch.int32 or actions.device != self.device or not actions.is_contiguous() \
                or actions.numel() != self.num_envs:
            raise ValueError("step_raw needs a contiguous int32 tensor of %d actions on %s (got %s %s on %s)"
                             % (self.num_envs, self.device, actions.dtype, tuple(actions.shape), actions.device))
        check(self._L.tg_step(self._h, _ptr(actions), _ptr(self._obs), _ptr(self._reward), _ptr(self._done),
                              _ptr(self._ran), None, self._stream()))
        return self._obs, self._reward, self._done, self._ran

    def primitive_step(self, actions: torch.Tensor):
        """``_TreasureGameImpl.step(action)`` (impl:290-359) for every env: one primitive action
        (``_actions.py:7-13``), no option layer.  Returns the raw buffers ``(obs, reward, done_bits)``."""
        if actions.dtype != torch.int32 or actions.device != self.device or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
        check(self._L.tg_primitive_step(self._h, _ptr(actions), _ptr(self._obs), _ptr(self._reward), _ptr(self._done),
                                        self._stream()))
        return self._obs, self._reward, self._done

    def step_frames(self, env_ids, actions, max_ticks: int = 128, out: Optional[torch.Tensor] = None):
        """The option layer with a drawer (``_option.py:20-36``, ``drawer.draw_domain()`` after every primitive tick):
        the envs ``env_ids`` (distinct) take one gym step with ``actions``, and the frame after tick ``t`` of env ``k``
        is ``frames[k, t-1]``; frames after the last tick repeat the final one.  Returns ``(frames (M, max_ticks, H, W,
        3) uint8, n_ticks (M,) int32, obs (M, obs_dim), reward (M,), done_bits (M,), ran (M,))``; the other envs are
        not stepped."""
        ids = torch.as_tensor(env_ids, dtype=torch.int64).to(self.device).contiguous()
        acts = torch.as_tensor(actions, dtype=torch.int32).to(self.device).contiguous()
        m = int(ids.numel())
        if acts.numel() != m or m < 1:
            raise ValueError("one action per env id")
        shape = (m, int(max_ticks)) + self.frame_shape
        if out is None:
            out = torch.empty(shape, dtype=torch.uint8, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != torch.uint8 or out.device != self.device or not out.is_contiguous() or out.data_ptr() % 16:
            raise ValueError("out must be a contiguous, 16-byte aligned uint8 tensor of shape %s on %s" % (shape, self.device))
        n_ticks = torch.empty((m,), dtype=torch.int32, device=self.device)
        obs = torch.empty((m, self.obs_dim), dtype=torch.float32, device=self.device)
        rew = torch.empty((m,), dtype=torch.float32, device=self.device)
        done = torch.empty((m,), dtype=torch.uint8, device=self.device)
        ran = torch.empty((m,), dtype=torch.uint8, device=self.device)
        check(self._L.tg_step_frames(self._h, _ptr(ids), m, _ptr(acts), int(max_ticks), _ptr(out), _ptr(n_ticks), _ptr(obs),
                                     _ptr(rew), _ptr(done), _ptr(ran), self._stream()))
        return out, n_ticks, obs, rew, done, ran

    def make_host_buffers(self):
        """Pinned host buffers for ``step_host`` (actions in; obs, reward, done, ran out)."""
        n = self.num_envs
        self._host = dict(
            actions=torch.empty((n,), dtype=torch.int32).pin_memory(),
            obs=torch.empty((n, self.obs_dim), dtype=torch.float32).pin_memory(),
            reward=torch.empty((n,), dtype=torch.float32).pin_memory(),
            done=torch.empty((n,), dtype=torch.uint8).pin_memory(),
            ran=torch.empty((n,), dtype=torch.uint8).pin_memory())
        return self._host

    def step_host(self, host: Optional[Dict[str, torch.Tensor]] = None):
        """End-to-end step through ``tg_step_host``: host actions in, host obs/reward/done/ran out,
        with the copies and the final synchronisation inside the call."""
        host = host or self._host or self.make_host_buffers()
        check(self._L.tg_step_host(self._h, _ptr(host["actions"]), _ptr(host["obs"]), _ptr(host["reward"]),
                                   _ptr(host["done"]), _ptr(host["ran"]), self._stream()))
        return host

    def step_host_sparse(self, host: Optional[Dict[str, torch.Tensor]] = None):
        """``step_host`` with only the changed envs crossing the bus (``tg_step_host_sparse``): the host arrays must
        be the ones the previous ``step_host`` / ``step_host_sparse`` call filled; the result is the same."""
        host = host or self._host or self.make_host_buffers()
        check(self._L.tg_step_host_sparse(self._h, _ptr(host["actions"]), _ptr(host["obs"]), _ptr(host["reward"]),
                                          _ptr(host["done"]), _ptr(host["ran"]), self._stream()))
        return host

    def step_host_sparse_begin(self, host: Optional[Dict[str, torch.Tensor]] = None, stream: Optional["torch.cuda.Stream"] = None):
        """First half of ``step_host_sparse`` (``tg_step_host_sparse_begin``): enqueue the action copy, the step kernels
        and the record copies on ``stream`` (default: the current stream) and return without waiting.  Until
        ``step_host_sparse_end`` the host arrays must stay untouched (the outputs still hold the previous step)."""
        host = host or self._host or self.make_host_buffers()
        st = self._stream() if stream is None else C.c_void_p(stream.cuda_stream)
        check(self._L.tg_step_host_sparse_begin(self._h, _ptr(host["actions"]), _ptr(host["obs"]), _ptr(host["reward"]),
                                                _ptr(host["done"]), _ptr(host["ran"]), st))
        return host

    def step_host_sparse_end(self) -> None:
        """Second half: wait for the records and patch the host arrays given to ``step_host_sparse_begin``."""
        check(self._L.tg_step_host_sparse_end(self._h))

    @property
    def available_mask(self) -> torch.Tensor:
        """``TreasureGame.available_mask`` (treasure_game.py:83-89): (N, 9) uint8."""
        if self._mask is None:
            self._mask = torch.empty((self.num_envs, 9), dtype=torch.uint8, device=self.device)
        check(self._L.tg_available_mask(self._h, _ptr(self._mask), self._stream()))
        return self._mask

    def render(self, mode: str = "rgb_array", first: int = 0, count: Optional[int] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``TreasureGame.render('rgb_array')`` (treasure_game.py:98-104): (count, H, W, 3) uint8."""
        if mode != "rgb_array":
            raise NotImplementedError("only mode='rgb_array' is supported (no on-screen viewer)")
        count = self.num_envs - first if count is None else count
        if out is None:
            out = torch.empty((count,) + self.frame_shape, dtype=torch.uint8, device=self.device)
        elif out.shape != (count,) + self.frame_shape or out.dtype != torch.uint8 or not out.is_contiguous():
            raise ValueError("out must be a contiguous uint8 tensor of shape %s" % ((count,) + self.frame_shape,))
        if out.device != self.device or out.data_ptr() % 16:
            raise ValueError("out must live on %s and start on a 16-byte boundary (the renderer writes with bulk stores)" % self.device)
        check(self._L.tg_render(self._h, first, count, _ptr(out), self._stream()))
        return out

    # ------------------------------------------------------------------ analysis helpers of the drawer
    def draw_background_to_surface(self, level: int = 0) -> torch.Tensor:
        """``_TreasureGameDrawer.draw_background_to_surface`` (drawer.py:165-182): the tile layer, (H, W, 3) uint8."""
        out = torch.empty(self.frame_shape, dtype=torch.uint8, device=self.device)
        check(self._L.tg_background(self._h, int(level), _ptr(out), self._stream()))
        return out

    def draw_to_surface(self, first: int = 0, count: Optional[int] = None) -> torch.Tensor:
        """``_TreasureGameDrawer.draw_to_surface`` (drawer.py:184-196): the pixels of ``draw_domain`` on fresh
        surfaces, (count, H, W, 3) uint8."""
        return self.render("rgb_array", first=first, count=count)

    def blend(self, surfaces: torch.Tensor, alpha_objs: float, alpha_player: float, first: int = 0,
              count: Optional[int] = None, accumulate: bool = False) -> torch.Tensor:
        """``_TreasureGameDrawer.blend(surf, alpha_objs, alpha_player)`` (drawer.py:207-231), in place on
        ``surfaces``: with ``accumulate=False`` surface k (``(count, H, W, 3)`` uint8) receives env ``first + k``;
        with ``accumulate=True`` one surface ``(H, W, 3)`` receives the envs ``first .. first+count-1`` one after
        the other (a picture of a set of states)."""
        count = self.num_envs - first if count is None else count
        want = self.frame_shape if accumulate else (count,) + self.frame_shape
        if tuple(surfaces.shape) != tuple(want) or surfaces.dtype != torch.uint8 or surfaces.device != self.device \
                or not surfaces.is_contiguous():
            raise ValueError("surfaces must be a contiguous uint8 tensor of shape %s on %s" % (want, self.device))
        ao, ap = int(255 * alpha_objs), int(255 * alpha_player)            # drawer.py:223, :227
        check(self._L.tg_blend(self._h, first, count, count if accumulate else 1, _ptr(surfaces), ao, ap, self._stream()))
        return surfaces

    def blit_alpha(self, target: torch.Tensor, source: torch.Tensor, location, opacity: int) -> torch.Tensor:
        """``_TreasureGameDrawer.blit_alpha(target, source, location, opacity)`` (drawer.py:198-205), in place on
        ``target`` (H, W, 3) uint8; ``source`` (h, w, 3 | 4) uint8, both on this env's device."""
        for t in (target, source):
            if t.dtype != torch.uint8 or t.device != self.device or not t.is_contiguous() or t.dim() != 3:
                raise ValueError("target and source must be contiguous (h, w, c) uint8 tensors on %s" % self.device)
        with torch.cuda.device(self.device):
            check(self._L.tg_blit_alpha(_ptr(target), target.shape[1], target.shape[0], _ptr(source), source.shape[1],
                                        source.shape[0], source.shape[2], int(location[0]), int(location[1]), int(opacity),
                                        self._stream()))
        return target

    def draw_to_file(self, fname: str, index: int = 0) -> None:
        """``_TreasureGameDrawer.draw_to_file`` (drawer.py:233-236): the frame of env ``index`` as an image file;
        like ``pygame.image.save`` the format follows the extension (``.png``, ``.bmp``; anything else is PNG)."""
        from .imageio import save_rgb
        save_rgb(fname, self.render("rgb_array", first=index, count=1)[0].cpu().numpy())

    # ------------------------------------------------------------------ state access / parity hooks
    def get_state(self) -> Dict[str, torch.Tensor]:
        n, d = self.num_envs, self.device
        t = dict(
            pos=torch.empty((n, 2), dtype=torch.int32, device=d), misc=torch.empty((n, 4), dtype=torch.int32, device=d),
            doors=torch.empty((n, _lib.MAX_DOORS), dtype=torch.uint8, device=d),
            handles=torch.empty((n, _lib.MAX_HANDLES), dtype=torch.uint8, device=d),
            bolts=torch.empty((n, _lib.MAX_BOLTS), dtype=torch.uint8, device=d),
            angles=torch.empty((n, _lib.MAX_HANDLES), dtype=torch.float64, device=d),
            items=torch.empty((n, _lib.MAX_ITEMS, 2), dtype=torch.int32, device=d),
            bag=torch.empty((n, _lib.MAX_ITEMS), dtype=torch.int32, device=d),
            acct=torch.empty((n, 3), dtype=torch.int64, device=d))
        v = TgStateView(**{k: t[k].data_ptr() for k in t})
        check(self._L.tg_get_state(self._h, C.byref(v), self._stream()))
        return t

    def set_state(self, state: Dict[str, torch.Tensor]) -> None:
        dt = dict(pos=torch.int32, misc=torch.int32, doors=torch.uint8, handles=torch.uint8, bolts=torch.uint8,
                  angles=torch.float64, items=torch.int32, bag=torch.int32, acct=torch.int64)
        keep = {k: state[k].to(device=self.device, dtype=dt[k]).contiguous() for k in state}
        v = TgStateView(**{k: keep[k].data_ptr() for k in keep})
        check(self._L.tg_set_state(self._h, C.byref(v), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()

    def init_with_state(self, states, mask: Optional[torch.Tensor] = None) -> None:
        """``_TreasureGameImpl.init_with_state`` (impl:447-481) for every env (or ``mask != 0``):
        ``states`` (N, obs_dim) float64 normalised vectors in ``Level.state_descriptors()`` order,
        ``-99`` keeps the current value.  Reference quirks are reproduced (see treasure_b200.h)."""
        st = torch.as_tensor(states, dtype=torch.float64).to(self.device).reshape(self.num_envs, self.obs_dim).contiguous()
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        check(self._L.tg_init_with_state(self._h, _ptr(st), _ptr(mask), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()

    def snapshot(self, i: int = 0, state: Optional[Dict[str, torch.Tensor]] = None) -> dict:
        """State of env ``i`` in the dict layout of the oracle snapshots (tests)."""
        s = state or self.get_state()
        lv = self.levels[int(self.level_ids[i]) if self.level_ids is not None else 0]
        info = self._compiled[int(self.level_ids[i]) if self.level_ids is not None else 0].info
        c = {k: v[i].cpu().numpy() for k, v in s.items()}
        ni = info.n_items
        return dict(
            px=int(c["pos"][0]), py=int(c["pos"][1]), facing=int(c["misc"][0]), ticker=int(c["misc"][1]),
            doors=[int(v) for v in c["doors"][: info.n_doors]], handles_up=[int(v) for v in c["handles"][: info.n_handles]],
            angles=[float(v) for v in c["angles"][: info.n_handles]], bolts=[int(v) for v in c["bolts"][: info.n_bolts]],
            items=[(int(x), int(y), int(np.trunc(x / 48)), int(np.trunc(y / 48))) for x, y in c["items"][:ni]],
            bag=[int(v) for v in c["bag"] if v >= 0], total_actions=int(c["misc"][2]),
            handles_pt=[(int(c["acct"][2]) >> (1 + h)) & 1 for h in range(info.n_handles)])

    def set_draw_tape(self, tapes: Optional[Sequence[Sequence[float]]]) -> None:
        """Parity mode: env i consumes ``tapes[i]`` (uniform draws recorded from the reference)
        instead of its Philox stream; ``None`` switches back.  Draw indices restart at 0."""
        if tapes is None:
            check(self._L.tg_set_draw_tape(self._h, None, None, self._stream()))
            self._tape = None
            return
        if len(tapes) != self.num_envs:
            raise ValueError("one tape per env")
        off = np.zeros(self.num_envs + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(t) for t in tapes])
        flat = np.concatenate([np.asarray(t, dtype=np.float64) for t in tapes]) if off[-1] else np.zeros(1)
        self._tape = (torch.from_numpy(flat).to(self.device), torch.from_numpy(off).to(self.device))
        check(self._L.tg_set_draw_tape(self._h, _ptr(self._tape[0]), _ptr(self._tape[1]), self._stream()))

    # ------------------------------------------------------------------ statistics
    def stats_tensor(self) -> torch.Tensor:
        """Local int64[8] statistics vector (device); see ``STAT_NAMES``."""
        check(self._L.tg_stats(self._h, _ptr(self._stats), self._stream()))
        return self._stats

    def stats(self, all_reduce: bool = False) -> Dict[str, int]:
        """Episode statistics; with ``all_reduce`` summed over the ranks of the default
        ``torch.distributed`` group (the only collective of this path)."""
        t = self.stats_tensor().clone()
        if all_reduce:
            all_reduce_stats(t)
        return dict(zip(STAT_NAMES, (int(v) for v in t.cpu())))

    def clear_stats(self) -> None:
        check(self._L.tg_stats_clear(self._h, self._stream()))

    def host_traffic(self):
        """(host->device, device->host) bytes copied so far by ``step_host`` / ``step_host_sparse``."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._L.tg_host_traffic(self._h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def set_step_tile(self, tile: int = 0) -> None:
        """Tuning / test hook (``tg_debug_set_step_tile``): envs per step-kernel CTA, 0 = automatic."""
        check(self._L.tg_debug_set_step_tile(self._h, int(tile)))

    @property
    def launch_count(self) -> int:
        return int(self._L.tg_launch_count(self._h))


class PipelinedHostEnv:
    """``num_envs`` environments as ``parts`` sub-batches kept in flight together -- the double-buffered vector env of an
    actor loop with host-side policies.  Sub-batch k owns a contiguous slice of the global env ids (same streams and
    results as one ``VectorTreasureGame`` over the whole range: Philox streams are keyed by global env id), its own
    CUDA stream and its own pinned host arrays ``hosts[k]``.  The loop::

        for k in range(env.parts): env.begin(k)          # actions in env.hosts[k]["actions"]
        while training:
            for k in range(env.parts):
                env.end(k)                               # hosts[k] now holds obs / reward / done / ran of its step
                ... choose hosts[k]["actions"] from hosts[k]["obs"] ...
                env.begin(k)                             # runs on the device while the host works on the other parts

    keeps the closed loop per sub-batch (the actions of a step may depend on that sub-batch's last observations)
    while the device and the bus work on one part and the host patches another."""

    def __init__(self, num_envs: int, parts: Optional[int] = None, first_env_id: int = 0, **kwargs):
        # measured on one B200 + 16 host cores: 1,048,576 envs 2 / 3 / 4 parts = 0.32 / 0.29 / 0.37 ms per step, 131,072 envs 0.12 / 0.15 / 0.18
        self.num_envs = int(num_envs)
        self.parts = int(parts) if parts else (3 if self.num_envs >= 6 * 131072 else 2)
        self.ranges = [shard_range(self.num_envs, k, self.parts) for k in range(self.parts)]
        self.envs = [VectorTreasureGame(hi - lo, first_env_id=first_env_id + lo, **kwargs) for lo, hi in self.ranges]
        self.streams = [torch.cuda.Stream(device=e.device) for e in self.envs]
        self.hosts = [e.make_host_buffers() for e in self.envs]

    def reset(self):
        """Reset every sub-batch and fill the host arrays (observations of the initial states, zero flags)."""
        for e, h in zip(self.envs, self.hosts):
            h["obs"].copy_(e.reset())
            h["reward"].zero_(); h["done"].zero_(); h["ran"].zero_()
        return self.hosts

    def begin(self, k: int) -> None:
        # whatever the caller queued for this sub-batch on the current stream (reset, set_state) comes first
        self.streams[k].wait_stream(torch.cuda.current_stream(self.envs[k].device))
        self.envs[k].step_host_sparse_begin(self.hosts[k], self.streams[k])

    def end(self, k: int):
        self.envs[k].step_host_sparse_end()
        return self.hosts[k]

    def stats(self) -> Dict[str, int]:
        out: Dict[str, int] = {}
        for e in self.envs:
            for key, v in e.stats().items():
                out[key] = out.get(key, 0) + v
        return out

    def host_traffic(self):
        t = [e.host_traffic() for e in self.envs]
        return sum(a for a, _ in t), sum(b for _, b in t)

    def close(self):
        for e in self.envs:
            e.close()


def all_reduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """Sum an int64[8] statistics vector over the ranks of the default process group, in place
    (NCCL for CUDA tensors, gloo for CPU tensors).  No-op outside ``torch.distributed``."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def shard_range(total_envs: int, rank: int, world_size: int):
    """Contiguous env-id range owned by ``rank``: [lo, hi)."""
    base, rem = divmod(total_envs, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
