"""gym_treasure_game_b200 -- B200-native batched Treasure Game simulator.

``make('treasure_game-v0')`` mirrors ``gym.make('treasure_game-v0')`` of the
reference (``gym_treasure_game/__init__.py:3-6``); when ``gym`` is importable the
id is also registered with gym's own registry under the reference's name.
"""
from .level import Level  # noqa: F401

ENV_ID = "treasure_game-v0"
_REGISTRY = {ENV_ID: "gym_treasure_game_b200.envs:TreasureGame"}

try:  # pragma: no cover - gym is optional
    from gym.envs.registration import register as _register

    try:
        _register(id=ENV_ID, entry_point=_REGISTRY[ENV_ID])
    except Exception:
        pass
except Exception:
    pass


def make(env_id=ENV_ID, **kwargs):
    """Local stand-in for ``gym.make`` (no TimeLimit wrapper: the reference registers none)."""
    if env_id not in _REGISTRY:
        raise KeyError("unknown environment id %r" % env_id)
    from .envs import TreasureGame
    return TreasureGame(**kwargs)


def __getattr__(name):      # lazy: importing the package must not require torch/CUDA
    if name in ("VectorTreasureGame", "PipelinedHostEnv", "shard_range", "all_reduce_stats", "OPTION_NAMES", "STAT_NAMES"):
        from . import vector_env
        return getattr(vector_env, name)
    if name in ("TreasureGame", "ObservationWrapper"):
        from . import envs
        return getattr(envs, name)
    raise AttributeError(name)
