from .treasure_game import ObservationWrapper, TreasureGame  # noqa: F401
