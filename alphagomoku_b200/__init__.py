"""B200-native lockstep self-play engine for AlphaGomoku's data-parallel hot path.

The product is the CUDA library libagb200.so behind the C ABI in include/agb200.h; this package is the thin Python
host mirror used by the tests and the benchmark. There is no CPU fallback: constructing an Engine without a CUDA
device raises."""
from .engine import (AgbError, Engine, GameConfig, GameRules, Sign, move_to_short, short_to_move)  # noqa: F401

__all__ = ["AgbError", "Engine", "GameConfig", "GameRules", "Sign", "move_to_short", "short_to_move"]
