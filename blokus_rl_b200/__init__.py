"""blokus_rl_b200 -- B200-native batched Blokus environment engine.

Drop-in for the env hot path of KubiakJakub01/Blokus-RL (reset / step / legal-action mask /
observation / winners, plus batched random rollouts), implemented as hand-written sm_100a CUDA
behind the C ABI in include/blokus_b200.h.  Importing this package does not touch the GPU; creating
a :class:`BlokusEngine` does and fails loudly without one (there is no CPU fallback).
"""
from . import tables
from ._lib import EngineError

__all__ = ["BlokusEngine", "EngineError", "tables"]


def __getattr__(name):  # lazy: keep `import blokus_rl_b200` torch-free for table-only users
    if name in ("BlokusEngine", "StepOut", "RolloutOut"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
