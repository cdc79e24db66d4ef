"""Boundary B0: a ``colosseumrl.envs.blokus``-compatible module backed by the GPU engine.

The reference imports five names from that (absent, un-vendored) package at
``blokus_rl/colossumrl/blokus_wrapper.py:8-14``: ``BlokusEnvironment``, ``Board``, ``action_to_string``,
``GAME_PIECE_VALUES``, ``PLAYER_TO_COLOR``.  This module provides exactly the surface those call sites
use, so ``ColosseumBlokusGameWrapper``, ``MCTS``, the players and the arena of
the reference run unchanged:

    from blokus_rl_b200 import colosseum_shim
    colosseum_shim.install()            # registers sys.modules["colosseumrl.envs.blokus"]
    from blokus_rl.colossumrl.blokus_wrapper import ColosseumBlokusGameWrapper   # the reference's own file

States are immutable values ``(board, round_count, players)`` (``blokus_wrapper.py:83,94``).
"""
from __future__ import annotations

import sys
import types

import numpy as np

from . import tables

GAME_PIECE_VALUES = {name: int(size) for name, size in zip(tables.PIECE_NAMES, tables.piece_sizes())}
PLAYER_TO_COLOR = {p: p for p in range(0, 5)}      # blokus_wrapper.py:295 only indexes [1]
action_to_string = tables.action_to_string

_backend = None
_backend_factory = None


def set_backend(backend) -> None:
    """Use an existing backend (an :class:`blokus_rl_b200.backend.EngineBackend`)."""
    global _backend
    _backend = backend


def _get_backend():
    global _backend
    if _backend is None:
        if _backend_factory is not None:
            _backend = _backend_factory()
        else:
            from .backend import EngineBackend      # GPU engine; raises loudly without a B200
            _backend = EngineBackend(20, 4)
    return _backend


class _Player:
    __slots__ = ("player_color", "index")

    def __init__(self, index: int):
        self.index = index
        self.player_color = index          # accepted back as valid_actions' `player` (blokus_wrapper.py:243-244)


class Board:
    """State-side view (``state[0]``) and, when constructed directly, the enumeration helper that
    ``_set_all_possible_moves`` drives (``blokus_wrapper.py:292-312``)."""

    def __init__(self, track_canonical: bool = False, *, _handle=None, _backend=None):
        self._h, self._b = _handle, _backend
        self._scratch = None
        if _handle is None:
            self._n = _get_backend().N
            self.reset_board()

    # -- state view ---------------------------------------------------------------------------------
    @property
    def player_color(self):
        return self._b.mover(self._h)

    @property
    def canonical_board(self):
        return self._b.observation(self._h)

    @property
    def board_contents(self):
        if self._h is None:
            return self._scratch
        return self._b.board_contents(self._h)

    # -- enumeration helper ---------------------------------------------------------------------------
    def get_all_valid_moves(self, round_count=0, player_color=1, player_pieces=None, define_states=True):
        n = self._n
        t = tables.action_table(n)
        ors = tables.orientations()
        out: dict = {}
        for a in range(t.num_actions):
            o = ors[int(t.action_orient[a])]
            idx = tables.action_index(n, o.local, int(t.action_y[a]), int(t.action_x[a]))
            out.setdefault(o.piece, {})[idx] = [o.local]
        return out

    def reset_board(self):
        self._scratch = np.zeros((self._n, self._n), dtype=np.int8)

    def update_board(self, color, piece_type, index, orientation, round_count=0, flag=True):
        n = self._n
        local, rem = divmod(int(index), n * n)
        y, x = divmod(rem, n)
        o = next(o for o in tables.orientations() if o.piece == piece_type and o.local == local)
        for dy, dx in o.cells:
            self._scratch[y + dy, x + dx] = color


def _handle_of(state):
    return state[0]._h


class BlokusEnvironment:
    def __init__(self):
        self._b = _get_backend()
        self._ids = tables.string_to_action(self._b.N)
        self._strings = tables.action_strings(self._b.N)

    def _wrap(self, h):
        b = self._b
        return (Board(_handle=h, _backend=b), b.ply(h), [_Player(b.mover(h))])

    def new_state(self, num_players=None):
        h = self._b.new_state()
        return self._wrap(h), [self._b.mover(h)]

    def next_state(self, state, players, actions):
        action = actions[0]
        aid = self._ids[action] if isinstance(action, str) else int(action)
        nh = self._b.next_state(_handle_of(state), aid)
        b = self._b
        terminal = b.done(nh)
        winners = b.winners(nh)
        rewards = [float(v) for v in b.terminal_values(nh)]
        return self._wrap(nh), [b.mover(nh)], rewards, terminal, winners

    def valid_actions(self, state, player=None):
        """Legal action strings of the state's side to move; ``[""]`` when there are none
        (``blokus_wrapper.py:126``).  The three reference call sites pass three different things as `player`
        (index, ``board.player_color``, ``players[0].player_color``); all mean "the mover"."""
        ids = self._b.legal_ids(_handle_of(state))
        if len(ids) == 0:
            return [""]
        return [self._strings[i] for i in ids]

    def get_winners(self, state):
        return self._b.winners(_handle_of(state))

    def is_valid_action(self, state, player, action):
        aid = self._ids.get(action, -1) if isinstance(action, str) else int(action)
        return aid >= 0 and bool(self._b.legal_mask(_handle_of(state))[aid])


def install(backend=None, backend_factory=None) -> types.ModuleType:
    """Register this module as ``colosseumrl.envs.blokus`` (and its parent packages) in ``sys.modules``."""
    global _backend_factory
    if backend is not None:
        set_backend(backend)
    if backend_factory is not None:
        _backend_factory = backend_factory
    me = sys.modules[__name__]
    root = sys.modules.setdefault("colosseumrl", types.ModuleType("colosseumrl"))
    envs = sys.modules.setdefault("colosseumrl.envs", types.ModuleType("colosseumrl.envs"))
    root.envs = envs
    envs.blokus = me
    sys.modules["colosseumrl.envs.blokus"] = me
    return me
