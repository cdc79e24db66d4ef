"""Boundary B1: the game API that the reference's search, players, arena and nets call.

Same 15 methods, argument meaning and return types as ``ColosseumBlokusGameWrapper``
(``blokus_rl/colossumrl/blokus_wrapper.py:21-324``), but talking to the engine by action *id* (no string
round trip, no 243 KB Python loop per mask) plus batched variants the GPU-native callers use.  The
reference's own wrapper also runs unchanged over :mod:`blokus_rl_b200.colosseum_shim` (boundary B0); this
class is the faster sibling with identical semantics.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from . import tables


class BlokusGameWrapper:
    def __init__(self, hparams=None, *, board_size: int | None = None, number_of_players: int | None = None,
                 backend=None, score_rule: int = 0, device=None):
        self.hparams = hparams
        self.board_size = board_size if board_size is not None else getattr(hparams, "board_size", 20)
        self.number_of_players = (number_of_players if number_of_players is not None
                                  else getattr(hparams, "number_of_players", 4))
        if backend is None:
            from .backend import EngineBackend          # GPU engine; no CPU fallback
            backend = EngineBackend(self.board_size, self.number_of_players, score_rule, device)
        self.backend = backend
        # id <-> string maps in the reference's cache format (blokus_wrapper.py:284-288)
        self._move_action_dict = tables.string_to_action(self.board_size)
        self.action_move_dict = dict(enumerate(tables.action_strings(self.board_size)))
        self.starter_won = self.last_won = self.games_played = 0
        states_dir = getattr(hparams, "states_dir", None)
        if states_dir is not None:
            fp = Path(states_dir) / f"colosseum_{self.board_size}_players_{self.number_of_players}.json"
            if not fp.exists():
                tables.write_action_json(fp, self.board_size)

    # ---- shape queries (blokus_wrapper.py:52-78) -----------------------------------------------------
    def get_board_size(self):
        return (self.board_size, self.board_size)

    def get_action_size(self) -> int:
        return len(self.action_move_dict)

    def get_observation_size(self):
        return [2 * self.number_of_players, self.board_size, self.board_size]

    def get_number_of_players(self) -> int:
        return self.number_of_players

    # ---- transitions (blokus_wrapper.py:80-106) --------------------------------------------------------
    def get_init_board(self):
        s = self.backend.new_state()
        return s, self.backend.mover(s)

    def get_next_state(self, current_state, current_player, action_id):
        aid = self._move_action_dict[action_id] if isinstance(action_id, str) else int(action_id)
        nxt = self.backend.next_state(current_state, aid)
        return nxt, self.backend.mover(nxt)

    # ---- masks / observations (blokus_wrapper.py:108-162) ------------------------------------------------
    def get_valid_moves(self, current_state, current_player=-1):
        # the reference hands `current_player` to env.valid_actions (blokus_wrapper.py:121-124); every call site passes the
        # state's own side to move (or -1), and the engine only ever evaluates that player
        if current_player not in (-1, self.backend.mover(current_state)):
            raise ValueError(f"get_valid_moves: player {current_player} is not the side to move "
                             f"({self.backend.mover(current_state)}) of this state")
        return self.backend.legal_mask(current_state).astype(np.float64)

    def get_observation(self, state, player):
        return self.backend.observation(state), self.get_valid_moves(state, player)

    def get_valid_actions_for_human_player(self, state, player):
        ids = self.backend.legal_ids(state)
        return [self.action_move_dict[int(i)] for i in ids] if len(ids) else [""]

    # ---- terminal (blokus_wrapper.py:164-206) ---------------------------------------------------------------
    def get_game_ended(self, state):
        if not self.backend.done(state):
            return None
        return np.asarray(self.backend.terminal_values(state), dtype=np.float64).copy()

    def get_scores(self, winners):
        out = -np.ones(self.number_of_players)
        sole = len(winners) == 1
        for w in winners:
            out[w] = 1 if sole else 0
        return out

    # ---- hashing / display / sampling (blokus_wrapper.py:208-279) ----------------------------------------------
    def string_representation(self, state):
        return hash(self.backend.board_key(state))

    def display(self, state) -> None:
        print(self.backend.board_contents(state))

    def get_sample_move(self, state):
        return self.backend.sample_move(state)

    def render(self, state):
        """RGB image of the board (uint8 HxWx3).  The reference draws with matplotlib
        (blokus_wrapper.py:248-279); this is a dependency-free raster of the same colouring."""
        colors = np.array([[211, 211, 211], [255, 0, 0], [0, 0, 255], [255, 255, 0], [0, 128, 0]], np.uint8)
        b = self.backend.board_contents(state)
        cell = 24
        img = colors[b][::-1].repeat(cell, 0).repeat(cell, 1)      # matplotlib's y axis points up
        img[::cell, :, :] = 128
        img[:, ::cell, :] = 128
        return img
