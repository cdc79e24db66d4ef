"""Single-state (functional) view over the batched GPU engine.

The reference's search code handles ONE immutable state at a time (``MCTS.simulate`` re-descends from the
same ``s`` many times: ``blokus_rl/alphazero/mcts.py:47``).  :class:`EngineBackend` gives that shape over
:class:`BlokusEngine`: every transition is one ``blk_step`` launch with n = 1 that also produces the new
state's legal mask, terminal vector and scores, which the handle caches.  The adapters
(``colosseum_shim``, ``game_wrapper``) are written against this small interface only, so the tests can
run them over a CPU-oracle backend inside the build container (no GPU there) — the product never does.
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import BlokusEngine


class StateHandle:
    """Immutable env state: device words + lazily fetched host views."""
    __slots__ = ("words", "host_words", "mask_dev", "_mask", "flags", "terminal", "scores", "_obs", "_board")

    def __init__(self, words, host_words, mask_dev, flags, terminal, scores):
        self.words, self.host_words, self.mask_dev = words, host_words, mask_dev
        self.flags, self.terminal, self.scores = flags, terminal, scores
        self._mask = self._obs = self._board = None


class EngineBackend:
    def __init__(self, board_size: int = 20, num_players: int = 4, score_rule: int = 0, device=None,
                 engine: BlokusEngine | None = None):
        self.eng = engine if engine is not None else BlokusEngine(board_size, num_players, score_rule, device)
        self.N, self.P, self.A = self.eng.board_size, self.eng.num_players, self.eng.num_actions
        self._meta = self.P * self.N + self.P

    # ---- construction ------------------------------------------------------------------------------
    def _finish(self, out) -> StateHandle:
        host = out.states.cpu().numpy().view(np.uint32)[0].copy()      # one small D2H + sync
        return StateHandle(out.states, host, out.mask, int(out.flags.item()), out.terminal.cpu().numpy()[0],
                           out.scores.cpu().numpy()[0])

    def new_state(self) -> StateHandle:
        s = self.eng.new_states(1)
        return self._finish(self.eng.step(s, None, mask="bytes"))

    def from_words(self, words: np.ndarray) -> StateHandle:
        s = torch.from_numpy(np.ascontiguousarray(words, np.uint32).view(np.int32).reshape(1, -1)).to(self.eng.device)
        return self._finish(self.eng.step(s, None, mask="bytes"))

    def next_state(self, h: StateHandle, action_id: int) -> StateHandle:
        act = torch.tensor([int(action_id)], dtype=torch.int32, device=self.eng.device)
        dst = torch.empty_like(h.words)
        out = self.eng.step(h.words, act, out_states=dst, mask="bytes")
        nh = self._finish(out)
        if nh.flags & 2:
            raise ValueError(f"illegal action {action_id} for player {self.mover(h)}")
        return nh

    # ---- queries -------------------------------------------------------------------------------------
    def mover(self, h: StateHandle) -> int:
        return int(h.host_words[self._meta] & 15)

    def done(self, h: StateHandle) -> bool:
        return bool((h.host_words[self._meta] >> 4) & 1)

    def ply(self, h: StateHandle) -> int:
        return int(h.host_words[self._meta] >> 16)

    def legal_mask(self, h: StateHandle) -> np.ndarray:
        if h._mask is None:
            h._mask = h.mask_dev[0].cpu().numpy().astype(np.uint8)
        return h._mask

    def legal_ids(self, h: StateHandle) -> np.ndarray:
        return np.flatnonzero(self.legal_mask(h))

    def winners(self, h: StateHandle) -> list[int]:
        if not self.done(h):
            return []
        return [p for p in range(self.P) if h.terminal[p] > 0]

    def terminal_values(self, h: StateHandle) -> np.ndarray:
        return h.terminal

    def scores(self, h: StateHandle) -> np.ndarray:
        return h.scores

    def observation(self, h: StateHandle) -> np.ndarray:
        if h._obs is None:
            h._obs = self.eng.observe(h.words).cpu().numpy()[0]
        return h._obs

    def board_contents(self, h: StateHandle) -> np.ndarray:
        if h._board is None:
            rows = h.host_words[: self.P * self.N].reshape(self.P, self.N)
            b = np.zeros((self.N, self.N), np.uint8)
            cols = np.arange(self.N, dtype=np.uint32)
            for q in range(self.P):
                b[((rows[q][:, None] >> cols[None, :]) & 1).astype(bool)] = q + 1
            h._board = b
        return h._board

    def board_key(self, h: StateHandle) -> bytes:
        """What ``hash(board.board_contents.tobytes())`` keys on (blokus_wrapper.py:217-218): the cells only."""
        return h.host_words[: self.P * self.N].tobytes()

    def sample_move(self, h: StateHandle, rng=np.random) -> int:
        ids = self.legal_ids(h)
        return int(rng.choice(ids))
