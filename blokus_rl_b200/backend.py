"""Single-state (functional) view over the batched GPU engine.

The reference's search code handles ONE immutable state at a time (``MCTS.simulate`` re-descends from the
same ``s`` many times: ``blokus_rl/alphazero/mcts.py:47``).  :class:`EngineBackend` gives that shape over
:class:`BlokusEngine`: every transition is one ``blk_step`` launch with n = 1 that also produces the new
state's legal mask, terminal vector and scores, which the handle caches.  The adapters
(``colosseum_shim``, ``game_wrapper``) are written against this small interface only, so the tests can
run them over a CPU-oracle backend inside the build container (no GPU there) — the product never does.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import BLK_FLAG_ILLEGAL, BLK_FLAG_TRUNCATED, BLK_MASK_INDICES
from .engine import BlokusEngine, StepOut


def _cudart():
    """libcudart as this process already has it (torch loads it), for cudaMemcpyAsync / cudaStreamSynchronize without the
    5-15 us of Python wrapping per call; None when it cannot be found by name (the torch calls are used then)."""
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            rt = C.CDLL(name)
            rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
            rt.cudaMemcpyAsync.restype = C.c_int
            rt.cudaStreamSynchronize.argtypes = [C.c_void_p]
            rt.cudaStreamSynchronize.restype = C.c_int
            return rt
        except (OSError, AttributeError):
            continue
    return None


class StateHandle:
    """Immutable env state: device words + the host copy of everything the step that produced it returned."""
    __slots__ = ("_pack", "_words", "_sw", "host_words", "ids", "_mask", "flags", "terminal", "scores", "_obs", "_board")

    def __init__(self, words, host_words, ids, flags, terminal, scores, obs=None, pack=None, sw=0):
        self._words, self._pack, self._sw = words, pack, sw
        self.host_words, self.ids = host_words, ids
        self.flags, self.terminal, self.scores = flags, terminal, scores
        self._mask = self._board = None
        self._obs = obs

    @property
    def words(self) -> torch.Tensor:
        """The state on the device, int32 [1, state_words] (the head of the buffer the transition wrote)."""
        if self._words is None:
            self._words = self._pack[: 4 * self._sw].view(torch.int32).view(1, self._sw)
        return self._words


class EngineBackend:
    def __init__(self, board_size: int = 20, num_players: int = 4, score_rule: int = 0, device=None,
                 engine: BlokusEngine | None = None, fuse_observation: bool = True):
        """``fuse_observation``: every transition also writes the new state's ``canonical_board`` planes (from the same
        launch, into the same buffer), so the ``get_observation`` that follows each expansion of the reference's search
        (mcts.py:60-66) costs no second launch.  12.8 KB more per device-to-host copy at 20x20 / 4 players."""
        self.eng = engine if engine is not None else BlokusEngine(board_size, num_players, score_rule, device)
        self.N, self.P, self.A = self.eng.board_size, self.eng.num_players, self.eng.num_actions
        self._meta = self.P * self.N + self.P
        # Every output of a transition lands in ONE device buffer, so a transition costs one launch and ONE device-to-host
        # copy (the reference's search makes 25-200 of them per move, blokus_rl/alphazero/mcts.py:47-71):
        #   state words | terminal f32[P] | scores i16[P] | legal_count i32 | flags u8 | legal ids u16[max_legal]
        # The mask travels in the sparse BLK_MASK_INDICES form (~350 B instead of 30 KB).
        a16 = lambda x: (x + 15) & ~15
        self._o_term = a16(4 * self.eng.state_words)
        self._o_score = self._o_term + 16
        self._o_count = self._o_score + 16
        self._o_flags = self._o_count + 4
        self._o_ids = a16(self._o_flags + 1)
        self._o_obs = a16(self._o_ids + 2 * self.eng.max_legal)
        self._obs_floats = 2 * self.P * self.N * self.N if fuse_observation else 0
        self._pack_bytes = self._o_obs + 4 * self._obs_floats
        # The per-move path (`next_state`: 25-200 calls per move of the reference's search) goes to the C ABI directly: the
        # argument block is filled once, the action is read by the kernel from pinned host memory (no H2D copy call), and
        # the outputs come back through one async copy into a pinned buffer + one stream synchronize.
        dev = self.eng.device
        self._direct = isinstance(self.eng, BlokusEngine)     # (the host-logic tests drive this class over a stand-in engine)
        if not self._direct:
            return
        self._act = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._act_np = self._act.numpy()
        self._host = torch.empty(self._pack_bytes, dtype=torch.uint8).pin_memory()
        self._host_np = self._host.numpy()
        self._args = _lib.BlkStepArgs(1, None, None, self._act.data_ptr(), None, BLK_MASK_INDICES, self.eng.max_legal,
                                      None, None, None, None, None, 0, 0, 0, None, None, None, None)
        self._args_ref = C.byref(self._args)
        self._step = self.eng._lib.blk_step
        self._dev_index = dev.index
        self._host_ptr = C.c_void_p(self._host.data_ptr())
        self._cudart = _cudart()
        with torch.cuda.device(dev):
            self._fetched = torch.cuda.Event()

    # ---- construction ------------------------------------------------------------------------------
    def _views(self, pack: torch.Tensor):
        P, SW = self.P, self.eng.state_words
        words = pack[: 4 * SW].view(torch.int32).view(1, SW)
        bufs = StepOut(None, None,
                       pack[self._o_count: self._o_count + 4].view(torch.int32),
                       pack[self._o_term: self._o_term + 4 * P].view(torch.float32).view(1, P),
                       pack[self._o_flags: self._o_flags + 1],
                       pack[self._o_score: self._o_score + 2 * P].view(torch.int16).view(1, P), None)
        ids = pack[self._o_ids: self._o_ids + 2 * self.eng.max_legal].view(torch.int16).view(1, self.eng.max_legal)
        return words, bufs, ids

    def _transition(self, src: torch.Tensor, action: torch.Tensor | None) -> StateHandle:
        """One ``blk_step`` launch (n = 1) whose outputs all land in one buffer, then one D2H copy of it."""
        pack = torch.empty(self._pack_bytes, dtype=torch.uint8, device=self.eng.device)
        words, bufs, ids = self._views(pack)
        if action is None:
            words.copy_(src)
            src = words
        self.eng.step(src, action, out_states=words, mask=ids, buffers=bufs)
        host = pack.cpu().numpy()                                          # the one synchronising copy
        flags = int(host[self._o_flags])
        n = int(host[self._o_count: self._o_count + 4].view(np.int32)[0])
        if flags & BLK_FLAG_TRUNCATED:                                     # more legal moves than the id row holds
            m = self.eng.step(words, None, mask="bytes", want_count=False, want_terminal=False, want_scores=False).mask
            hids = np.flatnonzero(m[0].cpu().numpy()).astype(np.int64)
        else:
            hids = host[self._o_ids: self._o_ids + 2 * n].view(np.uint16).astype(np.int64)
        return StateHandle(words, host[: 4 * self.eng.state_words].view(np.uint32).copy(), hids, flags,
                           host[self._o_term: self._o_term + 4 * self.P].view(np.float32).copy(),
                           host[self._o_score: self._o_score + 2 * self.P].view(np.int16).copy())

    def new_state(self) -> StateHandle:
        return self._transition(self.eng.new_states(1), None)

    def from_words(self, words: np.ndarray) -> StateHandle:
        s = torch.from_numpy(np.ascontiguousarray(words, np.uint32).view(np.int32).reshape(1, -1)).to(self.eng.device)
        return self._transition(s, None)

    def _launch_and_fetch(self, pack: torch.Tensor) -> None:
        stream = self.eng._stream()
        _lib.check(self._step(self.eng._h, self._args_ref, stream))
        rt = self._cudart
        if rt is not None:                               # the CUDA runtime torch already loaded, called directly
            if rt.cudaMemcpyAsync(self._host_ptr, C.c_void_p(pack.data_ptr()), self._pack_bytes, 2, stream) or \
                    rt.cudaStreamSynchronize(stream):
                raise _lib.EngineError("device-to-host copy of a transition failed")
            return
        self._host.copy_(pack, non_blocking=True)
        self._fetched.record()
        self._fetched.synchronize()

    def next_state(self, h: StateHandle, action_id: int) -> StateHandle:
        if not self._direct:
            act = torch.tensor([int(action_id)], dtype=torch.int32).to(self.eng.device, non_blocking=True)
            nh = self._transition(h.words, act)
            if nh.flags & BLK_FLAG_ILLEGAL:
                raise ValueError(f"illegal action {action_id} for player {self.mover(h)}")
            return nh
        eng, a, host = self.eng, self._args, self._host_np
        pack = torch.empty(self._pack_bytes, dtype=torch.uint8, device=eng.device)
        base = pack.data_ptr()
        self._act_np[0] = action_id
        a.state_in = h._pack.data_ptr() if h._pack is not None else h.words.data_ptr()
        a.state_out = base
        a.mask = base + self._o_ids
        a.legal_count = base + self._o_count
        a.terminal = base + self._o_term
        a.flags = base + self._o_flags
        a.scores = base + self._o_score
        a.obs = base + self._o_obs if self._obs_floats else None
        if torch.cuda.current_device() == self._dev_index:
            self._launch_and_fetch(pack)
        else:
            with torch.cuda.device(self._dev_index):
                self._launch_and_fetch(pack)
        flags = int(host[self._o_flags])
        if flags & BLK_FLAG_ILLEGAL:
            raise ValueError(f"illegal action {action_id} for player {self.mover(h)}")
        if flags & BLK_FLAG_TRUNCATED:                                     # more legal moves than the id row holds: general path
            return self._transition(pack[: 4 * eng.state_words].view(torch.int32).view(1, eng.state_words), None)
        n = int(host[self._o_count: self._o_count + 4].view(np.int32)[0])
        P = self.P
        obs = (host[self._o_obs: self._o_obs + 4 * self._obs_floats].view(np.float32).reshape(2 * P, self.N, self.N).copy()
               if self._obs_floats else None)
        return StateHandle(None, host[: 4 * eng.state_words].view(np.uint32).copy(),
                           host[self._o_ids: self._o_ids + 2 * n].view(np.uint16).astype(np.int64), flags,
                           host[self._o_term: self._o_term + 4 * P].view(np.float32).copy(),
                           host[self._o_score: self._o_score + 2 * P].view(np.int16).copy(), obs, pack, eng.state_words)

    # ---- queries -------------------------------------------------------------------------------------
    def mover(self, h: StateHandle) -> int:
        return int(h.host_words[self._meta] & 15)

    def done(self, h: StateHandle) -> bool:
        return bool((h.host_words[self._meta] >> 4) & 1)

    def ply(self, h: StateHandle) -> int:
        return int(h.host_words[self._meta] >> 16)

    def legal_mask(self, h: StateHandle) -> np.ndarray:
        if h._mask is None:
            h._mask = np.zeros(self.A, np.uint8)
            h._mask[h.ids] = 1
        return h._mask

    def legal_ids(self, h: StateHandle) -> np.ndarray:
        return h.ids

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
