"""Batched Blokus environment engine: torch-tensor front-end of the C ABI (include/blokus_b200.h).

PyTorch is plumbing only here (device memory, streams); every operation is a hand-written sm_100a
kernel in csrc/blk_engine.cu reached through ctypes.  The methods mirror the env-engine calls the
reference makes in blokus_rl/colossumrl/blokus_wrapper.py (cited per method), batched over envs.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import (BLK_FLAG_DONE, BLK_FLAG_ILLEGAL, BLK_MASK_BITS, BLK_MASK_BYTES, BLK_MASK_INDICES, BLK_MASK_NONE,
                   BLK_OPT_AUTO_RESET, BLK_OPT_WARP_KERNELS, EngineError)

__all__ = ["BlokusEngine", "StepOut", "RolloutOut", "EngineError", "BLK_FLAG_DONE", "BLK_FLAG_ILLEGAL"]


@dataclass
class StepOut:
    """Per-call outputs of :meth:`BlokusEngine.step` (all CUDA tensors, row i = env i)."""
    states: torch.Tensor                 # int32 [n, state_words]
    mask: torch.Tensor | None            # bool [n, A] (view of a 16 B-padded buffer) or int32 [n, mask_words]
    legal_count: torch.Tensor | None     # int32 [n]
    terminal: torch.Tensor | None        # float32 [n, P]   3/1/-1 when the game ended in this call, else 0
    flags: torch.Tensor | None           # uint8 [n]        BLK_FLAG_DONE | BLK_FLAG_ILLEGAL
    scores: torch.Tensor | None          # int16 [n, P]
    next_action: torch.Tensor | None     # int32 [n]        uniform random legal action of the new mover
    mask_raw: torch.Tensor | None = None # the padded buffer behind `mask`
    obs: torch.Tensor | None = None      # float32 [n, 2P, N, N] observation of the resulting states (fused output)
    csr_offset: torch.Tensor | None = None   # mask="csr": int64 [n], env i's ids are mask_raw[csr_offset[i] : + legal_count[i]]
    csr_cursor: torch.Tensor | None = None   # mask="csr": int64 [1], total entries requested by this call


@dataclass
class RolloutOut:
    final_scores: torch.Tensor           # int16 [n_roots, per_root, P]
    winners: torch.Tensor                # uint8 [n_roots, per_root] bitmask
    value_sum: torch.Tensor              # float32 [n_roots, P]
    plies: torch.Tensor                  # int32 [n_roots, per_root]
    action_log: torch.Tensor | None      # int32-viewable uint16 stored as int16 [n_roots, per_root, log_stride]


# the current stream's handle without building a torch.cuda.Stream object (9 us -> < 1 us per launch)
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None) or (lambda index: torch.cuda.current_stream(index).cuda_stream)


def _ptr(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


class BlokusEngine:
    """One engine per (board_size, num_players, score_rule, device).

    ``states`` are ``int32 [n, state_words]`` CUDA tensors in the layout documented in
    include/blokus_b200.h (the bits are uint32; torch's int32 is used as the carrier).
    """

    def __init__(self, board_size: int = 20, num_players: int = 4, score_rule: int = 0,
                 device: int | str | torch.device | None = None):
        if not torch.cuda.is_available():
            raise EngineError("CUDA is not available: the Blokus engine has no CPU fallback")
        self._lib = _lib.load()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise EngineError("device must be a CUDA device")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        cfg = _lib.BlkConfig(board_size, num_players, score_rule, self.device.index)
        self._h = C.c_void_p()
        _lib.check(self._lib.blk_create(C.byref(cfg), C.byref(self._h)))
        info = _lib.BlkInfo()
        _lib.check(self._lib.blk_get_info(self._h, C.byref(info)))
        self.info = info
        self.board_size, self.num_players = info.board_size, info.num_players
        self.num_actions, self.state_words = info.num_actions, info.state_words
        self.mask_words, self.mask_bytes = info.mask_words, info.mask_bytes
        self.sm_count = info.sm_count
        self.max_legal = 1024        # row width of the sparse 'indices' mask format (no reachable position is known to exceed ~800)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.blk_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(_raw_stream(self.device.index))

    def _check_states(self, states: torch.Tensor):
        if states.dtype != torch.int32 or states.dim() != 2 or states.shape[1] != self.state_words \
                or not states.is_contiguous() or states.device != self.device:
            raise ValueError(f"states must be a contiguous int32 [n, {self.state_words}] tensor on {self.device}")

    def action_to_cells(self, action: int):
        """(piece, orientation-within-piece, anchor_y, anchor_x), [(y, x), ...] of one action id
        (footprint view of the ids built at blokus_wrapper.py:300-316)."""
        meta = (C.c_int32 * 4)()
        cells = (C.c_uint8 * 10)()
        n = C.c_int32()
        _lib.check(self._lib.blk_action_to_cells(self._h, int(action), meta, cells, C.byref(n)))
        return tuple(meta), [(cells[2 * i], cells[2 * i + 1]) for i in range(n.value)]

    # ---- a1 new_state: blokus_wrapper.py:80-87 ---------------------------------------------------
    def new_states(self, n: int) -> torch.Tensor:
        states = torch.empty((n, self.state_words), dtype=torch.int32, device=self.device)
        self.reset(states)
        return states

    def reset(self, states: torch.Tensor) -> torch.Tensor:
        self._check_states(states)
        _lib.check(self._lib.blk_reset(self._h, _ptr(states), states.shape[0], self._stream()))
        return states

    # ---- buffers ------------------------------------------------------------------------------
    def alloc_mask(self, n: int, fmt: str = "bytes") -> torch.Tensor:
        """Raw mask buffer: uint8 [n, mask_bytes] (rows padded to 16 B) or int32 [n, mask_words]."""
        if fmt == "bytes":
            return torch.empty((n, self.mask_bytes), dtype=torch.uint8, device=self.device)
        if fmt == "bits":
            return torch.empty((n, self.mask_words), dtype=torch.int32, device=self.device)
        if fmt == "indices":      # ascending legal ids (uint16 carried as int16), legal_count[i] valid entries per row
            return torch.empty((n, self.max_legal), dtype=torch.int16, device=self.device)
        raise ValueError("fmt must be 'bytes', 'bits' or 'indices'")

    def mask_view(self, raw: torch.Tensor) -> torch.Tensor:
        """bool [n, A] view of a raw byte-mask buffer (what torch.masked_select consumes,
        blokus_rl/neural_network.py:169)."""
        return raw.view(torch.bool)[:, : self.num_actions]

    # ---- a2/a3/a4 next_state + valid_actions + get_winners: blokus_wrapper.py:89-132, 164-186 --------
    def step(self, states: torch.Tensor, actions: torch.Tensor | None = None, *, out_states: torch.Tensor | None = None,
             mask: str | torch.Tensor | None = "bytes", want_count: bool = True, want_terminal: bool = True,
             want_scores: bool = True, sample: bool = False, seed: int = 0, env_id_base: int = 0,
             auto_reset: bool = False, buffers: StepOut | None = None,
             obs: torch.Tensor | bool | None = None, warp_kernels: bool = False,
             state_index: torch.Tensor | None = None) -> StepOut:
        """Apply ``actions`` (or none: mask-only), resolve the next mover with auto-skip, detect the end
        of the game and emit the next mover's full legal mask.  In-place on ``states`` unless
        ``out_states`` is given (functional use, as MCTS needs: blokus_rl/alphazero/mcts.py:47).  ``obs`` (a
        float32 ``[n, 2P, N, N]`` tensor, or True to allocate one) additionally receives ``canonical_board`` of the
        resulting states from the same kernel (blokus_wrapper.py:144-146): leaf expansion in one launch.
        ``warp_kernels`` selects the warp-per-env kernels on boards that also have thread-per-env kernels (N <= 7):
        same results, for cross-checks."""
        self._check_states(states)
        n = states.shape[0]
        dev = self.device
        if state_index is not None:
            # env i reads states[state_index[i]] (a search tree stepping states out of its node pool: no gather pass)
            if state_index.dtype != torch.int32 or state_index.dim() != 1 or state_index.device != dev or not state_index.is_contiguous():
                raise ValueError("state_index must be a contiguous int32 [n] CUDA tensor")
            if out_states is None or out_states.data_ptr() == states.data_ptr():
                raise ValueError("state_index needs separate out_states")
            n = state_index.shape[0]
        if actions is not None:
            if actions.dtype != torch.int32 or actions.shape != (n,) or actions.device != dev or not actions.is_contiguous():
                raise ValueError("actions must be a contiguous int32 [n] CUDA tensor")
        out_states = states if out_states is None else out_states
        self._check_states(out_states)
        if out_states.shape[0] != n:
            raise ValueError("out_states must have one row per env")
        b = buffers
        raw_mask, fmt, stride = None, BLK_MASK_NONE, 0
        csr_cursor = csr_offset = None
        if isinstance(mask, str) and mask == "csr":
            # compact index lists: one flat uint16 array for the whole batch (blk_step_args.csr_cursor / csr_offset)
            raw_mask = getattr(b, "mask_raw", None) if b is not None else None
            if raw_mask is None:
                raw_mask = torch.empty(n * min(self.max_legal, 512), dtype=torch.int16, device=dev)
            if raw_mask.dtype != torch.int16 or raw_mask.dim() != 1 or not raw_mask.is_contiguous() or raw_mask.device != dev:
                raise ValueError("mask='csr' needs a flat contiguous int16 CUDA buffer in buffers.mask_raw")
            csr_cursor = getattr(b, "csr_cursor", None) if b is not None else None
            csr_cursor = torch.zeros(1, dtype=torch.int64, device=dev) if csr_cursor is None else csr_cursor.zero_()
            csr_offset = getattr(b, "csr_offset", None) if b is not None else None
            if csr_offset is None:
                csr_offset = torch.empty(n, dtype=torch.int64, device=dev)
            if csr_offset.dtype != torch.int64 or csr_offset.shape[0] < n or csr_offset.device != dev:
                raise ValueError("buffers.csr_offset must be an int64 [n] CUDA tensor")
            fmt, stride, want_count, mask = BLK_MASK_INDICES, raw_mask.numel(), True, None
        if isinstance(mask, torch.Tensor):
            raw_mask = mask
        elif mask is not None:
            want_dtype = {"bytes": torch.uint8, "bits": torch.int32, "indices": torch.int16}.get(mask)
            if want_dtype is None:
                raise ValueError("mask must be 'bytes', 'bits', 'indices', a tensor or None")
            raw_mask = b.mask_raw if b is not None and getattr(b, "mask_raw", None) is not None else None
            if raw_mask is not None and raw_mask.dtype != want_dtype:
                raise ValueError(f"buffers hold a {raw_mask.dtype} mask but mask={mask!r} was requested")
            if raw_mask is None:
                raw_mask = self.alloc_mask(n, mask)
        if raw_mask is not None and csr_cursor is None:
            if raw_mask.dtype == torch.int32:
                fmt, stride = BLK_MASK_BITS, raw_mask.stride(0)
            elif raw_mask.dtype in (torch.uint8, torch.bool):
                fmt, stride = BLK_MASK_BYTES, raw_mask.stride(0)
            elif raw_mask.dtype == torch.int16:
                fmt, stride = BLK_MASK_INDICES, raw_mask.stride(0)
                want_count = True
            else:
                raise ValueError("mask buffer must be uint8/bool (bytes), int32 (bits) or int16 (indices)")
            if raw_mask.shape[0] != n or raw_mask.stride(1) != 1:
                raise ValueError("mask buffer must have n rows with unit column stride")
            if fmt == BLK_MASK_BYTES and raw_mask.shape[1] < self.num_actions:
                raise ValueError("byte-mask rows must hold num_actions bytes")

        def buf(name, shape, dtype, want):
            if not want:
                return None
            t = getattr(b, name, None) if b is not None else None
            if t is None:
                return torch.empty(shape, dtype=dtype, device=dev)
            # a reused buffer made for a smaller batch would be written past its end by the kernel
            if t.dtype != dtype or t.device != dev or not t.is_contiguous() or t.dim() != len(shape) \
                    or t.shape[0] < n or tuple(t.shape[1:]) != tuple(shape[1:]):
                raise ValueError(f"buffers.{name} must be a contiguous {dtype} tensor of shape >= {tuple(shape)} on {dev}")
            return t

        P = self.num_players
        legal_count = buf("legal_count", (n,), torch.int32, want_count)
        terminal = buf("terminal", (n, P), torch.float32, want_terminal)
        flags = buf("flags", (n,), torch.uint8, True)
        scores = buf("scores", (n, P), torch.int16, want_scores)
        next_action = buf("next_action", (n,), torch.int32, sample)
        N = self.board_size
        if obs is True:
            obs = getattr(b, "obs", None) if b is not None else None
            if obs is None:
                obs = torch.empty((n, 2 * P, N, N), dtype=torch.float32, device=dev)
        elif obs is False:
            obs = None
        if obs is not None and (obs.dtype != torch.float32 or obs.device != dev or not obs.is_contiguous()
                                or obs.numel() < n * 2 * P * N * N):
            raise ValueError("obs must be a contiguous float32 [n, 2P, N, N] CUDA tensor")
        args = _lib.BlkStepArgs(
            n, states.data_ptr(), out_states.data_ptr(), None if actions is None else actions.data_ptr(),
            None if raw_mask is None else raw_mask.data_ptr(), fmt, stride,
            None if legal_count is None else legal_count.data_ptr(),
            None if terminal is None else terminal.data_ptr(), flags.data_ptr(),
            None if scores is None else scores.data_ptr(),
            None if next_action is None else next_action.data_ptr(),
            seed & 0xFFFFFFFFFFFFFFFF, env_id_base & 0xFFFFFFFF, (BLK_OPT_AUTO_RESET if auto_reset else 0) | (BLK_OPT_WARP_KERNELS if warp_kernels else 0),
            None if obs is None else obs.data_ptr(), None if state_index is None else state_index.data_ptr(),
            None if csr_cursor is None else csr_cursor.data_ptr(), None if csr_offset is None else csr_offset.data_ptr())
        _lib.check(self._lib.blk_step(self._h, C.byref(args), self._stream()))
        view = None
        if csr_cursor is not None:
            return StepOut(out_states, raw_mask, legal_count, terminal, flags, scores, next_action, raw_mask, obs, csr_offset, csr_cursor)
        if raw_mask is not None:
            view = self.mask_view(raw_mask) if fmt == BLK_MASK_BYTES and raw_mask.shape[1] >= self.num_actions and \
                raw_mask.dtype == torch.uint8 else raw_mask
        return StepOut(out_states, view, legal_count, terminal, flags, scores, next_action, raw_mask, obs)

    def legal_mask(self, states: torch.Tensor, fmt: str = "bytes", **kw) -> StepOut:
        """valid_actions for the side to move of every state (blokus_wrapper.py:108-132)."""
        return self.step(states, None, mask=fmt, **kw)

    def make_buffers(self, n: int, fmt: str | None = "bytes", sample: bool = False) -> StepOut:
        """Preallocate every per-step output once (rollout collection loops reuse them)."""
        P, dev = self.num_players, self.device
        raw = None if fmt is None else self.alloc_mask(n, fmt)
        b = StepOut(None, None, torch.empty(n, dtype=torch.int32, device=dev),
                    torch.empty((n, P), dtype=torch.float32, device=dev), torch.empty(n, dtype=torch.uint8, device=dev),
                    torch.empty((n, P), dtype=torch.int16, device=dev),
                    torch.empty(n, dtype=torch.int32, device=dev) if sample else None, raw)
        return b

    # ---- a5 canonical_board: blokus_wrapper.py:144-146 ---------------------------------------------
    def observe(self, states: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        self._check_states(states)
        n, P, N = states.shape[0], self.num_players, self.board_size
        if out is None:
            out = torch.empty((n, 2 * P, N, N), dtype=torch.float32, device=self.device)
        _lib.check(self._lib.blk_observe(self._h, _ptr(states), _ptr(out), n, self._stream()))
        return out

    # ---- a6 board_contents: blokus_wrapper.py:208-218 ------------------------------------------------
    def board_contents(self, states: torch.Tensor) -> torch.Tensor:
        self._check_states(states)
        n, N = states.shape[0], self.board_size
        out = torch.empty((n, N, N), dtype=torch.uint8, device=self.device)
        _lib.check(self._lib.blk_board_contents(self._h, _ptr(states), _ptr(out), n, self._stream()))
        return out

    # ---- a4 get_winners without stepping: blokus_wrapper.py:164-186 -----------------------------------
    def game_ended(self, states: torch.Tensor):
        self._check_states(states)
        n, P = states.shape[0], self.num_players
        flags = torch.empty(n, dtype=torch.uint8, device=self.device)
        terminal = torch.empty((n, P), dtype=torch.float32, device=self.device)
        scores = torch.empty((n, P), dtype=torch.int16, device=self.device)
        _lib.check(self._lib.blk_game_ended(self._h, _ptr(states), _ptr(flags), _ptr(terminal), _ptr(scores), n,
                                            self._stream()))
        return flags, terminal, scores

    # ---- kernel family 4: uniform-random playouts to the end of the game ---------------------------------
    def rollout(self, roots: torch.Tensor, per_root: int, seed: int = 0, rollout_id_base: int = 0,
                log_actions: bool = False, stop_player: int = -1, out_states: torch.Tensor | None = None,
                warp_kernels: bool = False) -> RolloutOut:
        """Uniform-random playouts.  ``stop_player = q`` stops each playout as soon as it is player q's turn (or
        the game is over) and ``out_states`` (may alias ``roots`` when ``per_root == 1``) receives the states
        reached: this is how the gym adapter plays the random-bot opponents inside one ``step``."""
        self._check_states(roots)
        n, P, dev = roots.shape[0], self.num_players, self.device
        total = n * per_root
        final_scores = torch.empty((n, per_root, P), dtype=torch.int16, device=dev)
        winners = torch.empty((n, per_root), dtype=torch.uint8, device=dev)
        value_sum = torch.zeros((n, P), dtype=torch.float32, device=dev)
        plies = torch.empty((n, per_root), dtype=torch.int32, device=dev)
        log_stride = 88
        log = torch.empty((n, per_root, log_stride), dtype=torch.int16, device=dev) if log_actions else None
        if out_states is not None:
            if out_states.dtype != torch.int32 or not out_states.is_contiguous() or out_states.device != dev \
                    or out_states.numel() != total * self.state_words:
                raise ValueError(f"out_states must be a contiguous int32 [n * per_root, state_words] tensor on {dev}")
            if per_root > 1 and out_states.untyped_storage().data_ptr() == roots.untyped_storage().data_ptr():
                raise ValueError("out_states may alias roots only when per_root == 1 (playouts of one root would overwrite it)")
        args = _lib.BlkRolloutArgs(n, roots.data_ptr(), per_root, seed & 0xFFFFFFFFFFFFFFFF, rollout_id_base & 0xFFFFFFFF,
                                   final_scores.data_ptr(), winners.data_ptr(), value_sum.data_ptr(),
                                   None if log is None else log.data_ptr(), log_stride, plies.data_ptr(),
                                   int(stop_player), None if out_states is None else out_states.data_ptr(),
                                   BLK_OPT_WARP_KERNELS if warp_kernels else 0)
        if total:
            _lib.check(self._lib.blk_rollout(self._h, C.byref(args), self._stream()))
        return RolloutOut(final_scores, winners, value_sum, plies, log)
