"""ctypes front-end of the C oracle (oracle/blokus_oracle.c).  TEST INFRASTRUCTURE ONLY.

Imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.  Parity is pinned only on
the reference's rendered games (tests/test_ref_render_golden.py); see the header of blokus_oracle.c for what stays
unpinned.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libblokus_oracle.so"
_lib = None

_BOARD_OFF, _TAIL_OFF = 8, 8 + 400        # orc_state: N, P | board[400] | inv, score, lastmono, mover, done, ply, game | rows


def build(force: bool = False) -> Path:
    src = _HERE / "blokus_oracle.c"
    mk = _HERE / "Makefile"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < max(src.stat().st_mtime, mk.stat().st_mtime):
        subprocess.check_call(["make", "-C", str(_HERE), "-B", "libblokus_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _load() -> C.CDLL:
    """The library is loaded in place, once per process (configurations live in orc_ctx objects, not in globals)."""
    global _lib
    if _lib is not None:
        return _lib
    if os.access(_HERE, os.W_OK) and (_HERE / "blokus_oracle.c").exists():
        try:
            build()
        except (OSError, subprocess.CalledProcessError):
            if not _LIB_PATH.exists():
                raise
    lib = C.CDLL(str(_LIB_PATH))
    vp, i32, u32, u64, i64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_int64
    lib.orc_create.restype = vp
    lib.orc_create.argtypes = [i32, i32, i32]
    lib.orc_destroy.argtypes = [vp]
    lib.orc_destroy.restype = None
    for name in ("orc_num_actions", "orc_num_orientations", "orc_num_fields", "orc_state_words"):
        getattr(lib, name).argtypes = [vp]
    lib.orc_piece_size.argtypes = [vp, i32]
    lib.orc_action_cells.argtypes = [vp, i32, vp, vp]
    lib.orc_reset.argtypes = [vp, vp, u32]
    lib.orc_reset.restype = None
    for name in ("orc_legal_mask", "orc_fast_legal_mask"):
        getattr(lib, name).argtypes = [vp, vp, i32, vp]
    for name in ("orc_step", "orc_fast_step"):
        getattr(lib, name).argtypes = [vp, vp, i32]
    lib.orc_final_score.argtypes = [vp, vp, i32]
    lib.orc_winners.argtypes = [vp, vp]
    lib.orc_rows_consistent.argtypes = [vp, vp]
    for name in ("orc_terminal_values", "orc_observe", "orc_board_contents", "orc_pack", "orc_unpack"):
        getattr(lib, name).argtypes = [vp, vp, vp]
        getattr(lib, name).restype = None
    lib.orc_philox.argtypes = [u32] * 6 + [vp]
    lib.orc_philox.restype = None
    lib.orc_sample_action.argtypes = [vp, vp, u64, u32, u32, vp, i32]
    lib.orc_random_play.restype = i64
    lib.orc_random_play.argtypes = [vp, vp, u64, u32, i32, i32, i32, vp, vp]
    lib.orc_play_many.restype = None
    lib.orc_play_many.argtypes = [vp, vp, i64, u64, u32, u32, i32, i32, vp, i64, vp, vp, vp, i32, vp, vp, vp]
    lib.orc_playout.argtypes = [vp, vp, u64, u32, i32, vp, vp, vp, vp, i32]
    lib.orc_playout_stream.argtypes = [vp, vp, u64, u32, u32, i32, vp, vp, vp, vp, i32]
    _lib = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data


class Oracle:
    """One oracle instance == one (N, P, score_rule) configuration (an immutable orc_ctx: thread-safe)."""

    def __init__(self, board_size: int = 20, num_players: int = 4, score_rule: int = 0):
        self.lib = lib = _load()
        self.N, self.P = board_size, num_players
        self.ctx = lib.orc_create(board_size, num_players, score_rule)
        if not self.ctx:
            raise ValueError("unsupported oracle configuration")
        self.A = lib.orc_num_actions(self.ctx)
        self.state_size = lib.orc_state_size()
        self.state_words = lib.orc_state_words(self.ctx)

    def __del__(self):  # pragma: no cover - best effort
        try:
            if self.ctx:
                self.lib.orc_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # --- states are opaque byte buffers -------------------------------------------------
    def new_state(self, game: int = 0):
        s = C.create_string_buffer(self.state_size)
        self.lib.orc_reset(self.ctx, s, game)
        return s

    def copy(self, s):
        return C.create_string_buffer(s.raw, self.state_size)

    def reset(self, s, game: int = 0):
        self.lib.orc_reset(self.ctx, s, game)

    def field(self, s, name: str):
        # mirrors orc_state layout
        off = _TAIL_OFF
        raw = s.raw
        if name == "board":
            return np.frombuffer(raw, np.uint8, 400, _BOARD_OFF)[: self.N * self.N].reshape(self.N, self.N).copy()
        if name == "inv":
            return np.frombuffer(raw, np.uint32, 4, off).copy()
        if name == "score":
            return np.frombuffer(raw, np.int16, 4, off + 16).copy()
        if name == "lastmono":
            return np.frombuffer(raw, np.uint8, 4, off + 24).copy()
        if name == "mover":
            return raw[off + 28]
        if name == "done":
            return raw[off + 29]
        if name == "ply":
            return int(np.frombuffer(raw, np.uint16, 1, off + 30)[0])
        if name == "game":
            return int(np.frombuffer(raw, np.uint32, 1, off + 32)[0])
        raise KeyError(name)

    # --- rules -----------------------------------------------------------------------------
    def legal_mask(self, s, player: int | None = None, fast: bool = False) -> np.ndarray:
        p = self.field(s, "mover") if player is None else player
        m = np.zeros(self.A, np.uint8)
        fn = self.lib.orc_fast_legal_mask if fast else self.lib.orc_legal_mask
        fn(self.ctx, s, p, m.ctypes.data)
        return m

    def step(self, s, action: int, fast: bool = False) -> int:
        return (self.lib.orc_fast_step if fast else self.lib.orc_step)(self.ctx, s, int(action))

    def winners(self, s) -> int:
        return self.lib.orc_winners(self.ctx, s)

    def final_scores(self, s) -> np.ndarray:
        return np.array([self.lib.orc_final_score(self.ctx, s, p) for p in range(self.P)], np.int16)

    def terminal_values(self, s) -> np.ndarray:
        v = np.zeros(self.P, np.float32)
        self.lib.orc_terminal_values(self.ctx, s, v.ctypes.data)
        return v

    def observe(self, s) -> np.ndarray:
        o = np.zeros((2 * self.P, self.N, self.N), np.float32)
        self.lib.orc_observe(self.ctx, s, o.ctypes.data)
        return o

    def board_contents(self, s) -> np.ndarray:
        b = np.zeros((self.N, self.N), np.uint8)
        self.lib.orc_board_contents(self.ctx, s, b.ctypes.data)
        return b

    def rows_consistent(self, s) -> bool:
        return bool(self.lib.orc_rows_consistent(self.ctx, s))

    def action_cells(self, a: int):
        cells = np.zeros(10, np.uint8)
        meta = np.zeros(4, np.int32)
        n = self.lib.orc_action_cells(self.ctx, int(a), cells.ctypes.data, meta.ctypes.data)
        return [(int(cells[2 * i]), int(cells[2 * i + 1])) for i in range(n)], meta

    # --- engine state format ------------------------------------------------------------------
    def pack(self, s) -> np.ndarray:
        w = np.zeros(self.state_words, np.uint32)
        self.lib.orc_pack(self.ctx, s, w.ctypes.data)
        return w

    def unpack(self, words: np.ndarray):
        w = np.ascontiguousarray(words, np.uint32)
        s = C.create_string_buffer(self.state_size)
        self.lib.orc_unpack(self.ctx, w.ctypes.data, s)
        return s

    # --- RNG / random play ------------------------------------------------------------------------
    def philox(self, ctr, key) -> np.ndarray:
        out = np.zeros(4, np.uint32)
        self.lib.orc_philox(*[int(c) for c in ctr], int(key[0]), int(key[1]), out.ctypes.data)
        return out

    def sample_action(self, s, seed: int, env_id: int, stream: int = 0, fast: bool = True) -> int:
        scratch = np.zeros(self.A, np.uint8)
        return self.lib.orc_sample_action(self.ctx, s, seed, env_id, stream, scratch.ctypes.data, int(fast))

    def random_play(self, s, seed: int, env_id: int, plies: int, auto_reset: bool = True, fast: bool = True,
                    log: bool = True):
        actions = np.full(plies, -1, np.int32) if log else None
        counters = np.zeros(4, np.int64)
        n = self.lib.orc_random_play(self.ctx, s, seed, env_id, plies, int(auto_reset), int(fast),
                                     actions.ctypes.data if log else None, counters.ctypes.data)
        return int(n), actions, counters

    # --- many envs at once, on all host threads (the CPU arm of bench.py and of the full-size parity tests) ------
    def new_states(self, n: int) -> np.ndarray:
        """``n`` fresh states as one uint8 [n, state_size] array (rows are orc_state structs)."""
        one = np.frombuffer(self.new_state().raw, np.uint8)
        return np.tile(one, (n, 1))

    def pack_many(self, states: np.ndarray) -> np.ndarray:
        out = np.zeros((states.shape[0], self.state_words), np.uint32)
        for i in range(states.shape[0]):
            self.lib.orc_pack(self.ctx, states[i].ctypes.data, out[i].ctypes.data)
        return out

    def unpack_many(self, words: np.ndarray) -> np.ndarray:
        w = np.ascontiguousarray(words).view(np.uint32)
        out = np.zeros((w.shape[0], self.state_size), np.uint8)
        for i in range(w.shape[0]):
            self.lib.orc_unpack(self.ctx, w[i].ctypes.data, out[i].ctypes.data)
        return out

    def play_many(self, states: np.ndarray, seed: int, plies: int, *, env_id0: int = 0, env_stride: int = 1,
                  auto_reset: bool = True, sel_plies=(), write_masks: bool = False, threads: int | None = None) -> dict:
        """Uniform-random legal play of every env in ``states`` (in place) for ``plies`` plies -- the engine's random-play
        workload, see orc_play_many.  Env i has global id ``env_id0 + i * env_stride``.  Returns the per-env trajectory
        hashes, the per-ply weighted legal-count sums, per-env mask checksums at ``sel_plies`` and the counters."""
        n = states.shape[0]
        assert states.dtype == np.uint8 and states.shape[1] == self.state_size and states.flags.c_contiguous
        threads = max(1, min(threads or len(os.sched_getaffinity(0)), n))
        sel = np.ascontiguousarray(sorted(sel_plies), np.int32)
        nsel = len(sel)
        traj = np.zeros(n, np.uint64)
        ids_sum = np.zeros((n, max(nsel, 1)), np.uint64)
        words_sum = np.zeros((n, max(nsel, 1)), np.uint64)
        cnt_parts = np.zeros((threads, plies + 1), np.uint64)
        ctr_parts = np.zeros((threads, 4), np.int64)
        masks = [np.zeros((1, self.A), np.uint8) for _ in range(threads)] if write_masks else None
        bounds = [n * t // threads for t in range(threads + 1)]
        cpus = sorted(os.sched_getaffinity(0))

        def work(t):
            lo, hi = bounds[t], bounds[t + 1]
            if hi == lo:
                return
            if threads > 1:
                # one worker per host core: left alone, some kernels keep the threads of one process on a few cores
                try:
                    os.sched_setaffinity(0, {cpus[t % len(cpus)]})
                except OSError:
                    pass
            # write_masks: the full byte mask of every step is written (that is the metric), into one row per thread
            # that stays in cache (stride 0)
            self.lib.orc_play_many(self.ctx, states[lo].ctypes.data, hi - lo, seed, env_id0 + lo * env_stride, env_stride,
                                   plies, int(auto_reset), masks[t].ctypes.data if write_masks else None, 0,
                                   traj[lo:].ctypes.data, cnt_parts[t].ctypes.data,
                                   _p(sel), nsel, ids_sum[lo].ctypes.data, words_sum[lo].ctypes.data,
                                   ctr_parts[t].ctypes.data)

        if threads == 1:
            work(0)
        else:
            ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
            [t.start() for t in ts]
            [t.join() for t in ts]
        with np.errstate(over="ignore"):
            cnt_sum = cnt_parts.sum(0, dtype=np.uint64)
        ctr = ctr_parts.sum(0)
        return {"traj": traj, "cnt_sum": cnt_sum, "sel_plies": sel, "ids_sum": ids_sum[:, :nsel], "words_sum": words_sum[:, :nsel],
                "steps": int(ctr[0]), "games": int(ctr[1]), "legal_sum": int(ctr[2]), "threads": threads}

    def playout(self, root, seed: int, game_index: int, stop_player: int = -1, log: bool = False, stream: int = 1):
        """One uniform-random playout with the engine's playout stream: (plies, final scores, winners bitmask,
        end state, action log or None)."""
        out = C.create_string_buffer(self.state_size)
        scores = np.zeros(4, np.int16)
        win = C.c_int32(0)
        alog = np.zeros(88, np.uint16) if log else None
        n = self.lib.orc_playout_stream(self.ctx, root, seed, game_index, stream, stop_player, out, scores.ctypes.data,
                                        C.byref(win), _p(alog), 88)
        return int(n), scores[: self.P].copy(), int(win.value), out, alog
