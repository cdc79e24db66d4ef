"""ctypes front-end of the C oracle (oracle/blokus_oracle.c).  TEST INFRASTRUCTURE ONLY.

Imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.  Parity is pinned only on
the reference's rendered games (tests/test_ref_render_golden.py); see the header of blokus_oracle.c for what stays
unpinned.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libblokus_oracle.so"


def build(force: bool = False) -> Path:
    src = _HERE / "blokus_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-B", "libblokus_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Oracle:
    """One oracle instance == one (N, P, score_rule) configuration (the C side keeps it in globals,
    so the library is loaded privately per instance)."""

    def __init__(self, board_size: int = 20, num_players: int = 4, score_rule: int = 0):
        build()
        # private copy of the mapping so two configurations can coexist in one process
        import shutil
        import tempfile
        self._tmp = tempfile.NamedTemporaryFile(suffix=".so", delete=False)
        self._tmp.close()
        shutil.copyfile(_LIB_PATH, self._tmp.name)
        self.lib = lib = C.CDLL(self._tmp.name)
        os.unlink(self._tmp.name)
        self.N, self.P = board_size, num_players
        lib.orc_random_play.restype = C.c_int64
        lib.orc_random_play.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p]
        lib.orc_sample_action.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
        lib.orc_reset.argtypes = [C.c_void_p, C.c_uint32]
        lib.orc_philox.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
        for name in ("orc_legal_mask", "orc_fast_legal_mask"):
            getattr(lib, name).argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        for name in ("orc_step", "orc_fast_step"):
            getattr(lib, name).argtypes = [C.c_void_p, C.c_int]
        lib.orc_action_cells.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        lib.orc_winners.argtypes = [C.c_void_p]
        lib.orc_final_score.argtypes = [C.c_void_p, C.c_int]
        lib.orc_terminal_values.argtypes = [C.c_void_p, C.c_void_p]
        lib.orc_observe.argtypes = [C.c_void_p, C.c_void_p]
        lib.orc_board_contents.argtypes = [C.c_void_p, C.c_void_p]
        lib.orc_pack.argtypes = [C.c_void_p, C.c_void_p]
        lib.orc_unpack.argtypes = [C.c_void_p, C.c_void_p]
        if lib.orc_init(board_size, num_players, score_rule) != 0:
            raise ValueError("unsupported oracle configuration")
        self.A = lib.orc_num_actions()
        self.state_size = lib.orc_state_size()
        self.state_words = lib.orc_state_words()

    # --- states are opaque byte buffers -------------------------------------------------
    def new_state(self, game: int = 0):
        s = C.create_string_buffer(self.state_size)
        self.lib.orc_reset(s, game)
        return s

    def copy(self, s):
        return C.create_string_buffer(s.raw, self.state_size)

    def reset(self, s, game: int = 0):
        self.lib.orc_reset(s, game)

    def field(self, s, name: str):
        # mirrors orc_state layout
        off = 8 + 400
        raw = s.raw
        if name == "board":
            return np.frombuffer(raw, np.uint8, 400, 8)[: self.N * self.N].reshape(self.N, self.N).copy()
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
        fn(s, p, m.ctypes.data)
        return m

    def step(self, s, action: int, fast: bool = False) -> int:
        return (self.lib.orc_fast_step if fast else self.lib.orc_step)(s, int(action))

    def winners(self, s) -> int:
        return self.lib.orc_winners(s)

    def final_scores(self, s) -> np.ndarray:
        return np.array([self.lib.orc_final_score(s, p) for p in range(self.P)], np.int16)

    def terminal_values(self, s) -> np.ndarray:
        v = np.zeros(self.P, np.float32)
        self.lib.orc_terminal_values(s, v.ctypes.data)
        return v

    def observe(self, s) -> np.ndarray:
        o = np.zeros((2 * self.P, self.N, self.N), np.float32)
        self.lib.orc_observe(s, o.ctypes.data)
        return o

    def board_contents(self, s) -> np.ndarray:
        b = np.zeros((self.N, self.N), np.uint8)
        self.lib.orc_board_contents(s, b.ctypes.data)
        return b

    def action_cells(self, a: int):
        cells = np.zeros(10, np.uint8)
        meta = np.zeros(4, np.int32)
        n = self.lib.orc_action_cells(int(a), cells.ctypes.data, meta.ctypes.data)
        return [(int(cells[2 * i]), int(cells[2 * i + 1])) for i in range(n)], meta

    # --- engine state format ------------------------------------------------------------------
    def pack(self, s) -> np.ndarray:
        w = np.zeros(self.state_words, np.uint32)
        self.lib.orc_pack(s, w.ctypes.data)
        return w

    def unpack(self, words: np.ndarray):
        w = np.ascontiguousarray(words, np.uint32)
        s = C.create_string_buffer(self.state_size)
        self.lib.orc_unpack(w.ctypes.data, s)
        return s

    # --- RNG / random play ------------------------------------------------------------------------
    def philox(self, ctr, key) -> np.ndarray:
        out = np.zeros(4, np.uint32)
        self.lib.orc_philox(*[int(c) for c in ctr], int(key[0]), int(key[1]), out.ctypes.data)
        return out

    def sample_action(self, s, seed: int, env_id: int, stream: int = 0, fast: bool = True) -> int:
        scratch = np.zeros(self.A, np.uint8)
        return self.lib.orc_sample_action(s, seed, env_id, stream, scratch.ctypes.data, int(fast))

    def random_play(self, s, seed: int, env_id: int, plies: int, auto_reset: bool = True, fast: bool = True,
                    log: bool = True):
        actions = np.full(plies, -1, np.int32) if log else None
        counters = np.zeros(4, np.int64)
        n = self.lib.orc_random_play(s, seed, env_id, plies, int(auto_reset), int(fast),
                                     actions.ctypes.data if log else None, counters.ctypes.data)
        return int(n), actions, counters
