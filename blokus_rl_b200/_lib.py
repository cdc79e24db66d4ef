"""ctypes binding of the C ABI in include/blokus_b200.h (libblokus_b200.so, built in-tree by
``__graft_entry__.build()`` / ``python -m blokus_rl_b200.build``).

There is NO fallback: if the shared library is missing or no B200 is visible, importing the engine
fails loudly.  The oracle under ``oracle/`` is test infrastructure and is never imported here.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
import os as _os
LIB_PATH = Path(_os.environ.get("BLOKUS_B200_LIB", _PKG / "libblokus_b200.so"))   # override: A/B kernel experiments only

BLK_MASK_NONE, BLK_MASK_BITS, BLK_MASK_BYTES, BLK_MASK_INDICES = 0, 1, 2, 3
BLK_OPT_AUTO_RESET = 1
BLK_OPT_WARP_KERNELS = 2
BLK_FLAG_DONE, BLK_FLAG_ILLEGAL, BLK_FLAG_TRUNCATED = 1, 2, 4
ABI_VERSION = 3

EXPORTS = (
    "blk_last_error", "blk_abi_version", "blk_create", "blk_destroy", "blk_get_info", "blk_action_to_cells",
    "blk_reset", "blk_step", "blk_observe", "blk_board_contents", "blk_game_ended", "blk_rollout",
    "blk_puct_last_error", "blk_puct_select", "blk_puct_expand", "blk_puct_backup", "blk_puct_best", "blk_puct_advance",
    "blk_puct_search", "blk_puct_reroot",
)


class BlkConfig(C.Structure):
    _fields_ = [("board_size", C.c_int32), ("num_players", C.c_int32), ("score_rule", C.c_int32),
                ("device", C.c_int32)]


class BlkInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "board_size", "num_players", "num_actions", "num_pieces", "num_orients", "num_fields",
        "state_words", "mask_words", "mask_bytes", "sm_count")]


class BlkStepArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("state_in", C.c_void_p), ("state_out", C.c_void_p), ("action", C.c_void_p),
        ("mask", C.c_void_p), ("mask_format", C.c_int32), ("mask_stride", C.c_int64),
        ("legal_count", C.c_void_p), ("terminal", C.c_void_p), ("flags", C.c_void_p), ("scores", C.c_void_p),
        ("next_action", C.c_void_p), ("seed", C.c_uint64), ("env_id_base", C.c_uint32), ("options", C.c_uint32),
        ("obs", C.c_void_p), ("state_index", C.c_void_p), ("csr_cursor", C.c_void_p), ("csr_offset", C.c_void_p),
    ]


class BlkRolloutArgs(C.Structure):
    _fields_ = [
        ("n_roots", C.c_int64), ("roots", C.c_void_p), ("per_root", C.c_int32), ("seed", C.c_uint64),
        ("rollout_id_base", C.c_uint32), ("final_scores", C.c_void_p), ("winners", C.c_void_p),
        ("value_sum", C.c_void_p), ("action_log", C.c_void_p), ("log_stride", C.c_int32), ("plies", C.c_void_p),
        ("stop_player", C.c_int32), ("state_out", C.c_void_p), ("options", C.c_uint32),
    ]


class BlkPuctForest(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("num_trees", "num_players", "num_actions", "mask_stride", "node_capacity",
                                          "edge_capacity", "max_depth")] +
                [(n, C.c_void_p) for n in ("node_edge0", "node_nedge", "node_state", "node_mover", "node_terminal",
                                           "node_term_value", "edge_action", "edge_child", "edge_n", "edge_q", "edge_p",
                                           "root", "path", "path_len", "status", "leaf_node", "leaf_edge", "src_slot",
                                           "step_action", "scores", "counters", "node_sum_n", "path_node", "node_uniform",
                                           "hash_table")] +
                [("hash_capacity", C.c_int32)] +
                [(n, C.c_void_p) for n in ("node_hash", "node_tree", "edge_vl", "node_front")])


class BlkPuctSearchArgs(C.Structure):
    _fields_ = [("num_sims", C.c_int32), ("cpuct", C.c_double), ("epsilon_fix", C.c_int32), ("pool", C.c_void_p),
                ("warps_per_tree", C.c_int32), ("playouts_per_leaf", C.c_int32), ("seed", C.c_uint64),
                ("virtual_loss", C.c_double)]


class BlkPuctExpandArgs(C.Structure):
    _fields_ = [("new_slot_base", C.c_int32), ("state_words", C.c_int32), ("meta_word", C.c_int32),
                ("attach_only", C.c_int32), ("new_states", C.c_void_p), ("pool", C.c_void_p), ("mask", C.c_void_p),
                ("mask_bits", C.c_int32), ("mask_stride_words", C.c_int32), ("flags", C.c_void_p),
                ("terminal", C.c_void_p), ("prior", C.c_void_p), ("prior_dtype", C.c_int32), ("prior_stride", C.c_int64),
                ("value", C.c_void_p), ("fuse_backup", C.c_int32)]


class EngineError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load libblokus_b200.so; raise EngineError (never fall back) when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise EngineError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    lib.blk_last_error.restype = C.c_char_p
    lib.blk_create.argtypes = [C.POINTER(BlkConfig), C.POINTER(C.c_void_p)]
    lib.blk_destroy.argtypes = [C.c_void_p]
    lib.blk_destroy.restype = None
    lib.blk_get_info.argtypes = [C.c_void_p, C.POINTER(BlkInfo)]
    lib.blk_action_to_cells.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
    lib.blk_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.blk_step.argtypes = [C.c_void_p, C.POINTER(BlkStepArgs), C.c_void_p]
    lib.blk_observe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.blk_board_contents.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.blk_game_ended.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.blk_rollout.argtypes = [C.c_void_p, C.POINTER(BlkRolloutArgs), C.c_void_p]
    lib.blk_puct_last_error.restype = C.c_char_p
    lib.blk_puct_select.argtypes = [C.POINTER(BlkPuctForest), C.c_double, C.c_int32, C.c_void_p]
    lib.blk_puct_expand.argtypes = [C.POINTER(BlkPuctForest), C.POINTER(BlkPuctExpandArgs), C.c_void_p]
    lib.blk_puct_backup.argtypes = [C.POINTER(BlkPuctForest), C.c_void_p]
    lib.blk_puct_best.argtypes = [C.POINTER(BlkPuctForest), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.blk_puct_advance.argtypes = [C.POINTER(BlkPuctForest), C.c_void_p, C.c_void_p]
    lib.blk_puct_search.argtypes = [C.c_void_p, C.POINTER(BlkPuctForest), C.POINTER(BlkPuctSearchArgs), C.c_void_p]
    lib.blk_puct_reroot.argtypes = [C.c_void_p, C.POINTER(BlkPuctForest), C.POINTER(BlkPuctSearchArgs), C.c_void_p, C.c_void_p]
    if lib.blk_abi_version() != ABI_VERSION and not _os.environ.get("BLOKUS_B200_ANY_ABI"):   # (A/B runs against an older build)
        raise EngineError("libblokus_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise EngineError(f"blokus_b200 error {rc}: {load().blk_last_error().decode()}")
