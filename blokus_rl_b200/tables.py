"""Piece, orientation and action tables for the Blokus engine (host side, numpy only).

This module is the *definition* of the canonical action order that the CUDA engine, the
C-ABI and the adapters share:

* 21 free polyominoes of size 1..5 (reference rule R4: the 30,433-action space quoted at
  ``blokus_rl/models/blokus_nnet.py:17`` is exactly the number of in-bounds footprints of
  this set on 20x20).
* every piece expands to its distinct fixed orientations (rotations + reflections),
  normalised to a (0,0) bounding-box origin; orientations of one piece are ordered by
  their sorted ``(dy, dx)`` cell tuples, lexicographically  ->  91 orientations.
* action id = running index over ``(piece, orientation, anchor_y, anchor_x)`` with
  ``anchor_y in [0, N-h]`` and ``anchor_x in [0, N-w]``  (reference a7,
  ``blokus_rl/colossumrl/blokus_wrapper.py:281-324``: one id per distinct footprint,
  legality-agnostic, no pass action).

``tools/gen_orient_inc.py`` turns the same tables into the X-macro list compiled into the
kernels; ``oracle/blokus_oracle.c`` re-derives them independently in C and the tests
compare all three.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache

import numpy as np

# Shape rows top->bottom, '#' = cell.  Order follows SURVEY.md Appendix B.
PIECE_SHAPES: tuple[tuple[str, ...], ...] = (
    ("#",),
    ("#", "#"),
    ("#", "#", "#"),
    ("##", "#."),
    ("#", "#", "#", "#"),
    ("##", "#.", "#."),
    ("#.", "##", "#."),
    ("##", "##"),
    ("#.", "##", ".#"),
    ("#", "#", "#", "#", "#"),
    ("##", "#.", "#.", "#."),
    ("#.", "##", "#.", "#."),
    ("##", "##", "#."),
    ("##", "#.", "##"),
    ("###", "#..", "#.."),
    ("#..", "###", "#.."),
    ("#.", "#.", "##", ".#"),
    ("#..", "###", ".#."),
    ("#..", "##.", ".##"),
    ("#..", "###", "..#"),
    (".#.", "###", ".#."),
)
PIECE_NAMES = ("I1", "I2", "I3", "V3", "I4", "L4", "T4", "O4", "S4", "I5", "L5", "Y5",
               "P5", "U5", "V5", "T5", "N5", "F5", "W5", "Z5", "X5")
NUM_PIECES = len(PIECE_SHAPES)
MAX_CELLS = 5


def _cells_of(shape: tuple[str, ...]) -> frozenset[tuple[int, int]]:
    return frozenset((y, x) for y, row in enumerate(shape) for x, c in enumerate(row) if c == "#")


def _normalise(cells) -> tuple[tuple[int, int], ...]:
    my = min(y for y, _ in cells)
    mx = min(x for _, x in cells)
    return tuple(sorted((y - my, x - mx) for y, x in cells))


def _orientations_of(cells) -> list[tuple[tuple[int, int], ...]]:
    out = set()
    cur = list(cells)
    for _ in range(4):
        cur = [(x, -y) for y, x in cur]          # rotate 90 degrees
        out.add(_normalise(cur))
        out.add(_normalise([(y, -x) for y, x in cur]))  # mirrored
    return sorted(out)


@dataclass(frozen=True)
class Orientation:
    index: int            # global orientation index 0..90
    piece: int            # piece index 0..20
    local: int            # index among this piece's orientations
    cells: tuple[tuple[int, int], ...]   # sorted (dy, dx)
    h: int
    w: int

    @property
    def ncells(self) -> int:
        return len(self.cells)


@lru_cache(maxsize=None)
def orientations() -> tuple[Orientation, ...]:
    out: list[Orientation] = []
    for p, shape in enumerate(PIECE_SHAPES):
        for j, cells in enumerate(_orientations_of(_cells_of(shape))):
            h = 1 + max(y for y, _ in cells)
            w = 1 + max(x for _, x in cells)
            out.append(Orientation(len(out), p, j, cells, h, w))
    return tuple(out)


def piece_sizes() -> np.ndarray:
    return np.array([len(_cells_of(s)) for s in PIECE_SHAPES], dtype=np.int32)


@dataclass(frozen=True)
class ActionTable:
    board_size: int
    num_actions: int
    orient_base: np.ndarray     # [n_orient+1] first action id of each orientation
    action_orient: np.ndarray   # [A] orientation index
    action_piece: np.ndarray    # [A]
    action_y: np.ndarray        # [A] anchor row
    action_x: np.ndarray        # [A] anchor column


@lru_cache(maxsize=None)
def action_table(board_size: int) -> ActionTable:
    n = board_size
    base = [0]
    ao, ap, ay, ax = [], [], [], []
    for o in orientations():
        rows = max(0, n - o.h + 1)
        cols = max(0, n - o.w + 1)
        for y in range(rows):
            for x in range(cols):
                ao.append(o.index)
                ap.append(o.piece)
                ay.append(y)
                ax.append(x)
        base.append(len(ao))
    return ActionTable(
        n, len(ao), np.array(base, dtype=np.int32), np.array(ao, dtype=np.int16),
        np.array(ap, dtype=np.int16), np.array(ay, dtype=np.int16), np.array(ax, dtype=np.int16))


def action_cells(board_size: int, action: int) -> list[tuple[int, int]]:
    """Absolute (row, col) cells covered by ``action`` on a ``board_size`` board."""
    t = action_table(board_size)
    o = orientations()[int(t.action_orient[action])]
    y0, x0 = int(t.action_y[action]), int(t.action_x[action])
    return [(y0 + dy, x0 + dx) for dy, dx in o.cells]


def action_index(board_size: int, orientation: int, y: int, x: int) -> int:
    """The opaque ``index`` of an action string: ``orientation * N*N + y * N + x``.  Orientation-major on
    purpose: the reference enumerates ``{piece: {index: [orientation...]}}`` in dict order and numbers
    the distinct footprints in first-seen order (``blokus_wrapper.py:300-316``); with this key that order
    IS the canonical id order, so the ids the reference would build equal the engine's."""
    return (orientation * board_size + y) * board_size + x


def action_to_string(piece_type: int, index: int, orientation: int) -> str:
    """Engine-defined action string, the analogue of ``colosseumrl.envs.blokus.action_to_string`` used at
    ``blokus_rl/colossumrl/blokus_wrapper.py:310-312``: ``"<piece>;<index>;<orientation-within-piece>"``.
    The real format is not recoverable (dependency absent); the wrapper treats the
    string as an opaque dict key, so any injective format is a drop-in."""
    return f"{piece_type};{index};{orientation}"


@lru_cache(maxsize=None)
def action_strings(board_size: int) -> tuple[str, ...]:
    t = action_table(board_size)
    ors = orientations()
    out = []
    for a in range(t.num_actions):
        loc = ors[int(t.action_orient[a])].local
        out.append(action_to_string(int(t.action_piece[a]),
                                    action_index(board_size, loc, int(t.action_y[a]), int(t.action_x[a])), loc))
    return tuple(out)


@lru_cache(maxsize=None)
def string_to_action(board_size: int) -> dict:
    return {s: i for i, s in enumerate(action_strings(board_size))}


def write_action_json(path, board_size: int) -> None:
    """``states/colosseum_{N}_players_{P}.json`` in the reference's cache format ``{action_string: id}``
    (``blokus_wrapper.py:45-50, 284-288, 321-322``); with it in place the wrapper skips its own enumeration."""
    import json
    from pathlib import Path
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w", encoding="utf-8") as f:
        json.dump(string_to_action(board_size), f)


def start_corners(board_size: int, num_players: int) -> list[tuple[int, int]]:
    """Start corner (row, col) per player (rule R3)."""
    n = board_size - 1
    if num_players == 2:
        return [(0, 0), (n, n)]
    return [(0, 0), (0, n), (n, 0), (n, n)][:num_players]
