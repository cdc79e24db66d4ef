"""Pure-Python, set-based Blokus rules.  TEST INFRASTRUCTURE ONLY (third, slowest oracle).

Parity pinned only on the reference's rendered games (tests/test_ref_render_golden.py): the reference env (colosseumrl / blokus-gym, /root/reference/setup.py:11,33) is
absent; this restates SURVEY.md Appendix A rules R1-R11 at the level of cell sets, with its own
polyomino enumeration (it does NOT import blokus_rl_b200.tables), so that it can check both the C
oracle and the table generator.  Small cases only.
"""
from __future__ import annotations

SHAPES = ["#", "#/#", "#/#/#", "##/#.", "#/#/#/#", "##/#./#.", "#./##/#.", "##/##", "#./##/.#",
          "#/#/#/#/#", "##/#./#./#.", "#./##/#./#.", "##/##/#.", "##/#./##", "###/#../#..",
          "#../###/#..", "#./#./##/.#", "#../###/.#.", "#../##./.##", "#../###/..#", ".#./###/.#."]


def _variants(shape: str):
    cells = [(y, x) for y, row in enumerate(shape.split("/")) for x, c in enumerate(row) if c == "#"]
    seen = set()
    for flip in (False, True):
        cur = [(y, -x) if flip else (y, x) for y, x in cells]
        for _ in range(4):
            cur = [(x, -y) for y, x in cur]
            my, mx = min(y for y, _ in cur), min(x for _, x in cur)
            seen.add(tuple(sorted((y - my, x - mx) for y, x in cur)))
    return sorted(seen)


ORIENTS = [(p, v) for p, s in enumerate(SHAPES) for v in _variants(s)]


def all_actions(n: int):
    """[(piece, frozenset(cells))] in canonical id order (piece, orientation, ay, ax)."""
    out = []
    for p, v in ORIENTS:
        h = 1 + max(y for y, _ in v)
        w = 1 + max(x for _, x in v)
        for ay in range(n - h + 1):
            for ax in range(n - w + 1):
                out.append((p, frozenset((ay + y, ax + x) for y, x in v)))
    return out


def corners(n: int, players: int):
    m = n - 1
    return [(0, 0), (m, m)] if players == 2 else [(0, 0), (0, m), (m, 0), (m, m)]


class NaiveGame:
    def __init__(self, n: int, players: int):
        self.n, self.P = n, players
        self.actions = all_actions(n)
        self.cells = [set() for _ in range(players)]     # cells owned per player
        self.hand = [set(range(21)) for _ in range(players)]
        self.mover, self.done, self.ply = 0, False, 0

    def legal(self, p: int):
        if self.done:
            return []
        own = self.cells[p]
        occupied = set().union(*self.cells)
        first = len(self.hand[p]) == 21
        cy, cx = corners(self.n, self.P)[p]
        out = []
        for a, (piece, cs) in enumerate(self.actions):
            if piece not in self.hand[p] or cs & occupied:
                continue
            if any((y + dy, x + dx) in own for y, x in cs for dy, dx in ((1, 0), (-1, 0), (0, 1), (0, -1))):
                continue
            if first:
                ok = (cy, cx) in cs
            else:
                ok = any((y + dy, x + dx) in own for y, x in cs for dy in (-1, 1) for dx in (-1, 1))
            if ok:
                out.append(a)
        return out

    def step(self, a: int):
        p = self.mover
        piece, cs = self.actions[a]
        assert a in self.legal(p)
        self.cells[p] |= cs
        self.hand[p].discard(piece)
        self.ply += 1
        for k in range(1, self.P + 1):
            q = (p + k) % self.P
            if self.legal(q):
                self.mover = q
                return
        self.done = True

    def scores(self):
        return [len(c) for c in self.cells]

    def board(self):
        b = [[0] * self.n for _ in range(self.n)]
        for p, cs in enumerate(self.cells):
            for y, x in cs:
                b[y][x] = p + 1
        return b
