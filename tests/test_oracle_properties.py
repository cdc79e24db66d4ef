"""Property tests of the rules restatement (CPU): invariances the Blokus rules must have, checked on random
reachable positions with hypothesis.  They guard the oracle itself (the reference's engine cannot be run)."""
import numpy as np
from hypothesis import given, settings, strategies as st


def _play(orc, seed, plies):
    s = orc.new_state()
    for _ in range(plies):
        if orc.field(s, "done"):
            break
        orc.step(s, orc.sample_action(s, seed, 0), fast=True)
    return s


def _legal_footprints(orc, s, player):
    m = orc.legal_mask(s, player, fast=True)
    return {frozenset(orc.action_cells(int(a))[0]) for a in np.flatnonzero(m)}


@settings(max_examples=12, deadline=None)
@given(seed=st.integers(0, 10_000), plies=st.integers(0, 60))
def test_rotating_the_board_180_degrees_permutes_players(oracle20, seed, plies):
    """R3/R5/R6 are symmetric under a half turn that swaps players 0<->3 and 1<->2 (their start corners swap too)."""
    orc = oracle20
    s = _play(orc, seed, plies)
    N = orc.N
    w = orc.pack(s).copy()
    rows = w[: 4 * N].reshape(4, N)
    rot = np.zeros_like(rows)
    for q in range(4):
        for y in range(N):
            v = int(rows[q, y])
            rot[3 - q, N - 1 - y] = sum(((v >> x) & 1) << (N - 1 - x) for x in range(N))
    w2 = w.copy()
    w2[: 4 * N] = rot.reshape(-1)
    w2[4 * N: 4 * N + 4] = w[4 * N: 4 * N + 4][::-1]                    # inventories follow their players
    s2 = orc.unpack(w2)
    for p in range(4):
        a = _legal_footprints(orc, s, p)
        b = {frozenset((N - 1 - y, N - 1 - x) for y, x in fp) for fp in _legal_footprints(orc, s2, 3 - p)}
        assert a == b


@settings(max_examples=10, deadline=None)
@given(seed=st.integers(0, 10_000), plies=st.integers(1, 70))
def test_every_legal_placement_obeys_the_rules_cell_by_cell(oracle20, seed, plies):
    """Independent numpy re-check of R5/R6 on the bit-parallel mask (which is what the GPU is compared with)."""
    orc = oracle20
    s = _play(orc, seed, plies)
    if orc.field(s, "done"):
        return
    p = orc.field(s, "mover")
    board = orc.board_contents(s)
    N = orc.N
    own = np.pad(board == p + 1, 1)
    occupied = board != 0
    first = orc.field(s, "inv")[p] == (1 << 21) - 1
    corner = [(0, 0), (0, N - 1), (N - 1, 0), (N - 1, N - 1)][p]
    m = orc.legal_mask(s, p, fast=True)
    for a in np.flatnonzero(m)[:: max(1, int(m.sum()) // 40)]:
        cells, meta = orc.action_cells(int(a))
        assert (orc.field(s, "inv")[p] >> int(meta[0])) & 1
        assert not any(occupied[y, x] for y, x in cells)
        assert not any(own[y + 1 + dy, x + 1 + dx] for y, x in cells for dy, dx in ((1, 0), (-1, 0), (0, 1), (0, -1)))
        if first:
            assert corner in cells
        else:
            assert any(own[y + 1 + dy, x + 1 + dx] for y, x in cells for dy in (-1, 1) for dx in (-1, 1))


@settings(max_examples=8, deadline=None)
@given(seed=st.integers(0, 10_000))
def test_stuck_players_stay_stuck_and_scores_count_squares(oracle20, seed):
    """Monotonicity used by the rollout kernel (a player without a move never gets one back) and R10."""
    orc = oracle20
    s = orc.new_state()
    stuck = [False] * 4
    while not orc.field(s, "done"):
        for p in range(4):
            has = orc.legal_mask(s, p, fast=True).any()
            assert not (stuck[p] and has)
            stuck[p] = stuck[p] or not has
        orc.step(s, orc.sample_action(s, seed, 1), fast=True)
    board = orc.board_contents(s)
    assert list(orc.field(s, "score")) == [(board == p + 1).sum() for p in range(4)]
    assert all(not orc.legal_mask(s, p).any() for p in range(4))
