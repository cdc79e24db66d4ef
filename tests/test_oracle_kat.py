"""CPU tests: known-answer vectors and cross-checks of the three oracles and the table generator.

The reference has no tests or golden vectors for this path; the executable pins
that exist are the sizes quoted in its sources, checked here.
"""
import numpy as np
import pytest

from blokus_rl_b200 import tables
from oracle import naive


def test_piece_set_known_answers():
    ors = tables.orientations()
    assert tables.NUM_PIECES == 21 and len(ors) == 91 and int(tables.piece_sizes().sum()) == 89
    by_size = np.bincount(tables.piece_sizes())
    assert list(by_size[1:]) == [1, 1, 2, 5, 12]
    assert max(max(o.h, o.w) for o in ors) == 5
    assert sum(o.h for o in ors) == 246      # kSumH in csrc/blk_engine.cu


@pytest.mark.parametrize("n,a", [(20, 30433), (14, 13729), (7, 2522), (6, 1649), (5, 958)])
def test_action_counts(n, a):
    # 30433: blokus_rl/models/blokus_nnet.py:17,97 and docs/README.md:128
    assert tables.action_table(n).num_actions == a
    assert len(naive.all_actions(n)) == a


def test_action_counts_by_piece_size_20():
    t = tables.action_table(20)
    sizes = tables.piece_sizes()[t.action_piece]
    assert [int((sizes == k).sum()) for k in range(1, 6)] == [400, 760, 2164, 6513, 20596]


def test_footprints_are_distinct_and_match_naive_order():
    for n in (7, 20):
        acts = naive.all_actions(n)
        assert len({c for _, c in acts}) == len(acts)        # one id per distinct footprint (a7)
        for a in range(0, len(acts), 13):
            assert frozenset(tables.action_cells(n, a)) == acts[a][1]
            assert int(tables.action_table(n).action_piece[a]) == acts[a][0]


def test_oracle_tables_match_generator(oracle20, oracle7):
    for orc in (oracle20, oracle7):
        t = tables.action_table(orc.N)
        assert orc.A == t.num_actions
        for a in list(range(0, orc.A, 11)) + [orc.A - 1]:
            cells, meta = orc.action_cells(a)
            assert sorted(cells) == sorted(tables.action_cells(orc.N, a))
            assert meta[0] == t.action_piece[a] and meta[2] == t.action_y[a] and meta[3] == t.action_x[a]


def test_generated_inc_is_current():
    import subprocess, sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    inc = root / "blokus_rl_b200" / "csrc" / "blk_orient.inc"
    before = inc.read_text()
    subprocess.check_call([sys.executable, str(root / "tools" / "gen_orient_inc.py")], stdout=subprocess.DEVNULL)
    assert inc.read_text() == before
    small = root / "blokus_rl_b200" / "csrc" / "blk_small_fields.inc"          # bit-stream code of the small-board kernels
    before = small.read_text()
    subprocess.check_call([sys.executable, str(root / "tools" / "gen_small_inc.py")], stdout=subprocess.DEVNULL)
    assert small.read_text() == before


def test_small_board_bit_stream_matches_action_table():
    """blk_small_fields.inc: the fields of every orientation, appended row by row, tile the action ids 0..A-1 exactly --
    (half, shift) is anchor row y of a stride-8 bitboard, W the number of anchor columns, fill the bit inside the mask word."""
    import re
    from pathlib import Path
    txt = (Path(__file__).resolve().parents[1] / "blokus_rl_b200" / "csrc" / "blk_small_fields.inc").read_text()
    for n in (5, 6, 7):
        sec = txt.split(f"#if BLK_SMALL_N == {n}\n")[1].split("#endif\n#endif")[0]
        t = tables.action_table(n)
        assert f"#define BLK_SMALL_A {t.num_actions}" in sec
        bit, orient = 0, -1
        for ln in sec.splitlines():
            m = re.match(r"SM_ORIENT\((\d+), (\d+), (\d+),", ln)
            if m:
                orient, y = int(m.group(1)), 0
                assert bit == t.orient_base[orient] and int(m.group(2)) == tables.orientations()[orient].piece
                continue
            m = re.match(r"SM_FIELD(_X)?\((\d), (\d+), (\d+), (\d+)", ln)
            if m:
                half, shift, w, fill = (int(m.group(k)) for k in (2, 3, 4, 5))
                o = tables.orientations()[orient]
                assert 4 * half + shift // 8 == y and shift % 8 == 0 and w == n - o.w + 1 and fill == bit % 32
                assert (m.group(1) is not None) == (fill + w > 32)
                bit, y = bit + w, y + 1
        assert bit == t.num_actions


def test_first_move_count_and_corners(oracle20, oracle7):
    for orc in (oracle20, oracle7):
        s = orc.new_state()
        for p in range(orc.P):
            m = orc.legal_mask(s, p)
            assert m.sum() == 58
            cy, cx = tables.start_corners(orc.N, orc.P)[p]
            for a in np.nonzero(m)[0][::7]:
                assert (cy, cx) in orc.action_cells(a)[0]  # R5


def test_naive_vs_bitparallel_vs_python_7x7(oracle7):
    rng = np.random.default_rng(0)
    for game in range(12):
        g = naive.NaiveGame(7, 2)
        s = oracle7.new_state()
        while not g.done:
            legal = g.legal(g.mover)
            mn, mf = oracle7.legal_mask(s), oracle7.legal_mask(s, fast=True)
            assert list(np.nonzero(mn)[0]) == legal and (mn == mf).all()
            a = int(rng.choice(legal))
            g.step(a)
            assert oracle7.step(s, a) == 0
            assert bool(oracle7.field(s, "done")) == g.done
            if not g.done:
                assert oracle7.field(s, "mover") == g.mover
        assert list(oracle7.field(s, "score")[:2]) == g.scores()
        assert oracle7.board_contents(s).tolist() == g.board()


def test_naive_vs_bitparallel_20x20_and_pack_roundtrip(oracle20):
    orc = oracle20
    for g in range(4):
        s = orc.new_state()
        while not orc.field(s, "done"):
            mn, mf = orc.legal_mask(s), orc.legal_mask(s, fast=True)
            assert (mn == mf).all()
            assert orc.unpack(orc.pack(s)).raw == s.raw
            a = orc.sample_action(s, 42, g)
            assert mn[a] == 1
            s2 = orc.copy(s)
            assert orc.step(s, a) == 0 and orc.step(s2, a, fast=True) == 0 and s.raw == s2.raw
        assert 40 <= orc.field(s, "ply") <= 84
        assert orc.legal_mask(s).sum() == 0


def test_python_naive_spot_check_20x20(oracle20):
    g = naive.NaiveGame(20, 4)
    s = oracle20.new_state()
    for ply in range(10):
        legal = g.legal(g.mover)
        assert list(np.nonzero(oracle20.legal_mask(s, fast=True))[0]) == legal
        a = legal[(7 * ply + 3) % len(legal)]
        g.step(a)
        oracle20.step(s, a)


def test_terminal_vector_and_winners(oracle20):
    # blokus_wrapper.py:177-185: -1 everywhere, 3 sole winner, 1 for each tied winner
    orc = oracle20
    seen_tie = seen_sole = False
    for g in range(60):
        s = orc.new_state()
        orc.random_play(s, 5, g, 100, auto_reset=False, log=False)
        assert orc.field(s, "done")
        w = orc.winners(s)
        v = orc.terminal_values(s)
        sc = orc.final_scores(s)
        best = sc.max()
        assert w == sum(1 << p for p in range(4) if sc[p] == best)
        nw = bin(w).count("1")
        for p in range(4):
            assert v[p] == ((3 if nw == 1 else 1) if (w >> p) & 1 else -1)
        seen_tie |= nw > 1
        seen_sole |= nw == 1
        if seen_tie and seen_sole and g > 20:
            break
    assert seen_sole


def test_running_game_has_no_winners_and_illegal_moves_rejected(oracle20):
    orc = oracle20
    s = orc.new_state()
    assert orc.winners(s) == 0 and (orc.terminal_values(s) == 0).all()
    keep = s.raw
    assert orc.step(s, 5000) == 1 and s.raw == keep          # not on the start corner
    assert orc.step(s, -1) == 1 and orc.step(s, orc.A) == 1
    assert orc.step(s, 0) == 0
    assert orc.field(s, "mover") == 1 and orc.field(s, "ply") == 1
    assert list(orc.field(s, "score")) == [1, 0, 0, 0]


def test_observation_layout(oracle20):
    orc = oracle20
    s = orc.new_state()
    orc.step(s, 0)
    obs = orc.observe(s)
    assert obs.shape == (8, 20, 20)                          # blokus_rl/models/blokus_nnet.py:99
    assert obs[0, 0, 0] == 1 and obs[0].sum() == 1 and obs[1:4].sum() == 0
    assert (obs[5] == 1).all() and obs[4].sum() == 0 and obs[6:].sum() == 0
    assert set(np.unique(orc.board_contents(s))) <= {0, 1, 2, 3, 4}   # blokus_wrapper.py:259


def test_philox_known_answer(oracle20):
    # Random123 known-answer test vectors for philox4x32_10
    assert list(oracle20.philox((0, 0, 0, 0), (0, 0))) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(oracle20.philox((0xffffffff,) * 4, (0xffffffff, 0xffffffff))) == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(oracle20.philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0))) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_bonus_score_rule():
    from oracle.oracle import Oracle
    o = Oracle(14, 2, score_rule=1)
    s = o.new_state()
    o.random_play(s, 3, 0, 200, auto_reset=False, log=False)
    base = o.field(s, "score")[:2]
    fin = o.final_scores(s)
    inv = o.field(s, "inv")[:2]
    for p in range(2):
        assert fin[p] - base[p] == ((15 + 5 * int(o.field(s, "lastmono")[p])) if inv[p] == 0 else 0)
