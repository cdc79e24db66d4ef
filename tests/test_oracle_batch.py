"""The oracle's fast many-env forms (orc_play_many, orc_playout: what the full-size GPU parity tests and bench.py's CPU arm
run) against its slow, obvious single-env forms (naive cell-by-cell legality, mask + linear scan sampler)."""
import numpy as np
import pytest

from oracle.oracle import Oracle

M64 = (1 << 64) - 1


@pytest.mark.parametrize("N,P,rule", [(20, 4, 0), (7, 2, 0), (14, 2, 1), (5, 4, 0)])
def test_play_many_equals_single_env_play(N, P, rule):
    o = Oracle(N, P, rule)
    n, plies, seed, base, stride = 6, 70, 0x5EED, 100, 3
    sel = [0, 5, plies - 1, plies]
    st = o.new_states(n)
    r = o.play_many(st, seed, plies, env_id0=base, env_stride=stride, sel_plies=sel, threads=3)
    steps = games = 0
    cnt_sum = [0] * (plies + 1)
    for i in range(n):
        s, h, gid = o.new_state(), 0, base + stride * i
        for t in range(plies + 1):
            m = o.legal_mask(s, fast=False)                      # the naive rules
            ids = np.flatnonzero(m)
            a = o.sample_action(s, seed, gid, fast=False)
            h = (h * 0x9E3779B97F4A7C15 + ((a + 2) & M64)) & M64
            cnt_sum[t] = (cnt_sum[t] + len(ids) * (2 * gid + 1)) & M64
            if t in sel:
                j = sel.index(t)
                assert int(r["ids_sum"][i, j]) == int((ids + 1).sum())
                assert int(r["words_sum"][i, j]) == sum((1 << (int(x) & 31)) * (2 * (int(x) >> 5) + 1) for x in ids) & M64
            if t == plies:
                break
            assert o.step(s, a) == 0
            steps += 1
            if o.field(s, "done"):
                games += 1
                o.reset(s, o.field(s, "game") + 1)
        assert int(r["traj"][i]) == h
        assert (o.pack(s) == o.pack_many(st[i:i + 1])[0]).all()
        assert o.rows_consistent(st[i].ctypes.data)
    assert [int(x) for x in r["cnt_sum"]] == cnt_sum
    assert (r["steps"], r["games"]) == (steps, games)


def test_play_many_is_partition_and_thread_invariant():
    o = Oracle(20, 4)
    a = o.new_states(24)
    ra = o.play_many(a, 7, 40, threads=1, sel_plies=(13,))
    b = o.new_states(24)
    rb = o.play_many(b, 7, 40, threads=5, sel_plies=(13,), write_masks=True)
    assert (a == b).all() and (ra["traj"] == rb["traj"]).all() and (ra["cnt_sum"] == rb["cnt_sum"]).all()
    assert (ra["words_sum"] == rb["words_sum"]).all()
    # every third env of a wider run == a strided run over the same global ids
    c = o.new_states(8)
    rc = o.play_many(c, 7, 40, env_id0=0, env_stride=3, threads=2)
    assert (rc["traj"] == ra["traj"][::3]).all() and (c == a[::3]).all()


def test_play_many_without_auto_reset_keeps_finished_games():
    o = Oracle(7, 2)
    st = o.new_states(5)
    r = o.play_many(st, 3, 40, auto_reset=False, threads=1, sel_plies=(40,))
    assert r["games"] == 5                                         # a 7x7 game is over long before ply 40
    for i in range(5):
        s = o.unpack(o.pack_many(st[i:i + 1])[0])
        assert o.field(s, "done") == 1 and o.field(s, "game") == 0
    assert (r["ids_sum"] == 0).all()


@pytest.mark.parametrize("N,P", [(20, 4), (7, 2)])
def test_playout_equals_stepping_with_the_playout_stream(N, P):
    o = Oracle(N, P)
    root = o.new_state()
    for _ in range(6):                                              # a mid-game root
        o.step(root, o.sample_action(root, 11, 0))
    for g in range(4):
        n, scores, win, end, log = o.playout(root, 99, 1000 + g, log=True)
        s = o.copy(root)
        acts = []
        while not o.field(s, "done"):
            a = o.sample_action(s, 99, 1000 + g, stream=1, fast=False)
            acts.append(a)
            assert o.step(s, a) == 0
        assert n == len(acts) and [int(x) for x in log[:n]] == acts and log[n] == 0xFFFF
        assert (scores == o.final_scores(s)).all() and win == o.winners(s)
        assert (o.pack(end) == o.pack(s)).all()
    # stop_player: the playout hands the state back at that player's turn
    q = (o.field(root, "mover") + 1) % P
    n, _, win, end, _ = o.playout(root, 99, 5, stop_player=q)
    assert n == 1 and o.field(end, "mover") == q and win == 0
