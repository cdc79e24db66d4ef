"""Batched self-play / arena drivers on the CPU stand-in and on the GPU."""
import pickle

import numpy as np
import pytest
import torch

from oracle_engine import OracleEngine


def _check_selfplay(eng, tmp_path):
    from blokus_rl_b200.selfplay import save_examples, self_play_batched
    data, stats = self_play_batched(eng, num_games=5, num_mcts_sims=4, rng=np.random.default_rng(0))
    assert stats["games"] == 5 and stats["examples"] == sum(len(d) for d in data)
    P, N, A = eng.num_players, eng.board_size, eng.num_actions
    for game in data:
        assert 2 <= len(game) <= 42
        for obs, mask, prob, scores in game:                          # alphazero/trainer.py:118-121
            assert obs.shape == (2 * P, N, N) and obs.dtype == np.float32
            assert mask.shape == (A,) and mask.dtype == np.float64
            assert prob.dtype == np.float32 and len(prob) == int(mask.sum()) and abs(prob.sum() - 1) < 1e-5
            assert scores.shape == (P,) and set(np.unique(scores)) <= {-1.0, 1.0, 3.0}
    files = save_examples(data, tmp_path, iteration=3)
    assert files[0].parent.name == "iteration_3" and files[0].suffix == ".examples"
    loaded = []
    for fp in sorted(tmp_path.rglob("*.examples")):                  # what AlphaZeroDataset.load_data does (dataset.py:29-36)
        with open(fp, "rb") as f:
            loaded.extend(pickle.Unpickler(f).load())
    assert len(loaded) == stats["examples"]
    item = loaded[0]                                                  # dataset.py:41-47
    torch.from_numpy(item[0]).float(), torch.from_numpy(item[1]).bool(), torch.from_numpy(item[2]).float(), torch.from_numpy(item[3]).float()


def _check_arena(eng):
    from blokus_rl_b200.selfplay import MCTSSeat, RandomSeat, RolloutSeat, play_match_batched
    seats = [MCTSSeat(simulations=3), RandomSeat()] if eng.num_players == 2 else \
        [MCTSSeat(simulations=2), RandomSeat(), RolloutSeat(2), RandomSeat()]
    scores, terminal, orders = play_match_batched(eng, seats, games_num=6, permute=True, seed=5)
    assert terminal.shape == (6, eng.num_players) and (np.abs(terminal).sum(1) > 0).all()
    assert set(np.unique(terminal)) <= {-1.0, 1.0, 3.0}
    expect = np.zeros(eng.num_players)
    for g in range(6):
        expect[list(orders[g])] += terminal[g]
    assert (scores == expect).all() and len({tuple(o) for o in orders}) > 1


def test_selfplay_examples_cpu(tmp_path):
    _check_selfplay(OracleEngine(7, 2), tmp_path)


def test_arena_cpu():
    _check_arena(OracleEngine(7, 2))


@pytest.mark.gpu
def test_selfplay_examples_gpu(tmp_path, engine7):
    _check_selfplay(engine7, tmp_path)


@pytest.mark.gpu
def test_selfplay_on_the_gpu_forest(tmp_path, engine7):
    from blokus_rl_b200.selfplay import self_play_gpu
    data, stats = self_play_gpu(engine7, num_games=32, num_mcts_sims=6, rng=np.random.default_rng(1))
    assert stats["examples"] == sum(len(d) for d in data) and all(2 <= len(g) <= 42 for g in data)
    for game in data[:4]:
        for obs, mask, prob, scores in game:
            assert obs.shape == (4, 7, 7) and len(prob) == int(mask.sum()) and abs(prob.sum() - 1) < 1e-5
            assert scores is not None and set(np.unique(scores)) <= {-1.0, 1.0, 3.0}


@pytest.mark.gpu
def test_arena_gpu(engine7, engine20):
    _check_arena(engine7)
    from blokus_rl_b200.selfplay import RandomSeat, RolloutSeat, play_match_batched
    scores, terminal, _ = play_match_batched(engine20, [RolloutSeat(8), RandomSeat(), RandomSeat(), RandomSeat()], games_num=16, seed=2)
    assert (np.abs(terminal).sum(1) > 0).all()
    assert scores[0] > scores[1:].mean()          # playouts beat uniform-random play


@pytest.mark.gpu
def test_sharded_self_play_is_partition_invariant(engine7):
    """Self-play games shard by GAME index (alphazero/trainer.py:152-154 plays them one after the other): game g draws its
    moves from its own generator and searches its own tree, so 3 ranks' games are the single rank's games."""
    from blokus_rl_b200.distributed import SELFPLAY_COUNTERS, Shard, self_play_shard
    whole_c, whole = self_play_shard(engine7, Shard(0, 1, 9), num_mcts_sims=8, seed=5)
    parts = [self_play_shard(engine7, Shard(r, 3, 9), num_mcts_sims=8, seed=5) for r in range(3)]
    games = [g for _, data in parts for g in data]
    assert len(games) == len(whole) == 9
    for a, b in zip(games, whole):
        assert len(a) == len(b)
        for (oa, ma, pa, sa), (ob, mb, pb, sb) in zip(a, b):
            assert (oa == ob).all() and (ma == mb).all() and (pa == pb).all() and (sa == sb).all()
    total = sum(c for c, _ in parts)
    whole_d, total_d = dict(zip(SELFPLAY_COUNTERS, whole_c.tolist())), dict(zip(SELFPLAY_COUNTERS, total.tolist()))
    assert total_d["games"] == 9 and total_d["examples"] == whole_d["examples"]
    assert [total_d[f"wins_p{q}"] for q in range(2)] == [whole_d[f"wins_p{q}"] for q in range(2)]
