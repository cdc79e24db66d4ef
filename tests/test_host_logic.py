"""CPU tests of the host-side layers (game wrapper, players, vector env, sharding) over the test-only oracle
engine; the same layers are run on the real engine by tests/test_gpu_adapters.py."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle_engine import OracleEngine

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def game7():
    from blokus_rl_b200.backend import EngineBackend
    from blokus_rl_b200.game_wrapper import BlokusGameWrapper
    return BlokusGameWrapper(board_size=7, number_of_players=2, backend=EngineBackend(engine=OracleEngine(7, 2)))


def test_game_wrapper_api_shape(game7):
    g = game7
    assert g.get_board_size() == (7, 7) and g.get_action_size() == 2522
    assert g.get_observation_size() == [4, 7, 7] and g.get_number_of_players() == 2
    s, p = g.get_init_board()
    m = g.get_valid_moves(s, p)
    assert m.dtype == np.float64 and m.sum() == 58 and (g.get_valid_moves(s, -1) == m).all()
    obs, m2 = g.get_observation(s, p)
    assert obs.shape == (4, 7, 7) and (m == m2).all()
    assert g.get_game_ended(s) is None
    s2, p2 = g.get_next_state(s, p, g.action_move_dict[0])       # action strings are accepted (blokus_wrapper.py:102)
    assert p2 == 1 and g.get_valid_moves(s, p).sum() == 58
    with pytest.raises(ValueError):
        g.get_next_state(s, p, 1500)
    assert g.string_representation(s) != g.string_representation(s2)
    assert g.render(s2).shape == (7 * 24, 7 * 24, 3)
    assert list(g.get_scores([1])) == [-1, 1] and list(g.get_scores([0, 1])) == [0, 0]
    assert g.get_valid_actions_for_human_player(s, p)[0] == "0;0;0"


def test_players_finish_games_through_the_player_interface(game7):
    from blokus_rl_b200.players import MCTSPlayer, RandomPlayer, RolloutPlayer
    np.random.seed(0)
    g = game7
    for players in ([RandomPlayer(g), MCTSPlayer(g, simulations=4)], [RolloutPlayer(g, per_move=2), RandomPlayer(g)]):
        for pl in players:
            pl.reset()
        s, cur = g.get_init_board()
        ended, plies = None, 0
        while ended is None:
            s, cur = players[cur].update_state(s, cur)
            ended = g.get_game_ended(s)
            plies += 1
        assert 2 <= plies <= 42 and set(np.unique(ended)) <= {-1.0, 1.0, 3.0}


def test_mcts_player_equals_reference_mcts_player(game7):
    """players/mcts_player.py:15-28 over the same game: same moves for a whole game (uniform net, quirk 4)."""
    import ref_stubs
    if not ref_stubs.available():
        pytest.skip("/root/reference not present")
    from fake_nets import UniformNet
    from blokus_rl_b200.players import MCTSPlayer
    RefMCTS = ref_stubs.load_reference_mcts().MCTS
    g = game7
    mine = MCTSPlayer(g, simulations=12)
    tree = RefMCTS(g, UniformNet(2))
    s, cur = g.get_init_board()
    for _ in range(6):
        for _ in range(12):
            tree.simulate(s, cur)
        dist = tree.get_distribution(s, 0)
        ref_action = int(dist[np.argmax(dist[:, 1]), 0][0])
        s_mine, cur_mine = mine.update_state(s, cur)
        s, cur = g.get_next_state(s, cur, ref_action)
        assert (s_mine.host_words == s.host_words).all() and cur_mine == cur
        if g.get_game_ended(s) is not None:
            break


def test_vector_env_gym_contract():
    from blokus_rl_b200.vector_env import BlokusVectorEnv
    env = BlokusVectorEnv(6, engine=OracleEngine(7, 2), seed=3)
    assert env.single_observation_space.shape == (7, 7) and env.single_action_space.n == 2522
    obs, info = env.reset()
    assert obs.shape == (6, 7, 7) and obs.sum() == 0 and info == {}
    rng = np.random.default_rng(0)
    finished, total_len = 0, 0
    for _ in range(40):
        poss = env.get_attr("ai_possible_indexes")          # ppo/trainer.py:385
        assert len(poss) == 6 and all(len(p) > 0 for p in poss)
        acts = np.array([rng.choice(p) for p in poss])
        obs, reward, terminated, truncated, info = env.step(acts)
        assert obs.shape == (6, 7, 7) and reward.shape == (6,) and not truncated.any()
        assert set(np.unique(reward)) <= {-1.0, 0.0, 1.0} and (reward[~terminated] == 0).all()
        for i in np.flatnonzero(terminated):
            ep = info["final_info"][i]["episode"]           # ppo/trainer.py:162-173
            assert ep["r"] == reward[i] and 1 <= ep["l"] <= 21
            assert obs[i].sum() <= 5 * 2                    # autoreset: already the next episode's first obs
            finished += 1
            total_len += ep["l"]
        if not terminated.any():
            assert "final_info" not in info
    assert finished >= 6
    # an illegal action for ONE env raises and leaves every env where it was (legal envs do not advance half-way)
    poss = env.get_attr("ai_possible_indexes")
    acts = np.array([p[0] for p in poss])
    acts[3] = next(a for a in range(2522) if a not in set(poss[3]))
    before = env.states.clone()
    with pytest.raises(ValueError):
        env.step(acts)
    assert (env.states == before).all() and env.get_attr("ai_possible_indexes") == poss
    acts[3] = poss[3][-1]
    env.step(acts)                                          # ... and the env is still usable
    assert not (env.states == before).all()
    env.close()


def test_shard_ranges_cover_everything():
    from blokus_rl_b200.distributed import Shard
    for total in (1, 7, 65536, 1048576):
        for world in (1, 2, 3, 8):
            sh = [Shard(r, world, total) for r in range(world)]
            assert sh[0].lo == 0 and sh[-1].hi == total and all(a.hi == b.lo for a, b in zip(sh, sh[1:]))
            assert max(s.n for s in sh) - min(s.n for s in sh) <= 1


WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch
from oracle_engine import OracleEngine
from blokus_rl_b200 import distributed as D
D.init("gloo")
shard = D.shard_from_env(10)
eng = OracleEngine(7, 2)
local = D.random_play_shard(eng, shard, plies=12, seed=99)
total = D.reduce_counters(local)
# playouts shard by ROOT index (global playout ids in the RNG key): 6 mid-game roots x 5 playouts
rc, _ = D.rollout_shard(eng, D.shard_from_env(6), 5, seed=7, root_plies=2)
total["rollouts"] = D.reduce_counters(rc, D.ROLLOUT_COUNTERS)
if shard.rank == 0:
    print("RESULT " + json.dumps(total))
'''


def _run_world(world):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29500 + world), WORLD_SIZE=str(world))
    procs = [subprocess.Popen([sys.executable, "-c", WORKER, str(ROOT)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT ")][0]
    import json
    return json.loads(line[7:])


def test_two_rank_gloo_run_equals_single_rank_run():
    """World-size-2 gloo run: env sharding + the final all-reduce give the same totals as one rank
    (global env ids make the trajectories partition-invariant)."""
    one, two = _run_world(1), _run_world(2)
    assert one == two and one["steps"] == 120 and one["games"] > 0 and one["illegal"] == 0
    r = one["rollouts"]
    assert r["playouts"] == 30 and r["plies"] > 30 and r["wins_p0"] + r["wins_p1"] >= 30


def test_vector_env_four_players_agent_not_first():
    """4-player board, the agent sits in seat 2: three random bots move around it inside reset/step."""
    from blokus_rl_b200.vector_env import BlokusVectorEnv
    env = BlokusVectorEnv(5, engine=OracleEngine(7, 4), seed=1, agent_player=2)
    obs, _ = env.reset()
    assert obs.shape == (5, 7, 7) and (obs > 0).any()              # seats 0 and 1 have already moved
    assert set(np.unique(obs)) <= {0.0, 1.0, 2.0}
    rng = np.random.default_rng(3)
    episodes = 0
    for _ in range(25):
        poss = env.get_attr("ai_possible_indexes")
        acts = np.array([rng.choice(p) if p else -1 for p in poss])
        obs, reward, terminated, truncated, info = env.step(acts)
        episodes += int(terminated.sum())
        assert (reward[~terminated] == 0).all()
    assert episodes >= 5
