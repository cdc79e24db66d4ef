"""The zero-change chain ON HARDWARE: the reference's own ``ColosseumBlokusGameWrapper``, ``MCTS``, players and arena
(blokus_rl/colossumrl/blokus_wrapper.py:21-324, alphazero/mcts.py, alphazero/arena.py, players/*.py -- imported
UNMODIFIED from the ``pip install --no-deps --target baseline/_ref`` copy that travels to the GPU box, or from
/root/reference) run over ``blokus_rl_b200.colosseum_shim`` on the CUDA engine.  The same calls through the CPU oracle
backend give the expected values, and the reference's MCTS must reproduce the committed golden vectors (which the same
unmodified file produced over the oracle in the build container) when the env underneath is the GPU."""
import json
import os
import types
from pathlib import Path

import numpy as np
import pytest

import ref_stubs

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_stubs.available(), reason="reference package not present")]

GOLDEN = json.loads((Path(__file__).parent / "golden" / "mcts_golden.json").read_text())


class UniformNet:
    """DumbNet (blokus_rl/models/dumbnet.py:14-21) through predict's contract (neural_network.py:92-110)."""

    def __init__(self, players):
        self.players = players

    def predict(self, obs, mask):
        n = int(mask.sum())
        return np.full(n, 1.0 / n), np.zeros(self.players)


def _reference_over(backend, work, N, P):
    from blokus_rl_b200 import colosseum_shim
    colosseum_shim.set_backend(backend)
    colosseum_shim.install()
    ref_stubs.install_stubs()
    from blokus_rl.alphazero.arena import play_match
    from blokus_rl.alphazero.mcts import MCTS
    from blokus_rl.colossumrl.blokus_wrapper import ColosseumBlokusGameWrapper
    from blokus_rl.players import MCTSPlayer, RandomPlayer
    game = ColosseumBlokusGameWrapper(types.SimpleNamespace(board_size=N, number_of_players=P, states_dir=work / "states"))
    return types.SimpleNamespace(game=game, MCTS=MCTS, play_match=play_match, MCTSPlayer=MCTSPlayer, RandomPlayer=RandomPlayer)


@pytest.fixture()
def chdir_tmp(tmp_path):
    cwd = os.getcwd()
    os.chdir(tmp_path)                      # the reference writes debug.log / states/ into the CWD
    yield tmp_path
    os.chdir(cwd)
    from blokus_rl_b200 import colosseum_shim
    colosseum_shim.set_backend(None)


@pytest.mark.parametrize("N,P,A", [(20, 4, 30433), (7, 2, 2522)])
def test_reference_wrapper_and_arena_over_the_gpu_engine(chdir_tmp, N, P, A):
    from blokus_rl_b200 import tables
    from blokus_rl_b200.backend import EngineBackend
    from oracle_backend import OracleBackend
    gpu = _reference_over(EngineBackend(N, P), chdir_tmp, N, P)
    game = gpu.game
    assert game.get_action_size() == A and game._move_action_dict == tables.string_to_action(N)
    assert game.get_observation_size() == [2 * P, N, N]
    # one random game, move by move, against the same reference wrapper over the CPU oracle backend
    cpu_b = OracleBackend(N, P)
    s, cur = game.get_init_board()
    o = cpu_b.new_state()
    rng = np.random.default_rng(5)
    plies = 0
    while True:
        mask = game.get_valid_moves(s, cur)
        assert mask.dtype == np.float64 and mask.shape == (A,)
        assert (mask == cpu_b.legal_mask(o)).all() and cur == cpu_b.mover(o)
        obs, _ = game.get_observation(s, cur)
        assert (obs == cpu_b.observation(o)).all()
        assert (s[0].board_contents == cpu_b.board_contents(o)).all()
        a = int(rng.choice(np.flatnonzero(mask)))
        s2, cur = game.get_next_state(s, cur, a)
        assert game.get_valid_moves(s, -1).sum() == mask.sum()           # the input state is untouched (functional)
        s, o = s2, cpu_b.next_state(o, a)
        plies += 1
        end = game.get_game_ended(s)
        if end is not None:
            assert (end == cpu_b.terminal_values(o)).all() and set(np.unique(end)) <= {-1.0, 1.0, 3.0}
            break
        assert not cpu_b.done(o)
    assert plies >= (6 if N == 7 else 30)
    # the reference's arena with its own players
    np.random.seed(0)
    players = [gpu.MCTSPlayer(game, UniformNet(P), simulations=4)] + [gpu.RandomPlayer(game) for _ in range(P - 1)]
    scores, items = gpu.play_match(game, players, games_num=1)
    assert len(items) == 1 and set(np.unique(items[0]["scores"])) <= {-1.0, 1.0, 3.0}


def test_reference_mcts_over_the_gpu_engine_reproduces_its_golden_vectors(chdir_tmp):
    """The unmodified reference MCTS, searching through the reference wrapper over the shim over the CUDA engine, gives
    the visit counts / Q values it gave over the oracle when the goldens were made (tests/golden/make_mcts_golden.py)."""
    from blokus_rl_b200.backend import EngineBackend
    done = 0
    refs = {}
    for c in GOLDEN["cases"]:
        if c["net"] != "uniform":
            continue
        N, P = c["board_size"], c["players"]
        if (N, P) not in refs:
            b = EngineBackend(N, P)
            refs[(N, P)] = (b, _reference_over(b, chdir_tmp, N, P))
        b, ref = refs[(N, P)]
        from blokus_rl_b200 import colosseum_shim
        colosseum_shim.set_backend(b)
        game = ref.game
        s = game.env._wrap(b.from_words(np.array(c["root_words"], dtype=np.uint32)))
        tree = ref.MCTS(game, UniformNet(P))
        per_sim = [np.asarray(tree.simulate(s, c["root_player"], c["cpuct"]), dtype=np.float64) for _ in range(c["sims"])]
        assert np.allclose(per_sim, np.array(c["scores"]), rtol=0, atol=1e-12)
        node = tree.tree[game.string_representation(s)]
        ids = [int(np.asarray(e[0]).reshape(-1)[0]) for e in node]
        assert ids == c["ids"] and [float(e[1]) for e in node] == c["N"]
        assert np.allclose([float(e[2]) for e in node], c["Q"], rtol=0, atol=1e-12)
        done += 1
    assert done >= 6


def test_gpu_mcts_player_plays_the_reference_mcts_players_moves(chdir_tmp):
    """players/mcts_player.py (unmodified, persistent dict, DumbNet-like uniform net) against this repo's MCTSPlayer, whose
    single tree lives on the GPU (fused search kernel, board-keyed nodes, reroot between moves): same moves, whole game."""
    from blokus_rl_b200.backend import EngineBackend
    from blokus_rl_b200.game_wrapper import BlokusGameWrapper
    from blokus_rl_b200.players import MCTSPlayer
    b = EngineBackend(7, 2)
    ref = _reference_over(b, chdir_tmp, 7, 2)
    ours = BlokusGameWrapper(board_size=7, number_of_players=2, backend=b)
    sims = 40
    ref_players = [ref.MCTSPlayer(ref.game, UniformNet(2), sims) for _ in range(2)]
    our_players = [MCTSPlayer(ours, simulations=sims) for _ in range(2)]
    assert all(p.gpu is not None and p.gpu.fused for p in our_players)
    s_ref, cur = ref.game.get_init_board()
    s_our, cur_our = ours.get_init_board()
    plies = 0
    while ref.game.get_game_ended(s_ref) is None:
        s_ref, nxt = ref_players[cur].update_state(s_ref, cur)
        s_our, nxt_our = our_players[cur].update_state(s_our, cur)
        assert nxt == nxt_our and (s_ref[0].board_contents == b.board_contents(s_our)).all(), f"ply {plies}: different move"
        cur = nxt
        plies += 1
    assert plies >= 6 and ours.get_game_ended(s_our) is not None


def test_reference_self_play_episode_on_the_gpu_equals_the_one_on_the_oracle(chdir_tmp):
    """AlphaZeroTrainer._self_play (alphazero/trainer.py:92-137), the reference's own method, unmodified: the same seeded
    episode over the CUDA engine and over the CPU oracle backend produces the same training examples, one by one."""
    from test_dropin_reference import _reference_self_play_episode
    from blokus_rl_b200.backend import EngineBackend
    from oracle_backend import OracleBackend
    on_gpu = _reference_self_play_episode(EngineBackend(7, 2), chdir_tmp, 7, 2, sims=6, seed=3)
    on_cpu = _reference_self_play_episode(OracleBackend(7, 2), chdir_tmp, 7, 2, sims=6, seed=3)
    assert len(on_gpu) == len(on_cpu) >= 4
    for (og, mg, pg, sg), (oc, mc, pc, sc) in zip(on_gpu, on_cpu):
        assert (og == oc).all() and (mg == mc).all() and (pg == pc).all() and (sg == sc).all()


def test_reference_ppo_rollout_collection_on_the_gpu_env(engine7, tmp_path):
    """ppo/trainer.py:128-175 (the reference's `_play_env`, unmodified, with its own CnnAgent / FilterLegalMoves / Memory):
    the rollout buffer it fills over BlokusVectorEnv on the CUDA engine (opponents, autoreset and legal ids on the device) is the
    buffer it fills over the CPU stand-in engine, step by step."""
    import torch
    from oracle_engine import OracleEngine
    from test_dropin_reference import _reference_ppo_rollouts
    gpu, adv_g = _reference_ppo_rollouts(engine7, tmp_path)
    cpu, adv_c = _reference_ppo_rollouts(OracleEngine(7, 2), tmp_path)
    for name in ("obs", "actions", "rewards", "dones", "logprobs", "values"):
        assert torch.equal(getattr(gpu.memory, name), getattr(cpu.memory, name)), name
    assert torch.equal(adv_g, adv_c) and gpu._total_episodes == cpu._total_episodes >= 2
