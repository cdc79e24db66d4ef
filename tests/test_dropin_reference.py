"""Drop-in integration in the build container: the reference's OWN wrapper, MCTS, players and arena
(imported unmodified from /root/reference) run over blokus_rl_b200.colosseum_shim.

There is no GPU here, so the shim is driven by the test-only oracle backend (tests/oracle_backend.py); the
GPU box has no /root/reference, so there the same adapters are checked against the oracle instead
(tests/test_gpu_adapters.py).  Skipped when /root/reference is absent."""
import os
import sys
import types

import numpy as np
import pytest

import ref_stubs
from oracle_backend import OracleBackend

pytestmark = pytest.mark.skipif(not ref_stubs.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref(tmp_path_factory):
    from blokus_rl_b200 import colosseum_shim
    backend = OracleBackend(20, 4)
    colosseum_shim.set_backend(backend)
    colosseum_shim.install()
    ref_stubs.install_stubs()
    cwd = os.getcwd()
    work = tmp_path_factory.mktemp("refrun")
    os.chdir(work)                      # the reference writes debug.log / states/ into the CWD
    try:
        from blokus_rl.colossumrl.blokus_wrapper import ColosseumBlokusGameWrapper
        from blokus_rl.alphazero.mcts import MCTS
        from blokus_rl.alphazero.arena import play_match
        from blokus_rl.players import MCTSPlayer, RandomPlayer
        hp = types.SimpleNamespace(board_size=20, number_of_players=4, states_dir=work / "states")
        game = ColosseumBlokusGameWrapper(hp)     # builds its action table through the shim's Board API
        yield types.SimpleNamespace(game=game, MCTS=MCTS, play_match=play_match, MCTSPlayer=MCTSPlayer,
                                    RandomPlayer=RandomPlayer, backend=backend, work=work)
    finally:
        os.chdir(cwd)
        colosseum_shim.set_backend(None)


class UniformNet:
    """DumbNet's behaviour (blokus_rl/models/dumbnet.py:14-21) through predict's contract
    (blokus_rl/neural_network.py:92-110): uniform prior over the valid actions, zero values."""

    def __init__(self, players):
        self.players = players

    def predict(self, obs, mask):
        n = int(mask.sum())
        return np.full(n, 1.0 / n), np.zeros(self.players)


def test_reference_wrapper_builds_the_canonical_action_table(ref):
    from blokus_rl_b200 import tables
    game = ref.game
    assert game.get_action_size() == 30433                   # models/blokus_nnet.py:17
    assert game.get_observation_size() == [8, 20, 20]
    # ids numbered by the reference's own enumeration loop == the engine's canonical ids
    assert game._move_action_dict == tables.string_to_action(20)
    assert (ref.work / "states" / "colosseum_20_players_4.json").exists()


def test_reference_wrapper_semantics_over_shim(ref):
    game, b = ref.game, ref.backend
    s, player = game.get_init_board()
    assert player == 0
    mask = game.get_valid_moves(s, player)
    assert mask.dtype == np.float64 and mask.shape == (30433,) and mask.sum() == 58
    assert (game.get_valid_moves(s, -1) == mask).all()
    obs, mask2 = game.get_observation(s, player)
    assert obs.shape == (8, 20, 20) and (mask2 == mask).all()
    assert game.get_game_ended(s) is None
    a = game.get_sample_move(s)
    assert mask[a] == 1
    s2, p2 = game.get_next_state(s, player, a)
    assert p2 == 1 and game.get_valid_moves(s, player).sum() == 58      # input state untouched (functional)
    assert game.string_representation(s) != game.string_representation(s2)
    assert isinstance(game.get_valid_actions_for_human_player(s2, p2)[0], str)


def test_reference_arena_random_players_finish_a_game(ref):
    np.random.seed(0)
    players = [ref.RandomPlayer(ref.game) for _ in range(4)]
    scores, items = ref.play_match(ref.game, players, games_num=2)
    assert len(items) == 2
    for it in items:
        v = it["scores"]
        assert set(np.unique(v)) <= {-1.0, 1.0, 3.0} and (v > 0).any()   # blokus_wrapper.py:177-185


def test_reference_mcts_player_over_shim(ref):
    np.random.seed(1)
    game = ref.game
    player = ref.MCTSPlayer(game, UniformNet(4), simulations=6)
    s, cur = game.get_init_board()
    for _ in range(3):
        s, cur = player.update_state(s, cur)
    assert ref.backend.ply(s[0]._h) == 3
    assert len(player.tree.tree) >= 6


def test_reference_wrapper_over_shim_at_7x7_two_players(tmp_path):
    """The PPO-config geometry (7x7, 2 players) through the reference's wrapper + a full arena game."""
    import os
    from blokus_rl_b200 import colosseum_shim, tables
    backend = OracleBackend(7, 2)
    colosseum_shim.set_backend(backend)
    colosseum_shim.install()
    ref_stubs.install_stubs()
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        from blokus_rl.colossumrl.blokus_wrapper import ColosseumBlokusGameWrapper
        from blokus_rl.alphazero.arena import play_match
        from blokus_rl.players import MCTSPlayer, RandomPlayer
        game = ColosseumBlokusGameWrapper(types.SimpleNamespace(board_size=7, number_of_players=2, states_dir=tmp_path / "states"))
        assert game.get_action_size() == 2522 and game._move_action_dict == tables.string_to_action(7)
        np.random.seed(4)
        scores, items = play_match(game, [MCTSPlayer(game, UniformNet(2), simulations=5), RandomPlayer(game)], games_num=2, permute=True)
        assert len(items) == 2 and all(set(np.unique(it["scores"])) <= {-1.0, 1.0, 3.0} for it in items)
    finally:
        os.chdir(cwd)
        colosseum_shim.set_backend(None)


def _reference_self_play_episode(backend, work, N, P, sims, seed):
    """AlphaZeroTrainer._self_play (blokus_rl/alphazero/trainer.py:92-137), the reference's own method, run unmodified on a
    stand-in `self` that carries what it reads: game, nnet, hparams.num_mcts_sims / cpuct."""
    from blokus_rl_b200 import colosseum_shim
    colosseum_shim.set_backend(backend)
    colosseum_shim.install()
    ref_stubs.install_stubs()
    from blokus_rl.alphazero.trainer import AlphaZeroTrainer
    from blokus_rl.colossumrl.blokus_wrapper import ColosseumBlokusGameWrapper
    game = ColosseumBlokusGameWrapper(types.SimpleNamespace(board_size=N, number_of_players=P, states_dir=work / "states"))
    me = types.SimpleNamespace(game=game, nnet=UniformNet(P), hparams=types.SimpleNamespace(num_mcts_sims=sims, cpuct=1.0))
    np.random.seed(seed)
    return AlphaZeroTrainer._self_play(me, 1)


def test_reference_self_play_episode_over_shim(tmp_path):
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        from blokus_rl_b200 import colosseum_shim
        data = _reference_self_play_episode(OracleBackend(7, 2), tmp_path, 7, 2, sims=5, seed=0)
    finally:
        os.chdir(cwd)
        colosseum_shim.set_backend(None)
    assert 4 <= len(data) <= 42
    for obs, mask, prob, scores in data:                              # trainer.py:118-121, what AlphaZeroDataset reads
        assert obs.shape == (4, 7, 7) and mask.shape == (2522,) and mask.dtype == np.float64
        assert prob.dtype == np.float32 and len(prob) == int(mask.sum()) and abs(float(prob.sum()) - 1) < 1e-5
        assert set(np.unique(scores)) <= {-1.0, 1.0, 3.0}


def _reference_ppo_rollouts(engine, tmp, num_envs=4, num_steps=24, seed=7):
    """PPOTrainer._play_env + _compute_gae (blokus_rl/ppo/trainer.py:128-175, 177-205), the reference's own methods with its own
    CnnAgent, FilterLegalMoves and Memory, run unmodified over BlokusVectorEnv (boundary B2) on a stand-in `self`."""
    import torch
    from blokus_rl_b200.vector_env import BlokusVectorEnv
    ref_stubs.install_stubs()
    from blokus_rl.hparams import PPOHparams
    from blokus_rl.ppo.agent import get_agent
    from blokus_rl.ppo.memory import Memory
    from blokus_rl.ppo.trainer import PPOTrainer
    hp = PPOHparams(num_envs=num_envs, num_steps=num_steps, agent_type="cnn", cuda=False, seed=seed)
    envs = BlokusVectorEnv(hp.num_envs, engine=engine, seed=hp.seed)
    torch.manual_seed(seed)
    agent = get_agent(hp.agent_type)(envs, hp).eval()
    me = types.SimpleNamespace(hparams=hp, device="cpu", envs=envs, agent=agent, memory=Memory(hp, envs, "cpu"), global_step=0,
                               _total_episodes=0, _total_episodes_reward=0)
    me.running_vals = PPOTrainer._reset_running_vals(me)
    for name in ("_get_valid_moves_mask", "_update_running_vals", "_compute_gae"):
        setattr(me, name, types.MethodType(getattr(PPOTrainer, name), me))
    obs, _ = envs.reset()
    next_obs, next_done = PPOTrainer._play_env(me, torch.Tensor(obs), torch.zeros(hp.num_envs))
    with torch.inference_mode():
        adv = me._compute_gae(agent.get_value(next_obs).reshape(1, -1), next_done)
    return me, adv


def test_reference_ppo_rollout_collection_over_vector_env(tmp_path):
    from oracle_engine import OracleEngine
    me, adv = _reference_ppo_rollouts(OracleEngine(7, 2), tmp_path)
    m = me.memory
    assert me.global_step == 4 * 24 and me._total_episodes >= 2 and adv.shape == (24, 4)
    assert m.obs.shape == (24, 4, 7, 7) and set(np.unique(m.rewards.numpy())) <= {-1.0, 0.0, 1.0}
    assert (m.actions >= 0).all() and (m.actions < 2522).all() and m.dones.sum() >= 2
