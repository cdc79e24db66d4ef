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
