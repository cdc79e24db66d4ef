"""GPU tests of the reference-shaped adapters over the real engine, checked against the CPU oracle
(the GPU box has no /root/reference; the reference's own files are exercised in test_dropin_reference.py)."""
import numpy as np
import pytest
import torch

from oracle_backend import OracleBackend
from oracle_engine import OracleEngine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair20(engine20):
    from blokus_rl_b200.backend import EngineBackend
    from blokus_rl_b200.game_wrapper import BlokusGameWrapper
    g_gpu = BlokusGameWrapper(board_size=20, number_of_players=4, backend=EngineBackend(engine=engine20))
    g_cpu = BlokusGameWrapper(board_size=20, number_of_players=4, backend=OracleBackend(20, 4))
    return g_gpu, g_cpu


def test_game_wrapper_gpu_equals_oracle_for_whole_games(pair20):
    g_gpu, g_cpu = pair20
    rng = np.random.default_rng(5)
    for game in range(2):
        sg, pg = g_gpu.get_init_board()
        sc, pc = g_cpu.get_init_board()
        while True:
            mg, mc = g_gpu.get_valid_moves(sg, pg), g_cpu.get_valid_moves(sc, pc)
            assert mg.dtype == np.float64 and (mg == mc).all() and pg == pc
            og, _ = g_gpu.get_observation(sg, pg)
            oc, _ = g_cpu.get_observation(sc, pc)
            assert (og == oc).all()
            assert (g_gpu.backend.board_contents(sg) == g_cpu.backend.board_contents(sc)).all()
            eg, ec = g_gpu.get_game_ended(sg), g_cpu.get_game_ended(sc)
            assert (eg is None) == (ec is None)
            if eg is not None:
                assert (eg == ec).all()
                break
            a = int(rng.choice(np.flatnonzero(mg)))
            sg, pg = g_gpu.get_next_state(sg, pg, a)
            sc, pc = g_cpu.get_next_state(sc, pc, a)


def test_colosseum_shim_surface_on_gpu(engine20, oracle20):
    from blokus_rl_b200 import colosseum_shim as shim
    from blokus_rl_b200.backend import EngineBackend
    shim.set_backend(EngineBackend(engine=engine20))
    try:
        env = shim.BlokusEnvironment()
        state, players = env.new_state()
        assert players == [0] and state[0].player_color == 0 and state[-1][0].player_color == 0
        o = oracle20.new_state()
        for ply in range(10):
            acts = env.valid_actions(state=state, player=players[0])
            legal = np.flatnonzero(oracle20.legal_mask(o, fast=True))
            assert [env._ids[a] for a in acts] == list(legal)
            assert not env.get_winners(state)
            assert state[0].canonical_board.shape == (8, 20, 20)
            assert (state[0].board_contents == oracle20.board_contents(o)).all()
            a = acts[(7 * ply) % len(acts)]
            nxt, players, *_ = env.next_state(state=state, players=players, actions=[a])
            assert env.valid_actions(state, players[0]) == acts           # input state is immutable
            state = nxt
            oracle20.step(o, env._ids[a])
            assert players == [oracle20.field(o, "mover")]
    finally:
        shim.set_backend(None)


@pytest.mark.parametrize("N,P,agent,narrow", [(7, 2, 0, False), (7, 4, 2, False), (7, 2, 0, True), (20, 4, 0, True)])
def test_vector_env_on_gpu_matches_cpu_stand_in(N, P, agent, narrow):
    """`step` on the CUDA engine (direct C-ABI launches, host-side episode statistics) against the same class over the CPU
    stand-in engine (generic path): observations, rewards, done flags, legal ids, episode statistics and final observations,
    step by step; `narrow` starts with id rows that are too short, so the staging buffers are replaced mid-step."""
    from blokus_rl_b200 import BlokusEngine
    from blokus_rl_b200.vector_env import BlokusVectorEnv
    E = 16 if N == 7 else 6
    envs = [BlokusVectorEnv(E, engine=BlokusEngine(N, P), seed=11, agent_player=agent),
            BlokusVectorEnv(E, engine=OracleEngine(N, P), seed=11, agent_player=agent)]
    assert envs[0]._direct and not envs[1]._direct
    if narrow:
        envs[0]._idx_stride = 8
        envs[0]._alloc_idx()
    obs = [e.reset()[0] for e in envs]
    assert (obs[0] == obs[1]).all()
    rng = np.random.default_rng(1)
    episodes = 0
    for _ in range(30 if N == 7 else 24):
        poss = [e.get_attr("ai_possible_indexes") for e in envs]
        assert poss[0] == poss[1]
        acts = np.array([rng.choice(p) for p in poss[0]])
        res = [e.step(acts) for e in envs]
        for k in range(4):
            assert res[0][k].dtype == res[1][k].dtype and (res[0][k] == res[1][k]).all()
        episodes += int(res[0][2].sum())
        assert ("final_info" in res[0][4]) == bool(res[0][2].any()) == ("final_info" in res[1][4])
        for i in np.flatnonzero(res[0][2]):
            assert res[0][4]["final_info"][i] == res[1][4]["final_info"][i]
            assert (res[0][4]["final_observation"][i] == res[1][4]["final_observation"][i]).all()
    assert episodes > (16 if N == 7 else 0)
    assert (envs[0].states.cpu() == envs[1].states.cpu()).all()
    if narrow:
        assert envs[0]._idx_stride > 8
    # an illegal action for ONE env raises and leaves every env where it was
    poss = envs[0].get_attr("ai_possible_indexes")
    acts = np.array([p[0] for p in poss])
    acts[3] = next(a for a in range(envs[0].A) if a not in set(poss[3]))
    before = envs[0].states.clone()
    with pytest.raises(ValueError):
        envs[0].step(acts)
    assert (envs[0].states == before).all() and envs[0].get_attr("ai_possible_indexes") == poss
    acts[3] = poss[3][-1]
    res = [e.step(acts) for e in envs]
    for k in range(4):
        assert (res[0][k] == res[1][k]).all()


def test_random_play_shards_match_cpu_stand_in_and_are_partition_invariant(engine7):
    from blokus_rl_b200.distributed import Shard, random_play_shard
    cpu = OracleEngine(7, 2)
    whole = random_play_shard(engine7, Shard(0, 1, 24), plies=20, seed=5)
    assert (whole.cpu() == random_play_shard(cpu, Shard(0, 1, 24), plies=20, seed=5)).all()
    parts = sum(random_play_shard(engine7, Shard(r, 3, 24), plies=20, seed=5) for r in range(3))
    assert (parts == whole).all()


def test_players_on_gpu(engine7):
    from blokus_rl_b200.backend import EngineBackend
    from blokus_rl_b200.game_wrapper import BlokusGameWrapper
    from blokus_rl_b200.mcts import RolloutEvaluator
    from blokus_rl_b200.players import MCTSPlayer, RandomPlayer, RolloutPlayer
    np.random.seed(3)
    g = BlokusGameWrapper(board_size=7, number_of_players=2, backend=EngineBackend(engine=engine7))
    for players in ([MCTSPlayer(g, RolloutEvaluator(8), simulations=6), RandomPlayer(g)],
                    [RolloutPlayer(g, per_move=16), MCTSPlayer(g, simulations=5)]):
        s, cur = g.get_init_board()
        ended = None
        while ended is None:
            s, cur = players[cur].update_state(s, cur)
            ended = g.get_game_ended(s)
        assert set(np.unique(ended)) <= {-1.0, 1.0, 3.0}


def test_leaf_expansion_tensors_for_the_policy_net(engine20):
    """BASELINE config 4: contiguous obs float32 [B,8,20,20] + bool mask [B,30433] + terminal vectors on the GPU."""
    eng = engine20
    B = 256
    s = eng.new_states(B)
    out = eng.step(s, None, mask="bytes", sample=True, seed=4)
    for _ in range(30):
        out = eng.step(s, out.next_action, mask="bytes", sample=True, seed=4, auto_reset=True)
    obs = eng.observe(s)
    assert obs.is_contiguous() and obs.dtype == torch.float32 and obs.shape == (B, 8, 20, 20) and obs.is_cuda
    assert out.mask.dtype == torch.bool and out.mask.shape == (B, 30433) and out.mask.stride(1) == 1
    logits = torch.randn((B, 30433), device=obs.device)
    sel = torch.masked_select(logits[0], out.mask[0])              # neural_network.py:169
    assert sel.numel() == int(out.legal_count[0])
    mover = (s[:, 84] & 15).long()
    assert (obs[torch.arange(B), 4 + mover] == 1).all() and obs[:, 4:].sum() == B * 400
