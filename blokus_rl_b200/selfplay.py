"""Batched self-play and arena drivers.

The reference plays ONE game at a time: ``AlphaZeroTrainer._self_play`` (``alphazero/trainer.py:92-137``) and
``play_match`` / ``play_single_match`` (``alphazero/arena.py:10-87``).  Here B games advance in lockstep: every
simulation level of all B searches is one ``blk_step`` launch and every expansion wave one batched evaluator
call (see :mod:`blokus_rl_b200.mcts`).  Outputs keep the reference's formats:

* self-play examples ``[obs float32 [2P,N,N], mask float64 [A], prob float32 [n_valid], scores float64 [P]]``
  per ply (``trainer.py:118-121``), pickled one list per file as ``*.examples`` (``trainer.py:287-292``) so
  ``AlphaZeroDataset`` (``alphazero/dataset.py:29-47``) reads them unchanged;
* arena scores: sum over games of the 3/1/-1 terminal vectors, re-indexed by player order (``arena.py:52-54``).
"""
from __future__ import annotations

from itertools import permutations
from pathlib import Path
from pickle import Pickler

import numpy as np
import torch

from .mcts import BatchedMCTS, UniformEvaluator


def self_play_batched(engine, evaluator=None, num_games: int = 8, num_mcts_sims: int = 25, cpuct: float = 1.0,
                      temperature: float = 1.0, dirichlet_alpha: float = 1.0, dirichlet_weight: float = 0.25,
                      rng: np.random.Generator | None = None, max_plies: int = 4 * 21 + 1):
    """Play ``num_games`` self-play games concurrently.  Returns ``(examples_per_game, stats)``."""
    rng = rng or np.random.default_rng()
    search = BatchedMCTS(engine, evaluator or UniformEvaluator())
    roots = search.add_roots(engine.new_states(num_games))
    A, P = engine.num_actions, engine.num_players
    data: list[list] = [[] for _ in range(num_games)]
    final: list = [None] * num_games
    active = [t for t in range(num_games)]
    first = True
    plies = 0
    while active and plies < max_plies:
        cur = [roots[t] for t in active]
        for _ in range(num_mcts_sims):                                   # trainer.py:104-105
            search.simulate(cur, cpuct, tree_ids=active)
        obs = engine.observe(search.rows(cur)).cpu().numpy()
        chosen = []
        for j, t in enumerate(active):
            ids, dist = search.get_distribution(t, cur[j], temperature)  # trainer.py:108
            if first:                                                    # Dirichlet noise on the first move only
                noise = rng.dirichlet(dirichlet_alpha * np.ones(len(ids), dtype=np.float32))
                dist = dist * (1 - dirichlet_weight) + noise * dirichlet_weight
            prob = dist.astype(np.float32)
            mask = np.zeros(A, dtype=np.float64)
            mask[ids] = 1
            data[t].append([obs[j], mask, prob, None])
            p64 = prob.astype(np.float64)
            chosen.append(int(ids[rng.choice(len(ids), p=p64 / p64.sum())]))
        first = False
        nxt = search.child_states(cur, chosen)
        still = []
        for j, t in enumerate(active):
            roots[t] = nxt[j]
            if nxt[j].terminal is not None:
                final[t] = np.asarray(nxt[j].terminal, dtype=np.float64)
                for ex in data[t]:
                    ex[-1] = final[t]
            else:
                still.append(t)
        active = still
        plies += 1
    stats = {"games": num_games, "examples": sum(len(d) for d in data), "launches": search.launches, "plies": plies}
    return data, stats


def self_play_gpu(engine, evaluator=None, num_games: int = 256, num_mcts_sims: int = 25, cpuct: float = 1.0,
                  temperature: float = 1.0, dirichlet_alpha: float = 1.0, dirichlet_weight: float = 0.25,
                  rng: np.random.Generator | None = None, max_plies: int = 4 * 21 + 1, mean_edges_per_node: int = 384):
    """Same contract as :func:`self_play_batched`, with the search trees resident on the GPU
    (:class:`blokus_rl_b200.gpu_puct.GpuPuct`): the host only samples one move per game per ply."""
    from .gpu_puct import GpuPuct
    # `rng` may be one generator per game (sharded runs: game g always draws from its own stream, whatever the partition)
    rngs = list(rng) if isinstance(rng, (list, tuple)) else [rng or np.random.default_rng()] * num_games
    A, P = engine.num_actions, engine.num_players
    search = GpuPuct(engine, evaluator or UniformEvaluator(), num_trees=num_games,
                     max_simulations=(num_mcts_sims + 1) * max_plies + 2, mean_edges_per_node=mean_edges_per_node)
    search.set_roots(engine.new_states(num_games))
    data: list[list] = [[] for _ in range(num_games)]
    done = np.zeros(num_games, bool)
    first = True
    plies = 0
    while not done.all() and plies < max_plies:
        search.run(num_mcts_sims, cpuct, chain=num_mcts_sims if num_games <= 64 else 1)
        roots = search.root_states()
        obs = engine.observe(roots).cpu().numpy()
        stats = search.root_stats()
        acts = np.full(num_games, -1, np.int32)
        for t in range(num_games):
            if done[t]:
                continue
            ids, n, _, _ = stats[t]
            if temperature == 0:
                dist = np.zeros(len(ids)); dist[int(np.argmax(n))] = 1.0
            else:
                dist = np.power(n, 1.0 / temperature)
            dist = dist / dist.sum() if dist.sum() > 0 else np.full(len(ids), 1.0 / len(ids))
            if first:
                dist = dist * (1 - dirichlet_weight) + rngs[t].dirichlet(dirichlet_alpha * np.ones(len(ids), np.float32)) * dirichlet_weight
            prob = dist.astype(np.float32)
            mask = np.zeros(A, dtype=np.float64)
            mask[ids] = 1
            data[t].append([obs[t], mask, prob, None])
            p64 = prob.astype(np.float64)
            acts[t] = int(ids[rngs[t].choice(len(ids), p=p64 / p64.sum())])
        first = False
        search.advance(torch.as_tensor(acts, device=engine.device))
        flags, term, _ = engine.game_ended(search.root_states())
        ended = (flags.cpu().numpy() & 1).astype(bool) & ~done
        term = term.cpu().numpy().astype(np.float64)
        for t in np.flatnonzero(ended):
            for ex in data[t]:
                ex[-1] = term[t]
        done |= ended
        plies += 1
    search.check()
    return data, {"games": num_games, "examples": sum(len(d) for d in data), "launches": search.launches, "plies": plies}


def save_examples(examples_per_game, save_dir, iteration: int = 0, prefix: str = "checkpoint") -> list[Path]:
    """One pickle per game under ``save_dir/iteration_{k}/`` (reference layout: trainer.py:287-292)."""
    out_dir = Path(save_dir) / f"iteration_{iteration}"
    out_dir.mkdir(parents=True, exist_ok=True)
    files = []
    for g, ex in enumerate(examples_per_game):
        fp = out_dir / f"{prefix}_{g}.examples"
        with open(fp, "wb+") as f:
            Pickler(f).dump(ex)
        files.append(fp)
    return files


# ------------------------------------------------------------------------------------------------
# arena
# ------------------------------------------------------------------------------------------------
class RandomSeat:
    """Uniform-random legal moves (players/random_player.py) from the engine's on-device sampler."""
    kind = "random"


class MCTSSeat:
    """players/mcts_player.py: ``simulations`` PUCT simulations, then the most visited action."""
    kind = "mcts"

    def __init__(self, evaluator=None, simulations: int = 10, cpuct: float = 1.0):
        self.evaluator, self.simulations, self.cpuct = evaluator or UniformEvaluator(), simulations, cpuct


class RolloutSeat:
    """Flat Monte-Carlo: the move with the best mean playout value (GPU rollouts)."""
    kind = "rollout"

    def __init__(self, per_move: int = 16):
        self.per_move = per_move


def play_match_batched(engine, seats, games_num: int, permute: bool = False, seed: int = 0, max_plies: int = 4 * 21 + 1):
    """``games_num`` concurrent games between ``seats`` (one policy per player index).  With ``permute`` game i
    uses the i-th permutation of the seats, as ``play_match`` does (arena.py:33-38, 47).  Returns
    ``(scores [P], per_game_terminal [games, P] in board-player order, orders [games, P])``."""
    P, dev = engine.num_players, engine.device
    assert len(seats) == P
    perms = list(permutations(range(P))) if permute else [tuple(range(P))]
    orders = np.array([perms[i % len(perms)] for i in range(games_num)])
    states = engine.new_states(games_num)
    meta = P * engine.board_size + P
    searches = {k: BatchedMCTS(engine, s.evaluator) for k, s in enumerate(seats) if s.kind == "mcts"}
    for srch in searches.values():
        srch.trees = [dict() for _ in range(games_num)]
    out = engine.step(states, None, mask="bytes", sample=True, seed=seed)
    terminal = np.zeros((games_num, P))
    finished = np.zeros(games_num, bool)
    rollout_calls = 0
    for ply in range(max_plies):
        host = states[:, meta].cpu().numpy()
        mover = host & 15
        done = ((host >> 4) & 1).astype(bool)
        if done.all():
            break
        seat_of = orders[np.arange(games_num), mover]                     # arena.py:75: p = order[current_player]
        actions = torch.full((games_num,), -1, dtype=torch.int32, device=dev)
        for k, seat in enumerate(seats):
            games = np.flatnonzero((seat_of == k) & ~done)
            if len(games) == 0:
                continue
            gidx = torch.as_tensor(games, device=dev)
            if seat.kind == "random":
                actions[gidx] = out.next_action[gidx]
            elif seat.kind == "mcts":
                srch = searches[k]
                roots = srch.register(states.index_select(0, gidx))
                for _ in range(seat.simulations):
                    srch.simulate(roots, seat.cpuct, tree_ids=list(games))
                pick = [int(srch.stats(int(g), r)[0][int(np.argmax(srch.stats(int(g), r)[1]))]) for g, r in zip(games, roots)]
                actions[gidx] = torch.tensor(pick, dtype=torch.int32, device=dev)
            else:                                                          # rollout seat: all (game, move) children at once
                nz = torch.nonzero(out.mask.index_select(0, gidx))
                src = states.index_select(0, gidx[nz[:, 0]])
                kids = torch.empty_like(src)
                engine.step(src, nz[:, 1].to(torch.int32).contiguous(), out_states=kids, mask=None, want_count=False)
                ro = engine.rollout(kids, seat.per_move, seed=seed, rollout_id_base=rollout_calls)
                rollout_calls += kids.shape[0] * seat.per_move
                val = ro.value_sum[torch.arange(kids.shape[0], device=dev), torch.as_tensor(mover[games], device=dev)[nz[:, 0]].long()]
                best = torch.full((len(games),), -1e30, device=dev).scatter_reduce(0, nz[:, 0], val, "amax")
                is_best = val >= best[nz[:, 0]]
                cand = torch.where(is_best, nz[:, 1], torch.full_like(nz[:, 1], 1 << 30))
                first_best = torch.full((len(games),), 1 << 30, device=dev, dtype=nz.dtype).scatter_reduce(0, nz[:, 0], cand, "amin")
                actions[gidx] = first_best.to(torch.int32)
        out = engine.step(states, actions, mask="bytes", sample=True, seed=seed + 1 + ply)
        fl = out.flags.cpu().numpy()
        if (fl & 2).any():
            raise RuntimeError("a seat produced an illegal action")
        newly = ((fl & 1) != 0) & ~finished
        if newly.any():
            terminal[newly] = out.terminal.cpu().numpy()[newly]
            finished |= newly
    scores = np.zeros(P)
    for g in range(games_num):
        scores[list(orders[g])] += terminal[g]                             # arena.py:52-54
    return scores, terminal, orders
