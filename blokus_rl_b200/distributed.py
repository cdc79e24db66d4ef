"""Multi-GPU: env batches shard by index range, one process per GPU, no collective on the hot path.

The reference is single-process (``blokus_rl/ppo/trainer.py:33-38``).  Here rank r of G owns the global envs
``[r*E/G, (r+1)*E/G)``; every RNG key uses the GLOBAL env id (``blk_step_args.env_id_base``), so trajectories
do not depend on the GPU count.  The only collective is one ``all_reduce(SUM)`` of a small int64 counter
vector at the end of a run (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist

COUNTERS = ("steps", "games", "illegal", "legal_actions_sum", "score_sum", "wins_p0", "wins_p1", "wins_p2", "wins_p3")


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    total_envs: int

    @property
    def lo(self) -> int:
        return self.rank * self.total_envs // self.world

    @property
    def hi(self) -> int:
        return (self.rank + 1) * self.total_envs // self.world

    @property
    def n(self) -> int:
        return self.hi - self.lo


def shard_from_env(total_envs: int) -> Shard:
    return Shard(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), total_envs)


def init(backend: str | None = None, device: torch.device | None = None) -> None:
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)


def random_play_shard(engine, shard: Shard, plies: int, seed: int, fmt: str | None = "bytes") -> torch.Tensor:
    """Uniform-random play of this rank's envs for ``plies`` plies with auto-reset; returns the local counter
    vector (int64, order = COUNTERS).  No communication."""
    n, dev, P = shard.n, engine.device, engine.num_players
    c = torch.zeros(len(COUNTERS), dtype=torch.int64, device=dev)
    if n == 0:
        return c
    states = engine.new_states(n)
    out = engine.step(states, None, mask=fmt, sample=True, seed=seed, env_id_base=shard.lo)
    for _ in range(plies):
        c[3] += out.legal_count.sum()
        out = engine.step(states, out.next_action, mask=fmt, sample=True, seed=seed, env_id_base=shard.lo,
                          auto_reset=True)
        done = (out.flags & 1).bool()
        c[0] += n
        c[1] += done.sum()
        c[2] += ((out.flags & 2) != 0).sum()
        c[4] += out.scores[done].sum()
        wins = (out.terminal[done] > 0).sum(0)
        c[5:5 + P] += wins
    return c


def reduce_counters(local: torch.Tensor, names=COUNTERS) -> dict:
    """The one collective of a run."""
    total = local.clone()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return dict(zip(names, (int(x) for x in total.cpu())))


# ---------------------------------------------------------------------------------------------------------------
# The other batch workloads shard the same way: by ROOT index (playouts, searches) or GAME index (self-play), with the
# global index in every RNG key, so the union of the ranks' results is the single-GPU result.  The reference loops they
# replace are sequential: blokus_rl/alphazero/trainer.py:152-154 (self-play games), alphazero/arena.py:45-58 (matches).
# ---------------------------------------------------------------------------------------------------------------
ROLLOUT_COUNTERS = ("playouts", "plies", "wins_p0", "wins_p1", "wins_p2", "wins_p3", "score_sum", "score_hash")
PUCT_COUNTERS = ("trees", "simulations", "nodes", "edges", "best_action_sum", "root_visits", "overflow")
SELFPLAY_COUNTERS = ("games", "plies", "examples", "wins_p0", "wins_p1", "wins_p2", "wins_p3")


def midgame_roots(engine, lo: int, n: int, plies: int = 24, seed: int = 24) -> torch.Tensor:
    """The fixed mid-game roots of BASELINE.json configs[2] for the global root ids ``[lo, lo + n)``: the state of env
    ``lo + i`` after ``plies`` random plies from reset (key = global id, so every partition builds the same roots)."""
    roots = engine.new_states(n)
    out = engine.step(roots, None, mask=None, sample=True, seed=seed, env_id_base=lo)
    for _ in range(plies):
        out = engine.step(roots, out.next_action, mask=None, sample=True, seed=seed, env_id_base=lo)
    return roots


def rollout_shard(engine, shard: Shard, per_root: int, seed: int, roots: torch.Tensor | None = None, root_plies: int = 24):
    """This rank's roots ``[shard.lo, shard.hi)`` x ``per_root`` playouts to the end of the game (global playout index
    ``root * per_root + j`` keys the RNG).  Returns (local counters int64 in ROLLOUT_COUNTERS order, RolloutOut)."""
    c = torch.zeros(len(ROLLOUT_COUNTERS), dtype=torch.int64, device=engine.device)
    if shard.n == 0:
        return c, None
    if roots is None:
        roots = midgame_roots(engine, shard.lo, shard.n, root_plies)
    out = engine.rollout(roots, per_root, seed=seed, rollout_id_base=shard.lo * per_root)
    P = engine.num_players
    c[0] = shard.n * per_root
    c[1] = out.plies.sum()
    w = out.winners.long()
    for q in range(P):
        c[2 + q] = ((w >> q) & 1).sum()
    sc = out.final_scores.long()
    c[6] = sc.sum()
    # order-independent fingerprint of (global playout id, final scores): equal for every partition of the roots
    gid = (torch.arange(shard.n * per_root, device=engine.device, dtype=torch.int64) + shard.lo * per_root).view(shard.n, per_root)
    mix = (sc * torch.tensor([1, 131, 17161, 2248091][:P], device=engine.device)).sum(-1)
    c[7] = ((2 * gid + 1) * (mix + 1)).sum()
    return c, out


def puct_shard(engine, shard: Shard, simulations: int, evaluator=None, cpuct: float = 1.0, root_plies: int = 24,
               mean_edges_per_node: int = 420, chain: int = 1):
    """This rank's searches: one PUCT tree per global root id in ``[shard.lo, shard.hi)``, ``simulations`` simulations
    each, trees resident on the GPU.  Returns (local counters in PUCT_COUNTERS order, the forest)."""
    from .gpu_puct import GpuPuct
    c = torch.zeros(len(PUCT_COUNTERS), dtype=torch.int64, device=engine.device)
    if shard.n == 0:
        return c, None
    roots = midgame_roots(engine, shard.lo, shard.n, root_plies)
    search = GpuPuct(engine, evaluator, num_trees=shard.n, max_simulations=simulations + 4,
                     mean_edges_per_node=mean_edges_per_node)
    search.set_roots(roots)
    search.run(simulations, cpuct, chain=chain)
    best = search.best_actions_device().long()
    ctr = search.t["counters"].long()
    c[0], c[1] = shard.n, shard.n * simulations
    c[2], c[3] = ctr[0], ctr[1]
    gid = torch.arange(shard.n, device=engine.device, dtype=torch.int64) + shard.lo
    c[4] = ((2 * gid + 1) * (best + 2)).sum()
    c[5] = search.t["node_sum_n"].index_select(0, search.t["root"].long()).sum().long()
    c[6] = ctr[2] + ctr[3]
    return c, search


def self_play_shard(engine, shard: Shard, num_mcts_sims: int = 25, evaluator=None, seed: int = 0, **kw):
    """This rank's self-play games ``[shard.lo, shard.hi)`` (alphazero/trainer.py:152-154 plays them one after the
    other).  The host-side move sampling of game g uses ``numpy.random.default_rng([seed, g])``.  Returns (local counters
    in SELFPLAY_COUNTERS order, examples per game)."""
    import numpy as np
    from .selfplay import self_play_gpu
    c = torch.zeros(len(SELFPLAY_COUNTERS), dtype=torch.int64, device=engine.device)
    if shard.n == 0:
        return c, []
    rngs = [np.random.default_rng([seed, g]) for g in range(shard.lo, shard.hi)]
    data, stats = self_play_gpu(engine, evaluator, num_games=shard.n, num_mcts_sims=num_mcts_sims, rng=rngs, **kw)
    c[0], c[1], c[2] = shard.n, stats["plies"], stats["examples"]
    for ex in data:
        if ex and ex[-1][-1] is not None:
            for q, v in enumerate(ex[-1][-1]):
                c[3 + q] += int(v > 0)
    return c, data
