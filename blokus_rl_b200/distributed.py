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


def reduce_counters(local: torch.Tensor) -> dict:
    """The one collective of a run."""
    total = local.clone()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return dict(zip(COUNTERS, (int(x) for x in total.cpu())))
