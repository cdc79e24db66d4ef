"""Batched PUCT search over the GPU engine — the sibling of ``blokus_rl/alphazero/mcts.py``.

The reference runs ONE recursive, dict-based search that calls the env and a batch-1 network once per
simulation (``mcts.py:13-71``, ``neural_network.py:92-110``).  :class:`BatchedMCTS` runs B independent
searches (one per game / root) in lockstep: every descent level of all searches becomes ONE ``blk_step``
launch over the gathered states, every expansion wave ONE batched evaluator call fed with contiguous GPU
tensors (``obs float32 [m,2P,N,N]``, ``mask bool [m,A]``).  Per tree the arithmetic is the reference's,
quirks included:

  1. backup uses the value of the player to move AFTER the action (``mcts.py:47,54``);
  2. only the root call sees ``cpuct``; deeper levels use 1 (``mcts.py:50-52``);
  3. ``U = cpuct * P * sqrt(sum(N) + 1e-6) / (1 + N)``, first maximum wins (``mcts.py:43-46``), float64;
  4. terminal states are never stored: each visit returns the 3/1/-1 vector again (``mcts.py:60-62``);
  5. the tree key is the board cells only, not the mover or inventories (``blokus_wrapper.py:217-218``).

Evaluators replace ``nn.predict``: :class:`UniformEvaluator` (== DumbNet, ``models/dumbnet.py:14-21``),
:class:`RolloutEvaluator` (GPU random playouts; new capability) and :class:`TorchNetEvaluator` (any
``model(obs) -> (logits [m,A], v [m,P])``, masked softmax == ``neural_network.py:159-173``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------
# evaluators:  evaluate(engine, states int32 [m, SW], mask bool [m, A]) -> (p [m, A] float, v [m, P] float)
#   An evaluator with ``uniform_prior = True`` also offers values(engine, states) -> v [m, P] or None (zeros): the device forest
#   then expands with P = 1/n itself and never materialises the dense prior.
#   An evaluator with ``wants_obs = True`` is also handed ``obs=`` (float32 [m, 2P, N, N]) when the caller already has
#   it from the step kernel's fused observation output (GpuPuct), and computes it itself otherwise.
# ------------------------------------------------------------------------------------------------
class UniformEvaluator:
    graph_safe = True          # no host work: GpuPuct may capture it into a CUDA graph
    uniform_prior = True       # prior = 1/n over the legal actions; values() is all GpuPuct needs (None = zeros)

    def values(self, engine, states):
        return None

    def evaluate(self, engine, states, mask):
        cnt = mask.sum(1, keepdim=True).clamp(min=1).to(torch.float64)
        return mask.to(torch.float64) / cnt, torch.zeros((states.shape[0], engine.num_players), dtype=torch.float64,
                                                         device=states.device)


class RolloutEvaluator:
    """Uniform priors, value = mean 3/1/-1 vector over ``per_leaf`` uniform-random GPU playouts."""
    uniform_prior = True       # GpuPuct asks only for values(): no dense [m, A] float64 prior is built per simulation

    def __init__(self, per_leaf: int = 32, seed: int = 0):
        self.per_leaf, self.seed, self._calls = per_leaf, seed, 0

    def values(self, engine, states):
        out = engine.rollout(states, self.per_leaf, seed=self.seed, rollout_id_base=self._calls)
        self._calls += states.shape[0] * self.per_leaf
        return (out.value_sum / self.per_leaf).to(torch.float64)

    def evaluate(self, engine, states, mask):
        cnt = mask.sum(1, keepdim=True).clamp(min=1).to(torch.float64)
        return mask.to(torch.float64) / cnt, self.values(engine, states)


class TorchNetEvaluator:
    graph_safe = True          # static-shape torch ops only
    wants_obs = True           # blk_step can write the observation in the same launch (blk_step_args.obs)

    def __init__(self, model: torch.nn.Module):
        self.model = model

    @torch.inference_mode()
    def evaluate(self, engine, states, mask, obs=None):
        self.model.eval()
        logits, v = self.model(engine.observe(states) if obs is None else obs)
        logits = logits.float().masked_fill(~mask, float("-inf"))
        return torch.softmax(logits, dim=-1).to(torch.float64), v.to(torch.float64)


# ------------------------------------------------------------------------------------------------
@dataclass
class _Node:
    ids: np.ndarray            # legal action ids, ascending
    N: np.ndarray              # float64 visit counts
    Q: np.ndarray              # float64 running means
    P: np.ndarray              # float64 priors over `ids`


@dataclass
class _State:
    slot: int                  # row in the device pool
    key: bytes                 # board cells only (quirk 5)
    mover: int
    terminal: np.ndarray | None   # 3/1/-1 vector when the game is over
    ids: np.ndarray | None        # legal ids of the mover
    child: dict = field(default_factory=dict)     # action id -> _State reached from THIS state (functional next_state)


class BatchedMCTS:
    def __init__(self, engine, evaluator=None, capacity: int = 1 << 14):
        self.eng = engine
        self.evaluator = evaluator or UniformEvaluator()
        self.P, self.A = engine.num_players, engine.num_actions
        self._nrow = engine.num_players * engine.board_size
        self._meta = self._nrow + engine.num_players
        self.pool = torch.empty((capacity, engine.state_words), dtype=torch.int32, device=engine.device)
        self.used = 0
        self.trees: list[dict] = []
        self.launches = 0

    # ---- state pool -----------------------------------------------------------------------------------
    def _grow(self, extra: int):
        if self.used + extra > self.pool.shape[0]:
            cap = max(2 * self.pool.shape[0], self.used + extra)
            new = torch.empty((cap, self.pool.shape[1]), dtype=torch.int32, device=self.pool.device)
            new[: self.used] = self.pool[: self.used]
            self.pool = new

    def _register(self, rows: torch.Tensor, out) -> list[_State]:
        """Store freshly produced states (device rows + the step outputs that came with them)."""
        m = rows.shape[0]
        self._grow(m)
        self.pool[self.used: self.used + m] = rows
        host = rows.cpu().numpy().view(np.uint32)
        flags = out.flags.cpu().numpy()
        term = out.terminal.cpu().numpy().astype(np.float64)
        nz = torch.nonzero(out.mask).cpu().numpy()
        splits = np.searchsorted(nz[:, 0], np.arange(m + 1))
        states = []
        for i in range(m):
            done = bool((host[i, self._meta] >> 4) & 1)
            states.append(_State(self.used + i, host[i, : self._nrow].tobytes(), int(host[i, self._meta] & 15),
                                 term[i] if done else None, nz[splits[i]: splits[i + 1], 1].astype(np.int64)))
        self.used += m
        return states

    def register(self, states: torch.Tensor) -> list[_State]:
        """Bring external states (int32 [B, SW]) into the pool; does not touch the trees."""
        states = states.contiguous()
        out = self.eng.step(states, None, mask="bytes")
        self.launches += 1
        return self._register(states, out)

    def add_roots(self, states: torch.Tensor) -> list[_State]:
        """Register root states and start one empty tree per root."""
        roots = self.register(states)
        self.trees = [dict() for _ in roots]
        return roots

    # ---- B simulations, one per tree, in lockstep ------------------------------------------------------------
    def simulate(self, roots: list[_State], cpuct: float = 1.0, epsilon_fix: bool = True,
                 tree_ids: list[int] | None = None) -> np.ndarray:
        """One simulation from ``roots[i]`` in tree ``tree_ids[i]`` (default: tree i) for every i, in lockstep.
        Returns the score vectors [B, P]."""
        B = len(roots)
        tid = list(range(B)) if tree_ids is None else tree_ids
        cur = list(roots)
        paths: list[list] = [[] for _ in range(B)]
        scores: list = [None] * B
        depth = 0
        active = list(range(B))
        expand: list[int] = []
        while active:
            need: list[tuple[int, int]] = []
            nxt_active = []
            for t in active:
                s = cur[t]
                node = self.trees[tid[t]].get(s.key)
                if node is None:                                   # leaf: expand (or terminal)
                    if s.terminal is not None:
                        scores[t] = s.terminal
                    else:
                        expand.append(t)
                    continue
                c = cpuct if depth == 0 else 1.0                   # quirk 2
                eps = 1e-6 if (epsilon_fix or depth > 0) else 0.0
                u = c * node.P * math.sqrt(node.N.sum() + eps) / (1.0 + node.N)
                a = int(np.argmax(node.Q + u))                     # first maximum (quirk 3)
                aid = int(node.ids[a])
                nxt = s.child.get(aid)
                paths[t].append([node, a, nxt])
                if nxt is None:
                    need.append((t, aid))
                else:
                    cur[t] = nxt
                nxt_active.append(t)
            if need:                                               # one launch for this level's new edges
                idx = torch.tensor([cur[t].slot for t, _ in need], device=self.pool.device)
                acts = torch.tensor([aid for _, aid in need], dtype=torch.int32, device=self.pool.device)
                src = self.pool.index_select(0, idx)
                dst = torch.empty_like(src)
                out = self.eng.step(src, acts, out_states=dst, mask="bytes")
                self.launches += 1
                new = self._register(dst, out)
                if int((out.flags & 2).sum().item()):
                    raise RuntimeError("BatchedMCTS selected an illegal action (engine/tree mismatch)")
                for (t, aid), st in zip(need, new):
                    cur[t].child[aid] = st
                    paths[t][-1][2] = st
                    cur[t] = st
            active = nxt_active
            depth += 1
        if expand:                                                 # one evaluator call for all leaves
            idx = torch.tensor([cur[t].slot for t in expand], device=self.pool.device)
            rows = self.pool.index_select(0, idx)
            mask = self.eng.step(rows, None, mask="bytes", want_count=False, want_terminal=False,
                                 want_scores=False).mask
            self.launches += 1
            p, v = self.evaluator.evaluate(self.eng, rows, mask)
            v = v.cpu().numpy().astype(np.float64)
            lens = [len(cur[t].ids) for t in expand]
            ridx = torch.from_numpy(np.repeat(np.arange(len(expand)), lens)).to(p.device)
            cidx = torch.from_numpy(np.concatenate([cur[t].ids for t in expand])).to(p.device)
            pv = p[ridx, cidx].cpu().numpy().astype(np.float64)
            off = 0
            for j, t in enumerate(expand):
                s = cur[t]
                n = lens[j]
                self.trees[tid[t]][s.key] = _Node(s.ids, np.zeros(n), np.zeros(n), pv[off: off + n].copy())
                off += n
                scores[t] = v[j]
        for t in range(B):                                         # backup, deepest edge first
            sc = scores[t]
            path = paths[t]
            for d in range(len(path) - 1, -1, -1):
                node, a, child = path[d]
                val = sc[child.mover]                              # quirk 1: the NEXT player's value
                n, q = node.N[a], node.Q[a]
                node.Q[a] = (n * q + val) / (n + 1)
                node.N[a] = n + 1
        return np.stack(scores)

    # ---- visit-count policy (mcts.py:73-99) ---------------------------------------------------------------------
    def get_distribution(self, tree_index: int, state: _State, temperature: float):
        node = self.trees[tree_index][state.key]
        counts = node.N
        if temperature == 0:
            raised = np.zeros_like(counts)
            raised[int(np.argmax(counts))] = 1.0
        else:
            raised = np.power(counts, 1.0 / temperature)
        total = raised.sum()
        if total == 0:
            raised = np.ones_like(counts)
            total = raised.sum()
        return node.ids, raised / total

    def stats(self, tree_index: int, state: _State):
        node = self.trees[tree_index][state.key]
        return node.ids, node.N, node.Q, node.P

    def child_states(self, states: list[_State], action_ids: list[int]) -> list[_State]:
        """Successors of ``states[i]`` under ``action_ids[i]`` (cached edges are reused, the rest is one launch)."""
        out: list = [s.child.get(int(a)) for s, a in zip(states, action_ids)]
        miss = [i for i, c in enumerate(out) if c is None]
        if miss:
            idx = torch.tensor([states[i].slot for i in miss], device=self.pool.device)
            acts = torch.tensor([int(action_ids[i]) for i in miss], dtype=torch.int32, device=self.pool.device)
            src = self.pool.index_select(0, idx)
            dst = torch.empty_like(src)
            res = self.eng.step(src, acts, out_states=dst, mask="bytes")
            self.launches += 1
            if int((res.flags & 2).sum().item()):
                raise ValueError("illegal action")
            for i, st in zip(miss, self._register(dst, res)):
                states[i].child[int(action_ids[i])] = st
                out[i] = st
        return out

    def rows(self, states: list[_State]) -> torch.Tensor:
        """Device rows (int32 [m, SW]) of pooled states."""
        return self.pool.index_select(0, torch.tensor([s.slot for s in states], device=self.pool.device))

    def reset(self):
        self.trees = [dict() for _ in self.trees]
        self.used = 0
