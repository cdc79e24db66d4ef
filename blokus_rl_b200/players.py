"""Players behind the reference's ``Player`` interface (``blokus_rl/players/player.py:5-14``):
``update_state(s, current_player) -> (s', next_player)`` and ``reset()``.

They take a game wrapper (boundary B1: :class:`blokus_rl_b200.game_wrapper.BlokusGameWrapper`, whose states are
engine state handles) and plug into the reference's arena (``alphazero/arena.py:66-87``) unchanged.

* :class:`RandomPlayer`   == ``players/random_player.py:11-17``
* :class:`MCTSPlayer`     == ``players/mcts_player.py:15-28`` on :class:`BatchedMCTS` (tree kept across moves)
* :class:`RolloutPlayer`  new: flat Monte-Carlo over GPU playouts (kernel family 4)
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np
import torch

from .mcts import BatchedMCTS, UniformEvaluator


class Player(ABC):
    @abstractmethod
    def update_state(self, s, current_player):
        raise NotImplementedError

    @abstractmethod
    def reset(self):
        raise NotImplementedError


class RandomPlayer(Player):
    def __init__(self, game):
        self.game = game

    def update_state(self, s, current_player):
        return self.game.get_next_state(s, current_player, self.game.get_sample_move(s))

    def reset(self):
        return

    def __str__(self):
        return "RandomPlayer"


class MCTSPlayer(Player):
    """``simulations`` PUCT simulations per move, then the most visited action (first maximum); the tree outlives the
    move, as the reference's dict does (players/mcts_player.py:15-25).

    With an evaluator that needs no network (the default :class:`UniformEvaluator` == the reference's DumbNet "mcts"
    player, compare_arena.py:87-95; or :class:`RolloutEvaluator`) the ONE tree lives on the GPU and a whole move's search
    is a single launch of the fused search kernel (``blk_puct_search``): ``warps_per_tree = 1`` plays exactly the
    reference's search (same visit counts), ``warps_per_tree > 1`` is the faster leaf-parallel search with virtual loss
    (``reference_compat=False`` selects 16 warps).  Network evaluators use the host-side :class:`BatchedMCTS`."""

    def __init__(self, game, evaluator=None, simulations: int = 10, cpuct: float = 1.0, reference_compat: bool = True,
                 warps_per_tree: int | None = None, max_moves: int = 4 * 21 + 2):
        self.game, self.simulations, self.cpuct = game, simulations, cpuct
        self.evaluator = evaluator or UniformEvaluator()
        eng = game.backend.eng
        on_gpu = getattr(getattr(eng, "device", None), "type", "cpu") == "cuda"
        fusable = type(self.evaluator).__name__ in ("UniformEvaluator", "RolloutEvaluator")
        self.gpu = None
        if on_gpu and fusable:
            from .gpu_puct import GpuPuct
            wpt = warps_per_tree if warps_per_tree is not None else (1 if reference_compat else 16)
            self.gpu = GpuPuct(eng, self.evaluator, num_trees=1, max_simulations=(simulations + 1) * max_moves,
                               mean_edges_per_node=min(eng.num_actions, 600), warps_per_tree=wpt)
            self._fresh = True
        else:
            self.search = BatchedMCTS(eng, self.evaluator)
            self.search.trees = [dict()]
            self._known = {}

    def _root(self, s):
        key = s.host_words.tobytes()
        st = self._known.get(key)
        if st is None:
            st = self.search.register(s.words)[0]
            self._known[key] = st
        return st

    def update_state(self, s, current_player):
        if self.gpu is not None:
            if self._fresh:
                self.gpu.set_roots(s.words)
                self._fresh = False
            else:
                self.gpu.reroot(s.words)
            self.gpu.run(self.simulations, self.cpuct)
            action = int(self.gpu.best_actions()[0])
            self.gpu.check()
            return self.game.get_next_state(s, current_player, action)
        root = self._root(s)
        for _ in range(self.simulations):
            self.search.simulate([root], self.cpuct)
        ids, dist = self.search.get_distribution(0, root, 0)
        return self.game.get_next_state(s, current_player, int(ids[int(np.argmax(dist))]))

    def reset(self):
        if self.gpu is not None:
            self._fresh = True
            return
        self.search.reset()
        self.search.trees = [dict()]
        self._known = {}

    def __str__(self):
        return "MCTSPlayer"


class RolloutPlayer(Player):
    """Evaluate every legal move by ``per_move`` uniform-random playouts on the GPU and play the move with the
    best mean terminal value for the mover."""

    def __init__(self, game, per_move: int = 32, seed: int = 0):
        self.game, self.per_move, self.seed = game, per_move, seed
        self._calls = 0

    def update_state(self, s, current_player):
        eng = self.game.backend.eng
        ids = self.game.backend.legal_ids(s)
        m = len(ids)
        src = s.words.expand(m, -1).contiguous()
        children = torch.empty_like(src)
        eng.step(src, torch.as_tensor(ids, dtype=torch.int32, device=src.device), out_states=children, mask=None,
                 want_count=False)
        out = eng.rollout(children, self.per_move, seed=self.seed, rollout_id_base=self._calls)
        self._calls += m * self.per_move
        best = int(torch.argmax(out.value_sum[:, self.game.backend.mover(s)]).item())
        return self.game.get_next_state(s, current_player, int(ids[best]))

    def reset(self):
        return

    def __str__(self):
        return "RolloutPlayer"
