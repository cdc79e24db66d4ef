"""Device-resident batched PUCT search: B trees live in GPU arrays and every
simulation of all B trees is three kernel launches (+ the evaluator) with no host synchronisation:

    blk_puct_select  ->  blk_step (opened edges, legal masks)  ->  evaluator  ->  blk_puct_expand (+ backup, fused)

Per tree the arithmetic is that of ``blokus_rl/alphazero/mcts.py`` in float64 (see csrc/blk_puct.cu for the
quirks it reproduces); ``tests/test_gpu_puct.py`` checks visit counts / Q / per-simulation score vectors against
golden vectors produced by the unmodified reference file.  Unlike the host-side :class:`BatchedMCTS`, nodes are
keyed by path, not by ``hash(board cells)``; the two coincide unless two move orders reach the same board inside
one search.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .mcts import UniformEvaluator


class GpuPuct:
    def __init__(self, engine, evaluator=None, num_trees: int = 256, max_simulations: int = 4096,
                 mean_edges_per_node: int = 384, max_depth: int = 96, use_cuda_graph: bool = True,
                 fused: bool | None = None, warps_per_tree: int = 1, virtual_loss: float = 1.0):
        """``max_simulations`` bounds nodes per tree (one per simulation + roots); ``mean_edges_per_node`` sizes the
        edge arrays (32 B per edge; 20x20 positions have 58-760 legal moves, ~170 on average over a game and
        300-450 around plies 8-24).  Overflow is detected on the device and reported by :meth:`check`.

        ``fused`` (default: on whenever the evaluator needs no network -- :class:`UniformEvaluator`, the reference's DumbNet,
        or :class:`RolloutEvaluator`): whole simulations run inside ONE kernel (``blk_puct_search``, csrc/blk_search.cuh), a
        warp per tree, nodes keyed by board cells like the reference's dict, so ``run(k)`` is one launch whatever k is.
        ``warps_per_tree > 1`` (fused only) searches each tree with that many warps at once under virtual loss:
        leaf-parallel, faster for few trees, NOT the reference's visit order."""
        self.eng = engine
        self.evaluator = evaluator or UniformEvaluator()
        ev = self.evaluator
        self.playouts_per_leaf = int(getattr(ev, "per_leaf", 0)) if type(ev).__name__ == "RolloutEvaluator" else 0
        can_fuse = type(ev).__name__ in ("UniformEvaluator", "RolloutEvaluator") and bool(getattr(ev, "uniform_prior", False))
        if fused and not can_fuse:
            raise ValueError("fused search needs an evaluator without a network (UniformEvaluator or RolloutEvaluator)")
        self.fused = can_fuse if fused is None else bool(fused)
        self.warps_per_tree = int(warps_per_tree) if self.fused else 1
        if not 1 <= self.warps_per_tree <= 16:
            raise ValueError("warps_per_tree must be in 1..16")
        self.virtual_loss = float(virtual_loss)
        self.B, self.P, self.A = num_trees, engine.num_players, engine.num_actions
        self._lib = _lib.load()
        dev = engine.device
        i32 = dict(dtype=torch.int32, device=dev)
        f64 = dict(dtype=torch.float64, device=dev)
        self.node_cap = num_trees * (max_simulations + 2)
        self.edge_cap = self.node_cap * mean_edges_per_node
        self.max_depth = max_depth
        # Memory: nodes 4 * state_words + 29 + 8 P bytes each (the state pool dominates), edges 32 B each (24 B with a
        # uniform-prior evaluator, which stores no per-edge P).  The C ABI carries capacities as int32.
        if self.node_cap >= 2 ** 31 or self.edge_cap >= 2 ** 31:
            raise ValueError(f"GpuPuct: num_trees * (max_simulations + 2) * mean_edges_per_node = {self.edge_cap} edges does "
                             "not fit the forest's int32 indices; search fewer trees per forest or fewer simulations")
        self.uniform = bool(getattr(self.evaluator, "uniform_prior", False))
        edge_bytes = self.edge_cap * (24 if self.uniform else 32)
        free = torch.cuda.mem_get_info(dev)[0] if dev.type == "cuda" else edge_bytes * 4
        if edge_bytes + self.node_cap * 4 * engine.state_words > 0.9 * free:
            raise ValueError(f"GpuPuct: the forest needs {(edge_bytes + self.node_cap * 4 * engine.state_words) / 2**30:.1f} GiB "
                             f"({self.edge_cap} edges, {self.node_cap} nodes) but {free / 2**30:.1f} GiB are free; lower "
                             "num_trees, max_simulations or mean_edges_per_node")
        t = self.t = {
            "node_edge0": torch.empty(self.node_cap, **i32), "node_nedge": torch.empty(self.node_cap, **i32),
            "node_state": torch.empty(self.node_cap, **i32),
            "node_mover": torch.empty(self.node_cap, dtype=torch.int8, device=dev),
            "node_terminal": torch.empty(self.node_cap, dtype=torch.int8, device=dev),
            "node_term_value": torch.empty((self.node_cap, self.P), **f64),
            "edge_action": torch.empty(self.edge_cap, **i32), "edge_child": torch.empty(self.edge_cap, **i32),
            "edge_n": torch.empty(self.edge_cap, **f64), "edge_q": torch.empty(self.edge_cap, **f64),
            "edge_p": torch.empty(1 if self.uniform else self.edge_cap, **f64),   # uniform prior: P = 1/n, never stored
            "root": torch.empty(self.B, **i32), "path": torch.empty((self.B, max_depth), **i32),
            "path_len": torch.zeros(self.B, **i32), "status": torch.zeros(self.B, **i32),
            "leaf_node": torch.zeros(self.B, **i32), "leaf_edge": torch.zeros(self.B, **i32),
            "src_slot": torch.zeros(self.B, **i32), "step_action": torch.zeros(self.B, **i32),
            "scores": torch.zeros((self.B, self.P), **f64), "counters": torch.zeros(6, **i32),
            "node_sum_n": torch.empty(self.node_cap, **f64), "path_node": torch.empty((self.B, max_depth), **i32),
            "node_uniform": torch.zeros(self.node_cap, dtype=torch.int8, device=dev),
        }
        # board-keyed node table of the fused search (the reference keys its dict by board cells: mcts.py:37)
        self.hash_cap = 1 << max(4, (2 * self.node_cap - 1).bit_length()) if self.fused else 0
        if self.fused:
            t["hash_table"] = torch.zeros(self.hash_cap, **i32)
            t["node_hash"] = torch.empty(self.node_cap, dtype=torch.int64, device=dev)
            t["node_tree"] = torch.empty(self.node_cap, **i32)
            t["edge_vl"] = torch.zeros(self.edge_cap if self.warps_per_tree > 1 else 1, **i32)
            t["node_front"] = torch.zeros(self.node_cap, **i32)
        self.pool = torch.empty((self.node_cap, engine.state_words), dtype=torch.int32, device=dev)
        self.used = 0
        self.forest = _lib.BlkPuctForest(self.B, self.P, self.A, engine.mask_bytes, self.node_cap, self.edge_cap, max_depth,
                                         *[t[n].data_ptr() for n in (
                                             "node_edge0", "node_nedge", "node_state", "node_mover", "node_terminal",
                                             "node_term_value", "edge_action", "edge_child", "edge_n", "edge_q", "edge_p",
                                             "root", "path", "path_len", "status", "leaf_node", "leaf_edge", "src_slot",
                                             "step_action", "scores", "counters", "node_sum_n", "path_node", "node_uniform")],
                                         t["hash_table"].data_ptr() if self.fused else None, self.hash_cap,
                                         t["node_hash"].data_ptr() if self.fused else None,
                                         t["node_tree"].data_ptr() if self.fused else None,
                                         t["edge_vl"].data_ptr() if self.fused and self.warps_per_tree > 1 else None,
                                         t["node_front"].data_ptr() if self.fused else None)
        self._seed = int(getattr(ev, "seed", 0))
        # a net needs the dense bool mask; the uniform prior only needs the legal ids, which the 8x smaller
        # bit-packed mask gives just as well
        self.mask_fmt = "bits" if self.uniform else "bytes"
        self.buf = engine.make_buffers(self.B, self.mask_fmt)
        # a net evaluator gets its input planes from the same launch that produces the new states and masks
        self.obs = (torch.empty((self.B, 2 * self.P, engine.board_size, engine.board_size), dtype=torch.float32, device=dev)
                    if getattr(self.evaluator, "wants_obs", False) else None)
        self.stage = torch.empty((self.B, engine.state_words), dtype=torch.int32, device=dev)   # blk_step output
        self.use_cuda_graph = use_cuda_graph and getattr(self.evaluator, "graph_safe", False)
        self._graphs: dict = {}         # (cpuct, epsilon_fix) -> captured simulation
        self._eager_runs = 0
        self.launches = 0
        self._meta = self.P * engine.board_size + self.P

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.eng.device).cuda_stream)

    def _check(self, rc):
        if rc != 0:
            raise _lib.EngineError(f"blk_puct error {rc}: {self._lib.blk_puct_last_error().decode()}")

    # ---- roots ------------------------------------------------------------------------------------------
    def _search_args(self, num_sims: int, cpuct: float = 1.0, epsilon_fix: bool = True):
        return _lib.BlkPuctSearchArgs(int(num_sims), float(cpuct), int(epsilon_fix), self.pool.data_ptr(), self.warps_per_tree,
                                      self.playouts_per_leaf, self._seed & 0xFFFFFFFFFFFFFFFF, self.virtual_loss)

    def reroot(self, states: torch.Tensor) -> None:
        """Fused search only.  The root of tree t becomes the node the tree has filed under the board of ``states[t]``
        (statistics and subtree kept), or a fresh node: tree reuse the way the reference gets it from a dict that
        outlives the move (players/mcts_player.py:15-25)."""
        if not self.fused:
            raise _lib.EngineError("reroot needs the fused search (board-keyed nodes)")
        states = states.contiguous()
        assert states.shape == (self.B, self.eng.state_words) and states.dtype == torch.int32
        if self.used + self.B > self.node_cap:
            raise _lib.EngineError("GpuPuct state pool exhausted: raise max_simulations")
        args = self._search_args(0)
        _lib.check(self._lib.blk_puct_reroot(self.eng._h, C.byref(self.forest), C.byref(args), states.data_ptr(), self._stream()))
        self.used += self.B
        self.launches += 1

    def set_roots(self, states: torch.Tensor) -> None:
        """Start B fresh trees at ``states`` (int32 [B, state_words])."""
        assert states.shape == (self.B, self.eng.state_words)
        t, B = self.t, self.B
        if self.fused:
            t["hash_table"].zero_()
            t["counters"].zero_()
            if self.warps_per_tree > 1:
                t["edge_vl"].zero_()
            self.used = 0
            self.reroot(states)
            return
        self.pool[:B] = states
        self.used = B
        flags, term, _ = self.eng.game_ended(states)
        ar = torch.arange(B, dtype=torch.int32, device=states.device)
        t["root"].copy_(ar)
        t["node_state"][:B] = ar
        t["node_edge0"][:B] = -1
        t["node_nedge"][:B] = 0
        t["node_sum_n"][:B] = 0
        t["node_mover"][:B] = (states[:, self._meta] & 15).to(torch.int8)
        t["node_terminal"][:B] = (flags & 1).to(torch.int8)
        t["node_term_value"][:B] = term.to(torch.float64)
        t["counters"].copy_(torch.tensor([B, 0, 0, 0, B, 0], dtype=torch.int32))

    # ---- one simulation of every tree ---------------------------------------------------------------------------
    def _step_and_expand(self, attach_only: bool):
        """blk_step on the requested transitions -> evaluator -> blk_puct_expand.  Every pointer and scalar passed
        to a kernel here is the same on every call (new states go through `stage`, their pool slots come from a
        device-side counter), so the sequence can be captured into a CUDA graph."""
        t, B, eng = self.t, self.B, self.eng
        # the parents' states are read straight out of the pool (blk_step_args.state_index): no gather pass
        out = eng.step(self.pool, t["step_action"], out_states=self.stage, state_index=t["src_slot"], buffers=self.buf,
                       mask=self.mask_fmt, want_count=False, want_scores=False, obs=None if attach_only else self.obs)
        prior, pd, ps, value = None, 0, 0, None
        if not attach_only and self.uniform:
            v = self.evaluator.values(eng, self.stage)       # uniform prior: expanded as P = 1/n on the device
            value = None if v is None else v.to(torch.float64).contiguous()
        elif not attach_only:
            if self.obs is not None:
                p, v = self.evaluator.evaluate(eng, self.stage, out.mask, obs=self.obs)
            else:
                p, v = self.evaluator.evaluate(eng, self.stage, out.mask)
            prior = p.contiguous()
            pd = 2 if prior.dtype == torch.float64 else 1
            if pd == 1:
                prior = prior.float()
            ps = prior.stride(0)
            value = v.to(torch.float64).contiguous()
        args = _lib.BlkPuctExpandArgs(-1, eng.state_words, self._meta, int(attach_only), self.stage.data_ptr(),
                                      self.pool.data_ptr(), out.mask_raw.data_ptr(), int(self.mask_fmt == "bits"),
                                      eng.mask_words, out.flags.data_ptr(),
                                      out.terminal.data_ptr(), None if prior is None else prior.data_ptr(), pd, ps,
                                      None if value is None else value.data_ptr(), 1)      # fused backup
        self._check(self._lib.blk_puct_expand(C.byref(self.forest), C.byref(args), self._stream()))
        self._keep = (prior, value)                      # keep graph-captured temporaries alive

    def _simulate_eager(self, cpuct: float, epsilon_fix: bool) -> None:
        self._check(self._lib.blk_puct_select(C.byref(self.forest), float(cpuct), int(epsilon_fix), self._stream()))
        self._step_and_expand(False)                     # the expansion launch also walks the paths back

    def simulate(self, cpuct: float = 1.0, epsilon_fix: bool = True) -> None:
        """One simulation of every tree.  After two eager runs the launch sequence is captured into a CUDA graph
        (one per (cpuct, epsilon_fix)) and replayed: small batches stop being launch-bound."""
        if self.fused:
            return self.run(1, cpuct, epsilon_fix)
        if self.used + self.B > self.node_cap:
            raise _lib.EngineError("GpuPuct state pool exhausted: raise max_simulations")
        key = (float(cpuct), bool(epsilon_fix))
        g = self._graphs.get(key)
        if g is not None:
            g.replay()
        elif self.use_cuda_graph and self._eager_runs >= 2:
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.eng.device)
            with torch.cuda.graph(g):
                self._simulate_eager(cpuct, epsilon_fix)
            self._graphs[key] = g
            g.replay()                                   # capture only records: run the simulation it stands for
        else:
            self._simulate_eager(cpuct, epsilon_fix)
            self._eager_runs += 1
        self.used += self.B
        self.launches += 3

    def run(self, simulations: int, cpuct: float = 1.0, epsilon_fix: bool = True, chain: int = 1) -> None:
        """``simulations`` simulations of every tree.  With ``chain = K > 1`` (and a graph-safe evaluator) K consecutive
        simulations are captured into ONE CUDA graph, so a small forest (the reference's single search,
        players/mcts_player.py:15-22, is B = 1) pays one graph launch per K simulations instead of 3 K kernel launches."""
        if self.fused:
            if simulations <= 0:
                return
            if self.used + self.B * simulations > self.node_cap:
                raise _lib.EngineError("GpuPuct state pool exhausted: raise max_simulations")
            args = self._search_args(simulations, cpuct, epsilon_fix)
            _lib.check(self._lib.blk_puct_search(self.eng._h, C.byref(self.forest), C.byref(args), self._stream()))
            self.used += self.B * simulations           # upper bound: a simulation creates at most one node per tree
            self.launches += 1
            return
        done = 0
        chain = max(1, min(int(chain), simulations))
        if chain > 1 and self.use_cuda_graph:
            while self._eager_runs < 2 and done < simulations:
                self.simulate(cpuct, epsilon_fix)
                done += 1
            key = (float(cpuct), bool(epsilon_fix), chain)
            while simulations - done >= chain:
                if self.used + chain * self.B > self.node_cap:
                    raise _lib.EngineError("GpuPuct state pool exhausted: raise max_simulations")
                g = self._graphs.get(key)
                if g is None:
                    g = torch.cuda.CUDAGraph()
                    torch.cuda.synchronize(self.eng.device)
                    with torch.cuda.graph(g):
                        for _ in range(chain):
                            self._simulate_eager(cpuct, epsilon_fix)
                    self._graphs[key] = g
                g.replay()
                self.used += chain * self.B
                self.launches += 3 * chain
                done += chain
        while done < simulations:
            self.simulate(cpuct, epsilon_fix)
            done += 1

    def advance(self, actions: torch.Tensor) -> None:
        """Make the child under ``actions[t]`` the root of tree t (``-1`` leaves a tree where it is)."""
        acts = actions.to(torch.int32).contiguous()
        if self.fused:
            # the env transition of the roots under the chosen actions, then the board-keyed lookup: the child the search
            # built (if it did) becomes the root with its statistics
            nxt = torch.empty((self.B, self.eng.state_words), dtype=torch.int32, device=self.eng.device)
            out = self.eng.step(self.root_states(), acts, out_states=nxt, mask=None, want_count=False, want_terminal=False,
                                want_scores=False)
            self._last_advance_flags = out.flags
            self.reroot(nxt)
            return
        if self.used + self.B > self.node_cap:
            raise _lib.EngineError("GpuPuct state pool exhausted: raise max_simulations")
        self._check(self._lib.blk_puct_advance(C.byref(self.forest), acts.data_ptr(), self._stream()))
        self._step_and_expand(True)                      # empty paths: the fused backup only takes the pool slots
        self.used += self.B

    # ---- results (host side; synchronises) ---------------------------------------------------------------------------
    def check(self) -> None:
        c = self.t["counters"].cpu().numpy()
        fl = getattr(self, "_last_advance_flags", None)
        if fl is not None and bool((fl & 2).any()):
            raise _lib.EngineError("GpuPuct.advance was given an illegal action")
        if c[2]:
            raise _lib.EngineError("GpuPuct capacity overflow (nodes / edges / depth): enlarge the forest")
        if c[3]:
            raise _lib.EngineError("GpuPuct produced an illegal action")

    def scores(self) -> np.ndarray:
        return self.t["scores"].cpu().numpy()

    def root_states(self) -> torch.Tensor:
        return self.pool.index_select(0, self.t["node_state"].index_select(0, self.t["root"].long()).long())

    def root_stats(self):
        """Per tree: (action ids, N, Q, P) of the root's edges as NumPy arrays."""
        self.check()
        root = self.t["root"].long()
        e0 = self.t["node_edge0"].index_select(0, root).cpu().numpy()
        n = self.t["node_nedge"].index_select(0, root).cpu().numpy()
        n = np.where(e0 < 0, 0, n)
        idx = np.concatenate([np.arange(a, a + b) for a, b in zip(e0, n)]) if n.sum() else np.zeros(0, np.int64)
        di = torch.as_tensor(idx, device=root.device, dtype=torch.long)
        cols = [self.t[k].index_select(0, di).cpu().numpy() for k in ("edge_action", "edge_n", "edge_q")]
        cols.append(np.zeros(len(idx)) if self.uniform else self.t["edge_p"].index_select(0, di).cpu().numpy())
        if self.fused and self.warps_per_tree > 1:      # leaf-parallel search keeps SUMS of backed-up values per edge
            cols[2] = np.where(cols[1] > 0, cols[2] / np.maximum(cols[1], 1.0), 0.0)
        parts = [np.split(c, np.cumsum(n)[:-1]) for c in cols]
        # nodes expanded with the uniform prior do not store P per edge: it is 1/nedge (the same float64 division)
        uni = self.t["node_uniform"].index_select(0, root).cpu().numpy()
        for t in range(self.B):
            if uni[t] and n[t]:
                parts[3][t] = np.full(n[t], np.float64(1.0) / np.float64(n[t]))
        return [tuple(col[t] for col in parts) for t in range(self.B)]

    def best_actions_device(self) -> torch.Tensor:
        """Most visited root action per tree, first maximum (players/mcts_player.py:19-20); -1 for finished games.
        Stays on the device and does not synchronise: feed it straight to :meth:`advance` / ``engine.step``."""
        out = torch.empty(self.B, dtype=torch.int32, device=self.eng.device)
        self._check(self._lib.blk_puct_best(C.byref(self.forest), out.data_ptr(), None, self._stream()))
        return out

    def best_actions(self) -> np.ndarray:
        return self.best_actions_device().cpu().numpy()
