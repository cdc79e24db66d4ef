"""Rows a8/a9: BatchedMCTS vs the REFERENCE's MCTS (blokus_rl/alphazero/mcts.py), via golden vectors that
tests/golden/make_mcts_golden.py produced by running the unmodified reference file in the build container.

Tolerance (north_star: "UCB/visit-count selection must match within a stated floating-point tolerance"):
the tree arithmetic is float64 on the host in both implementations, so visit counts and argmax choices must be
IDENTICAL and Q / P / per-simulation score vectors agree to 1e-12 absolute.  (A float32 policy net enters
only through its priors; see test_torch_net_evaluator_matches_reference_valid_dist: rtol 1e-5, atol 1e-6.)
"""
import json
from collections import defaultdict
from pathlib import Path

import numpy as np
import pytest
import torch

from fake_nets import HashEvaluator

GOLDEN = json.loads((Path(__file__).parent / "golden" / "mcts_golden.json").read_text())


def _run_cases(make_engine):
    from blokus_rl_b200.mcts import BatchedMCTS, UniformEvaluator
    groups = defaultdict(list)
    for c in GOLDEN["cases"]:
        groups[(c["board_size"], c["players"], c["net"], c["cpuct"], c["sims"])].append(c)
    checked = 0
    for (N, P, net, cpuct, sims), cases in groups.items():
        eng = make_engine(N, P)
        search = BatchedMCTS(eng, UniformEvaluator() if net == "uniform" else HashEvaluator())
        words = torch.from_numpy(np.array([c["root_words"] for c in cases], dtype=np.uint32).view(np.int32)).to(eng.device)
        roots = search.add_roots(words)               # B roots searched in lockstep
        per_sim = [search.simulate(roots, cpuct) for _ in range(sims)]
        for t, c in enumerate(cases):
            assert roots[t].mover == c["root_player"]
            got = np.array([per_sim[k][t] for k in range(sims)])
            assert np.allclose(got, np.array(c["scores"]), rtol=0, atol=1e-12), (N, P, net, cpuct, c["plies"])
            ids, n, q, p = search.stats(t, roots[t])
            assert list(ids) == c["ids"]
            assert list(n) == c["N"], "visit counts differ from the reference"
            assert np.allclose(q, c["Q"], rtol=0, atol=1e-12) and np.allclose(p, c["P"], rtol=0, atol=1e-12)
            assert len(search.trees[t]) == c["tree_size"]
            _, d1 = search.get_distribution(t, roots[t], 1)
            _, d0 = search.get_distribution(t, roots[t], 0)
            assert np.allclose(d1, c["dist_T1"], atol=1e-12) and list(d0) == c["dist_T0"]
            checked += 1
    assert checked == len(GOLDEN["cases"])


def test_batched_mcts_matches_reference_on_cpu_stand_in():
    """Host logic (selection, backup, batching, key semantics) over the test-only oracle engine."""
    from oracle_engine import OracleEngine
    cache = {}
    _run_cases(lambda n, p: cache.setdefault((n, p), OracleEngine(n, p)))


@pytest.mark.gpu
def test_batched_mcts_matches_reference_on_gpu():
    from blokus_rl_b200 import BlokusEngine
    cache = {}
    _run_cases(lambda n, p: cache.setdefault((n, p), BlokusEngine(n, p)))


def test_distribution_edge_cases():
    from oracle_engine import OracleEngine
    from blokus_rl_b200.mcts import BatchedMCTS
    eng = OracleEngine(7, 2)
    search = BatchedMCTS(eng)
    roots = search.add_roots(eng.new_states(1))
    search.simulate(roots)                              # expansion only: all N are zero
    ids, d = search.get_distribution(0, roots[0], 1)
    assert np.allclose(d, 1.0 / len(ids))               # mcts.py:94-96: unexplored -> uniform
    ids, d0 = search.get_distribution(0, roots[0], 0)
    assert d0[0] == 1 and d0.sum() == 1                 # argmax of all-zero counts = first action


def test_torch_net_evaluator_matches_reference_valid_dist():
    """TorchNetEvaluator's masked softmax == neural_network.py:159-173 (masked_select + log_softmax + exp)."""
    from oracle_engine import OracleEngine
    from blokus_rl_b200.mcts import TorchNetEvaluator

    class Net(torch.nn.Module):
        def __init__(self, a, p):
            super().__init__()
            self.a, self.p = a, p

        def forward(self, x):
            g = torch.Generator().manual_seed(int(x.sum().item()))
            return torch.randn((x.shape[0], self.a), generator=g), torch.randn((x.shape[0], self.p), generator=g)

    eng = OracleEngine(7, 2)
    s = eng.new_states(2)
    eng.step(s, torch.tensor([0, -1], dtype=torch.int32))
    mask = eng.step(s, None).mask
    net = Net(eng.num_actions, 2)
    p, v = TorchNetEvaluator(net).evaluate(eng, s, mask)
    logits, _ = net(eng.observe(s))
    for i in range(2):
        ref = torch.exp(torch.log_softmax(torch.masked_select(logits[i], mask[i]), dim=-1))
        assert torch.allclose(p[i][mask[i]].float(), ref, rtol=1e-5, atol=1e-6)
        assert p[i][~mask[i]].sum() == 0
