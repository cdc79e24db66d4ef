"""Deterministic stand-ins for the policy/value net, shared by the golden-vector generator (which drives the
REFERENCE's MCTS through ``nn.predict(obs, mask)``) and the tests (which drive BatchedMCTS through an
evaluator).  Both derive everything from the board cells, so the two searches see identical priors/values."""
import zlib

import numpy as np
import torch


def _hash_outputs(key: bytes, n_valid: int, players: int):
    rng = np.random.default_rng(zlib.crc32(key))
    p = rng.random(n_valid) + 0.05
    return p / p.sum(), rng.uniform(-1.0, 1.0, players)


def key_from_obs(obs: np.ndarray, players: int) -> bytes:
    n = obs.shape[-1]
    cols = (1 << np.arange(n, dtype=np.uint64))
    rows = (obs[:players].astype(np.uint64) * cols[None, None, :]).sum(-1).astype(np.uint32)
    return rows.tobytes()


class UniformNet:
    """DumbNet (blokus_rl/models/dumbnet.py:14-21) seen through predict (neural_network.py:92-110)."""
    def __init__(self, players):
        self.players = players

    def predict(self, obs, mask):
        n = int(mask.sum())
        return np.full(n, 1.0 / n), np.zeros(self.players)


class HashNet:
    def __init__(self, players):
        self.players = players

    def predict(self, obs, mask):
        return _hash_outputs(key_from_obs(obs, self.players), int(mask.sum()), self.players)


class HashEvaluator:
    """Same function as HashNet, in BatchedMCTS's evaluator shape."""
    def evaluate(self, engine, states, mask):
        P, nrow = engine.num_players, engine.num_players * engine.board_size
        w = states.cpu().numpy().view(np.uint32)
        m = mask.cpu().numpy()
        p = np.zeros(m.shape, np.float64)
        v = np.zeros((m.shape[0], P), np.float64)
        for i in range(m.shape[0]):
            ids = np.flatnonzero(m[i])
            p[i, ids], v[i] = _hash_outputs(w[i, :nrow].tobytes(), len(ids), P)
        return torch.from_numpy(p).to(states.device), torch.from_numpy(v).to(states.device)
