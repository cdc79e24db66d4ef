"""TEST-ONLY backend: the adapter interface of blokus_rl_b200.backend.EngineBackend implemented on the CPU
oracle, so that the adapters (and the reference's own wrapper/MCTS/arena on top of them) can be exercised in
the build container, which has /root/reference but no GPU.  Never imported by the product package."""
import numpy as np

from oracle.oracle import Oracle


class _H:
    __slots__ = ("s", "_mask")

    def __init__(self, s):
        self.s, self._mask = s, None


class OracleBackend:
    def __init__(self, board_size=20, num_players=4, score_rule=0):
        self.orc = Oracle(board_size, num_players, score_rule)
        self.N, self.P, self.A = board_size, num_players, self.orc.A

    def new_state(self):
        return _H(self.orc.new_state())

    def from_words(self, words):
        return _H(self.orc.unpack(words))

    def next_state(self, h, action_id):
        s = self.orc.copy(h.s)
        if self.orc.step(s, int(action_id), fast=True) != 0:
            raise ValueError(f"illegal action {action_id}")
        return _H(s)

    def mover(self, h):
        return self.orc.field(h.s, "mover")

    def done(self, h):
        return bool(self.orc.field(h.s, "done"))

    def ply(self, h):
        return self.orc.field(h.s, "ply")

    def legal_mask(self, h):
        if h._mask is None:
            h._mask = self.orc.legal_mask(h.s, fast=True)
        return h._mask

    def legal_ids(self, h):
        return np.flatnonzero(self.legal_mask(h))

    def winners(self, h):
        w = self.orc.winners(h.s)
        return [p for p in range(self.P) if (w >> p) & 1]

    def terminal_values(self, h):
        return self.orc.terminal_values(h.s)

    def scores(self, h):
        return self.orc.final_scores(h.s)[: self.P]

    def observation(self, h):
        return self.orc.observe(h.s)

    def board_contents(self, h):
        return self.orc.board_contents(h.s)

    def board_key(self, h):
        return self.orc.pack(h.s)[: self.P * self.N].tobytes()

    def words(self, h):
        return self.orc.pack(h.s)

    def sample_move(self, h, rng=np.random):
        return int(rng.choice(self.legal_ids(h)))
