"""TEST-ONLY stand-in for blokus_rl_b200.engine.BlokusEngine: the same tensor-level API computed by the CPU
oracle on CPU tensors.  It exists so that the HOST logic layered on the engine (batched MCTS, players, vector
env, sharding) can be tested in the build container, which has no GPU.  Never imported by the product."""
from __future__ import annotations

import numpy as np
import torch

from blokus_rl_b200.engine import RolloutOut, StepOut
from oracle.oracle import Oracle


class OracleEngine:
    def __init__(self, board_size=20, num_players=4, score_rule=0):
        self.orc = Oracle(board_size, num_players, score_rule)
        self.board_size, self.num_players = board_size, num_players
        self.num_actions, self.state_words = self.orc.A, self.orc.state_words
        self.mask_bytes = (self.num_actions + 127) // 128 * 128
        self.mask_words = ((self.num_actions + 31) // 32 + 3) // 4 * 4
        self.device = torch.device("cpu")
        self.max_legal = 1024

    def _unpack(self, states):
        w = states.numpy().view(np.uint32)
        return [self.orc.unpack(w[i]) for i in range(w.shape[0])]

    def _pack(self, sts, out):
        o = out.numpy().view(np.uint32)
        for i, s in enumerate(sts):
            o[i] = self.orc.pack(s)

    def new_states(self, n):
        t = torch.zeros((n, self.state_words), dtype=torch.int32)
        self._pack([self.orc.new_state() for _ in range(n)], t)
        return t

    def reset(self, states):
        self._pack([self.orc.new_state() for _ in range(states.shape[0])], states)
        return states

    def make_buffers(self, n, fmt="bytes", sample=False):
        return None

    def step(self, states, actions=None, *, out_states=None, mask="bytes", want_count=True, want_terminal=True,
             want_scores=True, sample=False, seed=0, env_id_base=0, auto_reset=False, buffers=None, obs=None, warp_kernels=False, state_index=None):
        if state_index is not None:
            states = states.index_select(0, state_index.long())
        orc, n, P = self.orc, states.shape[0], self.num_players
        sts = self._unpack(states)
        flags = np.zeros(n, np.uint8)
        term = np.zeros((n, P), np.float32)
        scores = np.zeros((n, P), np.int16)
        acts = None if actions is None else actions.numpy()
        for i, s in enumerate(sts):
            ended = was_done = bool(orc.field(s, "done"))
            if acts is not None and acts[i] != -1:
                if ended or orc.step(s, int(acts[i]), fast=True) != 0:
                    flags[i] |= 2
                ended = bool(orc.field(s, "done"))
            scores[i] = orc.final_scores(s)[:P]
            if ended:
                flags[i] |= 1
                term[i] = orc.terminal_values(s)
                if auto_reset and not was_done:
                    orc.reset(s, orc.field(s, "game") + 1)
        out_states = states if out_states is None else out_states
        self._pack(sts, out_states)
        m = np.stack([orc.legal_mask(s, fast=True) for s in sts]) if n else np.zeros((0, self.num_actions), np.uint8)
        nxt = None
        if sample:
            nxt = torch.tensor([orc.sample_action(s, seed, env_id_base + i) for i, s in enumerate(sts)], dtype=torch.int32)
        mt = None
        is_t = isinstance(mask, torch.Tensor)
        if mask == "indices" or (is_t and mask.dtype == torch.int16):
            width = mask.shape[1] if is_t else self.max_legal
            ids = np.zeros((n, width), np.int16)
            for i in range(n):
                nz = np.flatnonzero(m[i])
                if len(nz) > width:
                    flags[i] |= 4
                nz = nz[:width]
                ids[i, : len(nz)] = nz.astype(np.uint16).view(np.int16)
            mt = torch.from_numpy(ids)
            if is_t:                                   # like the engine: the caller's buffer is written in place
                mask.copy_(mt)
                mt = mask
        elif mask == "bits" or (is_t and mask.dtype == torch.int32):
            bits = np.zeros((n, self.mask_words * 32), np.uint8)
            bits[:, : self.num_actions] = m
            mt = torch.from_numpy(np.packbits(bits, axis=1, bitorder="little").view(np.int32).copy())
            if is_t:
                mask[:, : self.mask_words] = mt
                mt = mask
        elif mask == "bytes" or is_t:
            mt = torch.from_numpy(m.astype(bool))
            if is_t:
                mask[:, : self.num_actions] = torch.from_numpy(m).to(mask.dtype)
        ot = None
        if obs is not None and obs is not False:
            ot = torch.from_numpy(np.stack([orc.observe(s) for s in sts]))
            if isinstance(obs, torch.Tensor):
                obs.copy_(ot.view_as(obs))
                ot = obs
        res = {"legal_count": torch.from_numpy(m.sum(1).astype(np.int32)), "terminal": torch.from_numpy(term),
               "flags": torch.from_numpy(flags), "scores": torch.from_numpy(scores), "next_action": nxt}
        for name, val in res.items():                  # like the engine: caller-owned buffers are written in place
            dst = getattr(buffers, name, None) if buffers is not None else None
            if dst is not None and val is not None:
                dst[:n].copy_(val)
                res[name] = dst
        return StepOut(out_states, mt, res["legal_count"], res["terminal"], res["flags"], res["scores"], res["next_action"],
                       None, ot)

    def legal_mask(self, states, fmt="bytes", **kw):
        return self.step(states, None, mask=fmt, **kw)

    def observe(self, states, out=None):
        o = np.stack([self.orc.observe(s) for s in self._unpack(states)])
        return torch.from_numpy(o)

    def board_contents(self, states):
        return torch.from_numpy(np.stack([self.orc.board_contents(s) for s in self._unpack(states)]))

    def game_ended(self, states):
        sts = self._unpack(states)
        flags = torch.tensor([int(self.orc.field(s, "done")) for s in sts], dtype=torch.uint8)
        term = torch.from_numpy(np.stack([self.orc.terminal_values(s) for s in sts]))
        scores = torch.from_numpy(np.stack([self.orc.final_scores(s)[: self.num_players] for s in sts]))
        return flags, term, scores

    def rollout(self, roots, per_root, seed=0, rollout_id_base=0, log_actions=False, stop_player=-1, out_states=None,
                warp_kernels=False):
        orc, P = self.orc, self.num_players
        sts = self._unpack(roots)
        n = len(sts)
        fs = np.zeros((n, per_root, P), np.int16)
        win = np.zeros((n, per_root), np.uint8)
        vs = np.zeros((n, P), np.float32)
        plies = np.zeros((n, per_root), np.int32)
        finals = []
        for r, root in enumerate(sts):
            for j in range(per_root):
                s = orc.copy(root)
                gid = rollout_id_base + r * per_root + j
                k = 0
                while not orc.field(s, "done") and orc.field(s, "mover") != stop_player:
                    orc.step(s, orc.sample_action(s, seed, gid, stream=1), fast=True)
                    k += 1
                fs[r, j] = orc.final_scores(s)[:P]
                if orc.field(s, "done"):
                    win[r, j] = orc.winners(s)
                    vs[r] += orc.terminal_values(s)
                plies[r, j] = k
                finals.append(s)
        if out_states is not None:
            self._pack(finals, out_states.view(-1, self.state_words))
        return RolloutOut(torch.from_numpy(fs), torch.from_numpy(win), torch.from_numpy(vs), torch.from_numpy(plies), None)
