"""Shared helpers for the GPU-vs-oracle parity tests."""
from __future__ import annotations

import numpy as np


def unpack_bits(words: np.ndarray, num_actions: int) -> np.ndarray:
    """int32/uint32 [n, mask_words] -> uint8 [n, A] (bit b of word g = action 32g+b)."""
    w = np.ascontiguousarray(words).view(np.uint32)
    bits = np.unpackbits(w.view(np.uint8).reshape(w.shape[0], -1), axis=1, bitorder="little")
    return bits[:, :num_actions]


def lockstep(eng, orc, n: int, plies: int, seed: int, *, auto_reset: bool = True, env_id_base: int = 0,
             check_naive_every: int = 0, fmt_bits: bool = True):
    """Random play on the GPU (its own Philox sampler) mirrored move by move on the oracle.
    Every ply compares: sampled action, packed state words, flags, terminal vector, scores, legal count,
    byte mask and bit mask.  Returns (#steps, #games finished)."""
    import torch
    P, A = eng.num_players, eng.num_actions
    states = eng.new_states(n)
    out = eng.step(states, None, mask="bytes", sample=True, seed=seed, env_id_base=env_id_base)
    ost = [orc.new_state() for _ in range(n)]
    torch.cuda.synchronize()
    m0 = out.mask.cpu().numpy()
    nxt = out.next_action.cpu().numpy()
    for i in range(n):
        assert (orc.legal_mask(ost[i], fast=True) == m0[i]).all()
        assert orc.sample_action(ost[i], seed, env_id_base + i) == int(nxt[i])
    steps = games = 0
    alive = np.ones(n, bool)
    for ply in range(plies):
        actions = out.next_action.clone()
        acts = actions.cpu().numpy()
        out = eng.step(states, actions, mask="bytes", sample=True, seed=seed, env_id_base=env_id_base,
                       auto_reset=auto_reset)
        bits = eng.step(states, None, mask="bits").mask if fmt_bits else None
        torch.cuda.synchronize()
        words = states.cpu().numpy().view(np.uint32)
        masks = out.mask.cpu().numpy()
        cnt = out.legal_count.cpu().numpy()
        flags = out.flags.cpu().numpy()
        term = out.terminal.cpu().numpy()
        scores = out.scores.cpu().numpy()
        nxt = out.next_action.cpu().numpy()
        if fmt_bits:
            bmask = unpack_bits(bits.cpu().numpy(), A)
            assert (bmask == masks).all(), f"bit mask != byte mask at ply {ply}"
        for i in range(n):
            s = ost[i]
            if not alive[i]:     # finished without auto-reset: action -1 is a no-op on the done state
                assert flags[i] == 1 and masks[i].sum() == 0 and nxt[i] == -1
                assert (orc.pack(s) == words[i]).all()
                continue
            assert orc.step(s, int(acts[i])) == 0, f"env {i} ply {ply}: GPU action {acts[i]} illegal for oracle"
            steps += 1
            done = bool(orc.field(s, "done"))
            assert (flags[i] & 2) == 0
            assert bool(flags[i] & 1) == done, f"env {i} ply {ply}: done flag"
            assert (scores[i] == orc.final_scores(s)[:P]).all(), f"env {i} ply {ply}: scores"
            assert (term[i] == orc.terminal_values(s)).all(), f"env {i} ply {ply}: terminal vector"
            if done:
                games += 1
                if auto_reset:
                    orc.reset(s, orc.field(s, "game") + 1)
                else:
                    alive[i] = False
            assert (orc.pack(s) == words[i]).all(), f"env {i} ply {ply}: state words"
            om = orc.legal_mask(s, fast=True)
            assert (om == masks[i]).all(), f"env {i} ply {ply}: legal mask"
            assert cnt[i] == om.sum()
            if check_naive_every and (ply + i) % check_naive_every == 0:
                assert (orc.legal_mask(s, fast=False) == masks[i]).all()
            assert orc.sample_action(s, seed, env_id_base + i) == int(nxt[i]), f"env {i} ply {ply}: sampler"
    return steps, games
