"""Oracle parity AT the BASELINE.json configurations themselves (not just invariants at that size):

* configs[1]: 65,536 envs of 20x20 / 4 players x 64 plies of random-legal play with auto-reset, byte AND bit masks, seed 0x5EED
  -- every env's final state words, its whole action trajectory (running hash), every ply's legal counts and, at
  selected plies, a checksum of every env's full mask, against the C oracle run over the same global env ids on all host
  threads;
* configs[4]: the 1,048,576-env sweep at 7x7 / 2 players (and the 20x20 shard a rank of an 8-GPU run owns), checked on a
  65,536-env subsample / at a non-zero global env id base;
* configs[2]: 1,024 roots x 1,024 playouts: the first 16 playouts of every root replayed by the oracle.

The reference functions restated: blokus_rl/colossumrl/blokus_wrapper.py:89-132, 164-186, 233-246.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

K = 0x9E3779B97F4A7C15
K_I64 = K - (1 << 64)                      # the same multiplier as a (wrapping) int64


def gpu_random_play(eng, n, plies, seed, fmt, base=0, sel=(), keep=None):
    """Random play on the GPU (auto-reset, on-device sampler) with the checksums orc_play_many defines."""
    import torch
    dev = eng.device
    states = eng.new_states(n)
    buf = eng.make_buffers(n, fmt, sample=True)
    gid = torch.arange(n, device=dev, dtype=torch.int64) + base
    w = 2 * gid + 1
    traj = torch.zeros(n, dtype=torch.int64, device=dev)
    cnt_sum, ids_sum, words_sum, games, illegal = [], {}, {}, 0, 0
    A, MW = eng.num_actions, eng.mask_words
    wa = torch.arange(1, A + 1, device=dev, dtype=torch.int32)
    wg = 2 * torch.arange(MW, device=dev, dtype=torch.int64) + 1
    out = eng.step(states, None, buffers=buf, mask=fmt, sample=True, seed=seed, env_id_base=base)
    for t in range(plies + 1):
        traj = traj * K_I64 + (out.next_action.long() + 2)
        cnt_sum.append(int((out.legal_count.long() * w).sum().item()) & ((1 << 64) - 1))
        if t in sel:
            rows = slice(None) if keep is None else keep
            if fmt == "bytes":                                   # sum over legal ids of (id + 1), in chunks of envs
                m = out.mask[rows]
                parts = [(m[i:i + 2048].to(torch.int32) * wa).sum(1, dtype=torch.int64) for i in range(0, m.shape[0], 2048)]
                ids_sum[t] = torch.cat(parts).cpu().numpy().astype(np.uint64)
            else:                                                # sum over mask words of word * (2g + 1), wrapping
                m = out.mask_raw[rows]
                parts = [((m[i:i + 8192].long() & 0xFFFFFFFF) * wg).sum(1) for i in range(0, m.shape[0], 8192)]
                words_sum[t] = torch.cat(parts).cpu().numpy().view(np.uint64)
        if t == plies:
            break
        out = eng.step(states, buf.next_action, buffers=buf, mask=fmt, sample=True, seed=seed, env_id_base=base,
                       auto_reset=True)
        games += int((out.flags & 1).sum().item())
        illegal += int((out.flags & 2).sum().item())
    torch.cuda.synchronize()
    return {"states": states.cpu().numpy().view(np.uint32), "traj": traj.cpu().numpy().view(np.uint64),
            "cnt_sum": cnt_sum, "ids_sum": ids_sum, "words_sum": words_sum, "games": games, "illegal": illegal}


def compare(orc, g, r, ost, sel, fmt, rows=slice(None)):
    assert g["illegal"] == 0
    bad = np.flatnonzero(g["traj"][rows] != r["traj"])
    assert bad.size == 0, f"{bad.size} envs played a different action sequence, first: env {bad[:5]}"
    want = orc.pack_many(ost)
    assert (g["states"][rows] == want).all(), "final state words differ"
    for j, t in enumerate(sel):
        if fmt == "bytes":
            assert (g["ids_sum"][t] == r["ids_sum"][:, j]).all(), f"byte-mask checksum differs at ply {t}"
        else:
            assert (g["words_sum"][t] == r["words_sum"][:, j]).all(), f"bit-mask checksum differs at ply {t}"


@pytest.mark.parametrize("fmt", ["bytes", "bits"])
def test_baseline_config1_65536_envs_64_plies_vs_oracle(engine20, oracle20, fmt):
    n, plies, seed, sel = 65536, 64, 0x5EED, [0, 1, 17, 40, 63, 64]
    g = gpu_random_play(engine20, n, plies, seed, fmt, sel=sel)
    ost = oracle20.new_states(n)
    r = oracle20.play_many(ost, seed, plies, sel_plies=sel)
    assert g["cnt_sum"] == [int(x) for x in r["cnt_sum"]], "per-ply legal-count checksums differ"
    assert g["games"] == r["games"] and r["steps"] == n * plies and g["games"] > n // 2
    compare(oracle20, g, r, ost, sel, fmt)


def test_baseline_config4_rank_shard_of_the_1M_sweep_vs_oracle(engine20, oracle20):
    """What rank 5 of an 8-GPU run of the 1,048,576-env sweep computes: 131,072 envs with global ids from 655,360,
    128 plies; every 8th env is replayed by the oracle under the same global id."""
    n, plies, seed, base, stride, sel = 131072, 128, 0x5EED, 5 * 131072, 8, [0, 64, 128]
    keep = slice(0, n, stride)
    g = gpu_random_play(engine20, n, plies, seed, "bits", base=base, sel=sel, keep=keep)
    ost = oracle20.new_states(n // stride)
    r = oracle20.play_many(ost, seed, plies, env_id0=base, env_stride=stride, sel_plies=sel)
    compare(oracle20, g, r, ost, sel, "bits", rows=keep)


@pytest.mark.parametrize("fmt", ["bytes", "bits"])
def test_1M_envs_7x7_two_players_subsample_vs_oracle(engine7, oracle7, fmt):
    n, plies, seed, stride, sel = 1 << 20, 64, 0x5EED, 16, [0, 5, 33, 64]
    keep = slice(0, n, stride)
    g = gpu_random_play(engine7, n, plies, seed, fmt, sel=sel, keep=keep)
    ost = oracle7.new_states(n // stride)
    r = oracle7.play_many(ost, seed, plies, env_stride=stride, sel_plies=sel)
    assert g["games"] > 4 * n                                   # a 7x7 game lasts ~9 plies
    compare(oracle7, g, r, ost, sel, fmt, rows=keep)


def test_baseline_config2_1024_roots_x_1024_playouts_vs_oracle(engine20, oracle20):
    import torch
    eng, orc = engine20, oracle20
    roots = eng.new_states(1024)
    out = eng.step(roots, None, mask=None, sample=True, seed=24)
    for _ in range(24):                                          # SURVEY.md 8d workload 3: roots after 24 random plies
        out = eng.step(roots, out.next_action, mask=None, sample=True, seed=24)
    res = eng.rollout(roots, 1024, seed=7)
    torch.cuda.synchronize()
    fs, win, plies = res.final_scores.cpu().numpy(), res.winners.cpu().numpy(), res.plies.cpu().numpy()
    vsum = res.value_sum.cpu().numpy()
    hroots = orc.unpack_many(roots.cpu().numpy())
    check = 16                                                   # playouts per root replayed on the host
    import ctypes as C
    for r in range(1024):
        root = C.create_string_buffer(hroots[r].tobytes(), orc.state_size)
        for j in range(check):
            n, scores, w, _, _ = orc.playout(root, 7, r * 1024 + j)
            assert n == plies[r, j] and (scores == fs[r, j]).all() and w == win[r, j], (r, j)
    # all 1,048,576 games finished; the per-root value sums equal the sums over the winners masks
    assert (win != 0).all()
    bits = (win[:, :, None] >> np.arange(4)[None, None, :]) & 1
    nwin = bits.sum(2, keepdims=True)
    val = np.where(bits == 1, np.where(nwin == 1, 3.0, 1.0), -1.0).sum(1)
    assert np.array_equal(vsum, val.astype(np.float32))
