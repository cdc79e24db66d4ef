"""Boards up to 7x7 run on the thread-per-env kernels (csrc/blk_small.cu).  They are checked against the oracle move by
move (every (N, P) they are instantiated for, with and without auto-reset) and, at full batch size, against the
warp-per-env kernels, which stay selectable for these boards (BLK_OPT_WARP_KERNELS)."""
import numpy as np
import pytest

from helpers import lockstep, unpack_bits

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,p", [(5, 2), (5, 4), (6, 2), (6, 4), (7, 2), (7, 4)])
def test_lockstep_small_boards(n, p):
    from blokus_rl_b200 import BlokusEngine
    from oracle.oracle import Oracle
    eng, orc = BlokusEngine(n, p), Oracle(n, p)
    steps, games = lockstep(eng, orc, n=40, plies=45, seed=100 * n + p, check_naive_every=7)
    assert games > 40
    steps, games = lockstep(eng, orc, n=33, plies=40, seed=7 * n + p, auto_reset=False, env_id_base=5000)
    assert games == 33                           # every game ran into its terminal state and stayed there
    eng.close()


@pytest.mark.parametrize("count", [1, 31, 255, 256, 257, 1000])
def test_ragged_batches_small(count, engine7, oracle7):
    """Batch sizes around the 256-env block of the small kernel."""
    import torch
    s = engine7.new_states(count)
    out = engine7.step(s, None, mask="bytes", sample=True, seed=5)
    for _ in range(6):
        out = engine7.step(s, out.next_action, mask="bytes", sample=True, seed=5, auto_reset=True)
    torch.cuda.synchronize()
    words = s.cpu().numpy().view(np.uint32)
    masks = out.mask.cpu().numpy()
    for i in sorted({0, count // 2, count - 1}):
        o = oracle7.unpack(words[i])
        assert (oracle7.legal_mask(o, fast=True) == masks[i]).all()
        assert oracle7.sample_action(o, 5, i) == int(out.next_action[i])


def test_illegal_actions_small(engine7, oracle7):
    import torch
    eng, orc = engine7, oracle7
    s = eng.new_states(6)
    before = s.clone()
    first = int(np.flatnonzero(orc.legal_mask(orc.new_state()))[0])
    acts = torch.tensor([first, eng.num_actions, -5, 1000, -1, first + 1], dtype=torch.int32, device="cuda")
    legal_second = bool(orc.legal_mask(orc.new_state())[first + 1])
    out = eng.step(s, acts, mask="bits", want_count=True)
    torch.cuda.synchronize()
    flags = out.flags.cpu().numpy()
    assert flags[0] == 0 and flags[1] == 2 and flags[2] == 2 and flags[4] == 0
    assert (flags[3] == 2) == (not bool(orc.legal_mask(orc.new_state())[1000]))
    assert (flags[5] == 2) == (not legal_second)
    for i in (1, 2, 4):                           # rejected / skipped envs keep their state and get the mover's mask
        assert (s[i] == before[i]).all()
        assert int(out.legal_count[i]) == 58
    assert not (s[0] == before[0]).all()


def test_small_kernels_equal_warp_kernels_at_full_batch(engine7):
    """65,536 envs of random play: the thread-per-env kernels and the warp-per-env kernels (BLK_OPT_WARP_KERNELS) agree
    bit for bit on states, masks, counts, sampled actions, flags, terminal vectors, scores and observation planes."""
    import torch
    eng = engine7
    n, seed = 65536, 0xABCDEF
    a = eng.new_states(n)
    b = a.clone()
    oa = eng.step(a, None, mask="bits", sample=True, seed=seed)
    ob = eng.step(b, None, mask="bits", sample=True, seed=seed, warp_kernels=True)
    checked = 0
    for ply in range(30):
        fmt = "bytes" if ply % 2 else "bits"
        act = oa.next_action.clone()
        assert (oa.next_action == ob.next_action).all()
        oa = eng.step(a, act, mask=fmt, sample=True, seed=seed, auto_reset=True, obs=ply % 5 == 0)
        ob = eng.step(b, act, mask=fmt, sample=True, seed=seed, auto_reset=True, obs=ply % 5 == 0, warp_kernels=True)
        if ply % 5 == 0:
            assert (oa.obs == ob.obs).all() and (oa.obs == eng.observe(a)).all(), f"observations differ at ply {ply}"
        assert (a == b).all(), f"states differ at ply {ply}"
        assert (oa.mask == ob.mask).all(), f"masks differ at ply {ply}"
        for name in ("legal_count", "flags", "terminal", "scores", "next_action"):
            assert (getattr(oa, name) == getattr(ob, name)).all(), f"{name} differs at ply {ply}"
        checked += int((oa.flags & 1).sum())
    assert checked > n                            # every env finished at least one game on the way
    # size-independent invariants of the final states: cells on the board == squares scored, inventories consistent
    w = a.cpu().numpy().view(np.uint32)
    N, P = eng.board_size, eng.num_players
    rows = w[:, : P * N].reshape(n, P, N)
    cells = np.unpackbits(rows.view(np.uint8), axis=-1).reshape(n, P, -1).sum(-1)
    scores = w[:, P * N + P + 2: P * N + P + 4].copy().view(np.int16)[:, :P]
    assert (cells == scores).all()
    assert not (rows[:, 0] & rows[:, 1]).any()     # colours never overlap


def test_mask_only_and_finished_states_small(engine7, oracle7):
    import torch
    eng, orc = engine7, oracle7
    n = 300
    s = eng.new_states(n)
    out = eng.step(s, None, mask=None, sample=True, seed=2)
    for _ in range(25):                            # far beyond the longest 7x7 game: every env is finished
        out = eng.step(s, out.next_action, mask=None, sample=True, seed=2)
    torch.cuda.synchronize()
    assert bool((out.flags & 1).all()) and bool((out.next_action == -1).all())
    keep = s.clone()
    o2 = eng.step(s, None, mask="bytes", sample=True, seed=2)
    assert (s == keep).all() and int(o2.mask.sum()) == 0 and bool((o2.flags == 1).all())
    w = s.cpu().numpy().view(np.uint32)
    for i in (0, 150, 299):
        o = orc.unpack(w[i])
        assert (o2.terminal[i].cpu().numpy() == orc.terminal_values(o)).all()
        assert (o2.scores[i].cpu().numpy() == orc.final_scores(o)[:2]).all()
    # an action on a finished game is rejected and changes nothing
    o3 = eng.step(s, torch.zeros(n, dtype=torch.int32, device="cuda"), mask="bits")
    assert (s == keep).all() and bool((o3.flags == 3).all())
    bits = unpack_bits(o3.mask.cpu().numpy(), eng.num_actions)
    assert bits.sum() == 0


@pytest.mark.parametrize("n,p", [(7, 2), (7, 4), (5, 2), (6, 4)])
def test_small_playouts_equal_warp_playouts_and_oracle(n, p):
    """Thread-per-playout kernel vs the warp-per-playout kernel (BLK_OPT_WARP_KERNELS): identical action logs, plies, final
    scores, winners and value sums; a sample of the logs is also replayed through the oracle (same Philox stream)."""
    import torch
    from blokus_rl_b200 import BlokusEngine
    from oracle.oracle import Oracle
    eng, orc = BlokusEngine(n, p), Oracle(n, p)
    n_roots, per_root, seed = 96, 40, 0xFACE + n
    roots = eng.new_states(n_roots)
    out = eng.step(roots, None, mask=None, sample=True, seed=4)
    for _ in range(2):
        out = eng.step(roots, out.next_action, mask=None, sample=True, seed=4)
    a = eng.rollout(roots, per_root, seed=seed, rollout_id_base=77, log_actions=True)
    b = eng.rollout(roots, per_root, seed=seed, rollout_id_base=77, log_actions=True, warp_kernels=True)
    torch.cuda.synchronize()
    assert (a.plies == b.plies).all() and (a.final_scores == b.final_scores).all() and (a.winners == b.winners).all()
    assert torch.allclose(a.value_sum, b.value_sum)
    la, lb = a.action_log.cpu().numpy().view(np.uint16), b.action_log.cpu().numpy().view(np.uint16)
    pl = a.plies.cpu().numpy()
    for r in range(n_roots):
        for j in range(per_root):
            assert (la[r, j, : pl[r, j] + 1] == lb[r, j, : pl[r, j] + 1]).all()
    words = roots.cpu().numpy().view(np.uint32)
    fs, win = a.final_scores.cpu().numpy(), a.winners.cpu().numpy()
    for r in range(0, n_roots, 7):
        for j in (0, per_root - 1):
            o = orc.unpack(words[r])
            k = 0
            while not orc.field(o, "done"):
                act = int(la[r, j, k])
                assert act == orc.sample_action(o, seed, 77 + r * per_root + j, stream=1)
                assert orc.step(o, act, fast=True) == 0
                k += 1
            assert la[r, j, k] == 0xFFFF and pl[r, j] == k
            assert (fs[r, j] == orc.final_scores(o)[:p]).all() and win[r, j] == orc.winners(o)
    # stop_player: hand the state back at the agent's turn (the gym adapter's opponent moves), in place
    sa, sb = roots.clone(), roots.clone()
    eng.rollout(sa, 1, seed=9, stop_player=0, out_states=sa)
    eng.rollout(sb, 1, seed=9, stop_player=0, out_states=sb, warp_kernels=True)
    assert (sa == sb).all()
    eng.close()


def test_index_lists_small(engine7, oracle7):
    """BLK_MASK_INDICES on the thread-per-env kernel: ascending legal ids, counts, truncation flag; equal to the warp kernel."""
    import torch
    eng = engine7
    n = 777
    s = eng.new_states(n)
    out = eng.step(s, None, mask="bytes", sample=True, seed=6)
    for _ in range(5):
        out = eng.step(s, out.next_action, mask="bytes", sample=True, seed=6, auto_reset=True)
    a = eng.step(s, None, mask="indices")
    b = eng.step(s, None, mask="indices", warp_kernels=True)
    torch.cuda.synchronize()
    ids, cnt = a.mask.cpu().numpy().view(np.uint16), a.legal_count.cpu().numpy()
    idb = b.mask.cpu().numpy().view(np.uint16)
    dense = out.mask.cpu().numpy()
    assert (a.legal_count == b.legal_count).all() and (a.flags == b.flags).all()
    for i in range(n):
        assert ids[i, : cnt[i]].tolist() == np.flatnonzero(dense[i]).tolist() == idb[i, : cnt[i]].tolist()
    fresh = eng.new_states(3)
    small = torch.zeros((3, 10), dtype=torch.int16, device=s.device)
    o2 = eng.step(fresh, None, mask=small)
    first = np.flatnonzero(oracle7.legal_mask(oracle7.new_state()))
    assert (o2.flags.cpu().numpy() & 4).all() and (o2.legal_count.cpu().numpy() == 58).all()
    assert small[2].cpu().numpy().view(np.uint16).tolist() == first[:10].tolist()
