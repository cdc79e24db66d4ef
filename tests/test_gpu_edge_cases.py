"""GPU edge cases: empty and ragged batches, large ids/seeds, terminal roots, sampler uniformity."""
import numpy as np
import pytest
import torch

from helpers import unpack_bits

pytestmark = pytest.mark.gpu


def test_empty_and_ragged_batches(engine20, oracle20):
    eng = engine20
    s0 = eng.new_states(0)
    out = eng.step(s0, None, mask="bytes", sample=True)
    assert out.mask.shape == (0, eng.num_actions) and out.next_action.shape == (0,)
    assert eng.observe(s0).shape == (0, 8, 20, 20) and eng.rollout(s0, 4).final_scores.shape == (0, 4, 4)
    ref = oracle20.legal_mask(oracle20.new_state())
    for n in (1, 7, 9, 257, 3553):              # not multiples of the 8 envs a block holds / of the resident grid
        s = eng.new_states(n)
        out = eng.step(s, None, mask="bytes")
        torch.cuda.synchronize()
        assert (out.legal_count == 58).all()
        assert (out.mask[0].cpu().numpy() == ref).all() and (out.mask[n - 1].cpu().numpy() == ref).all()
        assert (out.mask.sum(1) == 58).all()


def test_large_env_ids_and_64bit_seeds(engine20, oracle20):
    eng, orc = engine20, oracle20
    seed, base = 0xFEDCBA9876543210, 0xFFFFFFF0       # env ids wrap modulo 2^32 like the oracle's uint32
    s = eng.new_states(24)
    out = eng.step(s, None, mask=None, sample=True, seed=seed, env_id_base=base)
    ost = [orc.new_state() for _ in range(24)]
    for ply in range(6):
        acts = out.next_action.cpu().numpy()
        for i in range(24):
            assert acts[i] == orc.sample_action(ost[i], seed, (base + i) & 0xFFFFFFFF)
            orc.step(ost[i], int(acts[i]))
        out = eng.step(s, out.next_action, mask=None, sample=True, seed=seed, env_id_base=base)


def test_sampler_is_uniform_over_legal_actions(engine20):
    """k = mulhi(u32, n): every one of the 58 first moves is drawn with probability ~1/58 (chi-square)."""
    eng = engine20
    n = 58 * 2000
    s = eng.new_states(n)
    out = eng.step(s, None, mask="bits", sample=True, seed=12345)
    acts = out.next_action.cpu().numpy()
    legal = np.flatnonzero(unpack_bits(out.mask[:1].cpu().numpy(), eng.num_actions)[0])
    counts = np.array([(acts == a).sum() for a in legal])
    assert counts.sum() == n
    chi2 = ((counts - 2000.0) ** 2 / 2000.0).sum()
    assert chi2 < 110.0                          # 57 dof: P(chi2 > 110) ~ 4e-5


def test_rollouts_from_terminal_and_late_roots(engine20, oracle20):
    eng, orc = engine20, oracle20
    roots = []
    for plies in (200, 60, 55, 50):              # 200: certainly finished
        o = orc.new_state()
        orc.random_play(o, 9, plies, plies, auto_reset=False, log=False)
        roots.append(o)
    words = torch.tensor(np.stack([orc.pack(o) for o in roots]).view(np.int32)).cuda()
    out = eng.rollout(words, 5, seed=3, log_actions=True)
    torch.cuda.synchronize()
    assert orc.field(roots[0], "done")
    assert (out.plies[0] == 0).all()
    assert (out.final_scores[0].cpu().numpy() == orc.final_scores(roots[0])).all()
    assert (out.winners[0].cpu().numpy() == orc.winners(roots[0])).all()
    assert (out.action_log[0, :, 0].cpu().numpy().view(np.uint16) == 0xFFFF).all()
    fs = out.final_scores.cpu().numpy()
    for r in range(1, 4):
        base = orc.field(roots[r], "score")
        assert (fs[r] >= base[None, :]).all()    # scores only grow during a playout
    assert eng.rollout(words, 0).final_scores.shape == (4, 0, 4)


def test_many_launches_reuse_the_work_queue(engine7):
    """The self-resetting ticket counters survive hundreds of back-to-back launches (64-slot ring wraps)."""
    eng = engine7
    s = eng.new_states(500)
    out = eng.step(s, None, mask="bits", sample=True, seed=1)
    total = 0
    for _ in range(300):
        out = eng.step(s, out.next_action, mask="bits", sample=True, seed=1, auto_reset=True)
        total += 1
    torch.cuda.synchronize()
    assert int((out.flags & 2).sum()) == 0 and (out.legal_count > 0).all()
    w = s.cpu().numpy().view(np.uint32)
    assert ((w[:, 16] >> 4) & 1).sum() == 0      # auto-reset: nobody is parked in a finished game


def test_mask_only_on_finished_states(engine20, oracle20):
    eng, orc = engine20, oracle20
    o = orc.new_state()
    orc.random_play(o, 4, 0, 200, auto_reset=False, log=False)
    words = torch.tensor(orc.pack(o).view(np.int32)[None]).cuda()
    out = eng.step(words, None, mask="bytes", sample=True)
    assert int(out.flags[0]) == 1 and int(out.legal_count[0]) == 0 and int(out.next_action[0]) == -1
    assert not out.mask.any() and (out.terminal[0].cpu().numpy() == orc.terminal_values(o)).all()
    out = eng.step(words, torch.zeros(1, dtype=torch.int32, device="cuda"), mask="bytes", auto_reset=True)
    assert int(out.flags[0]) == 3                # stepping a finished game: illegal + done, state unchanged
    assert (words.cpu().numpy().view(np.uint32)[0] == orc.pack(o)).all()


def test_graph_replay_overlapping_eager_launches_on_another_stream(engine20):
    """A captured step replayed on stream A while eager steps (and playouts) of the SAME engine run on stream B: every
    launch draws its envs from its own work-queue slot (one per stream, one per graph capture), so nothing is skipped or
    stepped twice.  250 + 250 overlapping launches must leave exactly the states a serial run leaves."""
    eng, n, K, seed = engine20, 16384, 250, 77
    dev = eng.device

    def fresh(sd):
        st = eng.new_states(n)
        buf = eng.make_buffers(n, "bits", sample=True)
        eng.step(st, None, buffers=buf, mask="bits", sample=True, seed=sd)
        return st, buf

    def one(st, buf, sd):
        return eng.step(st, buf.next_action, buffers=buf, mask="bits", sample=True, seed=sd, auto_reset=True)

    # serial reference on the default stream
    (ra, ba), (rb, bb) = fresh(seed), fresh(seed + 1)
    for _ in range(K):
        one(ra, ba, seed)
        one(rb, bb, seed + 1)
    torch.cuda.synchronize()
    # the same two runs, overlapping
    (sa, fa), (sb, fb) = fresh(seed), fresh(seed + 1)
    torch.cuda.synchronize()
    st_a, st_b = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st_a):
        one(sa, fa, seed)                               # capture only records
    roots = eng.new_states(64)
    for k in range(K):
        with torch.cuda.stream(st_a):
            g.replay()
        with torch.cuda.stream(st_b):
            one(sb, fb, seed + 1)
            if k % 50 == 0:
                eng.rollout(roots, 8, seed=k)           # the playout kernel shares the engine's queue slots too
    torch.cuda.synchronize()
    assert torch.equal(sa, ra) and torch.equal(sb, rb)
    assert torch.equal(fa.next_action, ba.next_action) and torch.equal(fb.mask_raw, bb.mask_raw)


def test_reused_buffers_are_validated(engine20):
    eng = engine20
    small = eng.make_buffers(8, "bytes", sample=True)
    s = eng.new_states(16)
    with pytest.raises(ValueError):
        eng.step(s, None, buffers=small, mask=None)                     # per-env outputs made for 8 envs, 16 stepped
    roots = eng.new_states(4)
    with pytest.raises(ValueError):
        eng.rollout(roots, 2, out_states=roots)                         # aliasing is only defined for per_root == 1
    eng.rollout(roots, 1, stop_player=1, out_states=roots)
    cur = torch.cuda.current_device()
    eng.step(s, None, mask="bits")
    assert torch.cuda.current_device() == cur                           # entry points leave the caller's device alone
