"""GPU parity tests: the CUDA engine (through the C ABI) vs the CPU oracle, bit-exact.

PARITY UNPINNED caveat: the oracle restates the reference's contract; the
reference's own env engine (colosseumrl) is absent and cannot be run.
"""
import numpy as np
import pytest

from helpers import lockstep, unpack_bits

pytestmark = pytest.mark.gpu


def test_action_table_matches_oracle(engine20, oracle20, engine7, oracle7):
    for eng, orc in ((engine20, oracle20), (engine7, oracle7)):
        assert eng.num_actions == orc.A
        step = 1 if orc.A < 5000 else 7
        for a in list(range(0, orc.A, step)) + [orc.A - 1]:
            meta, cells = eng.action_to_cells(a)
            ocells, ometa = orc.action_cells(a)
            assert sorted(cells) == sorted(ocells)
            assert meta[0] == ometa[0] and meta[2] == ometa[2] and meta[3] == ometa[3]


def test_info_and_known_sizes(engine20, engine7):
    assert engine20.num_actions == 30433          # blokus_rl/models/blokus_nnet.py:17
    assert engine20.state_words == 88 and engine20.mask_words == 952 and engine20.mask_bytes == 30464
    assert engine20.info.num_fields == 1665 and engine20.info.num_orients == 91
    assert engine7.num_actions == 2522


def test_reset_and_first_masks(engine20, oracle20):
    import torch
    s = engine20.new_states(5)
    out = engine20.step(s, None, mask="bytes")
    torch.cuda.synchronize()
    m = out.mask.cpu().numpy()
    assert (m.sum(1) == 58).all() and (out.legal_count.cpu().numpy() == 58).all()
    assert (m[0] == oracle20.legal_mask(oracle20.new_state())).all()
    assert (s.cpu().numpy().view(np.uint32)[0] == oracle20.pack(oracle20.new_state())).all()


def test_lockstep_20x20_autoreset(engine20, oracle20):
    steps, games = lockstep(engine20, oracle20, n=96, plies=150, seed=0x5EED, check_naive_every=97)
    assert steps == 96 * 150 and games >= 96


def test_lockstep_20x20_no_reset_runs_to_terminal(engine20, oracle20):
    steps, games = lockstep(engine20, oracle20, n=40, plies=90, seed=7, auto_reset=False, env_id_base=1000)
    assert games == 40


def test_lockstep_7x7_2p(engine7, oracle7):
    steps, games = lockstep(engine7, oracle7, n=64, plies=60, seed=3, check_naive_every=5)
    assert games > 64


@pytest.mark.parametrize("n,p", [(14, 2), (14, 4), (7, 4), (5, 2), (20, 2), (12, 4), (13, 2), (15, 2), (10, 2)])   # 12-15: five-field gather descriptors; 10: the generic gather
def test_lockstep_other_geometries(n, p):
    from blokus_rl_b200 import BlokusEngine
    from oracle.oracle import Oracle
    eng, orc = BlokusEngine(n, p), Oracle(n, p)
    lockstep(eng, orc, n=24, plies=70, seed=n * 10 + p, check_naive_every=11)
    eng.close()


def test_bonus_score_rule():
    from blokus_rl_b200 import BlokusEngine
    from oracle.oracle import Oracle
    eng, orc = BlokusEngine(14, 2, score_rule=1), Oracle(14, 2, score_rule=1)
    lockstep(eng, orc, n=48, plies=80, seed=99)
    eng.close()


def test_illegal_and_out_of_range_actions(engine20, oracle20):
    import torch
    s = engine20.new_states(4)
    before = s.clone()
    # 0: monomino at (0,0) is legal for player 0; others: far from the corner, out of range, negative
    acts = torch.tensor([0, 5000, 30433, -7], dtype=torch.int32, device=s.device)
    out = engine20.step(s, acts, mask="bytes")
    torch.cuda.synchronize()
    flags = out.flags.cpu().numpy()
    assert list(flags) == [0, 2, 2, 2]
    assert (s[1:] == before[1:]).all() and not (s[0] == before[0]).all()
    m = out.mask.cpu().numpy()
    assert (m[1:].sum(1) == 58).all()           # unchanged states still report the mover's mask
    o = oracle20.new_state(); oracle20.step(o, 0)
    assert (m[0] == oracle20.legal_mask(o)).all()
    # overlapping / edge-touching / non-corner placements on a mid-game state
    o2 = oracle20.new_state()
    for _ in range(9):
        oracle20.step(o2, oracle20.sample_action(o2, 1, 0))
    words = torch.tensor(oracle20.pack(o2).view(np.int32)).cuda().repeat(64, 1).contiguous()
    rng = np.random.default_rng(0)
    cand = rng.integers(0, oracle20.A, 64).astype(np.int32)
    legal = oracle20.legal_mask(o2)
    out = engine20.step(words, torch.tensor(cand).cuda(), mask=None)
    torch.cuda.synchronize()
    got_illegal = (out.flags.cpu().numpy() & 2) != 0
    assert (got_illegal == (legal[cand] == 0)).all()


def test_functional_step_keeps_input(engine20):
    import torch
    s = engine20.new_states(3)
    keep = s.clone()
    dst = torch.empty_like(s)
    out = engine20.step(s, torch.zeros(3, dtype=torch.int32, device=s.device), out_states=dst, mask=None)
    torch.cuda.synchronize()
    assert (s == keep).all() and not (dst == keep).all() and out.states is dst


def test_observation_and_board_contents(engine20, oracle20, engine7, oracle7):
    import torch
    from blokus_rl_b200 import BlokusEngine
    from oracle.oracle import Oracle
    extra = [(BlokusEngine(n, p), Oracle(n, p)) for n, p in ((20, 2), (14, 2), (14, 4))]   # other streaming-kernel variants
    for eng, orc in [(engine20, oracle20), (engine7, oracle7)] + extra:
        sts = []
        for g in range(6):
            o = orc.new_state()
            for _ in range(3 + 5 * g):
                if orc.field(o, "done"):
                    break
                orc.step(o, orc.sample_action(o, 5, g))
            sts.append(o)
        words = torch.tensor(np.stack([orc.pack(o) for o in sts]).view(np.int32)).cuda()
        obs = eng.observe(words).cpu().numpy()
        brd = eng.board_contents(words).cpu().numpy()
        flags, term, scores = eng.game_ended(words)
        for i, o in enumerate(sts):
            assert obs.shape[1:] == (2 * orc.P, orc.N, orc.N)   # blokus_wrapper.py:66-71
            assert (obs[i] == orc.observe(o)).all()
            assert (brd[i] == orc.board_contents(o)).all()
            assert bool(flags[i].item() & 1) == bool(orc.field(o, "done"))
            assert (scores[i].cpu().numpy() == orc.final_scores(o)[: orc.P]).all()
            assert (term[i].cpu().numpy() == orc.terminal_values(o)).all()


def test_fused_observation_output(engine20, oracle20, engine7, oracle7):
    """blk_step_args.obs: the step kernel writes canonical_board of the RESULTING state (after the move, the
    next-mover resolution and a possible auto-reset) -- equal to the oracle's observation of the oracle's next
    state and to the stand-alone blk_observe, for every kernel variant (mask formats, sampler, geometries)."""
    import torch
    from blokus_rl_b200 import BlokusEngine
    from oracle.oracle import Oracle
    extra = [(BlokusEngine(n, p), Oracle(n, p)) for n, p in ((20, 2), (14, 4), (9, 4))]
    for eng, orc in [(engine20, oracle20), (engine7, oracle7)] + extra:
        n, seed, P, N = 37, 21, eng.num_players, eng.board_size
        s = eng.new_states(n)
        ost = [orc.new_state() for _ in range(n)]
        out = eng.step(s, None, mask="bits", sample=True, seed=seed, obs=True)
        torch.cuda.synchronize()
        assert out.obs.shape == (n, 2 * P, N, N)
        assert (out.obs.cpu().numpy() == np.stack([orc.observe(o) for o in ost])).all()
        fmts = ["bytes", "bits", None, "indices"]
        for ply in range(40 if N == 20 else 14):
            acts = out.next_action.clone()
            obs_buf = torch.full((n, 2 * P, N, N), 7.0, device=s.device)
            out = eng.step(s, acts, mask=fmts[ply % 4], sample=True, seed=seed, auto_reset=True, obs=obs_buf)
            torch.cuda.synchronize()
            assert out.obs is obs_buf
            a = acts.cpu().numpy()
            for i, o in enumerate(ost):
                assert orc.step(o, int(a[i])) == 0
                if orc.field(o, "done"):
                    orc.reset(o, orc.field(o, "game") + 1)
            want = np.stack([orc.observe(o) for o in ost])
            assert (obs_buf.cpu().numpy() == want).all(), (N, P, ply)
            assert (eng.observe(s).cpu().numpy() == want).all()
        # functional form without the sampler (the PUCT forest's call): obs belongs to out_states, not to the inputs
        dst = torch.empty_like(s)
        o2 = eng.step(s, out.next_action, out_states=dst, mask="bytes", obs=True)
        torch.cuda.synchronize()
        assert (o2.obs == eng.observe(dst)).all() and not (o2.obs == eng.observe(s)).all()
    with pytest.raises(ValueError):
        engine20.step(engine20.new_states(2), None, obs=torch.empty((2, 8, 20, 19), device="cuda"))


def test_state_index_equals_gather(engine20, engine7):
    """blk_step_args.state_index: env i steps pool[state_index[i]] -- same outputs as gathering the rows first (what the
    PUCT forest used to do with torch.index_select), the pool itself stays untouched."""
    import torch
    for eng in (engine20, engine7):
        m, n = 500, 333
        pool = eng.new_states(m)
        out = eng.step(pool, None, mask=None, sample=True, seed=12)
        for _ in range(10 if eng.board_size == 20 else 3):
            out = eng.step(pool, out.next_action, mask=None, sample=True, seed=12)
        keep = pool.clone()
        idx = torch.randint(0, m, (n,), device=pool.device, dtype=torch.int32)
        acts = out.next_action.index_select(0, idx.long()).contiguous()
        a_dst, b_dst = torch.empty((n, eng.state_words), dtype=torch.int32, device=pool.device), None
        a = eng.step(pool, acts, out_states=a_dst, state_index=idx, mask="bytes", obs=True)
        gathered = pool.index_select(0, idx.long())
        b_dst = torch.empty_like(a_dst)
        b = eng.step(gathered, acts, out_states=b_dst, mask="bytes", obs=True)
        torch.cuda.synchronize()
        assert (pool == keep).all() and (a_dst == b_dst).all() and (a.mask == b.mask).all() and (a.obs == b.obs).all()
        for name in ("legal_count", "flags", "terminal", "scores"):
            assert (getattr(a, name) == getattr(b, name)).all()
        with pytest.raises(ValueError):
            eng.step(pool, acts, state_index=idx)                    # in place makes no sense with an index


def test_unaligned_and_contiguous_bool_mask(engine20, oracle20, engine7):
    """Caller-provided contiguous bool [n, A] buffers (row stride 30,433: every row starts at a different offset inside
    a 16 B chunk) and arbitrarily offset bases give the same masks as the padded layout, and neighbours stay intact."""
    import torch
    for eng in (engine20, engine7):
        n, A = 70, eng.num_actions
        s = eng.new_states(n)
        out = eng.step(s, None, mask="bytes", sample=True, seed=9)
        for _ in range(14 if eng.board_size == 20 else 3):
            out = eng.step(s, out.next_action, mask="bytes", sample=True, seed=9)
        want = out.mask.clone()
        assert (want.sum(1) > 0).all()
        for base_off, stride in ((0, A), (1, A), (7, A + 3), (13, A + 16), (16, A)):
            raw = torch.full((base_off + n * stride + 64,), 7, dtype=torch.uint8, device=s.device)
            view = raw[base_off: base_off + n * stride].view(n, stride)[:, :A]
            got = eng.step(s, None, mask=view)
            torch.cuda.synchronize()
            assert got.mask.shape == (n, A)
            assert (view.bool() == want).all(), (base_off, stride)
            assert (raw[:base_off] == 7).all() and (raw[base_off + n * stride:] == 7).all()      # nothing outside the rows
            if stride > A:
                assert (raw[base_off: base_off + n * stride].view(n, stride)[:, A:] == 7).all()   # nor between them


@pytest.mark.parametrize("N,P", [(14, 2), (14, 4), (12, 2), (20, 4)])
def test_playouts_do_not_depend_on_what_shared_memory_held(N, P, engine20):
    """Playouts on every SM right after kernels with another shared-memory layout ran there (the chunked field readers of the
    14x14 kernel read the padding behind the last field: it must be zeroed by the kernel itself, not by luck).  Enough
    playouts to occupy the whole GPU; a strided sample is replayed by the oracle, all of them must be complete games."""
    from blokus_rl_b200 import BlokusEngine
    from oracle.oracle import Oracle
    import ctypes as C
    import torch
    # dirty the shared memory of every SM: 20x20 byte-mask steps (fields + tables + LUT at other offsets)
    s20 = engine20.new_states(16384)
    o = engine20.step(s20, None, mask="bytes", sample=True, seed=1)
    for _ in range(6):
        o = engine20.step(s20, o.next_action, mask="bytes", sample=True, seed=1)
    eng, orc = BlokusEngine(N, P), Oracle(N, P)
    roots = eng.new_states(64)
    o = eng.step(roots, None, mask=None, sample=True, seed=3)
    for _ in range(6):
        o = eng.step(roots, o.next_action, mask=None, sample=True, seed=3)
    per_root, seed = 128, 99
    out = eng.rollout(roots, per_root, seed=seed)
    torch.cuda.synchronize()
    fs, pl, win = out.final_scores.cpu().numpy(), out.plies.cpu().numpy(), out.winners.cpu().numpy()
    assert (win != 0).all() and (pl > 0).all() and (pl <= 21 * P).all()
    host = orc.unpack_many(roots.cpu().numpy())
    for r in range(0, 64, 4):
        root = C.create_string_buffer(host[r].tobytes(), orc.state_size)
        for j in range(0, per_root, 16):
            n, scores, _, _, _ = orc.playout(root, seed, r * per_root + j)
            assert n == pl[r, j] and (scores == fs[r, j]).all()
    eng.close()


def test_rollouts_replay_through_oracle(engine20, oracle20):
    import torch
    orc = oracle20
    roots = []
    for g in range(6):
        o = orc.new_state()
        for _ in range(24):                      # SURVEY.md 8d workload 3: 24 random plies
            orc.step(o, orc.sample_action(o, 11, g))
        roots.append(o)
    words = torch.tensor(np.stack([orc.pack(o) for o in roots]).view(np.int32)).cuda()
    per_root, seed = 16, 0xABCDEF
    out = engine20.rollout(words, per_root, seed=seed, log_actions=True)
    torch.cuda.synchronize()
    log = out.action_log.cpu().numpy().view(np.uint16)
    fs = out.final_scores.cpu().numpy()
    win = out.winners.cpu().numpy()
    plies = out.plies.cpu().numpy()
    vsum = out.value_sum.cpu().numpy()
    for r, root in enumerate(roots):
        acc = np.zeros(4, np.float32)
        for j in range(per_root):
            o = orc.copy(root)
            gid = r * per_root + j
            k = 0
            while not orc.field(o, "done"):
                a = int(log[r, j, k])
                assert a == orc.sample_action(o, seed, gid, stream=1), "rollout sampler differs from oracle"
                assert orc.step(o, a) == 0
                k += 1
            assert log[r, j, k] == 0xFFFF and plies[r, j] == k
            assert (fs[r, j] == orc.final_scores(o)).all()
            assert win[r, j] == orc.winners(o)
            acc += orc.terminal_values(o)
        assert np.allclose(vsum[r], acc)


def test_full_size_properties_65536(engine20):
    """BASELINE config[1] size: 65,536 envs.  Size-independent properties only (no oracle at this size):
    bytes == unpack(bits), count == popcount, scores == occupied squares, inventories consistent."""
    import torch
    eng = engine20
    n = 65536
    s = eng.new_states(n)
    buf = eng.make_buffers(n, "bytes", sample=True)
    out = eng.step(s, None, buffers=buf, mask="bytes", sample=True, seed=1)
    for _ in range(40):
        out = eng.step(s, out.next_action.clone(), buffers=buf, mask="bytes", sample=True, seed=1, auto_reset=True)
    bits = eng.step(s, None, mask="bits")
    torch.cuda.synchronize()
    assert (out.flags & 2).sum().item() == 0
    idx = torch.randint(0, n, (512,), device=s.device)
    mb = out.mask[idx].cpu().numpy()
    assert (unpack_bits(bits.mask[idx].cpu().numpy(), eng.num_actions) == mb).all()
    assert (out.mask.sum(1) == out.legal_count).all()
    assert (bits.legal_count == out.legal_count).all()
    # every chosen action is legal under the mask it was sampled from
    assert out.mask.gather(1, out.next_action.clamp(min=0).long()[:, None]).all()
    w = s.cpu().numpy().view(np.uint32)
    rows = w[:, :80].reshape(n, 4, 20)
    occupied = np.zeros((n, 4), np.int64)
    for q in range(4):
        occupied[:, q] = np.unpackbits(rows[:, q].copy().view(np.uint8), axis=1).sum(1)
    scores = w[:, 86:88].copy().view(np.int16).reshape(n, 4)
    assert (scores == occupied).all()
    assert (rows[:, 0] & rows[:, 1]).sum() == 0 and (rows[:, 2] & rows[:, 3]).sum() == 0


def test_plain_c_program_runs(tmp_path):
    """examples/c_abi_smoke.c: the C ABI driven from plain C with cudaMalloc'd buffers (no torch, no Python)."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    exe = tmp_path / "c_abi_smoke"
    cmd = ["gcc", "-std=c99", "-I", str(root / "include"), "-I", "/usr/local/cuda/include",
           str(root / "examples" / "c_abi_smoke.c"), "-o", str(exe), "-L", str(root / "blokus_rl_b200"),
           "-lblokus_b200", "-L", "/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{root / 'blokus_rl_b200'}"]
    subprocess.check_call(cmd)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "c_abi_smoke ok" in out.stdout, out.stdout + out.stderr


def test_sparse_index_mask_format(engine20, engine7, oracle20):
    """BLK_MASK_INDICES: ascending legal ids (the reference's ai_possible_indexes lists) == nonzero of the dense mask."""
    import torch
    for eng in (engine20, engine7):
        n = 300
        s = eng.new_states(n)
        out = eng.step(s, None, mask="bytes", sample=True, seed=4)
        for _ in range(16 if eng.board_size == 20 else 4):
            out = eng.step(s, out.next_action, mask="bytes", sample=True, seed=4)
        sparse = eng.step(s, None, mask="indices")
        torch.cuda.synchronize()
        ids = sparse.mask.cpu().numpy().view(np.uint16)
        cnt = sparse.legal_count.cpu().numpy()
        dense = out.mask.cpu().numpy()
        assert ids.shape == (n, eng.max_legal) and (sparse.flags & 4).sum() == 0
        for i in range(n):
            assert ids[i, : cnt[i]].tolist() == np.flatnonzero(dense[i]).tolist()
        # the sampler draws the same action whatever the mask format (index lists come straight from the fields)
        a_idx = eng.step(s, None, mask="indices", sample=True, seed=4, env_id_base=77).next_action
        a_byt = eng.step(s, None, mask="bytes", sample=True, seed=4, env_id_base=77).next_action
        a_non = eng.step(s, None, mask=None, sample=True, seed=4, env_id_base=77).next_action
        assert torch.equal(a_idx, a_byt) and torch.equal(a_idx, a_non) and bool((a_idx >= 0).all())
    for (N, P) in ((14, 2), (9, 4)):                              # 14x14 specialisation and the runtime-dimension kernels
        from blokus_rl_b200 import BlokusEngine
        eng = BlokusEngine(N, P)
        s = eng.new_states(200)
        out = eng.step(s, None, mask="bytes", sample=True, seed=5)
        for _ in range(10):
            out = eng.step(s, out.next_action, mask="bytes", sample=True, seed=5)
        sp = eng.step(s, None, mask="indices", sample=True, seed=5)
        torch.cuda.synchronize()
        ids, cnt, dense = sp.mask.cpu().numpy().view(np.uint16), sp.legal_count.cpu().numpy(), out.mask.cpu().numpy()
        for i in range(200):
            assert ids[i, : cnt[i]].tolist() == np.flatnonzero(dense[i]).tolist()
        assert torch.equal(sp.next_action, out.next_action)
        eng.close()
    # compact form (blk_step_args.csr_cursor / csr_offset): one flat array, as many entries as there are legal moves
    for eng in (engine20, engine7):
        n = 777
        s = eng.new_states(n)
        out = eng.step(s, None, mask="bytes", sample=True, seed=6)
        for _ in range(14 if eng.board_size == 20 else 3):
            out = eng.step(s, out.next_action, mask="bytes", sample=True, seed=6)
        c = eng.step(s, None, mask="csr", sample=True, seed=6)
        torch.cuda.synchronize()
        flat, off, cnt = c.mask_raw.cpu().numpy().view(np.uint16), c.csr_offset.cpu().numpy(), c.legal_count.cpu().numpy()
        dense = out.mask.cpu().numpy()
        assert int(c.csr_cursor.item()) == int(cnt.sum()) == int(dense.sum()) and not (c.flags.cpu().numpy() & 4).any()
        spans = sorted((int(o), int(k)) for o, k in zip(off, cnt))
        assert spans[0][0] == 0 and all(a + k == b for (a, k), (b, _) in zip(spans, spans[1:]))      # the envs tile the array
        for i in range(n):
            assert flat[off[i]: off[i] + cnt[i]].tolist() == np.flatnonzero(dense[i]).tolist()
        assert torch.equal(c.next_action, out.next_action)
        # too small an array: envs that do not fit are flagged and write nothing; the others are complete
        from blokus_rl_b200.engine import StepOut
        cap = int(cnt.sum()) // 2
        small = StepOut(None, None, None, None, None, None, None, torch.full((cap,), -1, dtype=torch.int16, device=s.device))
        c2 = eng.step(s, None, mask="csr", buffers=small)
        torch.cuda.synchronize()
        fl, off2, flat2 = c2.flags.cpu().numpy(), c2.csr_offset.cpu().numpy(), c2.mask_raw.cpu().numpy().view(np.uint16)
        assert (fl & 4).any() and not (fl & 4).all() and int(c2.csr_cursor.item()) == int(cnt.sum())
        for i in range(n):
            if fl[i] & 4:
                assert off2[i] + cnt[i] > cap
            else:
                assert flat2[off2[i]: off2[i] + cnt[i]].tolist() == np.flatnonzero(dense[i]).tolist()
    # truncation is flagged, never silent: a 16-entry row cannot hold the 58 first moves
    s = engine20.new_states(2)
    small = torch.zeros((2, 16), dtype=torch.int16, device=s.device)
    o2 = engine20.step(s, None, mask=small)
    first = np.flatnonzero(oracle20.legal_mask(oracle20.new_state()))
    assert (o2.flags.cpu().numpy() & 4).all() and (o2.legal_count.cpu().numpy() == 58).all()
    assert small[0].cpu().numpy().view(np.uint16).tolist() == first[:16].tolist()
