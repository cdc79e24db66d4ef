"""Memory safety without compute-sanitizer (it is closed on this GPU pool): every output buffer of every entry point sits
between sentinel bands, batch sizes are ragged (not multiples of the 8 envs a block holds, of a warp, of the 16 B store
width), and after the launches the bands must be untouched; the device forest's arrays must be untouched behind the
nodes / edges / pool rows its counters say were used."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
BAND, FILL = 4096, 0xA5


class Guarded:
    def __init__(self, shape, dtype, device):
        self.shape, self.dtype = tuple(shape), dtype
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        self.raw = torch.full((BAND + n + BAND,), FILL, dtype=torch.uint8, device=device)
        self.t = self.raw[BAND:BAND + n].view(dtype).view(shape)

    def check(self, what):
        assert bool((self.raw[:BAND] == FILL).all()), f"{what}: wrote in front of the buffer"
        assert bool((self.raw[-BAND:] == FILL).all()), f"{what}: wrote behind the buffer"


@pytest.mark.parametrize("N,P", [(20, 4), (20, 2), (14, 4), (14, 2), (12, 2), (7, 2), (7, 4), (9, 2)])
def test_step_outputs_stay_inside_their_buffers(N, P):
    from blokus_rl_b200 import BlokusEngine
    from blokus_rl_b200.engine import StepOut
    eng = BlokusEngine(N, P)
    dev = eng.device
    for n in (1, 13, 259):
        states = Guarded((n, eng.state_words), torch.int32, dev)
        eng.reset(states.t)
        states.check("reset")
        act = Guarded((n,), torch.int32, dev)
        for fmt, mshape, mdtype in (("bytes", (n, eng.mask_bytes), torch.uint8), ("bits", (n, eng.mask_words), torch.int32),
                                    ("indices", (n, 40), torch.int16), ("unaligned", (n, eng.num_actions), torch.uint8)):
            g = {"mask": Guarded(mshape, mdtype, dev), "legal_count": Guarded((n,), torch.int32, dev),
                 "terminal": Guarded((n, P), torch.float32, dev), "flags": Guarded((n,), torch.uint8, dev),
                 "scores": Guarded((n, P), torch.int16, dev), "next_action": act,
                 "obs": Guarded((n, 2 * P, N, N), torch.float32, dev), "out": Guarded((n, eng.state_words), torch.int32, dev)}
            bufs = StepOut(None, None, g["legal_count"].t, g["terminal"].t, g["flags"].t, g["scores"].t, act.t, None)
            eng.step(states.t, None, out_states=g["out"].t, mask=g["mask"].t, buffers=bufs, sample=True, seed=3, obs=g["obs"].t)
            for _ in range(6):
                eng.step(g["out"].t, act.t, mask=g["mask"].t, buffers=bufs, sample=True, seed=3, obs=g["obs"].t, auto_reset=True)
            torch.cuda.synchronize()
            for k, v in g.items():
                v.check(f"{N}x{N}/{P}p n={n} {fmt}: {k}")
            states.check("state_in")
        # the small streaming kernels
        for name, shape, dtype, fn in (("observe", (n, 2 * P, N, N), torch.float32, lambda o: eng.observe(states.t, out=o)),):
            o = Guarded(shape, dtype, dev)
            fn(o.t)
            torch.cuda.synchronize()
            o.check(name)
    eng.close()


@pytest.mark.parametrize("N,P", [(20, 4), (7, 2)])
def test_rollout_outputs_stay_inside_their_buffers(N, P):
    import ctypes as C
    from blokus_rl_b200 import BlokusEngine, _lib
    eng = BlokusEngine(N, P)
    dev = eng.device
    n, per = 11, 7
    roots = eng.new_states(n)
    o = eng.step(roots, None, mask=None, sample=True, seed=2)
    for _ in range(3):
        o = eng.step(roots, o.next_action, mask=None, sample=True, seed=2)
    g = {"fs": Guarded((n * per, P), torch.int16, dev), "win": Guarded((n * per,), torch.uint8, dev),
         "vs": Guarded((n, P), torch.float32, dev), "log": Guarded((n * per, 88), torch.int16, dev),
         "plies": Guarded((n * per,), torch.int32, dev), "out": Guarded((n * per, eng.state_words), torch.int32, dev)}
    g["vs"].t.zero_()
    for stop in (-1, 1):
        args = _lib.BlkRolloutArgs(n, roots.data_ptr(), per, 5, 0, g["fs"].t.data_ptr(), g["win"].t.data_ptr(), g["vs"].t.data_ptr(),
                                   g["log"].t.data_ptr(), 88, g["plies"].t.data_ptr(), stop, g["out"].t.data_ptr(), 0)
        _lib.check(eng._lib.blk_rollout(eng._h, C.byref(args), None))
        torch.cuda.synchronize()
        for k, v in g.items():
            v.check(f"rollout {N}x{N}/{P}p stop={stop}: {k}")
    eng.close()


@pytest.mark.parametrize("wpt", [1, 8])
def test_search_forest_is_untouched_behind_what_its_counters_claim(engine20, wpt):
    from blokus_rl_b200.gpu_puct import GpuPuct
    eng, B, sims = engine20, 5, 160
    roots = eng.new_states(B)
    o = eng.step(roots, None, mask=None, sample=True, seed=4)
    for _ in range(14):
        o = eng.step(roots, o.next_action, mask=None, sample=True, seed=4)
    s = GpuPuct(eng, num_trees=B, max_simulations=sims + 40, mean_edges_per_node=700, warps_per_tree=wpt)
    sentinel = {"edge_action": 0x5A5A5A5A, "edge_child": 0x5A5A5A5A, "node_edge0": 0x5A5A5A5A, "node_state": 0x5A5A5A5A,
                "node_front": 0x5A5A5A5A, "node_tree": 0x5A5A5A5A}
    for k, v in sentinel.items():
        s.t[k].fill_(v)
    s.t["edge_n"].fill_(-7.0)
    s.pool.fill_(0x5A5A5A5A)
    s.set_roots(roots)
    s.run(sims)
    s.check()
    nodes, edges = int(s.t["counters"][0].item()), int(s.t["counters"][1].item())
    assert B < nodes <= B * (sims + 1) + (B * sims if wpt > 1 else 0) and edges > 0
    for k, v in sentinel.items():
        used = edges if k.startswith("edge") else nodes
        assert bool((s.t[k][used:] == v).all()), f"{k} written behind the {used} entries in use"
        assert not bool((s.t[k][:used] == v).any()), f"{k} has unwritten entries among the {used} in use"
    assert bool((s.t["edge_n"][edges:] == -7.0).all()) and bool((s.t["edge_n"][:edges] >= 0).all())
    assert bool((s.pool[nodes:] == 0x5A5A5A5A).all()) and not bool((s.pool[:nodes, 0] == 0x5A5A5A5A).any())
    # every link points at a live node, every node's edge range lies inside the used edges
    child = s.t["edge_child"][:edges]
    assert int(child.max().item()) < nodes and int(child.min().item()) >= -1
    e0, ne = s.t["node_edge0"][:nodes], s.t["node_nedge"][:nodes]
    live = e0 >= 0
    assert bool(((e0[live] + ne[live]) <= edges).all())
    filed = s.t["hash_table"]
    nfiled = int((filed > 0).sum().item())          # (leaf-parallel: a node that lost the race for its board is not filed)
    assert int(filed.max().item()) <= nodes and (nfiled == nodes if wpt == 1 else nodes - 16 <= nfiled <= nodes)
