#!/usr/bin/env python
"""Multi-GPU check of blokus_rl_b200.distributed on real GPUs (NCCL).   torchrun --nproc-per-node N tools/dist_check.py

The N-rank run's reduced counters must equal (a) the ORACLE's counters for the same global env ids (the C restatement
on the host cores: steps, finished games, sum of legal counts seen by the sampler) and (b) a single-rank run over the
same global range -- for random play, for sharded playouts (plus an oracle replay of a sample of them) and for the
sharded PUCT forest."""
import json
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from blokus_rl_b200 import BlokusEngine, distributed as D

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
D.init("nccl", dev)
eng = BlokusEngine(20, 4, device=dev)
ok = True

# ---- random play (BASELINE.json configs[4] in small) ----
total, plies, seed = 4096, 80, 2024
shard = D.shard_from_env(total)
got = D.reduce_counters(D.random_play_shard(eng, shard, plies, seed))
# ---- playouts sharded by root index ----
rshard = D.shard_from_env(256)
rc, rout = D.rollout_shard(eng, rshard, 64, 7)
rgot = D.reduce_counters(rc, D.ROLLOUT_COUNTERS)
# ---- PUCT searches sharded by root index ----
pshard = D.shard_from_env(64)
pc, _ = D.puct_shard(eng, pshard, 30)
pgot = D.reduce_counters(pc, D.PUCT_COUNTERS)

if shard.rank == 0:
    from oracle.oracle import Oracle
    orc = Oracle(20, 4)
    ost = orc.new_states(total)
    r = orc.play_many(ost, seed, plies)
    want_orc = {"steps": r["steps"], "games": r["games"], "legal_actions_sum": r["legal_sum"]}
    whole = D.random_play_shard(eng, D.Shard(0, 1, total), plies, seed)
    want = dict(zip(D.COUNTERS, (int(x) for x in whole.cpu())))
    print("random play, sharded :", json.dumps(got))
    print("random play, 1 rank  :", json.dumps(want))
    print("random play, oracle  :", json.dumps(want_orc))
    ok &= got == want and got["illegal"] == 0 and got["games"] > 0 and all(got[k] == v for k, v in want_orc.items())

    wc, wout = D.rollout_shard(eng, D.Shard(0, 1, 256), 64, 7)
    rwant = dict(zip(D.ROLLOUT_COUNTERS, (int(x) for x in wc.cpu())))
    print("playouts, sharded    :", json.dumps(rgot))
    print("playouts, 1 rank     :", json.dumps(rwant))
    ok &= rgot == rwant
    # oracle replay: the first 4 playouts of every 8th root
    import ctypes as C
    roots = orc.unpack_many(D.midgame_roots(eng, 0, 256).cpu().numpy())
    fs, pl = wout.final_scores.cpu().numpy(), wout.plies.cpu().numpy()
    for rr in range(0, 256, 8):
        root = C.create_string_buffer(roots[rr].tobytes(), orc.state_size)
        for j in range(4):
            n, scores, _, _, _ = orc.playout(root, 7, rr * 64 + j)
            ok &= bool(n == pl[rr, j] and (scores == fs[rr, j]).all())
    print("playouts, oracle replay of 128 games:", "equal" if ok else "DIFFERENT")

    qc, _ = D.puct_shard(eng, D.Shard(0, 1, 64), 30)
    pwant = dict(zip(D.PUCT_COUNTERS, (int(x) for x in qc.cpu())))
    print("puct, sharded        :", json.dumps(pgot))
    print("puct, 1 rank         :", json.dumps(pwant))
    ok &= pgot == pwant and pgot["overflow"] == 0
    print("DIST_CHECK", "OK" if ok else "MISMATCH", "world", shard.world)
if torch.distributed.is_initialized():
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
