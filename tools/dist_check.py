#!/usr/bin/env python
"""Multi-GPU check of blokus_rl_b200.distributed on real GPUs (NCCL): the sharded run's reduced counters must equal
a single-rank run over the same global env range.   torchrun --nproc-per-node N tools/dist_check.py"""
import json
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from blokus_rl_b200 import BlokusEngine, distributed as D

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
D.init("nccl", dev)
eng = BlokusEngine(20, 4, device=dev)
total, plies, seed = 4096, 80, 2024
shard = D.shard_from_env(total)
local_c = D.random_play_shard(eng, shard, plies, seed)
got = D.reduce_counters(local_c)
if shard.rank == 0:
    whole = D.random_play_shard(eng, D.Shard(0, 1, total), plies, seed)
    want = dict(zip(D.COUNTERS, (int(x) for x in whole.cpu())))
    print("sharded :", json.dumps(got))
    print("single  :", json.dumps(want))
    print("DIST_CHECK", "OK" if got == want and got["illegal"] == 0 and got["games"] > 0 else "MISMATCH", "world", shard.world)
if torch.distributed.is_initialized():
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
