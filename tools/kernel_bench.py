#!/usr/bin/env python
"""Device-resident micro-benchmarks of the hot kernels (CUDA events), one JSON line.  Used for A/B runs of build variants:
   BLOKUS_B200_LIB=build_exp/libvariant.so python tools/kernel_bench.py [--only step,rollout,small]"""
import argparse
import json
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from blokus_rl_b200 import BlokusEngine

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="step,rollout,small")
ap.add_argument("--tag", default=os.environ.get("BLOKUS_B200_LIB", "product"))
a = ap.parse_args()
only = set(a.only.split(","))


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


res = {"tag": a.tag}
if "step" in only or "rollout" in only:
    eng = BlokusEngine(20, 4)
if "step" in only:
    E = 65536
    for fmt in ("bytes", "bits", "indices"):
        st = eng.new_states(E)
        buf = eng.make_buffers(E, fmt, sample=True)
        eng.step(st, None, buffers=buf, mask=fmt, sample=True, seed=1)
        sec = timed(lambda: eng.step(st, buf.next_action, buffers=buf, mask=fmt, sample=True, seed=1, auto_reset=True), 300, 20)
        res[f"step_{fmt}_per_s"] = E / sec
        del st, buf
if "rollout" in only:
    roots = eng.new_states(1024)
    o = eng.step(roots, None, mask=None, sample=True, seed=24)
    for _ in range(24):
        o = eng.step(roots, o.next_action, mask=None, sample=True, seed=24)
    res["rollouts_per_s"] = 1024 * 1024 / timed(lambda: eng.rollout(roots, 1024, seed=7), 3, 2)
if "small" in only:
    e7 = BlokusEngine(7, 2)
    E = 1 << 20
    for fmt in ("bytes", "bits"):
        st = e7.new_states(E)
        buf = e7.make_buffers(E, fmt, sample=True)
        e7.step(st, None, buffers=buf, mask=fmt, sample=True, seed=1)
        sec = timed(lambda: e7.step(st, buf.next_action, buffers=buf, mask=fmt, sample=True, seed=1, auto_reset=True), 100, 10)
        res[f"small7_{fmt}_per_s"] = E / sec
        del st, buf
    roots = e7.new_states(1024)
    o = e7.step(roots, None, mask=None, sample=True, seed=3)
    for _ in range(2):
        o = e7.step(roots, o.next_action, mask=None, sample=True, seed=3)
    res["small7_rollouts_per_s"] = 1024 * 1024 / timed(lambda: e7.rollout(roots, 1024, seed=7), 5, 2)
print(json.dumps(res))
