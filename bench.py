#!/usr/bin/env python
"""bench.py -- env steps/s with full legal masks (20x20, 4 players), BASELINE.json's metric.

One "step" = one ply in every env of the batch: read state + action -> placement, inventory, score,
next-mover resolution with auto-skip, terminal detection with auto-reset, the next mover's FULL legal
mask written to HBM, and the next uniform-random legal action sampled on the device (Philox-4x32-10).
Workload = BASELINE.json configs[1]: 65,536 envs per GPU, random-legal play from reset (weak scaling:
every rank runs 65,536 envs with global env ids, no collective on the hot path, one NCCL all-reduce of
the counters at the end).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mask bytes|bits] [--envs E]
  python bench.py --impl reference ...   # the CPU arm: oracle port on all host cores
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "env steps/s w/ legal masks (20x20 4p)"
UNIT = "steps/s"
ALG_BYTES = {"bytes": 31154, "bits": 4529}     # SURVEY.md section 8d / DESIGN.md: algorithmic HBM bytes per env step


def alg_bytes(fmt: str, N: int, P: int, A: int) -> int:
    """Algorithmic HBM bytes of one env step: state read + write, action, reward/done/score, mask."""
    if (N, P) == (20, 4):
        return ALG_BYTES[fmt]
    state = 4 * (P * N + P + 4)
    return 2 * state + 4 + 13 + (A if fmt == "bytes" else 4 * ((A + 31) // 32))


def metric_name(N: int, P: int) -> str:
    return f"env steps/s w/ legal masks ({N}x{N} {P}p)"


def host_cores() -> int:
    return len(os.sched_getaffinity(0))


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (oracle/blokus_oracle.c, bit-parallel variant) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_random_play(total_plies: int, threads: int, seed: int = 0x5EED, N: int = 20, P: int = 4):
    """Uniform-random legal play with full byte masks on `threads` host threads (ctypes releases the GIL).
    Returns (plies, seconds)."""
    from oracle.oracle import Oracle
    orc = Oracle(N, P)
    per = max(1, total_plies // threads)
    states = [orc.new_state() for _ in range(threads)]
    done = [0] * threads

    def work(i):
        n, _, _ = orc.random_play(states[i], seed, i, per, auto_reset=True, fast=True, log=False)
        done[i] = n

    t0 = time.perf_counter()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return sum(done), time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    N, P = args.board, args.players
    per_step = (256 if N >= 20 else 4096) * cores      # bounded sample: 256 plies (20x20) per core per "step"
    cpu_random_play(per_step * max(1, args.warmup), cores, N=N, P=P)
    plies, secs = cpu_random_play(per_step * args.steps, cores, N=N, P=P)
    v = plies / secs
    line = {
        "impl": "reference", "metric": metric_name(N, P), "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"Blokus {N}x{N} {P}-player random-legal play with full byte masks, CPU",
                   "note": "the reference env engine (colosseumrl) is an absent un-vendored dependency; this arm times "
                           "this repo's C restatement (oracle port, bit-parallel variant), one env per host thread"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{plies} plies of random-legal play ({per_step} per step), one env per thread"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Polls NVML (SM clock + clock-event reasons) every 10 ms from a thread DURING the timed region."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, torch_index: int):
        self.sm, self.bits, self.max_mhz, self.h = [], 0, None, None
        self._stop = threading.Event()
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(torch_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception as e:                      # noqa: BLE001 - report, never fail the bench on telemetry
            self.err = repr(e)
            self.h = None

    def sample_now(self):
        nv = self.nv
        try:
            self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            pass

    def _poll(self):
        while not self._stop.is_set():
            self.sample_now()
            time.sleep(0.005)

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self._stop.set()
        self.t.join(timeout=1)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": [n for b, n in self.REASONS.items() if self.bits & b], "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from blokus_rl_b200 import BlokusEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly one JSON line: anything libraries print there while the bench runs (NCCL writes its
    # "NCCL version ..." banner to stdout when NCCL_DEBUG is set) is sent to stderr; the real stdout comes back for the line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    N, P = args.board, args.players
    headline = (N, P) == (20, 4)
    eng = BlokusEngine(N, P, device=dev)
    AB = alg_bytes(args.mask, N, P, eng.num_actions)
    n, fmt, seed = args.envs, args.mask, 0x5EED
    base = rank * n                                  # global env ids: results are partition-invariant
    states = eng.new_states(n)
    buf = eng.make_buffers(n, fmt, sample=True)
    act = buf.next_action                            # the step reads action[i] then writes next_action[i]: may alias

    def step():
        return eng.step(states, act, buffers=buf, mask=fmt, sample=True, seed=seed, env_id_base=base, auto_reset=True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    eng.step(states, None, buffers=buf, mask=fmt, sample=True, seed=seed, env_id_base=base)   # first masks + actions
    for _ in range(args.warmup):
        step()
    sync_all()

    sampler = ClockSampler(local) if rank == 0 else None
    # ---- device-resident throughput: K launches, CUDA events on the launching stream ----
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    games = torch.zeros((), dtype=torch.int64, device=dev)
    evs[0].record()
    for k in range(args.steps):
        out = step()
        evs[k + 1].record()
    if sampler is not None and sampler.h is not None:
        sampler.sample_now()                    # the queue is still draining: at least one sample under load, however short the run
    sync_all()
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())

    # ---- end to end through the public API with HOST buffers: pinned actions H2D, results D2H, every step ----
    # The batch is driven as `args.e2e_halves` independent half-batches on separate streams (the usual way to
    # keep a device env busy while the host consumes results): while one half's results travel to the host and
    # its next actions come back, the other half's step kernel runs.  Every step of every env still pays its
    # H2D action copy, its D2H result copies and a host synchronisation before the next actions are issued.
    H = max(1, args.e2e_halves)
    nh = n // H
    halves = []
    for k in range(H):
        st = states[k * nh:(k + 1) * nh]
        b = eng.make_buffers(nh, fmt, sample=True)
        b.next_action.copy_(act[k * nh:(k + 1) * nh])
        halves.append({
            "states": st, "buf": b, "stream": torch.cuda.Stream(device=dev), "event": torch.cuda.Event(),
            "h_act": torch.empty(nh, dtype=torch.int32).pin_memory(), "h_flags": torch.empty(nh, dtype=torch.uint8).pin_memory(),
            "h_term": torch.empty((nh, P), dtype=torch.float32).pin_memory(), "base": base + k * nh})
        halves[-1]["h_act"].copy_(b.next_action)
    e2e_steps = max(4, min(args.steps, 512))
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for hv in halves:
        hv["stream"].wait_event(e0)
    for _ in range(e2e_steps):
        for hv in halves:
            hv["event"].synchronize()                              # the host needs this half's results to act on them
            with torch.cuda.stream(hv["stream"]):
                b = hv["buf"]
                b.next_action.copy_(hv["h_act"], non_blocking=True)          # host policy's actions -> device
                o = eng.step(hv["states"], b.next_action, buffers=b, mask=fmt, sample=True, seed=seed,
                             env_id_base=hv["base"], auto_reset=True)
                hv["h_act"].copy_(o.next_action, non_blocking=True)          # sampled legal actions -> host
                hv["h_flags"].copy_(o.flags, non_blocking=True)              # done / illegal flags -> host
                hv["h_term"].copy_(o.terminal, non_blocking=True)            # terminal vectors (rewards) -> host
                hv["event"].record()
    for hv in halves:
        torch.cuda.current_stream().wait_stream(hv["stream"])
    e1.record()
    sync_all()
    e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    clocks = sampler.stop() if sampler else None

    # final counter reduction (the only collective): steps, finished games, illegal flags
    ctr = torch.tensor([n * args.steps, int((out.flags & 1).sum().item()), int((out.flags & 2).sum().item())],
                       dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ctr)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * n * args.steps / (total_ms * 1e-3)
    e2e_value = world * nh * H * e2e_steps / (e2e_ms * 1e-3)
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    mean_launch_ms = sum(per_launch_ms) / len(per_launch_ms)
    achieved = AB * n / (mean_launch_ms * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        traffic = json.loads(tf.read_text()).get(f"step_kernel_{fmt}_{n}" if headline else f"step_kernel_{N}_{P}_{fmt}_{n}")
    line = {
        "metric": metric_name(N, P), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"Blokus {N}x{N} {P}-player batched env: step + full legal mask ({fmt}) + on-device random "
                               f"policy, {n} envs per GPU, auto-reset" + (" (BASELINE.json configs[1])" if headline and n == 65536 else ""),
                   "envs_per_gpu": n, "mask_format": fmt, "parallelism": f"env-sharded x{world}, no hot-path collective",
                   "l2": f"per-step working set {(AB * n) / 1e6:.0f} MB vs 126 MB L2"
                         + (" (mask writes evict it every step)" if AB * n > 2 * 126e6 else " (not larger than L2: states stay L2-resident; see DESIGN.md)")},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * n * world,
                "d2h_bytes_per_step": (4 + 1 + 4 * P) * n * world, "steps": e2e_steps,
                "note": f"pinned host actions H2D -> blk_step -> sampled actions, flags, terminal vectors D2H, host sync before the next actions; {H} half-batches pipelined on {H} streams; masks stay on the device for the policy net"},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "step_kernel",
                     "alg_bytes_per_step": AB, "mean_launch_ms": mean_launch_ms},
        "clocks": clocks,
        "counters": {"steps": int(ctr[0]), "games_finished_last_step": int(ctr[1]), "illegal": int(ctr[2])},
    }
    if world == 1 and not args.no_extra and headline:
        line["extra"] = extra_workloads(eng, torch)
    if world == 1 and not args.no_cpu:
        cores = host_cores()
        target = 12.0                                         # seconds of CPU work
        p1, s1 = cpu_random_play(2000 * cores, cores, N=N, P=P)
        plies, secs = cpu_random_play(int(p1 / s1 * target), cores, N=N, P=P)
        from oracle.oracle import Oracle
        o7 = Oracle(7, 2)
        s7 = o7.new_state()
        t7 = time.perf_counter()
        n7, _, _ = o7.random_play(s7, 0, 0, 20000, auto_reset=True, fast=True, log=False)
        line.setdefault("extra", {})["cpu_7x7_2p_single_env_plies_per_s"] = n7 / (time.perf_counter() - t7)   # BASELINE configs[0]
        line["cpu_baseline"] = {"value": plies / secs, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{plies} plies of {N}x{N} {P}p random-legal play with full byte masks, one env per "
                                          f"host thread, {secs:.1f} s (oracle port, bit-parallel variant)"}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)                                   # teardown chatter, if any, stays off stdout too
    if world > 1:
        dist.destroy_process_group()


def extra_workloads(eng, torch):
    """BASELINE.json configs[2] and [3], device-resident, CUDA-event timed (reported next to the headline metric)."""
    def timed(fn, reps):
        for _ in range(3):                      # allocator + clocks settle (the first two playout launches run slow)
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    # configs[2]: 1,024 fixed mid-game roots (24 random plies from seeds 0..1023), 1,024 playouts per root to the end
    roots = eng.new_states(1024)
    out = eng.step(roots, None, mask=None, sample=True, seed=24)
    for _ in range(24):
        out = eng.step(roots, out.next_action, mask=None, sample=True, seed=24)
    res = {}
    sec = timed(lambda: res.__setitem__("r", eng.rollout(roots, 1024, seed=7)), 3)
    plies = float(res["r"].plies.float().mean().item())
    extra = {"mcts_rollouts_per_s": 1024 * 1024 / sec, "rollout_plies_per_s": 1024 * 1024 * plies / sec,
             "rollout_workload": "1024 roots after 24 random plies x 1024 uniform-random playouts to terminal",
             "rollout_mean_plies": plies}
    # SURVEY 8d workload 2: reset, mask and step kernels separately (the headline is the fused step + mask + sampler)
    E = 65536
    st = eng.new_states(E)
    o = eng.step(st, None, mask=None, sample=True, seed=11)
    for _ in range(16):
        o = eng.step(st, o.next_action, mask=None, sample=True, seed=11, auto_reset=True)
    mb = eng.make_buffers(E, "bytes")
    scratch_states = eng.new_states(E)
    extra["reset_states_per_s"] = E / timed(lambda: eng.reset(scratch_states), 20)
    extra["legal_mask_only_bytes_per_s"] = E / timed(lambda: eng.step(st, None, buffers=mb, mask="bytes"), 20)
    nb = eng.make_buffers(E, None, sample=True)
    nb.next_action.copy_(o.next_action)
    extra["step_without_mask_per_s"] = E / timed(
        lambda: eng.step(st, nb.next_action, buffers=nb, mask=None, sample=True, seed=11, auto_reset=True), 20)
    del mb, st, scratch_states
    # Everything through HOST buffers (states, actions in; states, masks, outputs back), for the record: this is what
    # a caller pays who keeps nothing on the device.  PCIe carries 30 KB of byte mask (or 3.8 KB of bit mask) per env
    # and step, so the link, not the kernel, sets the rate -- which is why the adapters keep states and masks on the GPU.
    for fmt in ("bytes", "bits"):
        E = 8192
        d_st = eng.new_states(E)
        o = eng.step(d_st, None, mask=None, sample=True, seed=13)
        for _ in range(16):
            o = eng.step(d_st, o.next_action, mask=None, sample=True, seed=13)
        hb = eng.make_buffers(E, fmt, sample=True)
        h_st = d_st.cpu().pin_memory()
        h_act = o.next_action.cpu().pin_memory()
        h_mask = torch.empty(hb.mask_raw.shape, dtype=hb.mask_raw.dtype).pin_memory()
        h_flags = torch.empty(E, dtype=torch.uint8).pin_memory()
        h_term = torch.empty((E, 4), dtype=torch.float32).pin_memory()
        d_act = torch.empty(E, dtype=torch.int32, device=d_st.device)

        def host_step():
            d_st.copy_(h_st, non_blocking=True)
            d_act.copy_(h_act, non_blocking=True)
            r = eng.step(d_st, d_act, buffers=hb, mask=fmt, sample=True, seed=13, auto_reset=True)
            h_st.copy_(d_st, non_blocking=True)
            h_mask.copy_(r.mask_raw, non_blocking=True)
            h_act.copy_(r.next_action, non_blocking=True)
            h_flags.copy_(r.flags, non_blocking=True)
            h_term.copy_(r.terminal, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        extra[f"e2e_all_host_buffers_{fmt}_per_s"] = E / timed(host_step, 10)
        del hb, h_mask
    # configs[3]: batched leaf expansion feeding the torch net: obs f32 [B,8,20,20] + bool mask [B,30433] + terminal vectors
    for B in (256, 4096, 65536):
        leaves = eng.new_states(B)
        o = eng.step(leaves, None, mask=None, sample=True, seed=3)
        for _ in range(20):
            o = eng.step(leaves, o.next_action, mask=None, sample=True, seed=3)
        buf = eng.make_buffers(B, "bytes")
        obs = torch.empty((B, 8, 20, 20), dtype=torch.float32, device=leaves.device)

        def expand():                       # one launch: legal mask + terminal vector + observation planes
            eng.step(leaves, None, buffers=buf, mask="bytes", obs=obs)
        sec = timed(expand, 20)
        extra[f"leaf_expansions_per_s_B{B}"] = B / sec

        def expand2():                      # the two-kernel form (blk_step + blk_observe), for comparison
            eng.step(leaves, None, buffers=buf, mask="bytes")
            eng.observe(leaves, out=obs)
        extra[f"leaf_expansions_two_kernels_per_s_B{B}"] = B / timed(expand2, 20)
    # the reference's "simplified Blokus" (config/ppo_blokus_7x7.yml): 7x7, two players, thread-per-env kernels
    from blokus_rl_b200 import BlokusEngine
    e7 = BlokusEngine(7, 2, device=eng.device)
    for fmt in ("bytes", "bits"):
        E = 1 << 20
        st = e7.new_states(E)
        b7 = e7.make_buffers(E, fmt, sample=True)
        e7.step(st, None, buffers=b7, mask=fmt, sample=True, seed=1)
        sec = timed(lambda: e7.step(st, b7.next_action, buffers=b7, mask=fmt, sample=True, seed=1, auto_reset=True), 30)
        extra[f"steps_per_s_7x7_2p_{fmt}_1M_envs"] = E / sec
        del st, b7
    e7.close()
    # device-resident PUCT forest (config/mcts_blokus.yml player: MCTS with the uniform DumbNet prior), B trees in lockstep
    from blokus_rl_b200.gpu_puct import GpuPuct
    B, sims = 4096, 50
    roots = eng.new_states(B)
    o = eng.step(roots, None, mask=None, sample=True, seed=5)
    for _ in range(24):
        o = eng.step(roots, o.next_action, mask=None, sample=True, seed=5)
    search = GpuPuct(eng, num_trees=B, max_simulations=2 * sims + 4, mean_edges_per_node=420)
    search.set_roots(roots)
    sec = timed(search.simulate, sims)
    search.check()
    extra["mcts_simulations_per_s"] = B / sec
    extra["mcts_workload"] = f"{B} PUCT searches in lockstep from 24-ply roots, uniform prior (DumbNet), tree on the GPU (blk_puct_*)"
    return extra


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mask", default="bytes", choices=["bytes", "bits"])
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--board", type=int, default=20, help="board size N (the headline metric is 20)")
    ap.add_argument("--players", type=int, default=4, choices=[2, 4])
    ap.add_argument("--e2e-halves", type=int, default=2, help="half-batches pipelined on separate streams in the e2e leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the rollout / leaf-expansion extras")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
