#!/usr/bin/env python
"""bench.py -- env steps/s with full legal masks (20x20, 4 players), BASELINE.json's metric.

One "step" = one ply in every env of the batch: read state + action -> placement, inventory, score,
next-mover resolution with auto-skip, terminal detection with auto-reset, the next mover's FULL legal
mask written to HBM, and the next uniform-random legal action sampled on the device (Philox-4x32-10).
Workload = BASELINE.json configs[1]: 65,536 envs per GPU, random-legal play from reset (weak scaling:
every rank runs 65,536 envs with global env ids, no collective on the hot path, one NCCL all-reduce of
the counters at the end).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mask bytes|bits] [--envs E]
  python bench.py --workload rollouts|puct ...      # BASELINE.json's second metric (MCTS rollouts/s) and the PUCT forest
  python bench.py --impl reference ...              # the CPU arm: oracle port on all host cores
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

UNIT = "steps/s"
ALG_BYTES = {"bytes": 31154, "bits": 4529}     # SURVEY.md section 8d / DESIGN.md: algorithmic HBM bytes per env step
E2E_MIN_STEPS = 200                             # the end-to-end window never shrinks below this, whatever --steps says
IDX_STRIDE = 1024                               # compact id array of the sparse-mask end-to-end leg: room for IDX_STRIDE / 2 ids per env on average


def alg_bytes(fmt: str, N: int, P: int, A: int) -> int:
    """Algorithmic HBM bytes of one env step: state read + write, action, reward/done/score, mask."""
    if (N, P) == (20, 4):
        return ALG_BYTES[fmt]
    state = 4 * (P * N + P + 4)
    return 2 * state + 4 + 13 + (A if fmt == "bytes" else 4 * ((A + 31) // 32))


def metric_name(N: int, P: int, workload: str = "step") -> str:
    if workload == "rollouts":
        return f"MCTS rollouts/s ({N}x{N} {P}p)"
    if workload == "puct":
        return f"MCTS (PUCT) simulations/s ({N}x{N} {P}p)"
    return f"env steps/s w/ legal masks ({N}x{N} {P}p)"


def unit_name(workload: str) -> str:
    return {"rollouts": "rollouts/s", "puct": "simulations/s"}.get(workload, UNIT)


def host_cores() -> int:
    return len(os.sched_getaffinity(0))


def pin_rank_to_its_cores(local: int, local_world: int) -> list[int]:
    """One rank per GPU share the box's host cores: give each its own slice, so the host loop of one rank (event waits,
    pinned copies, launches) is never descheduled by another rank's (the N = 8 end-to-end straggler of round 1)."""
    cpus = sorted(os.sched_getaffinity(0))
    if local_world <= 1 or len(cpus) < local_world:
        return cpus
    k = len(cpus) // local_world
    mine = cpus[local * k:(local + 1) * k]
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return cpus
    return mine


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (oracle/blokus_oracle.c, vectorised bit-parallel variant) on the host cores
# ------------------------------------------------------------------------------------------------
class CpuArm:
    """Uniform-random legal play with the FULL byte mask written every ply, `envs_per_core` envs per host core, every core
    busy (oracle/oracle.py:play_many -> orc_play_many; ctypes releases the GIL, workers are pinned one per core)."""

    def __init__(self, N: int, P: int, envs_per_core: int = 32, seed: int = 0x5EED):
        from oracle.oracle import Oracle
        self.orc = Oracle(N, P)
        self.cores = host_cores()
        self.states = self.orc.new_states(self.cores * envs_per_core)
        self.seed = seed

    def play(self, plies_per_env: int):
        """-> (env steps executed, seconds)."""
        t0 = time.perf_counter()
        r = self.orc.play_many(self.states, self.seed, plies_per_env, write_masks=True, threads=self.cores)
        return r["steps"], time.perf_counter() - t0

    def playouts(self, roots, per_root: int, seed: int = 7):
        """Uniform-random playouts to the end of the game on all cores -> (playouts, plies, seconds)."""
        import ctypes as C
        orc, n = self.orc, roots.shape[0]
        bufs = [C.create_string_buffer(roots[i].tobytes(), orc.state_size) for i in range(n)]
        plies = [0] * self.cores
        cpus = sorted(os.sched_getaffinity(0))

        def work(t):
            try:
                os.sched_setaffinity(0, {cpus[t % len(cpus)]})
            except OSError:
                pass
            for r in range(t, n, self.cores):
                for j in range(per_root):
                    plies[t] += orc.playout(bufs[r], seed, r * per_root + j)[0]

        t0 = time.perf_counter()
        ts = [threading.Thread(target=work, args=(t,)) for t in range(self.cores)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        return n * per_root, sum(plies), time.perf_counter() - t0


def reference_wrapper_rate(N: int, P: int, seconds: float = 3.0, backend=None, ours: bool = False):
    """the REFERENCE's own wrapper code path (blokus_rl/colossumrl/blokus_wrapper.py:233-246, 89-132,
    164-186: get_sample_move -> get_next_state -> get_valid_moves -> get_game_ended), unmodified, one process = how the
    reference runs, with the CPU oracle as the engine under the colosseumrl shim.  None when the reference's Python package
    is not visible (/root/reference or its pip --target copy under baseline/_ref)."""
    sys.path.insert(0, str(ROOT / "tests"))
    try:
        import ref_stubs
        if not ref_stubs.available():
            return None
        import tempfile
        import types
        from blokus_rl_b200 import colosseum_shim
        if backend is None:
            from oracle_backend import OracleBackend
            backend = OracleBackend(N, P)
        colosseum_shim.set_backend(backend)
        colosseum_shim.install()
        ref_stubs.install_stubs()
        cwd, work = os.getcwd(), Path(tempfile.mkdtemp())
        os.chdir(work)                          # the reference writes debug.log / states/ into the CWD
        try:
            import logging
            logging.disable(logging.WARNING)
            if ours:                                     # boundary B1: this repo's sibling of the wrapper (ids, no string round trip)
                from blokus_rl_b200.game_wrapper import BlokusGameWrapper
                game = BlokusGameWrapper(board_size=N, number_of_players=P, backend=backend)
            else:
                from blokus_rl.colossumrl.blokus_wrapper import ColosseumBlokusGameWrapper
                game = ColosseumBlokusGameWrapper(types.SimpleNamespace(board_size=N, number_of_players=P, states_dir=work / "states"))
            plies, t0 = 0, time.perf_counter()
            while time.perf_counter() - t0 < seconds:
                s, p = game.get_init_board()
                while True:
                    s, p = game.get_next_state(s, p, game.get_sample_move(s))
                    game.get_valid_moves(s, p)
                    plies += 1
                    if game.get_game_ended(s) is not None:
                        break
            dt = time.perf_counter() - t0
        finally:
            os.chdir(cwd)
            colosseum_shim.set_backend(None)
            logging.disable(logging.NOTSET)
        return {"value": plies / dt, "unit": "plies/s", "cores": 1, "plies": plies,
                "what": "the reference's unmodified ColosseumBlokusGameWrapper (get_sample_move -> get_next_state -> "
                        "get_valid_moves -> get_game_ended) over the colosseumrl shim on the CPU oracle, one process",
                "reference_from": str(ref_stubs.REFERENCE)}
    except Exception as e:                       # noqa: BLE001 - a reported extra, never a reason to lose the bench line
        return {"error": repr(e)[:200]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, P = args.board, args.players
    arm = CpuArm(N, P)
    cores, E = arm.cores, arm.states.shape[0]
    if args.workload == "rollouts":
        import numpy as np
        roots_o = arm.orc.new_states(cores * 4)
        arm.orc.play_many(roots_o, 24, 24, auto_reset=False, threads=cores)        # mid-game roots (24 random plies)
        per_root = 8
        arm.playouts(roots_o, 1)
        tot = pl = 0
        secs = 0.0
        for _ in range(args.steps):
            a, b, c = arm.playouts(roots_o, per_root)
            tot, pl, secs = tot + a, pl + b, secs + c
        v = tot / secs
        sample = f"{tot} uniform-random playouts to terminal ({pl} plies) from {roots_o.shape[0]} roots after 24 random plies"
        workload = f"Blokus {N}x{N} {P}-player uniform-random playouts to the end of the game, CPU"
    else:
        # Bounded sample: every step plays the same number of plies in each of the E envs; sized so that the whole run is
        # a few seconds of all-core work whatever --steps is (at least ~2 s timed: the round-1 arm timed 0.1 s)
        rate_guess = 6e4 * cores * (1 if N >= 14 else 12)
        plies_per_env = max(2, int(rate_guess * 6.0 / max(1, args.steps) / E))
        arm.play(max(2, plies_per_env * min(args.warmup, 3) // 3))
        tot, secs = 0, 0.0
        for _ in range(args.steps):
            a, b = arm.play(plies_per_env)
            tot, secs = tot + a, secs + b
        v = tot / secs
        sample = (f"{tot} env steps of random-legal play with the full byte mask written every step: {E} envs x "
                  f"{plies_per_env} plies per step x {args.steps} steps, {secs:.1f} s")
        workload = f"Blokus {N}x{N} {P}-player random-legal play with full byte masks, CPU"
    line = {
        "impl": "reference", "metric": metric_name(N, P, args.workload), "value": v, "unit": unit_name(args.workload),
        "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload,
                   "note": "the reference env engine (colosseumrl) is an absent un-vendored dependency; this arm times "
                           "this repo's C restatement (oracle port: incremental row bitboards, AVX-512/AVX2 field "
                           "evaluation), every host core busy"},
        "cpu_baseline": {"value": v, "unit": unit_name(args.workload), "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": unit_name(args.workload), "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.workload == "step" and not args.no_extra:
        w = reference_wrapper_rate(N, P, 2.0)
        if w is not None:
            line["cpu_baseline"]["wrapper_path"] = w
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Polls NVML (SM clock + clock-event reasons) every 5 ms from a thread DURING the timed region."""
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


def profile_record(key: str):
    """ncu-derived figures of one kernel from profiles/traffic.json -- only when they were captured on the kernels that
    are in the tree now (sha256 over csrc/ + the ABI header); a stale record reads as null, never as a number."""
    tf = ROOT / "profiles" / "traffic.json"
    if not tf.exists():
        return None, "no profiles/traffic.json"
    try:
        from blokus_rl_b200.build import kernel_source_hash
        rec = json.loads(tf.read_text())
        if rec.get("source_hash") != kernel_source_hash():
            return None, f"profiles/traffic.json was captured on other kernel sources ({rec.get('source_hash')}, git {rec.get('git_head')})"
        k = rec.get("kernels", {}).get(key)
        return k, (f"ncu --set full capture, git {rec.get('git_head')}, sources {rec.get('source_hash')}" if k else f"no entry {key}")
    except Exception as e:                           # noqa: BLE001
        return None, repr(e)[:120]


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(self.world)))
        self.cpus = pin_rank_to_its_cores(self.local, self.local_world)
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        # stdout carries exactly one JSON line: anything libraries print there while the bench runs (NCCL writes its
        # "NCCL version ..." banner to stdout when NCCL_DEBUG is set) is sent to stderr; the real stdout comes back for the line
        sys.stdout.flush()
        self.real_stdout = os.dup(1)
        os.dup2(2, 1)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_ms(self, ms: float):
        """-> (max over ranks, per-rank list)."""
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
        if self.world == 1:
            return ms, [ms]
        allv = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(allv, t)
        per = [float(x.item()) for x in allv]
        return max(per), per

    def emit(self, line: dict):
        sys.stdout.flush()
        os.dup2(self.real_stdout, 1)
        os.close(self.real_stdout)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)                                   # teardown chatter, if any, stays off stdout too

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def e2e_leg(D: Dist, eng, states, act, fmt, seed, base, steps: int, halves: int):
    """End to end through the public API with HOST buffers: pinned actions H2D, results D2H, every step.
    The batch is driven as `halves` independent part-batches on separate streams (the usual way to keep a device env
    busy while the host consumes results): while one part's results travel to the host and its next actions come back,
    the other part's step kernel runs.  Every step of every env still pays its H2D action copy, its D2H result copies
    and a host synchronisation before the next actions are issued.
    fmt 'csr': the mask itself comes back to the host every step, as the sorted legal ids of every env in one compact
    array (BLK_MASK_INDICES with csr_cursor: as many uint16 entries as there are legal moves, + offset and count per env)
    -- the reference's PPO contract (envs.get_attr("ai_possible_indexes"), ppo/trainer.py:385).  The host reads the
    total first, then copies exactly that many ids: two synchronisations per part and step."""
    torch = D.torch
    P, n = eng.num_players, states.shape[0]
    H = max(1, halves)
    nh = n // H
    sparse = fmt == "csr"
    pin = lambda shape, dtype: torch.empty(shape, dtype=dtype).pin_memory()
    parts = []
    for k in range(H):
        st = states[k * nh:(k + 1) * nh]
        if sparse:
            b = eng.make_buffers(nh, None, sample=True)
            b.mask_raw = torch.empty(nh * IDX_STRIDE // 2, dtype=torch.int16, device=D.dev)
            b.csr_cursor = torch.zeros(1, dtype=torch.int64, device=D.dev)
            b.csr_offset = torch.empty(nh, dtype=torch.int64, device=D.dev)
        else:
            b = eng.make_buffers(nh, fmt, sample=True)
        b.next_action.copy_(act[k * nh:(k + 1) * nh])
        part = {"states": st, "buf": b, "stream": torch.cuda.Stream(device=D.dev), "event": torch.cuda.Event(),
                "h_act": pin(nh, torch.int32), "h_flags": pin(nh, torch.uint8), "h_term": pin((nh, P), torch.float32),
                "base": base + k * nh}
        if sparse:
            part.update(h_ids=pin(b.mask_raw.numel(), torch.int16), h_cnt=pin(nh, torch.int32), h_off=pin(nh, torch.int64),
                        h_tot=pin(1, torch.int64), ev_tot=torch.cuda.Event())
        part["h_act"].copy_(b.next_action)
        parts.append(part)
    D.sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for hv in parts:
        hv["stream"].wait_event(e0)
    ids_bytes = 0
    for _ in range(steps):
        for hv in parts:
            hv["event"].synchronize()                              # the host needs this part's results to act on them
            with torch.cuda.stream(hv["stream"]):
                b = hv["buf"]
                b.next_action.copy_(hv["h_act"], non_blocking=True)          # host policy's actions -> device
                o = eng.step(hv["states"], b.next_action, buffers=b, mask=fmt, sample=True,
                             seed=seed, env_id_base=hv["base"], auto_reset=True)
                hv["h_act"].copy_(o.next_action, non_blocking=True)          # sampled legal actions -> host
                hv["h_flags"].copy_(o.flags, non_blocking=True)              # done / illegal flags -> host
                hv["h_term"].copy_(o.terminal, non_blocking=True)            # terminal vectors (rewards) -> host
                if sparse:
                    hv["h_tot"].copy_(o.csr_cursor, non_blocking=True)       # how many ids there are this step ...
                    hv["h_cnt"].copy_(o.legal_count, non_blocking=True)
                    hv["h_off"].copy_(o.csr_offset, non_blocking=True)
                    hv["ev_tot"].record()
                    hv["ev_tot"].synchronize()
                    tot = min(int(hv["h_tot"][0]), b.mask_raw.numel())
                    hv["h_ids"][:tot].copy_(b.mask_raw[:tot], non_blocking=True)   # ... and exactly those ids -> host
                    ids_bytes += 2 * tot
                hv["event"].record()
    for hv in parts:
        torch.cuda.current_stream().wait_stream(hv["stream"])
    e1.record()
    D.sync_all()
    ms, per_rank = D.max_ms(e0.elapsed_time(e1))
    d2h = (4 + 1 + 4 * P) * n + ((8 + 4 + 8) * n + ids_bytes // max(1, steps) if sparse else 0)
    truncated = int(sum(int((hv["h_flags"] & 4).sum()) for hv in parts)) if sparse else 0
    return {"value": D.world * nh * H * steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 4 * n * D.world,
            "d2h_bytes_per_step": d2h * D.world, "steps": steps, "ms": ms,
            "per_rank_ms": [round(x, 3) for x in per_rank], "straggler_rank": int(max(range(len(per_rank)), key=per_rank.__getitem__)),
            **({"ids_bytes_per_env_step": round(ids_bytes / max(1, steps) / n, 1), "truncated_rows_last_step": truncated} if sparse else {})}


def run_ours(args):
    D = Dist()
    torch, dist, dev, world, rank = D.torch, D.dist, D.dev, D.world, D.rank
    from blokus_rl_b200 import BlokusEngine
    N, P = args.board, args.players
    eng = BlokusEngine(N, P, device=dev)
    if args.workload != "step":
        return run_search_workload(args, D, eng)
    headline = (N, P) == (20, 4)
    AB = alg_bytes(args.mask, N, P, eng.num_actions)
    n, fmt, seed = args.envs, args.mask, 0x5EED
    base = rank * n                                  # global env ids: results are partition-invariant
    states = eng.new_states(n)
    buf = eng.make_buffers(n, fmt, sample=True)
    act = buf.next_action                            # the step reads action[i] then writes next_action[i]: may alias

    def step():
        return eng.step(states, act, buffers=buf, mask=fmt, sample=True, seed=seed, env_id_base=base, auto_reset=True)

    eng.step(states, None, buffers=buf, mask=fmt, sample=True, seed=seed, env_id_base=base)   # first masks + actions
    for _ in range(args.warmup):
        step()
    D.sync_all()

    sampler = ClockSampler(D.local) if rank == 0 else None
    # ---- device-resident throughput: K launches, CUDA events on the launching stream ----
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for k in range(args.steps):
        out = step()
        evs[k + 1].record()
    if sampler is not None and sampler.h is not None:
        sampler.sample_now()                    # the queue is still draining: at least one sample under load, however short the run
    D.sync_all()
    per_launch_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    total_ms, _ = D.max_ms(evs[0].elapsed_time(evs[-1]))

    e2e_steps = max(E2E_MIN_STEPS, min(args.steps, 2000))
    e2e = e2e_leg(D, eng, states, act, fmt, seed, base, e2e_steps, args.e2e_halves)
    try:                                        # a secondary leg must never cost the headline line
        e2e_sparse = e2e_leg(D, eng, states, act, "csr", seed, base, max(40, e2e_steps // 5), args.e2e_halves)
    except Exception as exc:                    # noqa: BLE001
        e2e_sparse = {"error": repr(exc)[:300]}
        D.sync_all()
    clocks = sampler.stop() if sampler else None

    # final counter reduction (the only collective): steps, finished games, illegal flags
    ctr = torch.tensor([n * args.steps, int((out.flags & 1).sum().item()), int((out.flags & 2).sum().item())],
                       dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ctr)
    if rank != 0:
        return D.finish()

    value = world * n * args.steps / (total_ms * 1e-3)
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    mean_launch_ms = sum(per_launch_ms) / len(per_launch_ms)
    achieved = AB * n / (mean_launch_ms * 1e-3) / 1e9
    prof, prof_note = profile_record(f"step_kernel_{N}_{P}_{fmt}_{n}")
    e2e_note = (f"pinned host actions H2D -> blk_step -> sampled actions, flags, terminal vectors D2H, host sync before the next "
                f"actions; {args.e2e_halves} part-batches pipelined on {args.e2e_halves} streams; masks stay on the device "
                f"for the policy net; each rank pinned to its own host cores {D.cpus}")
    line = {
        "metric": metric_name(N, P), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"Blokus {N}x{N} {P}-player batched env: step + full legal mask ({fmt}) + on-device random "
                               f"policy, {n} envs per GPU, auto-reset" + (" (BASELINE.json configs[1])" if headline and n == 65536 else ""),
                   "envs_per_gpu": n, "mask_format": fmt, "parallelism": f"env-sharded x{world}, no hot-path collective",
                   "l2": f"per-step working set {(AB * n) / 1e6:.0f} MB vs 126 MB L2"
                         + (" (mask writes evict it every step)" if AB * n > 2 * 126e6 else " (not larger than L2: states stay L2-resident; see DESIGN.md)")},
        "e2e": {**{k: v for k, v in e2e.items() if k != "ms"}, "note": e2e_note},
        "e2e_mask_to_host": {**{k: v for k, v in e2e_sparse.items() if k != "ms"},
                             "note": "the same loop with the legal mask itself returned to the host every step, as the sorted legal ids of "
                                     "every env in one compact array (BLK_MASK_INDICES + csr_cursor: the reference's ai_possible_indexes "
                                     "contract, ppo/trainer.py:385); the host reads the total, then copies exactly that many ids"},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": prof.get("traffic") if prof else None, "traffic_source": prof_note,
                     "peak_source": peak_src, "kernel": "step_kernel" if N > 7 else "small_step_kernel",
                     "alg_bytes_per_step": AB, "mean_launch_ms": mean_launch_ms},
        "clocks": clocks,
        "counters": {"steps": int(ctr[0]), "games_finished_last_step": int(ctr[1]), "illegal": int(ctr[2])},
    }
    if prof:      # north_star: "integer-ALU/issue-slot utilisation" for the modes that are issue-bound, from the same capture
        line["roofline_issue"] = {k: prof.get(k) for k in ("issue_slot_pct", "alu_pipe_pct", "xu_pipe_pct", "smem_pipe_pct",
                                                           "lsu_data_pipe_pct", "warp_instr_per_unit", "warps_active_pct",
                                                           "duration_us", "dram_pct") if k in prof}
    if world == 1 and not args.no_extra and headline:
        try:
            line["extra"] = extra_workloads(eng, torch)
        except Exception as exc:                # noqa: BLE001 - extras are reported next to the metric, they never replace it
            line["extra"] = {"error": repr(exc)[:300]}
    if world == 1 and not args.no_cpu:
        try:
            arm = CpuArm(N, P)
            arm.play(8)
            s0, t0 = arm.play(16)
            target = 12.0                                         # seconds of CPU work
            plies_per_env = max(16, int(s0 / t0 * target / arm.states.shape[0]))
            steps_done, secs = arm.play(plies_per_env)
            line["cpu_baseline"] = {"value": steps_done / secs, "unit": UNIT, "cores": arm.cores, "kind": "port",
                                    "sample": f"{steps_done} env steps of {N}x{N} {P}p random-legal play with the full byte mask written "
                                              f"every step, {arm.states.shape[0]} envs on {arm.cores} pinned host threads, {secs:.1f} s "
                                              f"(oracle port, vectorised bit-parallel variant)"}
            a7 = CpuArm(7, 2, envs_per_core=1)
            a7.states = a7.states[:1].copy()
            a7.cores = 1
            n7, t7 = a7.play(20000)
            line.setdefault("extra", {})["cpu_7x7_2p_single_env_plies_per_s"] = n7 / t7     # BASELINE configs[0]
            if not args.no_extra:
                w = reference_wrapper_rate(N, P, 3.0)
                if w is not None:
                    line["cpu_baseline"]["wrapper_path"] = w
        except Exception as exc:                # noqa: BLE001 - the CPU arm is a reported baseline; its failure must not lose the GPU line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "port", "sample": "failed: " + repr(exc)[:200]}
    D.emit(line)
    D.finish()


def run_search_workload(args, D: Dist, eng):
    """BASELINE.json's second metric (MCTS rollouts/s, configs[2]) and the PUCT forest, sharded by ROOT index over the
    ranks (weak scaling: every rank owns --roots roots with global root ids; RNG keys carry the global playout / root
    index, so the union over ranks is what one GPU computes for the same ids; one all-reduce of counters at the end)."""
    torch, dist, dev, world, rank = D.torch, D.dist, D.dev, D.world, D.rank
    from blokus_rl_b200 import distributed as BD
    N, P = args.board, args.players
    nroots = args.roots
    shard = BD.Shard(rank, world, nroots * world)
    sampler = ClockSampler(D.local) if rank == 0 else None
    if args.workload == "rollouts":
        per_root = args.per_root
        roots = BD.midgame_roots(eng, shard.lo, shard.n)
        for _ in range(max(args.warmup, 3)):
            BD.rollout_shard(eng, shard, per_root, 7, roots=roots)
        D.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.steps):
            res = eng.rollout(roots, per_root, seed=7 + k, rollout_id_base=shard.lo * per_root)
        e1.record()
        D.sync_all()
        ms, per_rank = D.max_ms(e0.elapsed_time(e1))
        units = nroots * per_root
        # end to end: roots come from pinned host memory every step, the per-root value sums and winners go back
        h_roots = roots.cpu().pin_memory()
        h_val = torch.empty((nroots, P), dtype=torch.float32).pin_memory()
        h_win = torch.empty((nroots, per_root), dtype=torch.uint8).pin_memory()
        d_roots = torch.empty_like(roots)
        esteps = max(3, min(args.steps, 20))
        D.sync_all()
        e0.record()
        for k in range(esteps):
            d_roots.copy_(h_roots, non_blocking=True)
            r = eng.rollout(d_roots, per_root, seed=7 + k, rollout_id_base=shard.lo * per_root)
            h_val.copy_(r.value_sum, non_blocking=True)
            h_win.copy_(r.winners, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e1.record()
        D.sync_all()
        ems, _ = D.max_ms(e0.elapsed_time(e1))
        local_c, _ = BD.rollout_shard(eng, shard, per_root, 7, roots=roots)
        names = BD.ROLLOUT_COUNTERS
        mean_plies = float(res.plies.float().mean().item())
        e2e = {"value": world * units * esteps / (ems * 1e-3), "unit": "rollouts/s", "h2d_bytes_per_step": roots.numel() * 4 * world,
               "d2h_bytes_per_step": (h_val.numel() * 4 + h_win.numel()) * world, "steps": esteps}
        launches = args.steps
        workload = (f"{nroots} fixed mid-game roots per GPU (24 random plies, global root ids) x {per_root} uniform-random playouts "
                    f"to terminal" + (" (BASELINE.json configs[2])" if (nroots, per_root) == (1024, 1024) else ""))
        prof, prof_note = profile_record(f"rollout_kernel_{N}_{P}")
    else:
        from blokus_rl_b200.gpu_puct import GpuPuct
        sims = args.sims
        roots = BD.midgame_roots(eng, shard.lo, shard.n)
        search = GpuPuct(eng, num_trees=shard.n, max_simulations=(args.steps + max(args.warmup, 3)) * sims + 8,
                         mean_edges_per_node=420, fused=not args.lockstep, warps_per_tree=1 if args.lockstep else args.warps_per_tree)
        search.set_roots(roots)
        for _ in range(max(args.warmup, 3)):
            search.run(sims, chain=args.chain)
        D.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            search.run(sims, chain=args.chain)
        e1.record()
        D.sync_all()
        ms, per_rank = D.max_ms(e0.elapsed_time(e1))
        search.check()
        units = nroots * sims
        # end to end: roots from pinned host memory, `sims` simulations, the chosen actions back on the host
        h_roots = roots.cpu().pin_memory()
        h_best = torch.empty(nroots, dtype=torch.int32).pin_memory()
        d_roots = torch.empty_like(roots)
        esteps = max(3, min(args.steps, 10))
        s2 = GpuPuct(eng, num_trees=shard.n, max_simulations=sims + 8, mean_edges_per_node=420, fused=not args.lockstep,
                     warps_per_tree=1 if args.lockstep else args.warps_per_tree)
        s2.set_roots(roots); s2.run(sims, chain=args.chain)         # graphs captured outside the timed region
        D.sync_all()
        e0.record()
        for _ in range(esteps):
            d_roots.copy_(h_roots, non_blocking=True)
            s2.set_roots(d_roots)
            s2.run(sims, chain=args.chain)
            h_best.copy_(s2.best_actions_device(), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e1.record()
        D.sync_all()
        ems, _ = D.max_ms(e0.elapsed_time(e1))
        ctr = search.t["counters"].long()
        local_c = torch.tensor([nroots, units * args.steps, int(ctr[0]), int(ctr[1]), 0, 0, int(ctr[2] + ctr[3])], dtype=torch.int64, device=dev)
        names = BD.PUCT_COUNTERS
        mean_plies = None
        e2e = {"value": world * units * esteps / (ems * 1e-3), "unit": "simulations/s", "h2d_bytes_per_step": roots.numel() * 4 * world,
               "d2h_bytes_per_step": 4 * nroots * world, "steps": esteps}
        launches = (3 * sims if args.lockstep else 2) * args.steps
        how = (f"lockstep kernels (blk_puct_select / blk_step / blk_puct_expand), {args.chain} simulations per CUDA graph" if args.lockstep
               else f"whole simulations inside one kernel (blk_puct_search), {args.warps_per_tree} warp(s) per tree, nodes keyed by board")
        workload = (f"{nroots} PUCT searches per GPU from 24-ply roots (global root ids), {sims} simulations per step, "
                    f"uniform prior (DumbNet, config/mcts_blokus.yml), trees on the GPU: {how}")
        prof, prof_note = None, "no single dominant kernel (select / blk_step / expand)"
    clocks = sampler.stop() if sampler else None
    totals = BD.reduce_counters(local_c, names)
    if rank != 0:
        return D.finish()
    value = world * units * args.steps / (ms * 1e-3)
    line = {"metric": metric_name(N, P, args.workload), "value": value, "unit": unit_name(args.workload), "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32" if args.workload == "rollouts" else "f64", "data": "synthetic",
            "config": {"workload": workload, "roots_per_gpu": nroots, "parallelism": f"root-sharded x{world}, no hot-path collective",
                       "l2": "compute-bound: state in registers, fields in shared memory; HBM traffic is the root read and the result write"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "counters": totals,
            "per_rank_ms": [round(x, 3) for x in per_rank]}
    if args.workload == "rollouts":
        line["rollout_mean_plies"] = mean_plies
        line["plies_per_s"] = value * mean_plies
        # issue-bound kernel: the roofline is the issue rate; achieved = warp instructions issued per second (ncu count x measured rate)
        if prof and prof.get("warp_instr_per_unit"):
            sm, sched, clk = eng.sm_count, 4, (clocks or {}).get("sm_mhz") or 1965
            peak = sm * sched * clk * 1e6 / 1e12                       # T warp-instructions/s at one issue per scheduler per clock
            ach = prof["warp_instr_per_unit"] * value / world / 1e12
            line["roofline"] = {"bound": "issue", "achieved": ach, "peak": peak, "unit": "T warp-instr/s", "frac": ach / peak,
                                "traffic": prof.get("traffic"), "traffic_source": prof_note, "kernel": "rollout_kernel",
                                "warp_instr_per_rollout": prof["warp_instr_per_unit"], "ncu_issue_slot_pct": prof.get("issue_slot_pct"),
                                # what actually binds it (ncu): the shared-memory pipe, with the ALU and XU pipes close behind
                                "ncu_smem_pipe_pct": prof.get("smem_pipe_pct"), "ncu_alu_pipe_pct": prof.get("alu_pipe_pct"),
                                "ncu_xu_pipe_pct": prof.get("xu_pipe_pct")}
        else:
            line["roofline"] = {"bound": "issue", "achieved": None, "peak": None, "unit": "T warp-instr/s", "frac": None,
                                "traffic": None, "traffic_source": prof_note, "kernel": "rollout_kernel"}
    D.emit(line)
    D.finish()


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
    bb = eng.make_buffers(E, "bits", sample=True)
    bb.next_action.copy_(nb.next_action)
    extra["steps_per_s_bit_masks"] = E / timed(
        lambda: eng.step(st, bb.next_action, buffers=bb, mask="bits", sample=True, seed=11, auto_reset=True), 30)
    del mb, bb, st, scratch_states
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
    # the PPO surface of the reference (ppo/trainer.py:146-173, 380-386) over the gym adapter, NumPy in / NumPy out
    import numpy as np
    from blokus_rl_b200.vector_env import BlokusVectorEnv
    for E in (4, 4096):
        env = BlokusVectorEnv(E, engine=e7, seed=1)
        env.reset()
        rng = np.random.default_rng(0)

        def agent_step():                           # a vectorised random masked policy on the host
            ids, cnt = env.legal_ids_padded()
            acts = ids[np.arange(E), (rng.random(E) * cnt).astype(np.int64)].astype(np.int64)
            env.step(acts)
        for _ in range(5):
            agent_step()
        t0 = time.perf_counter()
        reps = 200 if E == 4 else 50
        for _ in range(reps):
            agent_step()
        extra[f"ppo_numpy_surface_agent_steps_per_s_{E}_envs"] = E * reps / (time.perf_counter() - t0)
    e7.close()
    # boundaries B0 / B1 on the GPU, one state at a time (how the reference's own loops call the env): the reference's unmodified
    # wrapper over the colosseumrl shim over the CUDA engine, and this repo's BlokusGameWrapper; one launch + one D2H per ply
    from blokus_rl_b200.backend import EngineBackend
    eb = EngineBackend(engine=eng)
    w0 = reference_wrapper_rate(20, 4, 2.0, backend=eb)
    if w0 is not None:
        extra["b0_reference_wrapper_over_gpu_plies_per_s"] = w0.get("value", w0.get("error"))
    w1 = reference_wrapper_rate(20, 4, 2.0, backend=eb, ours=True)
    if w1 is not None:
        extra["b1_game_wrapper_over_gpu_plies_per_s"] = w1.get("value", w1.get("error"))
    # device-resident PUCT forest (config/mcts_blokus.yml player: MCTS with the uniform DumbNet prior)
    from blokus_rl_b200.gpu_puct import GpuPuct

    def roots_after(B, plies=24, seed=5):
        roots = eng.new_states(B)
        o = eng.step(roots, None, mask=None, sample=True, seed=seed)
        for _ in range(plies):
            o = eng.step(roots, o.next_action, mask=None, sample=True, seed=seed)
        return roots
    # (name, trees, simulations per move, fused kernel?, warps per tree, simulations per CUDA graph on the lockstep path)
    for name, B, sims, fused, wpt, chain in (("mcts_simulations_per_s", 4096, 50, False, 1, 1),
                                             ("mcts_simulations_per_s_fused_B4096", 4096, 50, True, 1, 1),
                                             ("mcts_simulations_per_s_lockstep_graph_B1", 1, 200, False, 1, 50),
                                             ("mcts_simulations_per_s_B1", 1, 200, True, 1, 1),
                                             ("mcts_simulations_per_s_B4", 4, 200, True, 1, 1),
                                             ("mcts_simulations_per_s_B16", 16, 200, True, 1, 1),
                                             ("mcts_simulations_per_s_B1_leaf_parallel_8_warps", 1, 200, True, 8, 1),
                                             ("mcts_simulations_per_s_B1_leaf_parallel_16_warps", 1, 200, True, 16, 1)):
        roots = roots_after(B)
        search = GpuPuct(eng, num_trees=B, max_simulations=5 * sims + 8, mean_edges_per_node=420, fused=fused, warps_per_tree=wpt)

        def move():                                    # one move's search: fresh tree, `sims` simulations, the chosen action
            search.set_roots(roots)
            search.run(sims, chain=chain)
            return search.best_actions_device()
        sec = timed(move, 3)
        search.check()
        extra[name] = B * sims / sec
        del search
    # configs[3]: AlphaZero self-play (config/alphazero_blokus_20x20.yml: 25 simulations per move) -- lockstep forest, fused leaf
    # expansion (state + mask + observation planes in one launch) feeding a torch net of the reference's ResNet shape
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("selfplay_bench", ROOT / "tools" / "selfplay_bench.py")
        sb = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(sb)
        from blokus_rl_b200.mcts import TorchNetEvaluator
        games, sims, plies = 256, 25, 3
        net = sb.PolicyValueNet().to(eng.device).eval()
        search = GpuPuct(eng, TorchNetEvaluator(net), num_trees=games, max_simulations=(sims + 1) * (plies + 3) + 2, mean_edges_per_node=400)
        search.set_roots(eng.new_states(games))

        def one_move():
            search.run(sims)
            search.advance(search.best_actions_device())
        one_move()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(plies):
            one_move()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        search.check()
        extra["selfplay_resnet_simulations_per_s_256_games"] = games * plies * sims / dt
        extra["selfplay_resnet_plies_per_s_256_games"] = games * plies / dt
        extra["selfplay_workload"] = ("256 concurrent self-play games, 25 simulations per move, random-init torch net of the reference's ResNet "
                                      "shape (24.7 M parameters, fp32), leaves expanded by blk_step with fused observation planes; net-bound")
        del search, net
    except Exception as e:                      # noqa: BLE001 - an extra, never a reason to lose the bench line
        extra["selfplay_error"] = repr(e)[:200]
    extra["mcts_workload"] = ("B PUCT searches from 24-ply roots, uniform prior (DumbNet), trees on the GPU.  mcts_simulations_per_s: "
                              "4096 trees in lockstep, 3 launches per simulation (blk_puct_select / blk_step / blk_puct_expand) -- the "
                              "path a torch net uses; *_fused_*, *_B1/B4/B16: whole simulations inside one kernel (blk_puct_search), a "
                              "warp per tree, 200 simulations per move as in players/mcts_player.py + compare_arena.py:87-95, "
                              "reference visit order; *_leaf_parallel_*: the same single tree searched by 8 / 16 warps at once with "
                              "virtual loss (not the reference's visit order).  Each figure includes set_roots and the final argmax.  "
                              "CPU context: the reference's mcts.py over the C oracle does ~2.9e3 simulations/s on one core "
                              "(tools/ref_mcts_cpu_rate.py)")
    return extra


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="step", choices=["step", "rollouts", "puct"],
                    help="step = the headline metric; rollouts = BASELINE.json's 'MCTS rollouts/s' (configs[2]); puct = PUCT forest")
    ap.add_argument("--mask", default="bytes", choices=["bytes", "bits"])
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--board", type=int, default=20, help="board size N (the headline metric is 20)")
    ap.add_argument("--players", type=int, default=4, choices=[2, 4])
    ap.add_argument("--roots", type=int, default=1024, help="roots / trees per GPU (workloads rollouts, puct)")
    ap.add_argument("--per-root", type=int, default=1024, help="playouts per root (workload rollouts)")
    ap.add_argument("--sims", type=int, default=25, help="simulations per tree and step (workload puct)")
    ap.add_argument("--chain", type=int, default=1, help="simulations captured per CUDA graph (workload puct --lockstep)")
    ap.add_argument("--lockstep", action="store_true", help="workload puct: the 3-launch lockstep kernels instead of the fused search")
    ap.add_argument("--warps-per-tree", type=int, default=1, help="workload puct: > 1 = leaf-parallel fused search with virtual loss")
    ap.add_argument("--e2e-halves", type=int, default=2, help="part-batches pipelined on separate streams in the e2e leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the rollout / leaf-expansion extras")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload != "step" and args.steps == 2000:
        args.steps = 10
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
