#!/usr/bin/env python
"""Re-capture the ncu evidence of EVERY hot kernel at the kernels that are in the tree now, and rebuild profiles/.

  on the GPU box (one GPU):   python tools/profile_all.py --capture [--tag r2] [--only step_bytes,rollout]
      runs THIS file's --driver under `ncu --set full --import-source on --clock-control none --profile-from-start off`;
      the driver warms each workload up, then brackets ONE launch of its kernel with cudaProfilerStart/Stop, so one ncu
      run captures one launch per target -> gpurun_out/<tag>_all.ncu-rep + gpurun_out/<tag>_profile_meta.json
  back in the build container: python tools/profile_all.py --summarise [--tag r2]
      reads the report (ncu -i ... --page raw --csv), writes profiles/<tag>_<target>_summary.txt for every target and
      regenerates profiles/traffic.json, stamped with the git head and the sha256 of the kernel sources the capture ran
      on (blokus_rl_b200.build.kernel_source_hash); bench.py refuses (null) entries whose hash is not the tree's.

Targets: step kernel 20x20/4p in all four mask formats, the 7x7/2p thread-per-env step kernel (bytes, bits), both playout
kernels, the observation kernel, the PUCT select / expand kernels, the fused search kernel.
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_op_write.sum', 'smsp__cycles_active.avg',
        'smsp__average_warp_latency_issue_stalled_barrier.pct', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']


# ------------------------------------------------------------------------------------------------------------------
# driver (runs under ncu on the GPU box)
# ------------------------------------------------------------------------------------------------------------------
def driver(only: set[str] | None, tag: str):
    import torch
    from blokus_rl_b200 import BlokusEngine
    from blokus_rl_b200.build import kernel_source_hash
    from blokus_rl_b200.gpu_puct import GpuPuct
    prof = torch.cuda.profiler
    meta = {"source_hash": kernel_source_hash(), "gpu": torch.cuda.get_device_name(0), "targets": []}

    def capture(name, kernel, units, unit, fn, note=""):
        """`fn` launches exactly the kernel(s) of interest; everything before it ran outside the window."""
        if only and name not in only:
            return
        torch.cuda.synchronize()
        prof.start()
        fn()
        torch.cuda.synchronize()
        prof.stop()
        meta["targets"].append({"name": name, "kernel": kernel, "units": units, "unit": unit, "note": note})
        print("captured", name, flush=True)

    def midgame(eng, n, plies, fmt, seed=0x5EED):
        st = eng.new_states(n)
        buf = eng.make_buffers(n, fmt if isinstance(fmt, str) else None, sample=True)
        mask = fmt
        out = eng.step(st, None, buffers=buf, mask=mask, sample=True, seed=seed)
        for _ in range(plies):
            out = eng.step(st, buf.next_action, buffers=buf, mask=mask, sample=True, seed=seed, auto_reset=True)
        return st, buf

    eng = BlokusEngine(20, 4)
    E = 65536
    A = eng.num_actions
    for name, fmt in (("step_bytes", "bytes"), ("step_bits", "bits"), ("step_indices", "indices")):
        if only and name not in only:
            continue
        st, buf = midgame(eng, E, 20, fmt)
        capture(name, "step_kernel", E, "env steps",
                lambda: eng.step(st, buf.next_action, buffers=buf, mask=fmt, sample=True, seed=0x5EED, auto_reset=True),
                f"20x20/4p, {E} envs, ply 21 of random play, mask format {fmt}, sampler on")
        del st, buf
    if not only or "step_unaligned" in only:
        st, buf = midgame(eng, E, 20, None)
        raw = torch.empty((E, A), dtype=torch.uint8, device="cuda")            # contiguous bool [n, A]: rows start anywhere
        capture("step_unaligned", "step_kernel", E, "env steps",
                lambda: eng.step(st, buf.next_action, buffers=buf, mask=raw, sample=True, seed=0x5EED, auto_reset=True),
                f"20x20/4p, {E} envs, byte mask into a contiguous [n, {A}] buffer (unaligned rows)")
        del st, buf, raw
    if not only or "leaf_expand" in only:
        st, buf = midgame(eng, E, 20, "bytes")
        obs = torch.empty((E, 8, 20, 20), dtype=torch.float32, device="cuda")
        capture("leaf_expand", "step_kernel", E, "leaves",
                lambda: eng.step(st, None, buffers=buf, mask="bytes", obs=obs),
                "mask-only + fused observation planes (AlphaZero leaf expansion), 65,536 leaves")
        capture("observe", "observe_rows_kernel", E, "observations", lambda: eng.observe(st, out=obs), "stand-alone observation kernel")
        del st, buf, obs
    if not only or "rollout" in only:
        roots, _ = midgame(eng, 1024, 23, None, seed=24)
        eng.rollout(roots, 64, seed=7)
        capture("rollout", "rollout_kernel", 1024 * 1024, "playouts", lambda: eng.rollout(roots, 1024, seed=7),
                "1,024 roots after 24 random plies x 1,024 playouts to terminal")
    # the fused search: whole simulations in one kernel (one warp per tree; 8 warps on one tree with virtual loss)
    for name, B, sims, wpt in (("search_b1", 1, 200, 1), ("search_b4096", 4096, 50, 1), ("search_b1_lp8", 1, 200, 8)):
        if only and name not in only:
            continue
        roots, _ = midgame(eng, B, 23, None, seed=5)
        search = GpuPuct(eng, num_trees=B, max_simulations=2 * sims + 8, mean_edges_per_node=420, warps_per_tree=wpt)
        search.set_roots(roots)
        search.run(sims)
        search.set_roots(roots)
        capture(name, "puct_search_kernel", B * sims, "simulations", lambda: search.run(sims),
                f"{B} tree(s) x {sims} simulations in ONE launch, {wpt} warp(s) per tree, 24-ply 20x20 roots, uniform prior")
        del search
    for name, B in (("puct", 16384), ("puct_b1", 1)):
        if only and name not in only:
            continue
        roots, _ = midgame(eng, B, 23, None, seed=5)
        search = GpuPuct(eng, num_trees=B, max_simulations=64, mean_edges_per_node=420, use_cuda_graph=False, fused=False)
        search.set_roots(roots)
        for _ in range(30):
            search.simulate()
        capture(name, "puct_select_kernel|step_kernel|puct_expand_kernel", B, "simulations", lambda: search.simulate(),
                f"one lockstep simulation of {B} trees after 30 (select, blk_step, expand+backup)")
        del search
    eng.close()

    # the other warp-per-env instantiations: the 14x14 specialisation (Blokus Duo board) and the runtime-dimension kernels
    for name, (N_, P_), E_ in (("step_14_2_bytes", (14, 2), 131072), ("step_runtime_10_2_bytes", (10, 2), 131072)):
        if only and name not in only:
            continue
        eo = BlokusEngine(N_, P_)
        st, buf = midgame(eo, E_, 10, "bytes", seed=9)
        capture(name, "step_kernel", E_, "env steps",
                lambda: eo.step(st, buf.next_action, buffers=buf, mask="bytes", sample=True, seed=9, auto_reset=True),
                f"{N_}x{N_}/{P_}p, {E_} envs, byte masks, sampler on")
        del st, buf
        eo.close()

    e7 = BlokusEngine(7, 2)
    E7 = 1 << 20
    for name, fmt in (("small7_bytes", "bytes"), ("small7_bits", "bits")):
        if only and name not in only:
            continue
        st, buf = midgame(e7, E7, 5, fmt, seed=1)
        capture(name, "small_step_kernel", E7, "env steps",
                lambda: e7.step(st, buf.next_action, buffers=buf, mask=fmt, sample=True, seed=1, auto_reset=True),
                f"7x7/2p, {E7} envs, mask format {fmt}, sampler on")
        del st, buf
    if not only or "small7_rollout" in only:
        roots, _ = midgame(e7, 1024, 2, None, seed=3)
        e7.rollout(roots, 64, seed=7)
        capture("small7_rollout", "small_rollout_kernel", 1024 * 1024, "playouts", lambda: e7.rollout(roots, 1024, seed=7),
                "7x7/2p: 1,024 roots x 1,024 playouts, one playout per thread")
    OUT.mkdir(exist_ok=True)
    (OUT / f"{tag}_profile_meta.json").write_text(json.dumps(meta, indent=1))
    print("driver done:", len(meta["targets"]), "targets")


def capture(tag: str, only: str | None):
    OUT.mkdir(exist_ok=True)
    rep = OUT / f"{tag}_all"
    cmd = ["ncu", "--set", "full", "--import-source", "on", "--clock-control", "none", "--profile-from-start", "off",
           "-f", "-o", str(rep), sys.executable, str(Path(__file__).resolve()), "--driver", "--tag", tag]
    if only:
        cmd += ["--only", only]
    print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    (OUT / f"{tag}_profile_capture.log").write_text(res.stdout[-20000:] + "\n--- stderr ---\n" + res.stderr[-20000:])
    print(res.stdout[-3000:])
    if res.returncode != 0:
        print(res.stderr[-3000:])
    return res.returncode


# ------------------------------------------------------------------------------------------------------------------
# summarise (build container)
# ------------------------------------------------------------------------------------------------------------------
def summarise(tag: str):
    from blokus_rl_b200.build import kernel_source_hash
    rep = OUT / f"{tag}_all.ncu-rep"
    meta = json.loads((OUT / f"{tag}_profile_meta.json").read_text())
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = rows[2:]
    kn = hdr.index("Kernel Name")
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    dirty = bool(subprocess.run(["git", "status", "--porcelain", "--", "blokus_rl_b200/csrc", "include"], capture_output=True,
                                text=True, cwd=ROOT).stdout.strip())
    here = kernel_source_hash()
    traffic = {"_comment": "per-launch figures of one `ncu --set full --clock-control none` capture per kernel (tools/profile_all.py); "
                           "traffic = dram__bytes_read.sum + dram__bytes_write.sum",
               "source_hash": meta["source_hash"], "git_head": head + ("+uncommitted kernel sources" if dirty else ""),
               "tree_hash_at_summary": here, "gpu": meta.get("gpu"), "kernels": {}}
    if here != meta["source_hash"]:
        print(f"WARNING: the tree's kernel sources ({here}) are not the ones the capture ran on ({meta['source_hash']})")
    # a capture may come in several parts (gpurun brings back at most 64 MiB per call): keep what another part of the SAME
    # sources already put into traffic.json
    old_tf = ROOT / "profiles" / "traffic.json"
    if old_tf.exists():
        try:
            old = json.loads(old_tf.read_text())
            if old.get("source_hash") == meta["source_hash"]:
                traffic["kernels"].update(old.get("kernels", {}))
        except ValueError:
            pass
    import re
    pos = 0
    prof_dir = ROOT / "profiles"
    key_of = {"step_bytes": "step_kernel_20_4_bytes_65536", "step_bits": "step_kernel_20_4_bits_65536",
              "step_indices": "step_kernel_20_4_indices_65536", "step_unaligned": "step_kernel_20_4_unaligned_65536",
              "leaf_expand": "leaf_expand_20_4_65536", "observe": "observe_20_4_65536", "rollout": "rollout_kernel_20_4",
              "small7_bytes": "step_kernel_7_2_bytes_1048576", "small7_bits": "step_kernel_7_2_bits_1048576",
              "small7_rollout": "rollout_kernel_7_2", "search_b1": "search_kernel_20_4_b1", "search_b4096": "search_kernel_20_4_b4096",
              "search_b1_lp8": "search_kernel_20_4_b1_lp8", "step_14_2_bytes": "step_kernel_14_2_bytes_131072",
              "step_runtime_10_2_bytes": "step_kernel_10_2_bytes_131072"}
    for t in meta["targets"]:
        pat = re.compile(t["kernel"])
        mine = []
        # the capture window of a target holds its kernels in launch order (one per alternative of the pattern)
        want_n = len(t["kernel"].split("|"))
        while pos < len(launches) and len(mine) < want_n:
            if pat.search(launches[pos][kn]):
                mine.append(launches[pos])
            pos += 1
        lines = [f"# {tag} {t['name']}: {t['note']}", f"# git {traffic['git_head']}, kernel sources {meta['source_hash']}, {meta.get('gpu')}",
                 "# ncu --set full --import-source on --clock-control none (one launch; cold-cache, serialised)"]
        for r in mine:
            def val(m, r=r):
                return float(r[hdr.index(m)].replace(",", "")) if m in hdr and r[hdr.index(m)] not in ("", "n/a") else None
            lines.append("--- " + r[kn][:110])
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    lines.append(f"  {w:80s} {r[i]:>16s} {units[i]}")
            dur = val('gpu__time_duration.sum')
            tunit = units[hdr.index('gpu__time_duration.sum')]
            dur_us = None if dur is None else dur * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
                                                     "s": 1e6, "second": 1e6}[tunit]
            tb = (val('dram__bytes_read.sum') or 0) + (val('dram__bytes_write.sum') or 0)
            bunit = units[hdr.index('dram__bytes_read.sum')] if 'dram__bytes_read.sum' in hdr else "byte"
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[bunit]
            inst = val('smsp__inst_executed.sum')
            rec = {"kernel": r[kn][:80], "traffic": int(tb * scale), "duration_us": dur_us,
                   "issue_slot_pct": val('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                   "alu_pipe_pct": val('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'),
                   "dram_pct": val('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                   "lsu_data_pipe_pct": val('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
                   "smem_pipe_pct": val('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
                   "xu_pipe_pct": val('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
                   "warps_active_pct": val('sm__warps_active.avg.pct_of_peak_sustained_active'),
                   "registers": val('launch__registers_per_thread'),
                   "warp_instr_per_unit": None if inst is None else inst / t["units"], "units": t["units"], "unit": t["unit"]}
            lines.append(f"  => {rec['traffic'] / 1e6:.1f} MB DRAM traffic, {rec['warp_instr_per_unit']:.1f} warp instructions per {t['unit'][:-1]}"
                         if inst is not None else f"  => {rec['traffic'] / 1e6:.1f} MB DRAM traffic")
            key = key_of.get(t["name"])
            if key and len(mine) == 1:
                traffic["kernels"][key] = rec
            else:
                traffic["kernels"][f"{t['name']}:{r[kn].split('(')[0][:40]}"] = rec
        (prof_dir / f"{tag}_{t['name']}_summary.txt").write_text("\n".join(lines) + "\n")
        print("wrote", f"profiles/{tag}_{t['name']}_summary.txt", f"({len(mine)} launch(es))")
    (prof_dir / "traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
    print("wrote profiles/traffic.json for sources", meta["source_hash"], "git", traffic["git_head"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--capture", action="store_true")
    ap.add_argument("--driver", action="store_true")
    ap.add_argument("--summarise", action="store_true")
    ap.add_argument("--tag", default="r2")
    ap.add_argument("--only", default=None, help="comma-separated target names")
    a = ap.parse_args()
    if a.driver:
        driver(set(a.only.split(",")) if a.only else None, a.tag)
    elif a.capture:
        sys.exit(capture(a.tag, a.only))
    elif a.summarise:
        summarise(a.tag)
    else:
        ap.print_help()


if __name__ == "__main__":
    main()
