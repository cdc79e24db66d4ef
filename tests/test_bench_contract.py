"""bench.py's JSON contract (the driver parses these lines): the CPU arm here, the GPU arm on the GPU box."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(*args, timeout=600):
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "3", "--warmup", "3")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["metric"] == "env steps/s w/ legal masks (20x20 4p)" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["value"] > 1e4 and d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run("--steps", "30", "--warmup", "3", "--no-extra", "--no-cpu")
    assert (BASE_KEYS - {"cpu_baseline"}) <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 30 and d["gpu_launches"] == 30 and d["dtype"] == "u32"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0.5 < r["frac"] <= 1.05 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] > 1e7
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"} and d["counters"]["illegal"] == 0
    assert d["value"] > 1e8 and "workload" in d["config"] and "l2" in d["config"]


def test_reference_arm_rollouts_line():
    d = _run("--impl", "reference", "--workload", "rollouts", "--steps", "2", "--warmup", "3")
    assert d["impl"] == "reference" and d["unit"] == "rollouts/s" and d["metric"].startswith("MCTS rollouts/s")
    assert d["value"] > 10 and d["cpu_baseline"]["kind"] == "port" and "playouts" in d["cpu_baseline"]["sample"]


@pytest.mark.gpu
def test_gpu_arm_search_workload_lines():
    r = _run("--workload", "rollouts", "--steps", "2", "--roots", "256", "--per-root", "256")
    assert r["metric"].startswith("MCTS rollouts/s") and r["unit"] == "rollouts/s" and r["value"] > 1e6 and r["gpu_launches"] == 2
    assert r["counters"]["playouts"] == 256 * 256 and r["e2e"]["h2d_bytes_per_step"] > 0 and r["roofline"]["bound"] == "issue"
    p = _run("--workload", "puct", "--steps", "2", "--roots", "64", "--sims", "20")
    assert p["unit"] == "simulations/s" and p["value"] > 1e4 and p["counters"]["overflow"] == 0 and p["e2e"]["d2h_bytes_per_step"] == 4 * 64
    q = _run("--workload", "puct", "--steps", "2", "--roots", "1", "--sims", "50", "--warps-per-tree", "8")
    assert q["value"] > 1e4 and "8 warp(s) per tree" in q["config"]["workload"]


def test_committed_ncu_traffic_is_that_of_the_current_kernel_sources():
    """`roofline.traffic` comes from profiles/traffic.json, which `tools/profile_all.py --summarise` stamps with the hash of the
    kernel sources it captured; bench.py nulls an entry with another stamp.  The committed record must be the current one:
    after an edit under csrc/ (or of the header), re-capture (tools/gpu/README.md) before committing."""
    from blokus_rl_b200.build import kernel_source_hash
    rec = json.loads((ROOT / "profiles" / "traffic.json").read_text())
    assert rec["source_hash"] == kernel_source_hash(), "profiles/traffic.json was captured from other kernel sources"
    assert {"step_kernel_20_4_bytes_65536", "step_kernel_20_4_bits_65536", "rollout_kernel_20_4", "step_kernel_7_2_bytes_1048576"} <= set(rec["kernels"])
