#!/usr/bin/env python
"""Attribute an ncu capture's per-SASS-instruction counters to CUDA source lines.

  python tools/ncu_lines.py <report.ncu-rep> <kernel-substring> [--so blokus_rl_b200/libblokus_b200.so]
                            [--units N]  (divide counts by N, e.g. envs per launch)

`ncu --page source --csv` gives per-instruction "Instructions Executed" and stall samples but no source
lines; `nvdisasm --print-line-info` gives the line of every instruction.  Both list the kernel's
instructions in address order, so they are joined by index.  Prints instruction totals per source line,
per opcode, and the stall-sample distribution.
"""
import argparse
import csv
import io
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict
from pathlib import Path


def sass_lines(so: Path, kernel: str):
    # every kernel specialisation is its own object (same cubin name inside the .so): disassemble the objects
    objs = sorted((so.resolve().parent / "build").glob("*.o")) or [so.resolve()]
    chunks = []
    for obj in objs:
        tmp = Path(tempfile.mkdtemp())
        subprocess.run(["cuobjdump", "-xelf", "all", str(obj)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for c in sorted(tmp.glob("*.cubin")):
            chunks.append(subprocess.run(["nvdisasm", "--print-line-info", str(c)], capture_output=True, text=True).stdout)
    txt = "\n".join(chunks)
    out, cur_line, active = [], ("?", 0), False
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            active = kernel in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "(.*?)", line (\d+)', ln)
        if m:
            cur_line = (Path(m.group(1)).name, int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), cur_line, m.group(2).strip()))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--so", default="blokus_rl_b200/libblokus_b200.so")
    ap.add_argument("--units", type=float, default=1.0)
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--launch", type=int, default=0, help="which captured launch of the kernel")
    ap.add_argument("--sass-name", default=None, help="substring of the mangled name in the cubin (default: kernel)")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    # split per kernel launch
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(raw)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            blocks.append(cur)
        elif cur is not None and row:
            cur["rows"].append(row)
    blocks = [b for b in blocks if a.kernel in b["name"]]
    if not blocks:
        sys.exit("kernel not found in report")
    b = blocks[a.launch]
    hdr, rows = b["rows"][0], b["rows"][1:]
    ci = hdr.index("Instructions Executed")
    cs = hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    sass = sass_lines(Path(a.so), a.sass_name or a.kernel)
    if len(sass) != len(rows):
        print(f"warning: {len(sass)} SASS instructions in the .so vs {len(rows)} in the report (rebuilt since capture?)")
    per_line, per_op, samp_line = Counter(), Counter(), Counter()
    stalls = Counter()
    total = 0
    for (off, line, text), row in zip(sass, rows):
        n = int(row[ci])
        total += n
        per_line[line] += n
        op = text.split()[0] if not text.startswith("@") else text.split()[1]
        per_op[op.split(".")[0]] += n
        samp_line[line] += int(row[cs])
        for i, h in stall_cols:
            stalls[h] += int(row[i])
    srcs = {}
    def text_of(key):
        f, line = key if isinstance(key, tuple) else ("?", 0)
        if f not in srcs:
            cands = list(Path("blokus_rl_b200/csrc").glob(f))
            srcs[f] = cands[0].read_text().splitlines() if cands else []
        src = srcs[f]
        return src[line - 1].strip()[:100] if 0 < line <= len(src) else "?"
    u = a.units
    print(f"kernel {b['name']}: {total} warp instructions executed ({total / u:.1f} per unit), {len(rows)} SASS instructions")
    print("\n-- by source line (warp instr / unit, stall samples) --")
    for line, n in per_line.most_common(a.top):
        tag = f"{line[0]}:{line[1]}" if isinstance(line, tuple) else str(line)
        print(f"{n / u:10.1f} {100 * n / total:5.1f}%  samples {samp_line[line]:6d}  {tag}: {text_of(line)}")
    print("\n-- by opcode --")
    for op, n in per_op.most_common(25):
        print(f"{n / u:10.1f} {100 * n / total:5.1f}%  {op}")
    ts = sum(stalls.values())
    print("\n-- stall samples --")
    for h, n in stalls.most_common(10):
        print(f"{100 * n / max(ts, 1):5.1f}%  {h}")


if __name__ == "__main__":
    main()
