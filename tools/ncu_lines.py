#!/usr/bin/env python
"""Per-source-line stall samples of one kernel: joins `ncu --page source` (SASS rows) with `nvdisasm -g` line info.

usage: ncu_lines.py <rep> <kernel-substring> [min_samples] [lib.so]
Prints samples per source line (file:line, samples, top stall reasons) and per-line executed instruction counts.
The library must be the build the profile was taken with (same SASS).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, pat = sys.argv[1], sys.argv[2]
min_s = int(sys.argv[3]) if len(sys.argv) > 3 else 10
lib = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "pixelspointspolygons_b200", "libp3p.so")

# ---- SASS with line info --------------------------------------------------------------------------------------
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
lines_of = None
for cub in sorted(os.listdir(tmp)):
    if cub.count("-") > 0:
        continue  # the merged cubin duplicates the per-file ones
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    cur_fn, cur_line, acc = None, None, {}
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur_fn = m.group(1)
            acc[cur_fn] = []
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur_fn:
            acc[cur_fn].append((cur_line, m.group(2).strip()))
    for fn, ins in acc.items():
        if pat in fn and ins:
            lines_of = ins
            print("kernel:", fn[:120], "sass:", len(ins))
if lines_of is None:
    sys.exit("kernel not found in " + lib)

# ---- ncu per-instruction samples ----------------------------------------------------------------------------------
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, out, k, take = None, [], 0, False
for r in rows:
    if r and r[0] in ("Function Name", "Kernel Name"):
        take = not out  # first kernel of the report
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr is not None and take and len(r) == len(hdr):
        out.append(r)
if len(out) != len(lines_of):
    print(f"WARNING: ncu has {len(out)} SASS rows, nvdisasm {len(lines_of)}: library differs from the profiled build")
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for (line, _), r in zip(lines_of, out):
    s = int(r[isamp] or 0)
    tot += s
    e = per[line]
    e[0] += s
    e[1] += int(r[iex] or 0)
    for j in stalls:
        v = int(r[j] or 0)
        if v:
            e[2][hdr[j][6:]] += v
print("total samples", tot)
srcs = {}
for (line, e) in sorted(per.items(), key=lambda kv: (kv[0] is None, kv[0])):
    if e[0] < min_s or line is None:
        continue
    f, n = line
    if f not in srcs:
        p = os.path.join(os.path.dirname(lib), "csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.isfile(p) else []
    text = srcs[f][n - 1].strip()[:90] if 0 < n <= len(srcs[f]) else ""
    top = ",".join(f"{k}={v}" for k, v in e[2].most_common(2))
    print(f"{f}:{n:<4d} {e[0]:5d} {100 * e[0] / tot:5.1f}%  inst={e[1]:>9d}  [{top}]  {text}")
