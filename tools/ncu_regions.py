#!/usr/bin/env python
"""Stall samples of the first kernel of an .ncu-rep aggregated over ranges of SASS rows.  usage: ncu_regions.py rep [rows_per_bin]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; step = int(sys.argv[2]) if len(sys.argv) > 2 else 100
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, out, take = None, [], False
for r in rows:
    if r and r[0] in ("Function Name", "Kernel Name"):
        take = not out
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr is not None and take and len(r) == len(hdr):
        out.append(r)
isamp, iex, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for b0 in range(0, len(out), step):
    blk = out[b0:b0 + step]
    s = sum(int(r[isamp] or 0) for r in blk); ex = sum(int(r[iex] or 0) for r in blk)
    c = collections.Counter()
    for r in blk:
        for j in stalls:
            v = int(r[j] or 0)
            if v: c[hdr[j][6:]] += v
    ops = collections.Counter(r[isrc].split()[0] if not r[isrc].startswith("@") else r[isrc].split()[1] for r in blk)
    print(f"#{b0:5d} samples={s:5d} inst={ex:9d}  {dict(c.most_common(4))}  {dict(ops.most_common(4))}")
