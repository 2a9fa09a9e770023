#!/usr/bin/env python
"""Phase timeline of the voxelize kernel (library built with P3P_EXTRA_NVCC_FLAGS=-DP3P_TIMELINE).
usage (GPU box): P3P_EXTRA_NVCC_FLAGS=-DP3P_TIMELINE python -m pixelspointspolygons_b200.build --force && python tools/timeline.py [B] [N]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth as po  # (neutral generators: measurement tools do not touch oracle/)
from pixelspointspolygons_b200 import PointPillarsEncoder, _lib, default_cfg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
dev = torch.device("cuda:0")
cfg = default_cfg(device="cuda:0")
enc = PointPillarsEncoder(cfg, voxel_encoder={"in_channels": 3, "feat_channels": [64, 384]},
                          scatter={"in_channels": 384, "output_shape": [28, 28]}).to(dev).eval()
enc.load_state_dict(po.synth_weights(0)[0])
tiles = [po.synth_tile(N, 1000 + i, clustered=(i % 2 == 1)) for i in range(B)]
x = torch.from_numpy(np.stack(tiles)).to(dev)
out = torch.empty(B, 784, 384, device=dev)
import ctypes
zero = (C.c_ulonglong * (16 * 1024))()
for _ in range(5):
    enc.encode_into(x, out, 1)
torch.cuda.synchronize()
n = 1024
buf = (C.c_ulonglong * (16 * n))()
l = _lib.lib()
rc = l.p3p_debug_timeline(buf, n)
assert rc == 0, rc
t = np.frombuffer(buf, dtype=np.uint64).reshape(n, 16).astype(np.int64)
used = t[:, 0] > 0
t = t[used]
t0 = t[:, 0].min()
names = ["start", "located", "hashed", "counted", "published", "acquired", "classified", "scattered", "crossing", "synced", "plan0", "plan1"]
print(f"{used.sum()} CTAs; times in us relative to the first CTA start")
for i, nm in enumerate(names):
    col = t[:, i]
    ok = col > 0
    if ok.sum() == 0:
        continue
    v = (col[ok] - t0) / 1e3
    line = f"{nm:10s} n={ok.sum():4d} min={v.min():7.2f} p50={np.median(v):7.2f} p90={np.percentile(v, 90):7.2f} max={v.max():7.2f}"
    if i > 0:
        both = ok & (t[:, i - 1] > 0) & (col >= t[:, i - 1])
        d = (col[both] - t[both, i - 1]) / 1e3
        if both.sum():
            line += f"   | phase: p10={np.percentile(d, 10):6.2f} p50={np.median(d):6.2f} p90={np.percentile(d, 90):6.2f} max={d.max():6.2f}"
    print(line)

# per chunk index inside the tile (slot 12 = chunk + 1, slot 13 = tile + 1): where the tail of the kernel comes from
if t.shape[1] > 13 and (t[:, 12] > 0).any():
    print("chunk: CTAs | end of phase, median over the tiles (us): located hashed counted published acquired classified scattered crossing synced | slowest end")
    for c in sorted(set(int(v) for v in t[:, 12] if v > 0)):
        rows = t[t[:, 12] == c]
        ends = [(np.median((rows[:, i][rows[:, i] > 0] - t0) / 1e3) if (rows[:, i] > 0).any() else float("nan")) for i in range(1, 10)]
        last = max(((rows[:, i][rows[:, i] > 0] - t0) / 1e3).max() for i in range(1, 10) if (rows[:, i] > 0).any())
        print(f"  c={c - 1:3d} n={len(rows):3d} | " + " ".join(f"{e:6.2f}" for e in ends) + f" | {last:6.2f}")
