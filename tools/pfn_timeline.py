#!/usr/bin/env python
"""Pipeline timeline of pfn_tc_kernel (library built with P3P_EXTRA_NVCC_FLAGS=-DP3P_TIMELINE).
usage (GPU box): P3P_EXTRA_NVCC_FLAGS=-DP3P_TIMELINE python -m pixelspointspolygons_b200.build --force && python tools/pfn_timeline.py [precision]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth as po  # (neutral generators: measurement tools do not touch oracle/)
from pixelspointspolygons_b200 import PointPillarsEncoder, _lib, default_cfg

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
B, N = 16, 100_000
dev = torch.device("cuda:0")
cfg = default_cfg(device="cuda:0", p3p_precision=prec)
enc = PointPillarsEncoder(cfg, voxel_encoder={"in_channels": 3, "feat_channels": [64, 384]},
                          scatter={"in_channels": 384, "output_shape": [28, 28]}).to(dev).eval()
enc.load_state_dict(po.synth_weights(0)[0])
tiles = [po.synth_tile(N, 1000 + i, clustered=(i % 2 == 1)) for i in range(B)]
x = torch.from_numpy(np.stack(tiles)).to(dev)
out = torch.empty(B, 784, 384, device=dev)
for _ in range(5):
    enc.encode_into(x, out, 1)
torch.cuda.synchronize()
CT, R, IT, S = 4, 12, 96, 4
buf = (C.c_longlong * (CT * R * IT * S))()
l = _lib.lib()
assert l.p3p_debug_pfn_timeline(buf) == 0
t = np.frombuffer(buf, dtype=np.int64).reshape(CT, R, IT, S)
for cta in range(2):
    a = t[cta]
    t0 = a[a > 0].min()
    a = np.where(a > 0, a - t0, -1)
    print(f"==== CTA {cta} ({prec}); cycles relative to the first stamp")
    print("MMA warp 0: pair: h_full_ok, t_empty[0]_ok, issued")
    for p in list(range(0, 8)) + list(range(36, 44)):
        print(f"  p={p:3d} ", a[0, p, :3].tolist())
    print("front end warp 0 / 1: unit: loop_top, h_empty_ok, computed")
    for j in list(range(0, 3)) + list(range(8, 11)):
        print(f"  j={j:3d} ", a[1, j, :3].tolist(), a[2, j, :3].tolist())
    print("epilogue group 0 / 1 / 2 lead warp: pillar: wait_begin, t_full_ok, drained")
    for gp in list(range(0, 10)) + list(range(76, 88)):
        print(f"  q={gp:3d} ", a[9, gp, :3].tolist(), a[10, gp, :3].tolist(), a[11, gp, :3].tolist())
    mm = a[0, :, 0]
    ok = mm > 0
    d = np.diff(mm[ok])
    print("MMA pair period: median", np.median(d), "p10", np.percentile(d, 10), "p90", np.percentile(d, 90), "pairs", ok.sum(), "last", mm[ok].max())
    wt = a[0, :, 1] - a[0, :, 0]
    iss = a[0, :, 2] - a[0, :, 1]
    print("MMA warp 0: wait t_empty[0] median", np.median(wt[ok]), " issue (both stages incl. t_empty[1] wait) median", np.median(iss[ok]))
    for r in range(1, 9):
        c = a[r, :, 2] - a[r, :, 1]
        w = a[r, :, 1] - a[r, :, 0]
        okr = a[r, :, 2] > 0
        gap = a[r, 1:, 0] - a[r, :-1, 2]
        print(f"front warp {r-1}: compute median {np.median(c[okr]):.0f}  wait h_empty median {np.median(w[okr]):.0f}  computed->next top median {np.median(gap[okr[1:]]):.0f}  iters {okr.sum()}")
    for r in range(9, 12):
        okr = a[r, :, 2] > 0
        w = a[r, :, 1] - a[r, :, 0]
        ld = a[r, :, 2] - a[r, :, 1]
        per = np.diff(a[r, :, 2][okr])
        print(f"epi group {r-9}: wait t_full median {np.median(w[okr]):.0f}  drain median {np.median(ld[okr]):.0f}  pillar period median {np.median(per):.0f}  pillars {okr.sum()}  last {a[r, :, 2].max()}")
