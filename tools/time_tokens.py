#!/usr/bin/env python
"""Device time of PointPillarsEncoder.forward_tokens vs forward (B = 16, N = 100k) -- usage (GPU box): python tools/time_tokens.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth as po  # (neutral generators: measurement tools do not touch oracle/)
from pixelspointspolygons_b200 import PointPillarsEncoder, default_cfg
dev = torch.device("cuda:0")
enc = PointPillarsEncoder(default_cfg(device="cuda:0"), voxel_encoder={"in_channels": 3, "feat_channels": [64, 384]},
                          scatter={"in_channels": 384, "output_shape": [28, 28]}).to(dev).eval()
enc.load_state_dict(po.synth_weights(0)[0])
x = torch.from_numpy(np.stack([po.synth_tile(100_000, 1000 + i, clustered=(i % 2 == 1)) for i in range(16)])).to(dev)
cls, pos = torch.randn(1, 1, 384, device=dev), torch.randn(1, 785, 384, device=dev)
def timeit(f, n=100):
    for _ in range(10): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
with torch.no_grad():
    print("forward (B, 784, C) rows: %.1f us/step" % timeit(lambda: enc(x)))
    print("forward_tokens (B, 785, C): %.1f us/step" % timeit(lambda: enc.forward_tokens(x, cls, pos)))
    print("forward + cat + add in torch: %.1f us/step" % timeit(lambda: torch.cat([cls.expand(16, -1, -1), enc(x)], 1) + pos))
