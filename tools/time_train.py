#!/usr/bin/env python
"""Training step of the LiDAR encoder: fused kernels (forward + backward) against the dense autograd
formulation the reference runs (nn.Linear + BatchNorm1d + ReLU + max over (V, M, *) tensors), CUDA events, same inputs.
usage (GPU box): python tools/time_train.py [B] [N]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixelspointspolygons_b200 import PointPillarsEncoder, default_cfg
from tools.synth import synth_tile, synth_weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
FUSED_ONLY = len(sys.argv) > 3 and sys.argv[3] == "fused-only"  # (for an ncu launch list of the fused step)
dev = torch.device("cuda:0")
enc = PointPillarsEncoder(default_cfg(device="cuda:0"), voxel_encoder={"in_channels": 3, "feat_channels": [64, 384]},
                          scatter={"in_channels": 384, "output_shape": [28, 28]}).to(dev).train()
enc.load_state_dict(synth_weights(0)[0])
x = torch.from_numpy(np.stack([synth_tile(N, 1000 + i, clustered=(i % 2 == 1)) for i in range(B)])).to(dev)
w = torch.randn(B, 784, 384, device=dev)


def step(fn):
    enc.zero_grad(set_to_none=True)
    out = fn(x)
    (out * w).sum().backward()


def timed(fn, iters):
    for _ in range(3):
        step(fn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step(fn)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if FUSED_ONLY:
    for _ in range(3):
        step(lambda t: enc(t))
    torch.cuda.synchronize()
    sys.exit(0)
fused = timed(lambda t: enc(t), 20)
g_fused = [p.grad.clone() for p in enc.parameters()]
torch.cuda.reset_peak_memory_stats()
step(lambda t: enc(t))
torch.cuda.synchronize()
mem_fused = torch.cuda.max_memory_allocated()
dense = timed(lambda t: enc.forward_dense_reference(t), 5)
g_dense = [p.grad.clone() for p in enc.parameters()]
torch.cuda.reset_peak_memory_stats()
step(lambda t: enc.forward_dense_reference(t))
torch.cuda.synchronize()
mem_dense = torch.cuda.max_memory_allocated()
err = max(((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item() for a, b in zip(g_fused, g_dense))
# the same step in float64 (the checker of tests/test_gpu_train.py) at this size
import copy
ve = copy.deepcopy(enc.voxel_encoder).double().train()
out64 = enc.forward_dense_reference(x, True, voxel_encoder=ve)
(out64 * w.double()).sum().backward()
g64 = [p.grad for p in ve.parameters()]
err_fused64 = max(((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item() for a, b in zip(g_fused, g64))
err_dense64 = max(((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item() for a, b in zip(g_dense, g64))
print(json.dumps({"what": "training step forward + backward of the LiDAR encoder", "tiles": B, "points_per_tile": N,
                  "fused_ms": round(fused, 3), "dense_autograd_ms": round(dense, 3), "speedup": round(dense / fused, 2),
                  "fused_peak_bytes": mem_fused, "dense_peak_bytes": mem_dense,
                  "max_rel_grad_diff_fused_vs_dense_fp32": err, "max_rel_grad_err_fused_vs_float64": err_fused64,
                  "max_rel_grad_err_dense_fp32_vs_float64": err_dense64}))
