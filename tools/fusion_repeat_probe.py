#!/usr/bin/env python
"""Run the early-fusion front end several times on the same inputs and report bitwise differences per half (image / LiDAR)
between the calls -- a determinism probe (meant to run under compute-sanitizer, whose serialisation changes the timing)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixelspointspolygons_b200 import default_cfg
from pixelspointspolygons_b200.fusion import EarlyFusionFrontEnd
from tools.synth import synth_tile, synth_weights
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
cfg = default_cfg(device="cuda:0", p3p_precision=prec)
fe = EarlyFusionFrontEnd(cfg).to(dev).eval()
fe.image_embed.precision = prec
sd, sdi = synth_weights(13)
fe.lidar_embed.load_state_dict(sd); fe.image_embed.load_state_dict(sdi)
tiles = [synth_tile(20000, 41), synth_tile(500, 42, clustered=True), np.zeros((0, 3), np.float32)]
x = torch.nested.nested_tensor([torch.from_numpy(t) for t in tiles], layout=torch.jagged).to(dev)
img = torch.rand(3, 3, 224, 224, generator=torch.Generator().manual_seed(7)).to(dev)
outs = []
with torch.no_grad():
    for i in range(4):
        if i == 2:
            fe.cfg.experiment.lidar_dropout = 0.0
        outs.append(fe(img, x).clone())
        torch.cuda.synchronize()
for i in range(1, 4):
    d = (outs[i] != outs[0])
    print(f"call {i} vs 0: image half differs in {int(d[:, :384].sum())} elements, lidar half in {int(d[:, 384:].sum())}; max |diff| image "
          f"{(outs[i][:, :384] - outs[0][:, :384]).abs().max().item():.3e} lidar {(outs[i][:, 384:] - outs[0][:, 384:]).abs().max().item():.3e}")
    if d.any():
        idx = d.nonzero()[:5].tolist()
        print("   first differing (b, c, y, x):", idx)
