#!/usr/bin/env python
"""One small call of each compile-time mode of the tensor-core PFN (rows, NCHW, tokens) -- a liveness + parity probe for
kernel variants (run under `timeout`; P3P_LIB selects the build)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixelspointspolygons_b200 import PointPillarsEncoder, default_cfg
from tools.synth import synth_tile, synth_weights
dev = torch.device("cuda:0")
enc = PointPillarsEncoder(default_cfg(device="cuda:0"), voxel_encoder={"in_channels": 3, "feat_channels": [64, 384]},
                          scatter={"in_channels": 384, "output_shape": [28, 28]}).to(dev).eval()
enc.load_state_dict(synth_weights(0)[0])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
x = torch.from_numpy(np.stack([synth_tile(20000, 10 + i, clustered=(i % 2 == 1)) for i in range(B)])).to(dev)
ref = torch.empty(B, 784, 384, device=dev)
enc.encode_into(x, ref, 1, precision="fp32")
torch.cuda.synchronize(); print("fp32 reference done", flush=True)
for prec in ("fp16", "tf32"):
    out = torch.empty(B, 784, 384, device=dev)
    enc.encode_into(x, out, 1, precision=prec); torch.cuda.synchronize()
    print(prec, "rows  err", ((out - ref).abs().max() / ref.abs().max()).item(), flush=True)
    o2 = torch.empty(B, 384, 28, 28, device=dev)
    enc.encode_into(x, o2, 0, c_total=384, c_offset=0, precision=prec); torch.cuda.synchronize()
    print(prec, "nchw  err", ((o2.flatten(2).transpose(1, 2) - ref).abs().max() / ref.abs().max()).item(), flush=True)
print("probe OK")
