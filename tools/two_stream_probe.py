"""Throughput of the LiDAR-only step with two batches in flight: two encoder instances (two workspaces), their CUDA
graphs replayed alternately on two streams, against the single-stream loop."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pixelspointspolygons_b200 import PointPillarsEncoder, default_cfg
from pixelspointspolygons_b200._lib import P3P_LAYOUT_NLC
from tools import synth

dev = torch.device("cuda:0")
B, N = int(os.environ.get("PB", 16)), int(os.environ.get("PN", 100_000))
SETS = 8
def make():
    cfg = default_cfg(device="cuda:0")
    enc = PointPillarsEncoder(cfg, voxel_encoder={"in_channels": 3, "feat_channels": [64, 384]}, scatter={"in_channels": 384, "output_shape": [28, 28]}).to(dev).eval()
    enc.load_state_dict(synth.synth_weights(0)[0])
    return enc
encs = [make(), make()]
offs = (torch.arange(B + 1, dtype=torch.int64) * N).to(dev)
vals = [torch.from_numpy(np.concatenate([synth.synth_tile(N, 1000 + 16 * s + i, clustered=(i % 2 == 1)) for i in range(B)])).to(dev) for s in range(SETS)]
xs = [torch.nested.nested_tensor_from_jagged(v, offs) for v in vals]
outs = [torch.empty(B, 784, 384, device=dev) for _ in range(SETS)]
streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
graphs = [[None] * SETS for _ in range(2)]
for k in range(2):
    with torch.cuda.stream(streams[k]):
        for s in range(SETS):
            encs[k].encode_into(xs[s], outs[s], P3P_LAYOUT_NLC)
        torch.cuda.synchronize()
        for s in range(SETS):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=streams[k]):
                encs[k].encode_into(xs[s], outs[s], P3P_LAYOUT_NLC)
            graphs[k][s] = g
torch.cuda.synchronize()
def run(n, two):
    for i in range(n):
        k = (i & 1) if two else 0
        with torch.cuda.stream(streams[k]):
            graphs[k][i % SETS].replay()
def timed(two, n=2000):
    run(100, two); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    if two: streams[1].wait_event(e0)
    run(n, two)
    if two: streams[0].wait_stream(streams[1])
    e1.record(streams[0]); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for rep in range(2):
    print("B=%d N=%d  one stream us/step %.2f   two streams us/step %.2f" % (B, N, timed(False), timed(True)))
