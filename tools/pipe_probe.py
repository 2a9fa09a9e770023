"""Where the HostPipeline's step time goes: compute-only graphs, copy-only, the shipped graphs, and a two-stream variant."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pixelspointspolygons_b200 import HostPipeline, PointPillarsEncoder, default_cfg
from tools import synth

dev = torch.device("cuda:0")
B, N = 16, 100_000
cfg = default_cfg(device="cuda:0")
enc = PointPillarsEncoder(cfg, voxel_encoder={"in_channels": 3, "feat_channels": [64, 384]}, scatter={"in_channels": 384, "output_shape": [28, 28]}).to(dev).eval()
enc.load_state_dict(synth.synth_weights(0)[0])
pipe = HostPipeline(enc, B, B * N, slots=4, host_result=lambda y: y.reshape(B, -1).sum(dim=1))
rng = np.random.default_rng(0)
for s in pipe.slots:
    s.deltas.copy_(torch.from_numpy(rng.integers(0, 56000, (B * N, 3)).astype(np.uint16)))
    s.base.copy_(torch.tensor([[0, 0, 400000]] * B, dtype=torch.int32))
    pipe.set_tiles(s, [dict(scales=(0.001,) * 3, offsets=(0.0, 0.0, 0.0), top_left=(0.0, 0.0), height=224, width=224)] * B)

def timed(fn, n=200, stream=None):
    stream = stream or torch.cuda.current_stream()
    fn(20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(stream); fn(n); e1.record(stream); t1 = time.perf_counter(); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, (t1 - t0) / n * 1e6

def full(n):
    for _ in range(n):
        pipe.staging(); pipe.submit()
print("shipped graphs      us/step %.1f (host issue %.1f)" % timed(full, stream=pipe.stream))
def comp(n):
    with torch.cuda.stream(pipe.stream):
        for i in range(n): pipe._graphs[i % pipe.n].replay()
print("compute-only graphs us/step %.1f (host issue %.1f)" % timed(comp, stream=pipe.stream))
def copy(n):
    with torch.cuda.stream(pipe.stream):
        for i in range(n): pipe._dev[i % pipe.n].copy_(pipe._host[i % pipe.n], non_blocking=True)
print("copy-only (eager)   us/step %.1f (host issue %.1f)" % timed(copy, stream=pipe.stream))
cs = torch.cuda.Stream(dev)
copied = [torch.cuda.Event() for _ in range(pipe.n)]; used = [torch.cuda.Event() for _ in range(pipe.n)]
def two(n):
    for i in range(n):
        k, nx = i % pipe.n, (i + 1) % pipe.n
        cs.wait_event(used[nx])
        with torch.cuda.stream(cs):
            pipe._dev[nx].copy_(pipe._host[nx], non_blocking=True)
        copied[nx].record(cs)
        pipe.stream.wait_event(copied[k])
        with torch.cuda.stream(pipe.stream):
            pipe._graphs[k].replay()
        used[k].record(pipe.stream)
    pipe.stream.wait_stream(cs)
for e in used: e.record(pipe.stream)
for e in copied: e.record(cs)
print("two streams + graph us/step %.1f (host issue %.1f)" % timed(two, stream=pipe.stream))

# ---- stage by stage (eager, events) ----
d = pipe._dviews[0]
x = pipe._x[0]
def ev(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
out = pipe.slots[0].out
def graphed(fn):
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            fn()
    torch.cuda.synchronize()
    return ev(g.replay)
print("las front end   us %.1f" % graphed(lambda: pipe._fe(d["deltas"], d["base"], d["offsets"], meta=d["meta"])))
print("encode_into     us %.1f" % graphed(lambda: enc.encode_into(x, out, 1)))
print("row sums        us %.1f" % graphed(lambda: out.reshape(B, -1).sum(dim=1)))
print("whole compute   us %.1f" % graphed(lambda: pipe._compute(0)))
