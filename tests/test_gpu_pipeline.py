"""HostPipeline (pinned staging slots -> one CUDA-graph launch per batch) against the direct calls it replaces:
`las_packed_to_pixels` + the module on the same batch, bit for bit, and against the CPU oracle's loader arithmetic +
encoder within the fp32 contract.  Several ring cycles, jagged splits that change from batch to batch."""
import numpy as np
import pytest
import torch

from oracle import pillars_oracle as po
from pixelspointspolygons_b200 import (EarlyFusionFrontEnd, HostPipeline, PointPillarsEncoder, default_cfg, las_packed_to_pixels,
                                       pack_las)
from tools import synth

pytestmark = pytest.mark.gpu

B, TOTAL = 4, 60_000


def las_batch(seed):
    """A jagged batch of raw LAS integers (1 mm scale, 56 m tiles): per-tile packed deltas / base / header, and the split."""
    rng = np.random.default_rng(seed)
    cuts = np.sort(rng.integers(1, TOTAL, B - 1))
    lens = np.diff(np.concatenate([[0], cuts, [TOTAL]]))
    deltas, bases, metas, xyz = [], [], [], []
    for i, n in enumerate(lens):
        left, top = 2_600_000.0 + 56.0 * i + seed, 1_200_000.0 - 56.0 * i
        X = rng.integers(0, 56_000, n).astype(np.int32)
        Y = rng.integers(0, 56_000, n).astype(np.int32)
        Z = rng.integers(400_000, 430_000, n).astype(np.int32)
        d, b = pack_las(X, Y, Z)
        deltas.append(d); bases.append(b); xyz.append((X, Y, Z))
        metas.append(dict(scales=(0.001, 0.001, 0.001), offsets=(left, top, 0.0), top_left=(left, top), height=224, width=224))
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return np.concatenate(deltas), np.stack(bases), metas, offs, xyz


def fill(pipe, slot, batch, image=None):
    d, b, metas, offs, _ = batch
    slot.deltas.copy_(torch.from_numpy(d))
    slot.base.copy_(torch.from_numpy(b))
    slot.offsets.copy_(torch.from_numpy(offs))
    pipe.set_tiles(slot, metas)
    if image is not None:
        slot.image.copy_(image)


def encoder(dev):
    cfg = default_cfg(device=str(dev))
    enc = PointPillarsEncoder(cfg, voxel_encoder={"in_channels": 3, "feat_channels": [64, 384]},
                              scatter={"in_channels": 384, "output_shape": [28, 28]}).to(dev).eval()
    sd, sdi = synth.synth_weights(3)
    enc.load_state_dict(sd)
    return cfg, enc, sd, sdi


def direct_points(batch, dev):
    d, b, metas, offs, _ = batch
    offs_t = torch.from_numpy(offs).to(dev)
    vals = las_packed_to_pixels(torch.from_numpy(d).to(dev), torch.from_numpy(b).to(dev), offs_t, metas)
    return torch.nested.nested_tensor_from_jagged(vals, offs_t)


def test_lidar_pipeline_equals_direct_calls_and_oracle(cuda_device):
    cfg, enc, sd, _ = encoder(cuda_device)
    pipe = HostPipeline(enc, B, TOTAL, slots=3, host_result=lambda y: y.reshape(B, -1).sum(dim=1))
    batches = [las_batch(s) for s in range(7)]  # more batches than slots: the ring wraps twice
    results = []
    for batch in batches:
        fill(pipe, pipe.staging(), batch)
        prev = pipe.submit()
        if prev is not None:
            prev.done.synchronize()
            results.append((prev.out.clone(), prev.host_value.clone()))
    last = pipe.flush()
    last.done.synchronize()
    results.append((last.out.clone(), last.host_value.clone()))
    assert len(results) == len(batches) and pipe.launches == len(batches)
    for batch, (out, hv) in zip(batches, results):
        want = enc(direct_points(batch, cuda_device), return_flattened=True)
        assert torch.equal(out, want)
        assert torch.equal(hv, want.reshape(B, -1).sum(dim=1).cpu())
    # the oracle's loader arithmetic + encoder on the first batch (fp32 contract of the fp16-operand path: 1e-3 of scale)
    d, b, metas, offs, xyz = batches[0]
    ref_enc = po.OraclePointPillarsEncoder(po.GridSpec()).eval()
    ref_enc.load_state_dict(sd)
    tiles = [po.las_points_to_pixels(X, Y, Z, m["scales"], m["offsets"], top_left=m["top_left"], variant="dataset")
             for (X, Y, Z), m in zip(xyz, metas)]
    with torch.no_grad():
        ref = ref_enc([torch.from_numpy(t) for t in tiles], return_flattened=True)
    got = results[0][0].cpu()
    assert (got - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()


@pytest.mark.parametrize("mode", ["concat", "tokens"])
def test_fusion_pipeline_equals_direct_calls(cuda_device, mode):
    cfg, _, sd, sdi = encoder(cuda_device)
    fus = EarlyFusionFrontEnd(cfg).to(cuda_device).eval()
    fus.lidar_embed.load_state_dict(sd)
    fus.image_embed.load_state_dict(sdi)
    pipe = HostPipeline(fus, B, TOTAL, slots=2, mode=mode, host_result="full")
    g = torch.Generator().manual_seed(5)
    batches = [(las_batch(10 + s), torch.rand(B, 3, 224, 224, generator=g)) for s in range(5)]
    done = []
    for batch, img in batches:
        fill(pipe, pipe.staging(), batch, img)
        prev = pipe.submit()
        if prev is not None:
            prev.done.synchronize()
            done.append(prev.out.clone())
    last = pipe.flush()
    last.host_done.synchronize()
    done.append(last.out.clone())
    assert torch.equal(last.host_out, last.out.cpu())  # the D2H branch delivers the same bytes
    for (batch, img), out in zip(batches, done):
        x = direct_points(batch, cuda_device)
        if mode == "concat":
            want = fus(img.to(cuda_device), x)
        else:
            want = fus.forward_tokens(img.to(cuda_device), x, lidar_zero=False)
        assert torch.equal(out, want)


def test_pipeline_refuses_cpu_modules_and_train_mode(cuda_device):
    cfg, enc, _, _ = encoder(cuda_device)
    with pytest.raises(RuntimeError, match="eval"):
        HostPipeline(enc.train(), B, TOTAL)
    enc.eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        HostPipeline(enc.cpu(), B, TOTAL)


def test_batches_in_flight_on_two_lanes_equal_the_serial_result(cuda_device):
    """Two streams, two workspaces (`lane`): many alternating calls with different batches, every output equal to the one
    a plain single-stream call gives."""
    from pixelspointspolygons_b200._lib import P3P_LAYOUT_NLC

    cfg, enc, _, _ = encoder(cuda_device)
    xs = [direct_points(las_batch(30 + s), cuda_device) for s in range(6)]
    want = [enc(x, return_flattened=True).clone() for x in xs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(cuda_device), torch.cuda.Stream(cuda_device)]
    outs = [torch.empty_like(w) for w in want]
    for rep in range(4):
        for i, x in enumerate(xs):
            with torch.cuda.stream(streams[i % 2]):
                enc.encode_into(x, outs[i], P3P_LAYOUT_NLC, lane=i % 2)
    torch.cuda.synchronize()
    for o, w in zip(outs, want):
        assert torch.equal(o, w)
