"""GPU tests of the fused training step (csrc/pfn_train.cu, pixelspointspolygons_b200/train.py).

Checker: the dense autograd PillarFeatureNet written the way the reference has it (nn.Linear + nn.BatchNorm1d in train
mode + ReLU + max, encoder.forward_dense_reference) evaluated in float64 on the same pillars -- gradients of the six PFN
parameters, the forward output and the running-statistics update.  Tolerances: forward 1e-4 of scale (exact fp32);
gradients per tensor: 99 % of the elements within 2e-4 of the tensor's scale and rms error <= 1e-3 of its rms, maximum
<= 2e-2 -- the loss is not smooth where two rows tie for a maximum or a relu input crosses zero, and an fp32 and an fp64
evaluation fall on different sides of a handful of those among the ~10^6 (pillar, channel) pairs of the larger cases
(torch's own fp32 dense step differs from its fp64 one by the same amount: tools/time_train.py reports both)."""
import copy
import threading

import numpy as np
import pytest
import torch

from oracle import pillars_oracle as po
from pixelspointspolygons_b200 import PointPillarsEncoder, default_cfg
from pixelspointspolygons_b200 import train as p3p_train

pytestmark = pytest.mark.gpu

PARAMS = ("pfn_layers.0.linear.weight", "pfn_layers.0.norm.weight", "pfn_layers.0.norm.bias",
          "pfn_layers.1.linear.weight", "pfn_layers.1.norm.weight", "pfn_layers.1.norm.bias")


def build(dev, M=64, C=384, seed=0, center_alias=True):
    cfg = default_cfg(device=str(dev), max_num_points_per_voxel=M, patch_feature_dim=C, p3p_center_alias=center_alias)
    enc = PointPillarsEncoder(cfg, voxel_encoder={"in_channels": 3, "feat_channels": [64, C]},
                              scatter={"in_channels": C, "output_shape": [28, 28]}).to(dev)
    sd, _ = po.synth_weights(seed, feat_channels=(64, C))
    enc.load_state_dict(sd)
    return enc.train()


def nested(tiles, dev):
    return torch.nested.nested_tensor([torch.from_numpy(np.ascontiguousarray(t)) for t in tiles], layout=torch.jagged).to(dev)


def rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)


def dense_step(enc, x, weight, dtype=torch.float64):
    """dense autograd step on a copy of the module (float64: the checker): (out, grads by name, running stats by name)."""
    ve = copy.deepcopy(enc.voxel_encoder).to(dtype).train()
    out = enc.forward_dense_reference(x, True, voxel_encoder=ve)
    (out * weight.to(dtype)).sum().backward()
    grads = {n: dict(ve.named_parameters())[n].grad for n in PARAMS}
    bufs = {n: b.clone() for n, b in ve.named_buffers()}
    return out.detach(), grads, bufs


def grad_errors(got, ref):
    got, ref = got.double().flatten(), ref.double().flatten()
    err = (got - ref).abs() / max(ref.abs().max().item(), 1e-30)
    q99 = torch.quantile(err[:: max(1, err.numel() // 1_000_000)], 0.99).item()
    rms = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-30)).item()
    return q99, rms, err.max().item()


def assert_grad_close(got, ref, torch_fp32, what):
    """Strict bound, or -- where fp32 and fp64 fall on different sides of a tie / a relu zero -- no worse than twice what
    torch's own fp32 dense autograd step loses against the same fp64 step."""
    q99, rms, mx = grad_errors(got, ref)
    if q99 <= 2e-4 and rms <= 1e-3 and mx <= 2e-2:
        return
    tq, tr, tm = grad_errors(torch_fp32, ref)
    assert mx <= 2e-2 and q99 <= 2 * tq and rms <= 2 * tr, (what, (q99, rms, mx), (tq, tr, tm))


@pytest.mark.parametrize("case", ["sparse", "dense", "m16", "m512", "no_alias", "c128"])
def test_fused_training_step_matches_dense_autograd(cuda_device, case):
    M, C, alias = 64, 384, True
    tiles = [po.synth_tile(3000, 7), po.synth_tile(800, 8)]
    if case == "dense":
        tiles = [po.synth_tile(40000, 3, clustered=True), po.synth_tile(20000, 4)]
    if case == "m16":
        M, tiles = 16, [po.synth_tile(9000, 5), po.synth_tile(500, 6), po.synth_tile(20, 9)]
    if case == "m512":  # the far end of the reference's ablation axis (R:config/experiment/lidar_density_ablation512.yaml)
        M, tiles = 512, [po.synth_tile(30000, 5, clustered=True), po.synth_tile(700, 6)]
    if case == "no_alias":
        alias = False
    if case == "c128":
        C = 128
    enc = build(cuda_device, M=M, C=C, seed=21, center_alias=alias)
    x = nested(tiles, cuda_device)
    g = torch.Generator().manual_seed(17)
    weight = torch.randn(len(tiles), 784, C, generator=g).to(cuda_device)
    ref_out, ref_grads, ref_bufs = dense_step(enc, x, weight)
    _, t32_grads, _ = dense_step(enc, x, weight, torch.float32)

    out = enc(x)
    assert out.shape == (len(tiles), 784, C) and out.requires_grad
    (out * weight).sum().backward()
    assert rel(out.detach(), ref_out) <= 1e-4, ("forward", rel(out.detach(), ref_out))
    named = dict(enc.voxel_encoder.named_parameters())
    for n in PARAMS:
        assert named[n].grad is not None, n
        assert_grad_close(named[n].grad, ref_grads[n], t32_grads[n], (case, n))
    for n, b in enc.voxel_encoder.named_buffers():
        if b.dtype.is_floating_point:
            assert torch.allclose(b.double(), ref_bufs[n], rtol=1e-4, atol=1e-6), n
        else:
            assert int(b) == int(ref_bufs[n]), n


def test_training_nchw_view_and_second_step(cuda_device):
    """return_flattened=False goes through autograd's transpose; a second step accumulates into .grad like any module."""
    enc = build(cuda_device, seed=5)
    x = nested([po.synth_tile(5000, 1), po.synth_tile(5000, 2)], cuda_device)
    out = enc(x, return_flattened=False)
    assert out.shape == (2, 384, 28, 28)
    out.sum().backward()
    g1 = enc.voxel_encoder.pfn_layers[1].linear.weight.grad.clone()
    rm1 = enc.voxel_encoder.pfn_layers[1].norm.running_mean.clone()
    enc(x, return_flattened=False).sum().backward()
    assert rel(enc.voxel_encoder.pfn_layers[1].linear.weight.grad, 2 * g1) <= 1e-5  # (fp32 shared-memory atomics: order varies)
    assert int(enc.voxel_encoder.pfn_layers[1].norm.num_batches_tracked) == 2
    assert not torch.equal(enc.voxel_encoder.pfn_layers[1].norm.running_mean, rm1)
    with torch.no_grad():  # train mode without autograd: batch statistics, no graph
        o = enc(x)
    assert not o.requires_grad
    enc.zero_grad()
    out = enc(x)
    out.sum().backward(retain_graph=True)  # the node may run twice
    g_once = enc.voxel_encoder.pfn_layers[0].linear.weight.grad.clone()
    out.sum().backward()
    assert rel(enc.voxel_encoder.pfn_layers[0].linear.weight.grad, 2 * g_once) <= 1e-5
    enc.voxel_encoder.pfn_layers[0].norm.eval()
    with pytest.raises(NotImplementedError):
        enc(x)


def test_training_two_simulated_ranks_match_whole_batch(cuda_device):
    """SyncBatchNorm protocol: two 'ranks' (threads on one GPU, each with half of the batch) exchange the four packed
    sums through a reducer; their outputs concatenate to the whole-batch output and their parameter gradients add up to
    the whole-batch gradients (DDP's all-reduce would then average them)."""
    tiles = [po.synth_tile(6000, 31), po.synth_tile(1500, 32), po.synth_tile(12000, 33, clustered=True), po.synth_tile(300, 34)]
    g = torch.Generator().manual_seed(3)
    weight = torch.randn(4, 784, 384, generator=g).to(cuda_device)

    whole = build(cuda_device, seed=9)
    out_w = whole(nested(tiles, cuda_device))
    (out_w * weight).sum().backward()
    torch.cuda.synchronize()

    ranks = [build(cuda_device, seed=9), build(cuda_device, seed=9)]
    barrier = threading.Barrier(2)
    slots = {}
    lock = threading.Lock()

    def reducer(rank):
        def red(t, what):
            torch.cuda.synchronize()
            with lock:
                slots[(what, rank)] = t.clone()
            barrier.wait()
            total = slots[(what, 0)] + slots[(what, 1)]
            barrier.wait()
            t.copy_(total)
        return red

    results, errors = {}, []

    def run(rank):
        try:
            with torch.cuda.device(cuda_device):
                enc = ranks[rank]
                x = nested(tiles[2 * rank:2 * rank + 2], cuda_device)
                values, offsets, B = enc._pack(x)
                tensors = [dict(enc.voxel_encoder.named_parameters())[n] for n in PARAMS]
                out, ctx = p3p_train.train_forward(enc, values, offsets, B, tensors, reduce_fn=reducer(rank))
                grads = p3p_train.train_backward(ctx, weight[2 * rank:2 * rank + 2], reduce_fn=reducer(rank))
                torch.cuda.synchronize()
                results[rank] = (out, grads)
        except Exception as e:  # pragma: no cover
            errors.append(e)
            barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    out_s = torch.cat([results[0][0], results[1][0]], 0)
    assert rel(out_s, out_w.detach()) <= 1e-5
    named = dict(whole.voxel_encoder.named_parameters())
    for i, n in enumerate(PARAMS):
        e = rel(results[0][1][i] + results[1][1][i], named[n].grad)
        assert e <= 1e-5, (n, e)
    for (n, b), (_, b0) in zip(whole.voxel_encoder.named_buffers(), ranks[0].voxel_encoder.named_buffers()):
        assert torch.allclose(b.double(), b0.double(), rtol=1e-6, atol=1e-9), n


def test_training_rejects_unsupported_shapes(cuda_device):
    enc = build(cuda_device, M=1024, seed=1)
    x = nested([po.synth_tile(2000, 1)], cuda_device)
    with pytest.raises(Exception, match="max_points"):
        enc(x)
