"""Two-GPU test of the training step under the reference's wrapping: `DistributedDataParallel` over
`SyncBatchNorm.convert_sync_batchnorm(model)` (R:pixelspointspolygons/models/pix2poly/model_pix2poly.py:326-328), NCCL.
Each rank encodes its half of the batch; the ranks exchange the packed BatchNorm sums of pixelspointspolygons_b200/train.py
(four small all-reduces per step) and DDP averages the parameter gradients.  The result must equal the single-process
whole-batch step: same output rows, same running statistics, gradients = whole-batch gradients / world size.
Skipped on boxes with one GPU (run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(dev):
    from oracle import pillars_oracle as po
    from pixelspointspolygons_b200 import PointPillarsEncoder, default_cfg

    enc = PointPillarsEncoder(default_cfg(device=str(dev)), voxel_encoder={"in_channels": 3, "feat_channels": [64, 384]},
                              scatter={"in_channels": 384, "output_shape": [28, 28]}).to(dev)
    enc.load_state_dict(po.synth_weights(4)[0])
    return enc.train()


def _tiles():
    from oracle import pillars_oracle as po

    return [po.synth_tile(20000, 41), po.synth_tile(3000, 42), po.synth_tile(9000, 43, clustered=True), po.synth_tile(500, 44)]


def _nested(tiles, dev):
    return torch.nested.nested_tensor([torch.from_numpy(np.ascontiguousarray(t)) for t in tiles], layout=torch.jagged).to(dev)


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(_build(dev))
        assert isinstance(model.voxel_encoder.pfn_layers[0].norm, torch.nn.SyncBatchNorm)
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[rank])
        tiles = _tiles()
        per = len(tiles) // world
        g = torch.Generator().manual_seed(3)
        weight = torch.randn(len(tiles), 784, 384, generator=g)
        out = ddp(_nested(tiles[rank * per:(rank + 1) * per], dev))
        (out * weight[rank * per:(rank + 1) * per].to(dev)).sum().backward()
        torch.cuda.synchronize()
        outs = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(outs, out.detach())
        if rank == 0:
            torch.save({"out": torch.cat(outs).cpu(), "grads": {n: p.grad.cpu() for n, p in model.named_parameters()},
                        "bufs": {n: b.cpu() for n, b in model.named_buffers()}}, out_path)
    finally:
        dist.destroy_process_group()


def test_ddp_syncbn_training_step_equals_whole_batch(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    path = str(tmp_path / "ddp.pt")
    mp.spawn(_worker, args=(world, _free_port(), path), nprocs=world, join=True)
    got = torch.load(path)

    dev = torch.device("cuda", 0)
    whole = _build(dev)
    tiles = _tiles()
    g = torch.Generator().manual_seed(3)
    weight = torch.randn(len(tiles), 784, 384, generator=g).to(dev)
    out = whole(_nested(tiles, dev))
    (out * weight).sum().backward()

    def rel(a, b):
        return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)

    assert rel(got["out"], out.detach().cpu()) <= 1e-5
    for n, p in whole.named_parameters():
        assert rel(got["grads"][n] * world, p.grad.cpu()) <= 1e-4, n   # DDP averages over the ranks
    for n, b in whole.named_buffers():
        if b.dtype.is_floating_point:
            assert torch.allclose(got["bufs"][n].double(), b.cpu().double(), rtol=1e-5, atol=1e-8), n
