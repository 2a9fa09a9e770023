"""CPU tests of the multi-GPU host logic: tiles are sharded batch-wise, every rank encodes its own tiles,
no collective on the data path.  World-size-2 (and 3) `gloo` processes stand in for the GPUs; each rank runs the CPU
oracle on its shard (the checker standing in for the kernels, which need a GPU) and the gathered result must equal
the single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pillars_oracle as po
from pixelspointspolygons_b200 import shard


def test_shard_bounds_cover_the_batch_exactly_once():
    for B in (0, 1, 7, 16, 64):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard.shard_bounds(4, 2, 2)


def test_shard_jagged_rebases_offsets_and_slices_without_copy():
    lens = [5, 0, 3, 9, 1]
    vals = torch.arange(sum(lens) * 3, dtype=torch.float32).view(-1, 3)
    offs = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int64)
    got = []
    for r in range(3):
        v, o = shard.shard_jagged(vals, offs, r, 3)
        assert o[0] == 0 and o[-1] == v.shape[0]
        assert v.numel() == 0 or v.data_ptr() == vals[int(offs[shard.shard_bounds(5, r, 3)[0]])].data_ptr()
        got.append(torch.diff(o).tolist())
    assert sum(got, []) == lens
    nt = torch.nested.nested_tensor_from_jagged(vals, offs)
    parts = [shard.shard_lidar(nt, r, 2) for r in range(2)]
    assert [p.shape[0] for p in parts] == [3, 2]
    assert torch.equal(torch.cat([p.values() for p in parts]), vals)
    dense = torch.rand(5, 7, 3)
    assert torch.equal(torch.cat([shard.shard_lidar(dense, r, 4) for r in range(4)]), dense)
    assert sum((shard.shard_lidar([1, 2, 3], r, 2) for r in range(2)), []) == [1, 2, 3]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, num_tiles, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        tiles = [po.synth_tile(400 + 37 * i, 70 + i, clustered=(i % 2 == 1)) for i in range(num_tiles)]
        tiles[1] = np.zeros((0, 3), np.float32)
        x = torch.nested.nested_tensor([torch.from_numpy(t) for t in tiles], layout=torch.jagged)
        enc = po.OraclePointPillarsEncoder(po.GridSpec()).eval()
        enc.load_state_dict(po.synth_weights(3)[0])
        mine = shard.shard_lidar(x, rank, world)
        with torch.no_grad():
            local = enc(mine, return_flattened=True).contiguous() if mine.shape[0] else torch.zeros(0, 784, 384)
        full = shard.gather_tiles(local, num_tiles)
        if rank == 0:
            with torch.no_grad():
                ref = enc(x, return_flattened=True)
            torch.save({"equal": bool(torch.equal(full, ref)), "shape": tuple(full.shape)}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,num_tiles", [(2, 5), (3, 4)])
def test_sharded_ranks_reproduce_the_single_process_batch(tmp_path, world, num_tiles):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), num_tiles, out), nprocs=world, join=True)
    res = torch.load(out)
    assert res["equal"] and res["shape"] == (num_tiles, 784, 384)
