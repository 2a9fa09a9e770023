"""Randomised parity of the integer surface: seeded random grids, limits and point sets -- the
cases nobody wrote down -- voxelized on the GPU through the C ABI and compared bit for bit with the CPU oracle: per-point
hash, voxel order and coordinates, counts, the kept point indices of every pillar, gathered points, canvas owners.

Each draw mixes uniform points, tight blobs (pillars far beyond M), points snapped onto cell borders and onto the range's
max faces (hash aliasing, Appendix A.1), exact duplicates, and out-of-range / NaN / Inf rows; grids other than the shipped
one (4 / 8 / 16-px cells, 96 .. 224-px tiles; at most 6144 pillar keys per tile), M from 1 to 100, `max_voxels` cuts, both readings of ledger U1."""
import numpy as np
import pytest
import torch

from oracle import pillars_oracle as po
from pixelspointspolygons_b200 import PointPillarsEncoder, default_cfg

pytestmark = pytest.mark.gpu

F = np.float32
SHAPES = [(224, 8.0), (224, 16.0), (160, 4.0), (112, 8.0), (96, 8.0), (128, 16.0)]


def draw_case(seed):
    rng = np.random.default_rng(1000 + seed)
    size, v = SHAPES[rng.integers(len(SHAPES))]
    n_side = int(size // v)
    M = int(rng.choice([1, 3, 8, 64, 100]))
    cells = n_side * n_side
    vmax = int(rng.choice([cells, cells, max(1, cells // 3), 5]))
    grid = po.GridSpec(in_width=float(size), in_height=float(size), voxel_size=(v, v, 100.0), max_num_points=M,
                       max_voxels=(vmax, max(1, vmax - 1)), output_shape=(n_side, n_side), drop_overflow=bool(rng.integers(2)))
    tiles = []
    for _ in range(int(rng.integers(1, 5))):
        n = int(rng.choice([0, 1, 7, 500, 6000, 40000]))
        pts = np.empty((n, 3), F)
        pts[:, :2] = rng.uniform(0.0, size, (n, 2))
        pts[:, 2] = rng.uniform(0.0, 100.0, n)
        if n:
            k = n // 4                                                    # blobs: a few pillars far beyond M
            c = rng.uniform(v, size - v, (3, 2))
            pts[:k, :2] = c[rng.integers(0, 3, k)] + rng.uniform(-0.4 * v, 0.4 * v, (k, 2))
            s = rng.integers(0, n, n // 10)                               # snapped onto cell borders
            pts[s, :2] = np.round(pts[s, :2] / v) * v
            f = rng.integers(0, n, max(1, n // 50))                       # on the max faces: x == size, y == size, z == 100
            which = rng.integers(0, 3, len(f))
            pts[f[which == 0], 0] = size
            pts[f[which == 1], 1] = size
            pts[f[which == 2], 2] = 100.0
            d = rng.integers(0, n, n // 20)                               # exact duplicates of other points
            pts[d] = pts[rng.integers(0, n, len(d))]
            b = rng.integers(0, n, n // 40)                               # rows the range filter must drop
            bad = np.array([-1e-3, size + 1e-3, np.nan, np.inf, -np.inf], F)
            pts[b, rng.integers(0, 3, len(b))] = bad[rng.integers(0, len(bad), len(b))]
            pts = pts[rng.permutation(n)]
        tiles.append(np.ascontiguousarray(pts))
    return grid, tiles, n_side


def build(dev, grid, n_side):
    cfg = default_cfg(device=str(dev), in_size=int(grid.in_width), voxel=grid.voxel_size, max_num_points_per_voxel=grid.max_num_points,
                      max_num_voxels=grid.max_voxels, patch_size=int(grid.voxel_size[0]), p3p_drop_overflow=grid.drop_overflow)
    enc = PointPillarsEncoder(cfg, voxel_encoder={"in_channels": 3, "feat_channels": [64, 384]},
                              scatter={"in_channels": 384, "output_shape": [n_side, n_side]}).to(dev).eval()
    sd, _ = po.synth_weights(1)
    enc.load_state_dict(sd)
    ref = po.OraclePointPillarsEncoder(grid).eval()
    ref.load_state_dict(sd)
    return enc, ref


@pytest.mark.parametrize("seed", range(24))
def test_voxelizer_fuzz_is_bit_exact(cuda_device, seed):
    grid, tiles, n_side = draw_case(seed)
    enc, ref = build(cuda_device, grid, n_side)
    x = torch.nested.nested_tensor([torch.from_numpy(t) for t in tiles], layout=torch.jagged).to(cuda_device)
    for training in (False, True):
        enc.train(training)
        ref.train(training)
        raw = enc.voxelize_raw(x)
        gv, gn, gc = enc.voxelize(x)
        rv, rn, rc, rd = ref.voxelize(tiles)
        assert torch.equal(gc.cpu(), rc) and torch.equal(gn.cpu(), rn)
        assert torch.equal(gv.cpu().view(torch.int32), rv.view(torch.int32))       # bit patterns: NaN-safe
        hashes = np.concatenate([po.voxelize_c(t, grid, 1)["point_hash"] for t in tiles]) if sum(len(t) for t in tiles) else np.zeros(0)
        assert np.array_equal(raw["point_hash"].cpu().numpy(), hashes)
        V = raw["pillar_coords"].shape[1]
        mask = torch.arange(V).view(1, -1) < raw["num_pillars"].cpu().view(-1, 1)
        assert torch.equal(raw["pillar_point_idx"].cpu()[mask].long(), rd)
        owner = torch.full((len(tiles), n_side * n_side), -1, dtype=torch.int32)
        counts = [int((rc[:, 0] == b).sum()) for b in range(len(tiles))]
        start = 0
        for b, cnt in enumerate(counts):
            cb = rc[start:start + cnt]
            for r in range(cnt):
                owner[b, cb[r, 2] * n_side + cb[r, 3]] = r
            start += cnt
        assert torch.equal(raw["cell_owner"].cpu(), owner)
        assert raw["num_pillars"].cpu().tolist() == counts
    # and the features on top of it (eval mode, exact-fp32 route and the default tensor-core one)
    enc.eval(); ref.eval()
    with torch.no_grad():
        r = ref(tiles, return_flattened=True)
        scale = max(r.abs().max().item(), 1e-6)
        for prec, tol in (("fp32", 1e-4), ("fp16", 1e-3)):
            out = torch.empty(len(tiles), n_side * n_side, 384, device=cuda_device)
            enc.encode_into(x, out, 1, precision=prec)
            assert torch.isfinite(out).all()
            err = (out.cpu() - r).abs().max().item() / scale
            assert err <= tol, (seed, prec, err)
