"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI (libp3p.so) via the
drop-in module and is compared with the CPU oracle on the same seeded inputs.

Bars: integer / index outputs bit exact; features within 1e-3 (fp32 contract: exact-fp32 and tf32 tensor-core
paths) or 1e-2 (bf16 contract) of the oracle, measured as max |a - b| / max |b| over the tensor, and additionally
per element with rtol = atol = tol * scale."""
import numpy as np
import pytest
import torch

import p3p_cases as cases
from oracle import pillars_oracle as po
from pixelspointspolygons_b200 import PointPillarsEncoder, default_cfg

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "tf32": 1e-3, "bf16": 1e-2}


def build(dev, grid: po.GridSpec, seed=0, C=384):
    cfg = default_cfg(device=str(dev), max_num_points_per_voxel=grid.max_num_points, max_num_voxels=grid.max_voxels,
                      patch_feature_dim=C, p3p_drop_overflow=grid.drop_overflow)
    enc = PointPillarsEncoder(cfg, voxel_encoder={"in_channels": 3, "feat_channels": [64, C]},
                              scatter={"in_channels": C, "output_shape": [28, 28]}).to(dev).eval()
    sd, _ = po.synth_weights(seed, feat_channels=(64, C))
    enc.load_state_dict(sd)
    g2 = po.GridSpec(**{**grid.__dict__, "feat_channels": (64, C)})
    ref = po.OraclePointPillarsEncoder(g2).eval()
    ref.load_state_dict(sd)
    return enc, ref


def to_nested(tiles, dev):
    return torch.nested.nested_tensor([torch.from_numpy(np.ascontiguousarray(t)) for t in tiles], layout=torch.jagged).to(dev)


def assert_close(out, ref, tol, what):
    out, ref = out.float().cpu(), ref.float()
    assert out.shape == ref.shape, (what, out.shape, ref.shape)
    if ref.numel() == 0:
        return
    scale = max(ref.abs().max().item(), 1e-6)
    err = (out - ref).abs().max().item() / scale
    assert err <= tol, f"{what}: max|err|/max|ref| = {err:.3e} > {tol}"
    assert torch.allclose(out, ref, rtol=tol, atol=tol * scale), what


@pytest.mark.parametrize("name", sorted(cases.edge_cases()))
def test_voxelizer_is_bit_exact_on_edge_cases(cuda_device, name):
    tiles, kw = cases.edge_cases()[name]
    grid = cases.grid_for(kw)
    enc, ref = build(cuda_device, grid)
    for training in (False, True):
        enc.train(training)
        ref.train(training)
        x = to_nested(tiles, cuda_device)
        raw = enc.voxelize_raw(x)
        gv, gn, gc = enc.voxelize(x)
        rv, rn, rc, rd = ref.voxelize(tiles)
        assert torch.equal(gc.cpu(), rc), name
        assert torch.equal(gn.cpu(), rn), name
        assert torch.equal(gv.cpu(), rv), name
        # per-point hash and the dense index matrix (ragged_to_dense, -1 padded)
        hashes = np.concatenate([po.voxelize_c(t, grid, 1)["point_hash"] for t in tiles]) if sum(len(t) for t in tiles) else np.zeros(0)
        assert np.array_equal(raw["point_hash"].cpu().numpy(), hashes), name
        V = raw["pillar_coords"].shape[1]
        mask = (torch.arange(V).view(1, -1) < raw["num_pillars"].cpu().view(-1, 1))
        assert torch.equal(raw["pillar_point_idx"].cpu()[mask].long(), rd), name
        # scatter indices: owner of every canvas cell = last pillar (voxel order) that maps to it
        owner = torch.full((len(tiles), 784), -1, dtype=torch.int32)
        counts = [int((rc[:, 0] == b).sum()) for b in range(len(tiles))]
        start = 0
        for b, cnt in enumerate(counts):
            cb = rc[start:start + cnt]
            for r in range(cnt):
                owner[b, cb[r, 2] * 28 + cb[r, 3]] = r
            start += cnt
        assert torch.equal(raw["cell_owner"].cpu(), owner), name
        assert raw["num_pillars"].cpu().tolist() == counts, name


@pytest.mark.parametrize("prec", ["fp32", "tf32", "bf16"])
@pytest.mark.parametrize("name", ["occupancy_M", "z100_overwrites_cell", "ragged_batch", "empty_tile_between", "demo_shaped",
                                  "stability_shuffled", "x224_then_alias", "only_empty_tiles", "vmax_cut"])
def test_features_and_canvas_match_oracle(cuda_device, name, prec):
    tiles, kw = cases.edge_cases()[name]
    grid = cases.grid_for(kw)
    enc, ref = build(cuda_device, grid, seed=3)
    x = to_nested(tiles, cuda_device)
    with torch.no_grad():
        rfeat, rcoors, _ = ref.pillar_features(tiles)
        rout = ref(tiles, return_flattened=False)
    feats, coors = enc.pillar_features(x, precision=prec)
    assert torch.equal(coors.cpu(), rcoors)
    assert_close(feats, rfeat, TOL[prec], f"{name}/{prec}/pillar features")
    enc.precision = prec
    with torch.no_grad():
        nchw = enc(x, return_flattened=False)
        nlc = enc(x, return_flattened=True)
    assert nchw.shape == (len(tiles), 384, 28, 28) and nlc.shape == (len(tiles), 784, 384)
    assert_close(nchw, rout, TOL[prec], f"{name}/{prec}/NCHW")
    assert torch.equal(nlc.transpose(1, 2).reshape(nchw.shape), nchw), "NLC and NCHW outputs must hold identical values"
    # empty cells are exact zeros
    assert torch.equal((nchw.cpu() == 0).all(1), (rout == 0).all(1))


@pytest.mark.parametrize("name", ["M4", "M16", "M128", "M512"])
def test_density_ablation_shapes(cuda_device, name):
    tiles, kw = cases.edge_cases()[name]
    grid = cases.grid_for(kw)
    enc, ref = build(cuda_device, grid, seed=5)
    x = to_nested(tiles, cuda_device)
    with torch.no_grad():
        rout = ref(tiles, return_flattened=True)
        out = enc(x, return_flattened=True)
    assert_close(out, rout, 1e-3, name)


def test_dense_input_equals_jagged(cuda_device):
    grid = po.GridSpec()
    enc, ref = build(cuda_device, grid, seed=6)
    tiles = [po.synth_tile(5000, 1), po.synth_tile(5000, 2)]
    dense = torch.from_numpy(np.stack(tiles)).to(cuda_device)
    with torch.no_grad():
        a = enc(dense)
        b = enc(to_nested(tiles, cuda_device))
        c = enc([torch.from_numpy(t).to(cuda_device) for t in tiles])
        r = ref(tiles)
    assert torch.equal(a, b) and torch.equal(a, c)
    assert_close(a, r, 1e-3, "dense")


def test_point_stride_4_ignores_the_fourth_lane(cuda_device):
    grid = po.GridSpec()
    enc, ref = build(cuda_device, grid, seed=7)
    t = po.synth_tile(4000, 3)
    t4 = np.concatenate([t, np.random.default_rng(0).uniform(0, 1, (len(t), 1)).astype(np.float32)], 1)
    with torch.no_grad():
        a = enc(torch.from_numpy(t)[None].to(cuda_device))
        b = enc(torch.from_numpy(t4)[None].to(cuda_device))
    assert torch.equal(a, b)


def test_concat_offset_bf16_output_and_dropout(cuda_device):
    grid = po.GridSpec()
    enc, ref = build(cuda_device, grid, seed=8)
    tiles = [po.synth_tile(9000, 4), po.synth_tile(200, 5)]
    x = to_nested(tiles, cuda_device)
    with torch.no_grad():
        rout = ref(tiles, return_flattened=False)
    buf = torch.full((2, 768, 28, 28), 7.0, device=cuda_device)
    enc.encode_into(x, buf, 0, c_total=768, c_offset=384)
    assert torch.all(buf[:, :384] == 7.0)
    assert_close(buf[:, 384:], rout, 1e-3, "concat offset")
    b16 = torch.empty(2, 384, 28, 28, dtype=torch.bfloat16, device=cuda_device)
    enc.encode_into(x, b16, 0, c_total=384, c_offset=0, precision="bf16")
    assert_close(b16, rout, 1.5e-2, "bf16 out")
    enc.encode_into(x, buf, 0, c_total=768, c_offset=384, lidar_zero=True)
    assert torch.all(buf[:, 384:] == 0) and torch.all(buf[:, :384] == 7.0)


@pytest.mark.parametrize("C", [128, 256, 200])
def test_other_channel_widths(cuda_device, C):
    grid = po.GridSpec()
    enc, ref = build(cuda_device, grid, seed=9, C=C)
    tiles = [po.synth_tile(6000, 6)]
    with torch.no_grad():
        r = ref(tiles)
        a = enc(to_nested(tiles, cuda_device))
    assert_close(a, r, 1e-3, f"C={C}")


def test_full_size_batch_properties(cuda_device):
    """BASELINE config 2 (B=16, N=100k): properties that do not need the oracle at full size + oracle on 2 tiles."""
    grid = po.GridSpec()
    enc, ref = build(cuda_device, grid, seed=10)
    tiles = [po.synth_tile(100_000, 2000 + i, clustered=(i % 2 == 1)) for i in range(16)]
    x = to_nested(tiles, cuda_device)
    with torch.no_grad():
        out = enc(x, return_flattened=False)
        again = enc(x, return_flattened=False)
        assert torch.equal(out, again), "the path must be deterministic"
        # tile independence: any sub-batch gives the same tiles
        sub = enc(to_nested(tiles[5:7], cuda_device), return_flattened=False)
        assert torch.equal(sub, out[5:7])
        # permuting whole tiles permutes the output
        perm = [3, 0, 15, 7]
        p = enc(to_nested([tiles[i] for i in perm], cuda_device), return_flattened=False)
        assert torch.equal(p, out[perm])
        r = ref(tiles[:2], return_flattened=False)
    assert_close(out[:2], r, 1e-3, "full size")
    raw = enc.voxelize_raw(x, want_points=False)
    idx = raw["pillar_point_idx"]
    n = raw["pillar_num_points"]
    # kept indices are strictly ascending inside every pillar and counts are capped at M
    valid = torch.arange(64, device=cuda_device).view(1, 1, -1) < n.unsqueeze(-1)
    d = idx[..., 1:] - idx[..., :-1]
    assert torch.all(d[valid[..., 1:]] > 0) and int(n.max()) <= 64
    assert raw["num_pillars"].max() <= 784


def test_training_path_runs_and_matches_train_mode_oracle(cuda_device):
    grid = po.GridSpec()
    enc, ref = build(cuda_device, grid, seed=11)
    tiles = [po.synth_tile(3000, 7), po.synth_tile(800, 8)]
    enc.train(); ref.train()
    out = enc(to_nested(tiles, cuda_device), return_flattened=False)
    r = ref(tiles, return_flattened=False)
    assert_close(out.detach(), r.detach(), 1e-3, "train-mode forward")
    out.sum().backward()
    assert enc.voxel_encoder.pfn_layers[1].linear.weight.grad is not None
    assert torch.allclose(enc.voxel_encoder.pfn_layers[0].norm.running_mean.cpu(), ref.voxel_encoder.pfn_layers[0].norm.running_mean, atol=1e-4)
