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

TOL = {"fp32": 1e-4, "tf32": 1e-3, "fp16": 1e-3, "bf16": 1e-2}


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


REL_EPS = 0.25  # floor of the element-wise relative metric, as a fraction of the tensor's scale (max |ref|)


def assert_close(out, ref, tol, what):
    """Four checks against the oracle, tol = the contract's relative tolerance (1e-3 fp32 / 1e-2 bf16):
      1. norm-wise: max |a - b| <= tol * max |b|;
      2. element-wise mixed: |a - b| <= tol * (|b| + max |b|)                  (allclose with atol = tol * scale);
      3. root-mean-square: rms(a - b) <= tol / 2 * rms(b);
      4. element-wise relative with a floor: |a - b| <= tol * max(|b|, REL_EPS * max |b|) for all but 1 % of the
         elements (a sum of 10^1..10^4 products of operands rounded to 8 / 11 mantissa bits carries the rounding error
         of its TERMS, about tol / 3 of the tensor's scale whatever the sum cancels to, so a pure |a - b| / |b| bound
         is not meaningful for small elements; REL_EPS = 0.25 is where that error meets the bound)."""
    out, ref = out.float().cpu(), ref.float()
    assert out.shape == ref.shape, (what, out.shape, ref.shape)
    if ref.numel() == 0:
        return
    scale = max(ref.abs().max().item(), 1e-6)
    diff = (out - ref).abs()
    err = diff.max().item() / scale
    assert err <= tol, f"{what}: max|err|/max|ref| = {err:.3e} > {tol}"
    assert torch.allclose(out, ref, rtol=tol, atol=tol * scale), what
    rms_ref = ref.pow(2).mean().sqrt().item()
    if rms_ref > 0:
        rms = diff.pow(2).mean().sqrt().item() / rms_ref
        assert rms <= tol / 2, f"{what}: rms(err)/rms(ref) = {rms:.3e} > {tol / 2}"
    bad = (diff > tol * torch.clamp(ref.abs(), min=REL_EPS * scale)).float().mean().item()
    assert bad <= 1e-2, f"{what}: {100 * bad:.2f} % of the elements exceed the relative bound {tol} (floor {REL_EPS} of scale)"


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


@pytest.mark.parametrize("prec", ["fp32", "tf32", "fp16", "bf16"])
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
        # every precision and both layouts: M > 64 runs on the tensor cores as 2 / 4 / 8 blocks of 64 rows per pillar
        rn = ref(tiles, return_flattened=False)
        for prec, tol in (("tf32", 1e-3), ("bf16", 1e-2), ("fp32", 1e-4)):
            o2 = enc.encode_into(x, torch.empty(len(tiles), 384, 28, 28, device=cuda_device), 0, c_total=384, c_offset=0, precision=prec)
            torch.cuda.synchronize()
            assert_close(o2, rn, tol, f"{name} {prec} nchw")
        feats, coors = enc.pillar_features(x)
        rv, rnum, rc, _ = ref.voxelize(tiles)
        rf = ref.voxel_encoder(rv, rnum, rc)
        assert_close(feats, rf, 1e-3, f"{name} pillar features")


@pytest.mark.parametrize("M", [100, 128, 256])
def test_pillars_around_block_boundaries(cuda_device, M):
    """M > 64: pillars whose point counts sit on the 64-row block boundaries (63, 64, 65, 127, 128, 129, M - 1, M, M + 1, 1)
    and a pillar beyond M -- the padded slot of the reference lives in a block of its own when n is a multiple of 64."""
    counts = [63, 64, 65, 127, 128, 129, M - 1, M, M + 1, 1, 3 * M]
    tile = np.concatenate([cases.pillar_block(2 * i % 28, 2 * (2 * i // 28), n, 70 + i) for i, n in enumerate(counts)])
    tile = tile[np.random.default_rng(3).permutation(len(tile))]
    grid = po.GridSpec(max_num_points=M)
    enc, ref = build(cuda_device, grid, seed=6)
    x = to_nested([tile], cuda_device)
    with torch.no_grad():
        rout = ref([tile], return_flattened=True)
        for prec, tol in (("fp16", 1e-3), ("tf32", 1e-3), ("fp32", 1e-4)):
            out = enc.encode_into(x, torch.empty(1, 784, 384, device=cuda_device), 1, precision=prec)
            torch.cuda.synchronize()
            assert_close(out, rout, tol, f"M={M} {prec}")


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


@pytest.mark.parametrize("prec", ["tf32", "fp16"])
def test_odd_canvas_takes_the_general_path(cuda_device, prec):
    """216 px tiles -> 27 x 27 = 729 cells: units of 8 cells straddle tiles and the last unit is partial, so the
    tensor-core kernel runs its general mode (per-item bounds, scalar stores) in both layouts."""
    grid = po.GridSpec(in_width=216.0, in_height=216.0, output_shape=(27, 27), max_voxels=(729, 729))
    C = 384
    cfg = default_cfg(device=str(cuda_device), in_size=216, max_num_voxels=grid.max_voxels, p3p_precision=prec)
    enc = PointPillarsEncoder(cfg, voxel_encoder={"in_channels": 3, "feat_channels": [64, C]},
                              scatter={"in_channels": C, "output_shape": [27, 27]}).to(cuda_device).eval()
    sd, _ = po.synth_weights(21, feat_channels=(64, C))
    enc.load_state_dict(sd)
    ref = po.OraclePointPillarsEncoder(grid).eval()
    ref.load_state_dict(sd)
    tiles = []
    for i, n in enumerate((9000, 40, 20000)):
        t = po.synth_tile(n, 700 + i, clustered=(i == 2))
        t[:, :2] *= np.float32(216.0 / 224.0)
        tiles.append(t)
    with torch.no_grad():
        rv, rn, rc, _ = ref.voxelize(tiles)
        gv, gn, gc = enc.voxelize(to_nested(tiles, cuda_device))
        assert torch.equal(gc.cpu(), rc) and torch.equal(gn.cpu(), rn) and torch.equal(gv.cpu(), rv)
        r_rows = ref(tiles)
        r_nchw = ref(tiles, return_flattened=False)
        a_rows = enc(to_nested(tiles, cuda_device))
        a_nchw = enc(to_nested(tiles, cuda_device), return_flattened=False)
    assert_close(a_rows, r_rows, TOL[prec], f"rows {prec}")
    assert_close(a_nchw, r_nchw, TOL[prec], f"nchw {prec}")


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
        r = ref(tiles, return_flattened=False)  # all 16 tiles of BASELINE configs[1] against the oracle
    assert_close(out, r, 1e-3, "full size")
    # integer surface of the full batch: bit exact against the oracle
    gv, gn, gc = enc.voxelize(x)
    rv, rn, rc, rd = ref.voxelize(tiles)
    assert torch.equal(gc.cpu(), rc) and torch.equal(gn.cpu(), rn) and torch.equal(gv.cpu(), rv)
    raw = enc.voxelize_raw(x, want_points=False)
    idx = raw["pillar_point_idx"]
    n = raw["pillar_num_points"]
    # kept indices are strictly ascending inside every pillar and counts are capped at M
    valid = torch.arange(64, device=cuda_device).view(1, 1, -1) < n.unsqueeze(-1)
    d = idx[..., 1:] - idx[..., :-1]
    assert torch.all(d[valid[..., 1:]] > 0) and int(n.max()) <= 64
    assert raw["num_pillars"].max() <= 784


@pytest.mark.parametrize("name,B,N", [("ffl_b32", 32, 100_000), ("density_400k", 2, 400_000), ("density_10k", 16, 10_000)])
def test_baseline_config_sizes_against_oracle(cuda_device, name, B, N):
    """BASELINE configs[3] (FFL shape, B = 32) and the ends of the density sweep (configs[4]): every tile against the oracle,
    integers bit exact, canvas within the fp32 contract (default fp16 operands, and tf32)."""
    grid = po.GridSpec()
    enc, ref = build(cuda_device, grid, seed=21)
    tiles = [po.synth_tile(N, 5000 + i, clustered=(i % 2 == 1)) for i in range(B)]
    x = to_nested(tiles, cuda_device)
    with torch.no_grad():
        gv, gn, gc = enc.voxelize(x)
        rv, rn, rc, rd = ref.voxelize(tiles)
        assert torch.equal(gc.cpu(), rc) and torch.equal(gn.cpu(), rn) and torch.equal(gv.cpu(), rv), name
        r = ref(tiles, return_flattened=True)
        for prec in ("fp16", "tf32"):
            out = enc.encode_into(x, torch.empty(B, 784, 384, device=cuda_device), 1, precision=prec)
            torch.cuda.synchronize()
            assert_close(out, r, 1e-3, f"{name} {prec}")


def test_full_size_fusion_against_oracle(cuda_device):
    """BASELINE configs[2]'s per-GPU share (8 tiles, image + 100k points): the whole concat buffer against the oracle."""
    from pixelspointspolygons_b200.fusion import EarlyFusionFrontEnd

    cfg = default_cfg(device=str(cuda_device))
    fe = EarlyFusionFrontEnd(cfg).to(cuda_device).eval()
    sd, sdi = po.synth_weights(31)
    fe.lidar_embed.load_state_dict(sd)
    fe.image_embed.load_state_dict(sdi)
    ref_enc = po.OraclePointPillarsEncoder(po.GridSpec()).eval()
    ref_enc.load_state_dict(sd)
    ref_pe = po.OraclePatchEmbed().eval()
    ref_pe.load_state_dict(sdi)
    tiles = [po.synth_tile(100_000, 6000 + i, clustered=(i % 2 == 1)) for i in range(8)]
    imgs = torch.rand(8, 3, 224, 224, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        want = po.early_fusion_front(ref_pe, ref_enc, imgs, tiles)
        got = fe(imgs.to(cuda_device), to_nested(tiles, cuda_device))
    torch.cuda.synchronize()
    assert_close(got[:, :384], want[:, :384], 1e-3, "full-size fusion, image half")
    assert_close(got[:, 384:], want[:, 384:], 1e-3, "full-size fusion, LiDAR half")


def test_lidar_dropout_draw_uses_the_device_generator(cuda_device):
    """early_fusion_vit.py:115 draws `torch.rand(1, device=x_lidar.device)`: a seeded run must drop the same batches."""
    from pixelspointspolygons_b200.fusion import EarlyFusionFrontEnd

    fe = EarlyFusionFrontEnd(default_cfg(device=str(cuda_device), lidar_dropout=0.5))
    torch.cuda.manual_seed(123)
    ours = [fe._dropout_now(cuda_device) for _ in range(16)]
    torch.cuda.manual_seed(123)
    theirs = [bool(torch.rand(1, device=cuda_device).item() <= 0.5) for _ in range(16)]
    assert ours == theirs and any(ours) and not all(ours)


@pytest.mark.parametrize("prec", ["fp32", "tf32", "fp16"])
def test_vit_token_sequence_matches_oracle(cuda_device, prec):
    """class token + positional embedding fused into the encoder's store (p3p_encode_tokens)."""
    grid = po.GridSpec()
    enc, ref = build(cuda_device, grid, seed=13)
    g = torch.Generator().manual_seed(5)
    cls = torch.randn(1, 1, 384, generator=g) * 0.5
    pos = torch.randn(1, 785, 384, generator=g) * 0.2
    tiles = [po.synth_tile(7000, 41), np.zeros((0, 3), np.float32), po.synth_tile(30000, 42, clustered=True)]
    with torch.no_grad():
        r = po.vit_tokens(ref(tiles), cls, pos)
        a = enc.forward_tokens(to_nested(tiles, cuda_device), cls.to(cuda_device), pos.to(cuda_device), precision=prec)
    assert a.shape == (3, 785, 384)
    assert torch.equal(a[:, 0].cpu(), (cls + pos[:, :1]).reshape(1, 384).expand(3, -1))  # class-token rows: exact
    assert torch.equal(a[1, 1:].cpu(), pos[0, 1:])  # the empty tile: positional embedding only
    assert_close(a, r, TOL[prec], f"tokens {prec}")


@pytest.mark.parametrize("variant", ["dataset", "predict"])
def test_las_front_end_is_bit_exact(cuda_device, variant):
    """raw LAS integers -> pixel-space fp32 points, bit for bit the numpy / sklearn loader."""
    from pixelspointspolygons_b200 import las_to_pixels

    rng = np.random.default_rng(11)
    metas, Xs, Ys, Zs, refs = [], [], [], [], []
    for i, n in enumerate((30_000, 1, 77_777, 5)):
        left, top = 2_600_000.0 + 56.0 * i, 1_200_000.0 + 56.0 * i
        sc, of = (0.001, 0.001, 0.001 if i != 2 else 0.01), (left - 3.0, top - 7.0, 400.0)
        X = rng.integers(2_900, 59_100, n).astype(np.int32)   # a little beyond the 56 m tile on both sides
        Y = rng.integers(6_900, 63_100, n).astype(np.int32)
        Z = rng.integers(-5_000, 60_000, n).astype(np.int32) if i != 3 else np.full(n, 777, np.int32)
        m = dict(scales=sc, offsets=of, top_left=(left, top), height=224, width=224)
        r = po.las_points_to_pixels(X, Y, Z, sc, of, top_left=(left, top), variant=variant)
        if variant == "dataset":  # training loader: the replayed D4 element follows (one of the 8 per tile)
            m["d4"] = ("r90", "hvt", "v", "t")[i]
            r = po.apply_d4_to_lidar(r, m["d4"])
        metas.append(m); Xs.append(X); Ys.append(Y); Zs.append(Z)
        refs.append(r)
    offs = torch.tensor(np.concatenate([[0], np.cumsum([len(x) for x in Xs])]), dtype=torch.int64, device=cuda_device)
    cat = lambda a: torch.from_numpy(np.concatenate(a)).to(cuda_device)
    got = las_to_pixels(cat(Xs), cat(Ys), cat(Zs), offs, metas, z_hi=100.0, variant=variant).cpu().numpy()
    ref = np.concatenate(refs)
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), np.abs(got - ref).max()
    if variant == "dataset":
        assert got[:, :2].min() >= 0.0 and got[:, :2].max() <= 224.0
    # the packed transfer format (uint16 deltas + int32 base per tile, 6 bytes per point): the same bits
    from pixelspointspolygons_b200 import las_packed_to_pixels, pack_las

    packed = [pack_las(X, Y, Z) for X, Y, Z in zip(Xs, Ys, Zs)]
    deltas = torch.from_numpy(np.concatenate([d for d, _ in packed])).to(cuda_device)
    base = torch.from_numpy(np.stack([b for _, b in packed])).to(cuda_device)
    assert deltas.dtype == torch.uint16 and deltas.shape == (ref.shape[0], 3)
    got16 = las_packed_to_pixels(deltas, base, offs, metas, z_hi=100.0, variant=variant).cpu().numpy()
    assert np.array_equal(got16.view(np.uint32), ref.view(np.uint32))
    with pytest.raises(ValueError, match="65535"):
        pack_las(np.array([0, 70_000], np.int32), np.array([0, 1], np.int32), np.array([0, 1], np.int32))


def test_las_front_end_d4_elements(cuda_device):
    """all eight D4 elements against the reference's statements (p3_coco.py:114-160)"""
    from pixelspointspolygons_b200 import las_to_pixels

    rng = np.random.default_rng(12)
    names = ["e", "r90", "r180", "r270", "v", "hvt", "h", "t"]
    n = 4000
    X = rng.integers(0, 56_000, n).astype(np.int32); Y = rng.integers(0, 56_000, n).astype(np.int32)
    Z = rng.integers(0, 30_000, n).astype(np.int32)
    sc, of, tl = (0.001, 0.001, 0.001), (2_600_000.0, 1_200_000.0, 0.0), (2_600_000.0, 1_200_000.0)
    base = po.las_points_to_pixels(X, Y, Z, sc, of, top_left=tl)
    metas = [dict(scales=sc, offsets=of, top_left=tl, height=224, width=224, d4=g) for g in names]
    offs = torch.arange(0, (len(names) + 1) * n, n, dtype=torch.int64, device=cuda_device)
    rep = lambda a: torch.from_numpy(np.tile(a, len(names))).to(cuda_device)
    got = las_to_pixels(rep(X), rep(Y), rep(Z), offs, metas).cpu().numpy().reshape(len(names), n, 3)
    for k, g in enumerate(names):
        ref = po.apply_d4_to_lidar(base, g)
        assert np.array_equal(got[k].view(np.uint32), ref.view(np.uint32)), g


def test_multi_wave_batch_with_dependent_launch(cuda_device):
    """40 tiles x 60k points = about 1000 ranking chunks, more than two waves of the voxelizer grid: the PFN grid is a
    programmatic dependent launch that may start while the voxelizer's last wave drains -- no deadlock, same results."""
    grid = po.GridSpec()
    enc, ref = build(cuda_device, grid, seed=12)
    tiles = [po.synth_tile(60_000, 3000 + i, clustered=(i % 3 == 0)) for i in range(40)]
    x = to_nested(tiles, cuda_device)
    with torch.no_grad():
        out = enc(x)
        torch.cuda.synchronize()
        for _ in range(3):  # back to back: the next voxelizer follows the dependent launch of the previous step
            again = enc(x)
        torch.cuda.synchronize()
        assert torch.equal(out, again)
        sub = enc(to_nested(tiles[37:40], cuda_device))
        assert torch.equal(sub, out[37:40])
        r = ref(tiles[38:40])
    assert_close(out[38:40], r, 1e-3, "multi-wave batch")


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


# ---------------------------------------------------------------------------------------------------------------
# image patch embedding + early-fusion concat
# ---------------------------------------------------------------------------------------------------------------
def build_fusion(dev, seed=0, lidar_dropout=None, precision="tf32"):
    from pixelspointspolygons_b200.fusion import EarlyFusionFrontEnd

    cfg = default_cfg(device=str(dev), lidar_dropout=lidar_dropout, p3p_precision=precision)
    fe = EarlyFusionFrontEnd(cfg).to(dev).eval()
    fe.image_embed.precision = precision
    sd, sdi = po.synth_weights(seed)
    fe.lidar_embed.load_state_dict(sd)
    fe.image_embed.load_state_dict(sdi)
    ref_l = po.OraclePointPillarsEncoder(po.GridSpec()).eval()
    ref_l.load_state_dict(sd)
    ref_i = po.OraclePatchEmbed().eval()
    ref_i.load_state_dict(sdi)
    return fe, ref_i, ref_l


@pytest.mark.parametrize("prec", ["fp32", "tf32", "fp16", "bf16"])
def test_patch_embed_matches_conv2d(cuda_device, prec):
    fe, ref_i, _ = build_fusion(cuda_device, seed=12, precision=prec)
    g = torch.Generator().manual_seed(5)
    img = torch.rand(3, 3, 224, 224, generator=g)
    img[1] = (img[1] - 0.5) * 4.0  # normalised images are not confined to [0, 1]
    with torch.no_grad():
        r = ref_i(img)
    fe.image_embed.precision = prec
    out = fe.image_embed(img.to(cuda_device))
    assert out.shape == (3, 384, 28, 28)
    assert_close(out, r, TOL[prec], f"patch embed {prec}")
    # bf16 output buffer, channel offset inside a wider buffer
    buf = torch.full((3, 768, 28, 28), 3.0, dtype=torch.bfloat16, device=cuda_device)
    fe.image_embed.forward_into(img.to(cuda_device), buf, 768, 0, precision=prec)
    assert torch.all(buf[:, 384:] == 3.0)
    assert_close(buf[:, :384], r, max(TOL[prec], 8e-3), f"patch embed {prec} bf16 out")


@pytest.mark.parametrize("prec", ["tf32", "fp16", "bf16"])
def test_patch_embed_prepared_weights_equal_raw_weights(cuda_device, prec):
    """p3p_patch_embed_prepared (weights stored once as operand tiles, copied by TMA) against p3p_patch_embed (every CTA
    converts the raw weights): the same bits, both layouts, a channel count that is not a multiple of 128, and a parameter
    update that must invalidate the cached tiles."""
    import ctypes as C

    from pixelspointspolygons_b200 import _lib
    from pixelspointspolygons_b200.fusion import PatchEmbed

    g = torch.Generator().manual_seed(21)
    for dim in (384, 200):
        pe = PatchEmbed(224, 8, 3, dim, precision=prec).to(cuda_device).eval()
        with torch.no_grad():
            pe.proj.weight.copy_(torch.randn(pe.proj.weight.shape, generator=g) * 0.1)
            pe.proj.bias.copy_(torch.randn(dim, generator=g) * 0.1)
        img = ((torch.rand(5, 3, 224, 224, generator=g) - 0.5) * 3.0).to(cuda_device)
        for layout, shape in ((0, (5, dim, 28, 28)), (1, (5, 784, dim))):
            raw = torch.zeros(shape, device=cuda_device)
            w, b = pe.proj.weight.detach().contiguous(), pe.proj.bias.detach().clone()
            rc = _lib.lib().p3p_patch_embed(img.data_ptr(), 5, 3, 224, 224, 8, w.data_ptr(), b.data_ptr(), dim, _lib.P3P_PRECISION[prec],
                                            raw.data_ptr(), 0, layout, dim, 0, torch.cuda.current_stream().cuda_stream)
            _lib.check(rc, "p3p_patch_embed")
            got = pe.forward_into(img, torch.zeros(shape, device=cuda_device), dim, 0, layout=layout)
            assert torch.equal(got, raw), (prec, dim, layout)
        with torch.no_grad():
            pe.proj.weight.mul_(2.0)  # in-place update: new version counter, the tiles are re-made
            pe.proj.bias.zero_()
        again = pe.forward_into(img, torch.zeros(shape, device=cuda_device), dim, 0, layout=1)
        assert torch.allclose(again, got * 2.0 - 2.0 * b.view(1, 1, -1), rtol=1e-5, atol=1e-5)


def test_patch_embed_other_shapes_take_the_exact_route(cuda_device):
    from pixelspointspolygons_b200.fusion import PatchEmbed

    g = torch.Generator().manual_seed(6)
    for (size, patch, chans, dim) in [(64, 16, 3, 96), (48, 4, 1, 40), (224, 8, 3, 200), (128, 8, 3, 384)]:
        pe = PatchEmbed(size, patch, chans, dim).to(cuda_device).eval()
        ref = torch.nn.Conv2d(chans, dim, patch, patch)
        with torch.no_grad():
            ref.weight.copy_(torch.randn(ref.weight.shape, generator=g) * 0.1)
            ref.bias.copy_(torch.randn(dim, generator=g) * 0.1)
        pe.proj.load_state_dict(ref.state_dict())
        img = torch.rand(2, chans, size, size, generator=g)
        with torch.no_grad():
            r = ref(img)
        assert_close(pe(img.to(cuda_device)), r, 1e-3, f"patch embed {size}/{patch}/{chans}/{dim}")


@pytest.mark.parametrize("prec", ["tf32", "fp16", "bf16"])
def test_early_fusion_concat_matches_oracle(cuda_device, prec):
    fe, ref_i, ref_l = build_fusion(cuda_device, seed=13, precision=prec)
    tiles = [po.synth_tile(20000, 41), po.synth_tile(500, 42, clustered=True), np.zeros((0, 3), np.float32)]
    img = torch.rand(3, 3, 224, 224, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        r = po.early_fusion_front(ref_i, ref_l, img, tiles)
        rz = po.early_fusion_front(ref_i, ref_l, img, tiles, apply_lidar_dropout=True)
        out = fe(img.to(cuda_device), to_nested(tiles, cuda_device))
    assert out.shape == (3, 768, 28, 28)
    assert_close(out[:, :384], r[:, :384], TOL[prec], "image half")
    assert_close(out[:, 384:], r[:, 384:], TOL[prec], "lidar half")
    # LiDAR dropout with p = 1.0 (what the trainer forces during validation): x_lidar * 0.0
    fe.cfg.experiment.lidar_dropout = 1.0
    with torch.no_grad():
        z = fe(img.to(cuda_device), to_nested(tiles, cuda_device))
    assert torch.all(z[:, 384:] == 0)
    assert_close(z[:, :384], rz[:, :384], TOL[prec], "image half under dropout")
    fe.cfg.experiment.lidar_dropout = 0.0  # rand <= 0.0 practically never: LiDAR half present
    with torch.no_grad():
        k = fe(img.to(cuda_device), to_nested(tiles, cuda_device))
    assert torch.equal(k, out)


def test_sharded_batch_equals_single_gpu_batch(cuda_device):
    """encoding the shards of a batch one by one == encoding the whole batch (bit identical)."""
    from pixelspointspolygons_b200 import shard

    enc, _ = build(cuda_device, po.GridSpec(), seed=14)
    tiles = [po.synth_tile(3000 + 997 * i, 50 + i, clustered=(i % 3 == 0)) for i in range(11)]
    x = to_nested(tiles, cuda_device)
    with torch.no_grad():
        whole = enc(x, return_flattened=False)
        for world in (2, 4, 8):
            parts = [enc(shard.shard_lidar(x, r, world), return_flattened=False) for r in range(world)]
            assert torch.equal(torch.cat(parts, 0), whole), world


def test_fp16_range_guard_falls_back_to_tf32(cuda_device):
    """fp16 operands are only used while the layer-0 activations provably stay inside the fp16 range."""
    enc, ref = build(cuda_device, po.GridSpec(), seed=15)
    enc.precision = "fp16"
    assert enc._resolve_precision(None) == "fp16"
    tiles = [po.synth_tile(4000, 9)]
    with torch.no_grad():
        enc.voxel_encoder.pfn_layers[0].linear.weight.mul_(400.0)
        ref.voxel_encoder.pfn_layers[0].linear.weight.mul_(400.0)
        assert enc._resolve_precision(None) == "tf32"
        out = enc(to_nested(tiles, cuda_device))
        r = ref(tiles)
    assert torch.isfinite(out).all()
    assert_close(out, r, 1e-3, "fp16 guard -> tf32")
