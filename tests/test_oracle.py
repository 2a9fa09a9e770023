"""CPU tests: pin the oracle (oracle/) to hand-computed vectors of SURVEY.md Appendix A/B and cross-check its two
independent restatements (plain C vs numpy).  The reference ships no tests or golden vectors for this path and
open3d==0.19.0 cannot be imported here, so parity stays "unpinned" (DESIGN.md); these tests pin the *stated*
semantics instead."""
import numpy as np
import pytest
import torch

from oracle import pillars_oracle as po
import p3p_cases as cases

F = np.float32
G = po.GridSpec()


def vox(pts, grid=G, mv=784, impl=po.voxelize_c):
    return impl(np.asarray(pts, F).reshape(-1, 3), grid, mv)


def test_grid_constants():
    a = vox([[1, 1, 1]])
    assert a["extents"] == (28, 28, 1)  # ceil(100 * 0.01f) == 1
    assert a["num_cells"] == 784


def test_hash_is_fp32_multiply_by_reciprocal():
    # z: 100 * fp32(0.01) rounds to 1.0f -> cell 1; nextafter(100, 0) stays in cell 0 (Appendix B.2)
    a = vox([[0, 0, 100.0], [0, 0, np.nextafter(F(100), F(0))], [8.0, 16.0, 0.0], [7.9999995, 15.999999, 0.0]])
    assert a["point_hash"].tolist() == [784, 0, 1 + 2 * 28, 0 + 1 * 28]


def test_inclusive_range_and_invalids():
    tiles, _ = cases.edge_cases()["out_of_range"]
    a = vox(tiles[0])
    assert a["point_hash"].tolist() == [-1, -1, -1, -1, -1, -1, -1, -1, 0]
    assert a["voxel_coords"].tolist() == [[0, 0, 0]]
    assert a["voxel_point_indices"].tolist() == [8]


def test_x224_aliases_next_row_and_first_point_decides_coords():
    tiles, _ = cases.edge_cases()["x224_then_alias"]
    a = vox(tiles[0])
    # (cx=28, cy=2) hashes to 28 + 2*28 = 84 == hash of (0, 3): one merged run, coords from the lowest index
    assert a["point_hash"].tolist() == [84, 84, 84]
    assert a["voxel_coords"].tolist() == [[28, 2, 0]]
    v, c, n, d = po.voxelization_forward(tiles[0], G)
    assert len(c) == 0  # x cell 28 -> filtered by the x/y bound check
    tiles, _ = cases.edge_cases()["alias_then_x224"]
    v, c, n, d = po.voxelization_forward(tiles[0], G)
    assert c.tolist() == [[0, 3, 0]] and n.tolist() == [3] and d[0, :3].tolist() == [0, 1, 2]


def test_z100_pillar_sorts_last_and_overwrites_its_cell():
    tiles, _ = cases.edge_cases()["z100_overwrites_cell"]
    v, c, n, d = po.voxelization_forward(tiles[0], G)
    assert c.tolist() == [[0, 6, 5], [1, 6, 5]] and n.tolist() == [10, 1]
    enc = po.OraclePointPillarsEncoder(G).eval()
    enc.load_state_dict(po.synth_weights(3)[0])
    with torch.no_grad():
        feats, coors, _ = enc.pillar_features(tiles)
        out = enc(tiles, return_flattened=False)
    assert torch.equal(out[0, :, 6, 5], feats[1])  # the later (z-cell-1) pillar wins the cell
    out[0, :, 6, 5] = 0
    assert out.abs().max() == 0  # every other cell is zero


def test_first_M_by_original_index():
    tiles, _ = cases.edge_cases()["stability_shuffled"]
    pts = tiles[0]
    a = vox(pts)
    h = a["point_hash"]
    for r, (s, e) in enumerate(zip(a["voxel_point_row_splits"][:-1], a["voxel_point_row_splits"][1:])):
        idx = a["voxel_point_indices"][s:e]
        key = h[idx[0]]
        expect = np.flatnonzero(h == key)[:64]
        assert np.array_equal(idx, expect)


def test_max_voxels_keeps_first_runs_in_hash_order():
    tiles, kw = cases.edge_cases()["vmax_cut"]
    g = cases.grid_for(kw)
    v, c, n, d = po.voxelization_forward(tiles[0], g, training=False)
    assert len(c) == 40 and c[:, 1].max() == 1 and c[-1].tolist() == [0, 1, 11]
    v, c, n, d = po.voxelization_forward(tiles[0], g, training=True)
    assert len(c) == 50


@pytest.mark.parametrize("name", sorted(cases.edge_cases()))
def test_c_and_numpy_restatements_agree(name):
    tiles, kw = cases.edge_cases()[name]
    g = cases.grid_for(kw)
    for pts in tiles:
        for mv in set(g.max_voxels):
            a, b = po.voxelize_c(pts, g, mv), po.voxelize_numpy(pts, g, mv)
            for k in a:
                if isinstance(a[k], np.ndarray):
                    assert np.array_equal(a[k], b[k]), (name, k)
                else:
                    assert a[k] == b[k], (name, k)


def test_pfn_closed_form_matches_module():
    """Appendix A.4 eval-mode closed form == the literal module (padded slots take part in both maxes)."""
    g = po.GridSpec()
    enc = po.OraclePointPillarsEncoder(g).eval()
    sd, _ = po.synth_weights(5)
    enc.load_state_dict(sd)
    tiles = [np.concatenate([cases.pillar_block(3, 4, 5, 1), cases.pillar_block(9, 9, 64, 2)])]
    with torch.no_grad():
        feats, coors, nums = enc.pillar_features(tiles)
        voxels, _, _, _ = enc.voxelize(tiles)
    W0, W1 = sd["voxel_encoder.pfn_layers.0.linear.weight"], sd["voxel_encoder.pfn_layers.1.linear.weight"]

    def fold(i):
        p = f"voxel_encoder.pfn_layers.{i}.norm."
        a = sd[p + "weight"] / torch.sqrt(sd[p + "running_var"] + 1e-3)
        return a, sd[p + "bias"] - sd[p + "running_mean"] * a

    (a0, b0), (a1, b1) = fold(0), fold(1)
    for v in range(2):
        n = int(nums[v])
        pts = voxels[v, :n]
        mean = pts.sum(0) / n
        ctr = torch.tensor([coors[v, 3] * 8.0 + 4.0, coors[v, 2] * 8.0 + 4.0])
        d = torch.cat([pts[:, :2] - ctr, pts[:, 2:3], pts - mean, pts[:, :2] - ctr], 1)
        h = torch.relu(a0 * (d @ W0.t()) + b0)
        if n < 64:
            h = torch.cat([h, torch.relu(b0)[None]], 0)
        hmax = h.max(0)[0]
        o = torch.relu(a1 * (h @ W1[:, :32].t() + W1[:, 32:] @ hmax) + b1)
        assert torch.allclose(o.max(0)[0], feats[v], rtol=1e-5, atol=1e-5)


def test_state_dict_keys_match_reference_names():
    enc = po.OraclePointPillarsEncoder(G)
    keys = set(enc.state_dict())
    for i, (cin, cout) in enumerate([(8, 32), (64, 384)]):
        p = f"voxel_encoder.pfn_layers.{i}."
        assert enc.state_dict()[p + "linear.weight"].shape == (cout, cin)
        for s in ("norm.weight", "norm.bias", "norm.running_mean", "norm.running_var", "norm.num_batches_tracked"):
            assert p + s in keys
    assert len(keys) == 12


def test_flatten_is_row_major_tokens():
    enc = po.OraclePointPillarsEncoder(G).eval()
    enc.load_state_dict(po.synth_weights(1)[0])
    tiles = [cases.pillar_block(5, 2, 3, 1)]
    with torch.no_grad():
        a = enc(tiles, return_flattened=True)
        b = enc(tiles, return_flattened=False)
    assert a.shape == (1, 784, 384) and b.shape == (1, 384, 28, 28)
    assert torch.equal(a[0, 2 * 28 + 5], b[0, :, 2, 5]) and a[0, 0].abs().sum() == 0


def test_early_fusion_concat_order():
    enc = po.OraclePointPillarsEncoder(G).eval()
    sd, sdi = po.synth_weights(2)
    enc.load_state_dict(sd)
    pe = po.OraclePatchEmbed().eval()
    pe.load_state_dict(sdi)
    img = torch.rand(1, 3, 224, 224)
    tiles = [cases.pillar_block(1, 1, 9, 1)]
    with torch.no_grad():
        x = po.early_fusion_front(pe, enc, img, tiles)
        z = po.early_fusion_front(pe, enc, img, tiles, apply_lidar_dropout=True)
    assert x.shape == (1, 768, 28, 28)
    assert torch.equal(x[:, :384], pe(img)) and torch.equal(x[:, 384:], enc(tiles, return_flattened=False))
    assert z[:, 384:].abs().sum() == 0 and torch.equal(z[:, :384], x[:, :384])


def test_vit_tokens_is_cat_cls_plus_pos():
    import torch

    x = torch.arange(2 * 3 * 4, dtype=torch.float32).reshape(2, 3, 4)
    cls = torch.full((1, 1, 4), 100.0)
    pos = torch.arange(4 * 4, dtype=torch.float32).reshape(1, 4, 4) * 0.5
    t = po.vit_tokens(x, cls, pos)
    assert t.shape == (2, 4, 4)
    assert torch.equal(t[:, 0], (cls[0, 0] + pos[0, 0]).expand(2, -1))
    assert torch.equal(t[1, 2], x[1, 1] + pos[0, 2])


def test_las_loader_restatement_matches_sklearn_and_hand_values():
    from sklearn.preprocessing import MinMaxScaler

    rng = np.random.default_rng(3)
    n = 500
    X = rng.integers(100_000, 156_000, n).astype(np.int32)
    Y = rng.integers(200_000, 256_000, n).astype(np.int32)
    Z = rng.integers(40_000, 75_000, n).astype(np.int32)
    scales, offs = (0.001, 0.001, 0.001), (2_600_000.0, 1_200_000.0, 0.0)
    top_left = (2_600_100.0, 1_200_200.0)
    got = po.las_points_to_pixels(X, Y, Z, scales, offs, top_left=top_left)
    # literal reference code (p3_coco.py:79-96) on laspy-style float64 coordinates
    pts = np.vstack((X * scales[0] + offs[0], Y * scales[1] + offs[1], Z * scales[2] + offs[2])).transpose()
    pts[:, :2] = (pts[:, :2] - top_left) / 0.25
    pts[:, 1] = 224 - pts[:, 1]
    pts[:, -1] = MinMaxScaler(feature_range=(0, 100)).fit_transform(pts[:, -1].reshape(-1, 1)).squeeze()
    pts = pts.astype(np.float32)
    pts[:, 0] = np.clip(pts[:, 0], 0, 224)
    pts[:, 1] = np.clip(pts[:, 1], 0, 224)
    assert np.array_equal(got, pts)
    assert got[:, 2].min() == 0.0 and abs(got[:, 2].max() - 100.0) < 1e-4
    # predict variant: the tile's own minimum is the origin, no clipping
    got2 = po.las_points_to_pixels(X, Y, Z, scales, offs, variant="predict")
    assert got2[:, 0].min() == 0.0 and got2[:, 1].max() == 224.0
    # degenerate z range: sklearn maps everything to feature_range[0]
    flat = po.las_points_to_pixels(X[:5], Y[:5], np.full(5, 1234, np.int32), scales, offs, top_left=top_left)
    assert np.all(flat[:, 2] == 0.0)


def test_d4_replay_restatement_is_a_group_action():
    pts = np.array([[10.0, 20.0, 3.0], [200.5, 7.25, 50.0]], dtype=np.float32)
    r90 = po.apply_d4_to_lidar(pts, "r90")
    assert np.allclose(r90[0, :2], [112 + (20 - 112), 112 - (10 - 112)])
    four = pts
    for _ in range(4):
        four = po.apply_d4_to_lidar(four, "r90")
    assert np.array_equal(four, pts)
    assert np.array_equal(po.apply_d4_to_lidar(po.apply_d4_to_lidar(pts, "t"), "t"), pts)
    assert np.array_equal(po.apply_d4_to_lidar(pts, "e"), pts)
    assert np.array_equal(po.apply_d4_to_lidar(pts, "r180"), po.apply_d4_to_lidar(po.apply_d4_to_lidar(pts, "h"), "v"))

