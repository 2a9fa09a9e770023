"""Generate tests/golden/*.npz from the CPU oracle (run from the repo root: `python tests/golden/make_golden.py`).

The reference ships no golden vectors for this path and open3d==0.19.0 (which holds its arithmetic) cannot be
imported offline, so these fixtures pin the ORACLE (against drift of oracle/ and of torch/numpy versions) and give
the GPU tests a second, file-based target.  They do not pin the oracle against Open3D: parity stays "unpinned"
(DESIGN.md section 2).  Inputs are stored in the file, so the fixtures are self-contained.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import p3p_cases as cases  # noqa: E402
from oracle import pillars_oracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
C = 128  # narrow feature width keeps the files small; the arithmetic is the same


def fixture(name, tiles, kw, seed):
    grid = po.GridSpec(**{**kw, "feat_channels": (64, C)})
    enc = po.OraclePointPillarsEncoder(grid).eval()
    sd, sdi = po.synth_weights(seed, feat_channels=(64, C))
    enc.load_state_dict(sd)
    with torch.no_grad():
        voxels, nums, coors, dense = enc.voxelize(tiles)
        feats, _, _ = enc.pillar_features(tiles)
        canvas = enc(tiles, return_flattened=False)
    hashes = np.concatenate([po.voxelize_c(t, grid, 1)["point_hash"] for t in tiles]) if sum(len(t) for t in tiles) else np.zeros(0, np.int64)
    out = dict(
        points=np.concatenate(tiles).astype(np.float32), offsets=np.cumsum([0] + [len(t) for t in tiles]).astype(np.int64),
        grid_max_num_points=grid.max_num_points, grid_max_voxels=np.asarray(grid.max_voxels), grid_drop_overflow=int(grid.drop_overflow),
        weight_seed=seed, channels=C, point_hash=hashes.astype(np.int32), coors=coors.numpy().astype(np.int32),
        num_points=nums.numpy().astype(np.int32), dense_idx=dense.numpy().astype(np.int32),
        pillar_features=feats.numpy().astype(np.float32), canvas_nonzero_cells=np.argwhere((canvas != 0).any(1).numpy()).astype(np.int32),
        canvas_checksum=canvas.double().sum(dim=(2, 3)).numpy(),
    )
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "pillars", len(coors), "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


def main():
    ec = cases.edge_cases()
    fixture("edge_occupancy_M", *ec["occupancy_M"], seed=21)
    fixture("edge_alias_z100", [np.concatenate(ec["x224_then_alias"][0] + ec["z100_overwrites_cell"][0] + ec["corner_224_224_100"][0])], {}, seed=22)
    fixture("edge_vmax_cut", *ec["vmax_cut_with_alias"], seed=23)
    fixture("ragged_small", [po.synth_tile(1500, 91), np.zeros((0, 3), np.float32), po.synth_tile(400, 92, clustered=True)], {}, seed=24)
    # patch embed + concat (8 output channels of a 384-wide embed would not exercise the tiles; keep C = 128, 1 image)
    g = torch.Generator().manual_seed(25)
    img = torch.rand(1, 3, 224, 224, generator=g)
    pe = po.OraclePatchEmbed(embed_dim=C).eval()
    _, sdi = po.synth_weights(25, feat_channels=(64, C))
    pe.load_state_dict(sdi)
    with torch.no_grad():
        y = pe(img)
    np.savez_compressed(os.path.join(HERE, "patch_embed.npz"), image_seed=25, weight_seed=25, channels=C,
                        out=y.numpy().astype(np.float32))
    print("patch_embed bytes", os.path.getsize(os.path.join(HERE, "patch_embed.npz")))


if __name__ == "__main__":
    main()
