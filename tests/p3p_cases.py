"""Edge-case inputs of SURVEY.md Appendix B, shared by the CPU (oracle) and GPU (parity) tests."""
from __future__ import annotations

import numpy as np

from oracle import pillars_oracle as po

F = np.float32


def _t(rows):
    return np.asarray(rows, dtype=F).reshape(-1, 3)


def pillar_block(cx, cy, n, seed, z_hi=99.0):
    """n points strictly inside cell (cx, cy) of the default 8-px grid."""
    rng = np.random.default_rng(seed)
    xy = rng.uniform(0.05, 7.95, (n, 2)) + np.array([cx * 8.0, cy * 8.0])
    z = rng.uniform(0.5, z_hi, (n, 1))
    return np.concatenate([xy, z], 1).astype(F)


def edge_cases():
    """name -> (list of tiles, GridSpec kwargs)."""
    c = {}
    # B.1 inclusive max on x / y (cell 28) and hash aliasing
    c["x224_alone"] = ([_t([[224.0, 20.0, 5.0]])], {})
    c["y224_alone"] = ([_t([[20.0, 224.0, 5.0]])], {})
    c["x224_then_alias"] = ([_t([[224.0, 20.0, 5.0], [3.0, 28.0, 7.0], [4.0, 29.0, 8.0]])], {})  # (28,2) aliases (0,3)
    c["alias_then_x224"] = ([_t([[3.0, 28.0, 7.0], [224.0, 20.0, 5.0], [4.0, 29.0, 8.0]])], {})
    c["corner_224_224_100"] = ([_t([[224.0, 224.0, 100.0], [1.0, 1.0, 1.0]])], {})
    # B.2 z == 100 -> z-cell 1; just below stays in cell 0
    c["z100"] = ([_t([[12.0, 20.0, 100.0], [13.0, 21.0, 50.0], [100.0, 100.0, np.nextafter(F(100), F(0))]])], {})
    c["z100_overwrites_cell"] = ([np.concatenate([pillar_block(5, 6, 10, 1), _t([[44.0, 52.0, 100.0]])])], {})
    c["z100_first_in_file"] = ([np.concatenate([_t([[44.0, 52.0, 100.0]]), pillar_block(5, 6, 10, 2)])], {})
    # B.3 / B.7 out of range, NaN, Inf
    c["out_of_range"] = ([_t([[-0.001, 5.0, 5.0], [5.0, -1e-6, 5.0], [5.0, 5.0, -0.5], [224.01, 5.0, 5.0],
                              [5.0, 5.0, 100.01], [np.nan, 5.0, 5.0], [5.0, np.inf, 5.0], [5.0, 5.0, -np.inf],
                              [6.0, 6.0, 6.0]])], {})
    c["all_invalid"] = ([_t([[-1.0, 5.0, 5.0], [np.nan, np.nan, np.nan]])], {})
    # B.4 cell borders
    c["borders"] = ([_t([[8.0 * k, 8.0 * (27 - k), 10.0] for k in range(28)] + [[7.9999995, 8.0000005, 0.0]])], {})
    # B.5 pillar occupancies around M, empty tile, single pillar, several tiles
    c["occupancy_M"] = ([np.concatenate([pillar_block(0, 0, 64, 3), pillar_block(1, 0, 63, 4), pillar_block(2, 0, 65, 5),
                                         pillar_block(3, 0, 1, 6), pillar_block(27, 27, 200, 7)])], {})
    c["empty_tile_between"] = ([pillar_block(2, 2, 5, 8), np.zeros((0, 3), F), pillar_block(3, 3, 70, 9)], {})
    c["only_empty_tiles"] = ([np.zeros((0, 3), F), np.zeros((0, 3), F)], {})
    c["one_pillar_many"] = ([pillar_block(13, 14, 5000, 10)], {})
    # B.6 more runs than max_voxels
    c["vmax_cut"] = ([np.concatenate([pillar_block(x, y, 3, 100 + x + 28 * y) for y in range(4) for x in range(28)])],
                     dict(max_voxels=(50, 40)))
    c["vmax_cut_with_alias"] = ([np.concatenate([_t([[224.0, 0.5, 1.0]]),
                                                 np.concatenate([pillar_block(x, 0, 2, 300 + x) for x in range(28)]),
                                                 pillar_block(0, 1, 2, 400), pillar_block(1, 1, 2, 401)])],
                                dict(max_voxels=(29, 29)))
    # B.8 stability: interleaved pillars, shuffled order, far more than M points each
    rng = np.random.default_rng(42)
    blk = np.concatenate([pillar_block(4, 4, 700, 11), pillar_block(5, 4, 700, 12), pillar_block(4, 5, 30, 13)])
    c["stability_shuffled"] = ([blk[rng.permutation(len(blk))]], {})
    # keys that cross M in a LATER ranking chunk of the tile (chunks are 1024 .. 4096 points): the kept set is the M lowest
    # indices over chunk boundaries, whatever order the hardware serves the chunk's atomics in
    filler = po.synth_tile(12000, 51)
    filler = filler[(filler[:, 0] > 80) | (filler[:, 1] > 80)]  # keep cells (0..9, 0..9) free for the planted pillars
    a_early, a_late = pillar_block(3, 3, 40, 14), pillar_block(3, 3, 100, 15)
    b_early, b_late = pillar_block(4, 3, 63, 16), pillar_block(4, 3, 5, 17)
    c_late = pillar_block(5, 3, 300, 18)  # 300 same-key points inside one chunk, far beyond M
    k = len(filler) // 3
    c["crossing_in_later_chunk"] = ([np.concatenate([a_early, b_early, filler[:k], a_late[:30], filler[k:2 * k], a_late[30:], b_late,
                                                     c_late, filler[2 * k:]])], {})
    # B.9 extremes of the density ablation
    c["M4"] = ([po.synth_tile(3000, 21), po.synth_tile(100, 22)], dict(max_num_points=4))
    c["M512"] = ([po.synth_tile(30000, 23, clustered=True)], dict(max_num_points=512))
    c["M16"] = ([po.synth_tile(9000, 24)], dict(max_num_points=16))
    c["M128"] = ([po.synth_tile(60000, 25, clustered=True)], dict(max_num_points=128))
    # ragged batch like a real loader batch (truncated last batch: B = 3)
    c["ragged_batch"] = ([po.synth_tile(7000, 31), po.synth_tile(1, 32), po.synth_tile(15000, 33, clustered=True)], {})
    # demo-shaped tile (39 641 points, extent 223.92 px, one point at y == 224)
    d = po.synth_tile(39641, 34)
    d[:, :2] *= F(223.92 / 224.0)
    d[17, 1] = 224.0
    c["demo_shaped"] = ([d], {})
    # the alternative reading of the invalid-hash guard (ledger U1)
    c["drop_overflow"] = ([np.concatenate([pillar_block(5, 6, 10, 1), _t([[44.0, 52.0, 100.0], [224.0, 3.0, 3.0], [3.0, 224.0, 3.0]])])],
                          dict(drop_overflow=True))
    return c


def grid_for(kwargs) -> po.GridSpec:
    return po.GridSpec(**kwargs)
