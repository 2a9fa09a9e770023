"""CPU side of the randomised parity cases (tests/test_gpu_fuzz.py): the two restatements of the reference voxelizer -- plain
C (oracle/voxelize_ref.c) and numpy (oracle/pillars_oracle.py) -- must agree on every draw, and the oracle module must run
them (what the GPU test then compares the kernels with)."""
import numpy as np
import pytest

from oracle import pillars_oracle as po
from test_gpu_fuzz import draw_case


@pytest.mark.parametrize("seed", range(24))
def test_c_and_numpy_restatements_agree_on_fuzz_cases(seed):
    grid, tiles, n_side = draw_case(seed)
    for t in tiles:
        for mv in grid.max_voxels:
            c = po.voxelize_c(t, grid, mv)
            n = po.voxelize_numpy(t, grid, mv)
            for k in c:
                assert np.array_equal(np.asarray(c[k]), np.asarray(n[k])), (seed, k)
    ref = po.OraclePointPillarsEncoder(grid).eval()
    rv, rn, rc, rd = ref.voxelize(tiles)
    assert rv.shape[0] == rn.shape[0] == rc.shape[0] == rd.shape[0]
    assert int(rn.max() if rn.numel() else 0) <= grid.max_num_points
