/*
 * oracle/voxelize_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the CPU voxelizer that the reference's LiDAR encoder
 * calls through `open3d.ml.torch.ops.voxelize`.
 *
 * PARITY UNPINNED: the arithmetic lives in the third-party wheel
 * open3d==0.19.0 (R:pyproject.toml:23), which is neither vendored under
 * /root/reference nor installable offline.  This file restates the published
 * algorithm of upstream Open3D `cpp/open3d/ml/impl/misc/Voxelize.h`
 * (`VoxelizeCPU<T, NDIM=3>`), anchored on the reference's call site
 *   R:pixelspointspolygons/models/pointpillars/pointpillars_o3d.py:92
 *   (self.voxelize(x_lidar) -> Open3D-ML PointPillarsVoxelization.forward ->
 *    voxelize(points, row_splits=[0,N], voxel_size, range_min, range_max,
 *             max_num_points, max_voxels))
 * and on SURVEY.md Appendix A.1 / B.  The reference itself has no tests or
 * golden vectors for this path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library.
 *
 * Semantics restated (Appendix A.1):
 *   inv_i      = fp32(1) / voxel_size_i
 *   extents_i  = int32(ceil((max_i - min_i) * inv_i))
 *   strides    = (1, extents_0, extents_0 * extents_1)
 *   batch_hash = strides_2 * extents_2  (= number of regular cells, 784 at the default grid)
 *   valid(p)   = all_i(min_i <= p_i && p_i <= max_i)       (inclusive, NaN -> false)
 *   c_i        = (int64) trunc((p_i - min_i) * inv_i)      (fp32 sub, fp32 mul, no fma)
 *   hash       = sum_i c_i * strides_i (+ batch * batch_hash)
 *   order      = ascending (hash, original index)          (pair sort == stable sort by hash)
 *   run r of equal hash is voxel r iff r < max_voxels; first
 *   min(count, max_points_per_voxel) indices of the run are kept;
 *   voxel_coords[r] = c(point with the lowest index of the run), order (x, y, z).
 *
 * Hashes >= batch_hash.  A point sitting exactly on range_max_i has c_i == extents_i, so its
 * hash can reach sum_i extents_i * strides_i (1596 at the default grid) -- beyond upstream's
 * `invalid_hash = batch_hash * batch_size` (784).  SURVEY Appendix A.1 / B.1 / B.2 (our
 * blueprint) treat those as ordinary runs that sort after the regular cells (a z == 100 point
 * forms a z-cell-1 pillar; x == 224 aliases cell (0, y+1)); out-of-range points are a separate
 * class that never collides with them.  That is the default here (flags == 0): the invalid
 * marker is INT64_MAX and `out_point_hash` reports -1.  flags & P3P_ORACLE_DROP_OVERFLOW gives
 * the alternative reading (upstream's walk stops at the first hash >= invalid_hash, so every
 * hash >= batch_hash is dropped like an out-of-range point).  Which one open3d 0.19.0 really
 * does cannot be settled offline -- DESIGN.md "Uncertainty ledger" U1.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; no fast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define P3P_ORACLE_DROP_OVERFLOW 1
#define P3P_ORACLE_INVALID INT64_MAX

typedef struct {
    int64_t hash;
    int64_t index;
} hash_index_t;

static int cmp_hash_index(const void* a, const void* b) {
    const hash_index_t* x = (const hash_index_t*)a;
    const hash_index_t* y = (const hash_index_t*)b;
    if (x->hash != y->hash) return x->hash < y->hash ? -1 : 1;
    if (x->index != y->index) return x->index < y->index ? -1 : 1;
    return 0;
}

/* Cell coordinates of one point; returns 0 when the point is out of range. */
static int point_cell(const float* p, const float* vmin, const float* vmax,
                      const float* inv, int64_t c[3]) {
    for (int i = 0; i < 3; ++i) {
        /* written so that NaN fails the test, like `(p >= min && p <= max).all()` */
        if (!(p[i] >= vmin[i] && p[i] <= vmax[i])) return 0;
    }
    for (int i = 0; i < 3; ++i) {
        volatile float d = p[i] - vmin[i]; /* volatile: forbid contraction / excess precision */
        volatile float s = d * inv[i];
        c[i] = (int64_t)s;
    }
    return 1;
}

/*
 * One sample (batch_size == 1), exactly how the reference path calls the op
 * (once per tile with row_splits = [0, N]).
 *
 * points             (num_points, point_stride) fp32, xyz in the first 3 lanes
 * flags              0 or P3P_ORACLE_DROP_OVERFLOW
 * out_point_hash     optional (num_points) int64: hash per point, -1 if out of range / dropped
 * out_voxel_coords   (>= max_voxels, 3) int32 (x, y, z)
 * out_point_indices  (>= num_points) int64
 * out_row_splits     (>= max_voxels + 1) int64
 * out_counts[0] = number of voxels, out_counts[1] = number of kept indices,
 * out_counts[2] = batch_hash (regular cells), out_counts[3..5] = extents
 */
int p3p_oracle_voxelize(const float* points, int64_t num_points, int64_t point_stride,
                        const float* voxel_size, const float* range_min, const float* range_max,
                        int64_t max_points_per_voxel, int64_t max_voxels, int64_t flags,
                        int64_t* out_point_hash, int32_t* out_voxel_coords,
                        int64_t* out_point_indices, int64_t* out_row_splits, int64_t* out_counts) {
    float inv[3];
    int64_t extents[3], strides[3];
    for (int i = 0; i < 3; ++i) {
        volatile float one_over = 1.0f / voxel_size[i];
        inv[i] = one_over;
        volatile float span = range_max[i] - range_min[i];
        volatile float cells = span * inv[i];
        extents[i] = (int64_t)(int32_t)ceilf(cells);
    }
    strides[0] = 1;
    strides[1] = extents[0];
    strides[2] = extents[0] * extents[1];
    const int64_t batch_hash = strides[2] * extents[2];
    const int64_t invalid_hash = P3P_ORACLE_INVALID;

    hash_index_t* hi = (hash_index_t*)malloc((size_t)(num_points > 0 ? num_points : 1) * sizeof(hash_index_t));
    if (!hi) return -1;
    for (int64_t idx = 0; idx < num_points; ++idx) {
        int64_t c[3];
        int64_t h = invalid_hash;
        if (point_cell(points + idx * point_stride, range_min, range_max, inv, c)) {
            h = c[0] * strides[0] + c[1] * strides[1] + c[2] * strides[2];
            if ((flags & P3P_ORACLE_DROP_OVERFLOW) && h >= batch_hash) h = invalid_hash;
        }
        hi[idx].hash = h;
        hi[idx].index = idx;
        if (out_point_hash) out_point_hash[idx] = (h == invalid_hash) ? -1 : h;
    }
    qsort(hi, (size_t)num_points, sizeof(hash_index_t), cmp_hash_index);

    int64_t num_voxels = 0, num_indices = 0;
    out_row_splits[0] = 0;
    int64_t i = 0;
    while (i < num_points && hi[i].hash != invalid_hash) {
        int64_t j = i;
        while (j < num_points && hi[j].hash == hi[i].hash) ++j;
        if (num_voxels < max_voxels) {
            int64_t keep = j - i;
            if (keep > max_points_per_voxel) keep = max_points_per_voxel;
            for (int64_t k = 0; k < keep; ++k) out_point_indices[num_indices + k] = hi[i + k].index;
            num_indices += keep;
            int64_t c[3];
            point_cell(points + hi[i].index * point_stride, range_min, range_max, inv, c);
            out_voxel_coords[num_voxels * 3 + 0] = (int32_t)c[0];
            out_voxel_coords[num_voxels * 3 + 1] = (int32_t)c[1];
            out_voxel_coords[num_voxels * 3 + 2] = (int32_t)c[2];
            ++num_voxels;
            out_row_splits[num_voxels] = num_indices;
        }
        i = j;
    }
    free(hi);
    out_counts[0] = num_voxels;
    out_counts[1] = num_indices;
    out_counts[2] = batch_hash;
    out_counts[3] = extents[0];
    out_counts[4] = extents[1];
    out_counts[5] = extents[2];
    return 0;
}

/*
 * Dense gather of Open3D-ML PointPillarsVoxelization.forward after the op
 * (SURVEY Appendix A.2): ragged_to_dense(indices, row_splits, M, -1) + 1 into
 * feats = cat([zeros(1,3), points]); coords reordered (z, y, x); per-voxel
 * count; x/y out-of-bounds filter with num_voxels = int32((max - min) / vs).
 * Returns the number of pillars that survive the filter.
 *
 * out_voxels (>= max_voxels, M, 3) fp32, out_coords (>= max_voxels, 3) int32
 * (z, y, x), out_num_points (>= max_voxels) int64,
 * out_dense_idx optional (>= max_voxels, M) int64 with -1 padding.
 */
int64_t p3p_oracle_pillarize(const float* points, int64_t num_points, int64_t point_stride,
                             const float* voxel_size, const float* range_min, const float* range_max,
                             int64_t max_points_per_voxel, int64_t max_voxels, int64_t flags,
                             float* out_voxels, int32_t* out_coords, int64_t* out_num_points,
                             int64_t* out_dense_idx) {
    const int64_t M = max_points_per_voxel;
    int32_t* coords = (int32_t*)malloc((size_t)(max_voxels > 0 ? max_voxels : 1) * 3 * sizeof(int32_t));
    int64_t* indices = (int64_t*)malloc((size_t)(num_points > 0 ? num_points : 1) * sizeof(int64_t));
    int64_t* splits = (int64_t*)malloc((size_t)(max_voxels + 1) * sizeof(int64_t));
    int64_t counts[6];
    if (!coords || !indices || !splits) return -1;
    if (p3p_oracle_voxelize(points, num_points, point_stride, voxel_size, range_min, range_max, M,
                            max_voxels, flags, NULL, coords, indices, splits, counts) != 0)
        return -1;
    int32_t nv[3];
    for (int i = 0; i < 3; ++i) {
        volatile float span = range_max[i] - range_min[i];
        volatile float q = span / voxel_size[i];
        nv[i] = (int32_t)q;
    }
    int64_t kept = 0;
    for (int64_t v = 0; v < counts[0]; ++v) {
        const int32_t cx = coords[v * 3 + 0], cy = coords[v * 3 + 1], cz = coords[v * 3 + 2];
        if (!(cy < nv[1] && cx < nv[0])) continue;
        const int64_t n = splits[v + 1] - splits[v];
        for (int64_t s = 0; s < M; ++s) {
            float* dst = out_voxels + (kept * M + s) * 3;
            if (s < n) {
                const float* src = points + indices[splits[v] + s] * point_stride;
                dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
                if (out_dense_idx) out_dense_idx[kept * M + s] = indices[splits[v] + s];
            } else {
                dst[0] = 0.0f; dst[1] = 0.0f; dst[2] = 0.0f;
                if (out_dense_idx) out_dense_idx[kept * M + s] = -1;
            }
        }
        out_coords[kept * 3 + 0] = cz;
        out_coords[kept * 3 + 1] = cy;
        out_coords[kept * 3 + 2] = cx;
        out_num_points[kept] = n;
        ++kept;
    }
    free(coords); free(indices); free(splits);
    return kept;
}
