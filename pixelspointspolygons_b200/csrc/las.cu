// las.cu -- LiDAR input front end on the GPU: raw LAS integer coordinates -> the pixel-space (N, 3) fp32 points the
// encoder consumes (sm_100a).
//
// Replaces the numpy / scikit-learn body of `P3Dataset.load_lidar_points`
// (R:pixelspointspolygons/datasets/p3_coco.py:74-101) and of `Predictor.load_lidar_from_file`
// (R:pixelspointspolygons/predict/predictor.py:116-137) -- SURVEY 8a row a1 / 8f rank 3 -- for a whole jagged batch:
//   xyz64   = XYZ_int * las.header.scales + las.header.offsets               (laspy's ScaledArrayView, float64)
//   x       = (x64 - left) / res;   y = height - (y64 - top) / res           (float64)
//   z       = z64 * scale_ + min_   with sklearn's MinMaxScaler(feature_range = (0, z_hi)) fitted on the tile:
//             scale_ = z_hi / (zmax - zmin) (range below 10 eps -> 1), min_ = 0 - zmin * scale_
//   points  = float32(x, y, z);  dataset variant: x, y clipped to [0, width] / [0, height]; optionally the replayed D4
//             element of the training augmentation on x, y about the tile centre (float32, p3_coco.py:114-160)
// Every float64 operation is a single IEEE operation in the reference's order (no FMA contraction), so the result is
// bit-identical to the numpy path.  The tile's z extremes (and x / y minima for the predictor variant, whose origin is
// the tile's minimum) are taken over the integers: the int -> float64 map is monotone for positive scales.
#include "p3p_internal.cuh"

namespace p3p {
namespace {

constexpr int kLasThreads = 256;

// mm: [B][4] = min X, min Y, min Z, max Z
__global__ void las_init_kernel(int32_t* mm, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 4) mm[i] = ((i & 3) == 3) ? INT32_MIN : INT32_MAX;
}

__device__ __forceinline__ int tile_of(const int64_t* __restrict__ offsets, int B, int64_t i) {
    int lo = 0, hi = B;  // offsets[lo] <= i < offsets[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (offsets[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// Where the raw integer coordinates of point i (tile b) come from: three int32 arrays (las.X / las.Y / las.Z), or the
// packed transfer format -- per tile an int32 base and per point three uint16 deltas (6 bytes per point instead of 12: a
// 56 m tile at the usual 1 mm .. 1 cm LAS scale spans < 65536 steps) -- which halves the host -> device copy.
struct LasSource {
    const int32_t* X;
    const int32_t* Y;
    const int32_t* Z;
    const uint16_t* d;     // (total, 3) deltas, or NULL
    const int32_t* base;   // (B, 3)
    __device__ __forceinline__ void get(int64_t i, int b, int& x, int& y, int& z) const {
        if (d) {
            const uint16_t* p = d + i * 3;
            x = base[b * 3 + 0] + (int)p[0]; y = base[b * 3 + 1] + (int)p[1]; z = base[b * 3 + 2] + (int)p[2];
        } else {
            x = X[i]; y = Y[i]; z = Z[i];
        }
    }
};

__global__ void __launch_bounds__(kLasThreads)
las_minmax_kernel(LasSource src, const int64_t* __restrict__ offsets, int B, int64_t total, int32_t* mm) {
    for (int64_t i0 = (int64_t)blockIdx.x * kLasThreads; i0 < total; i0 += (int64_t)gridDim.x * kLasThreads) {
        const int64_t i = i0 + threadIdx.x;
        int b = -1, x = INT32_MAX, y = INT32_MAX, z0 = INT32_MAX, z1 = INT32_MIN;
        if (i < total) {
            b = tile_of(offsets, B, i);
            src.get(i, b, x, y, z0);
            z1 = z0;
        }
        // a warp's 32 consecutive points mostly share a tile: reduce over the lanes of the first lane's tile, the
        // others (a tile boundary inside the warp) go straight to the atomics
        const int b0 = __shfl_sync(0xffffffffu, b, 0);
        const bool same = (b == b0);
        const int rx = __reduce_min_sync(0xffffffffu, same ? x : INT32_MAX), ry = __reduce_min_sync(0xffffffffu, same ? y : INT32_MAX);
        const int rz0 = __reduce_min_sync(0xffffffffu, same ? z0 : INT32_MAX), rz1 = __reduce_max_sync(0xffffffffu, same ? z1 : INT32_MIN);
        if ((threadIdx.x & 31) == 0 && b0 >= 0) {
            atomicMin(mm + b0 * 4 + 0, rx); atomicMin(mm + b0 * 4 + 1, ry);
            atomicMin(mm + b0 * 4 + 2, rz0); atomicMax(mm + b0 * 4 + 3, rz1);
        }
        if (b >= 0 && !same) {
            atomicMin(mm + b * 4 + 0, x); atomicMin(mm + b * 4 + 1, y);
            atomicMin(mm + b * 4 + 2, z0); atomicMax(mm + b * 4 + 3, z1);
        }
    }
}

__global__ void __launch_bounds__(kLasThreads)
las_pixels_kernel(LasSource src, const int64_t* __restrict__ offsets, int B, int64_t total, const p3p_las_tile* __restrict__ tiles,
                  double z_hi, const int32_t* __restrict__ mm, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * kLasThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kLasThreads) {
        const int b = tile_of(offsets, B, i);
        const p3p_las_tile t = tiles[b];
        int Xi, Yi, Zi;
        src.get(i, b, Xi, Yi, Zi);
        const double x64 = __dadd_rn(__dmul_rn((double)Xi, t.scale[0]), t.offset[0]);
        const double y64 = __dadd_rn(__dmul_rn((double)Yi, t.scale[1]), t.offset[1]);
        const double z64 = __dadd_rn(__dmul_rn((double)Zi, t.scale[2]), t.offset[2]);
        double left = t.left, top = t.top;
        if (t.origin_from_min) {  // predictor.py:126: the tile's own minimum is the origin
            left = __dadd_rn(__dmul_rn((double)mm[b * 4 + 0], t.scale[0]), t.offset[0]);
            top = __dadd_rn(__dmul_rn((double)mm[b * 4 + 1], t.scale[1]), t.offset[1]);
        }
        const double px = __ddiv_rn(__dsub_rn(x64, left), t.res);
        const double py = __dsub_rn(t.height, __ddiv_rn(__dsub_rn(y64, top), t.res));
        // MinMaxScaler(feature_range = (0, z_hi)).fit_transform on the tile's z
        const double zmin = __dadd_rn(__dmul_rn((double)mm[b * 4 + 2], t.scale[2]), t.offset[2]);
        const double zmax = __dadd_rn(__dmul_rn((double)mm[b * 4 + 3], t.scale[2]), t.offset[2]);
        double range = __dsub_rn(zmax, zmin);
        if (range < 10.0 * 2.220446049250313e-16) range = 1.0;  // sklearn _handle_zeros_in_scale
        const double scale = __ddiv_rn(__dsub_rn(z_hi, 0.0), range);
        const double zoff = __dsub_rn(0.0, __dmul_rn(zmin, scale));
        const double pz = __dadd_rn(__dmul_rn(z64, scale), zoff);
        float fx = __double2float_rn(px), fy = __double2float_rn(py);
        const float fz = __double2float_rn(pz);
        if (t.clip) {  // p3_coco.py:95-96 (np.clip on the float32 array)
            fx = fminf(fmaxf(fx, 0.f), (float)t.width);
            fy = fminf(fmaxf(fy, 0.f), (float)t.height);
        }
        if (t.d4 != P3P_D4_NONE) {  // apply_d4_augmentations_to_lidar (p3_coco.py:114-160), float32 like the reference
            const float cx = (float)t.center_x, cy = (float)t.center_y;
            const float ax = fx - cx, ay = fy - cy;
            float bx = ax, by = ay;
            switch (t.d4) {
                case P3P_D4_R90: bx = ay; by = -ax; break;    // swap, then y = -y
                case P3P_D4_R180: bx = -ax; by = -ay; break;
                case P3P_D4_R270: bx = -ay; by = ax; break;   // swap, then x = -x
                case P3P_D4_V: by = -ay; break;
                case P3P_D4_HVT: bx = -ay; by = -ax; break;   // swap, then both negated
                case P3P_D4_H: bx = -ax; break;
                case P3P_D4_T: bx = ay; by = ax; break;
                default: break;
            }
            fx = bx + cx; fy = by + cy;
        }
        out[i * 3 + 0] = fx; out[i * 3 + 1] = fy; out[i * 3 + 2] = fz;
    }
}

}  // namespace

int launch_las_to_pixels(const int32_t* X, const int32_t* Y, const int32_t* Z, const uint16_t* deltas, const int32_t* base,
                         const int64_t* offsets, int B, int64_t total, const p3p_las_tile* tiles, double z_hi, int32_t* mm, float* out,
                         cudaStream_t st) {
    if (B <= 0 || total <= 0) return P3P_OK;
    LasSource src;
    src.X = X; src.Y = Y; src.Z = Z; src.d = deltas; src.base = base;
    las_init_kernel<<<(B * 4 + 127) / 128, 128, 0, st>>>(mm, B);
    P3P_CUDA_CHECK(cudaGetLastError());
    const int sms = device_sm_count();
    int64_t grid = (total + kLasThreads - 1) / kLasThreads;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    las_minmax_kernel<<<(unsigned)grid, kLasThreads, 0, st>>>(src, offsets, B, total, mm);
    P3P_CUDA_CHECK(cudaGetLastError());
    las_pixels_kernel<<<(unsigned)grid, kLasThreads, 0, st>>>(src, offsets, B, total, tiles, z_hi, mm, out);
    P3P_CUDA_CHECK(cudaGetLastError());
    return P3P_OK;
}

}  // namespace p3p
