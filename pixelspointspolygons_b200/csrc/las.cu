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
//
// Layout of the work: a CTA takes segments of 2048 consecutive points, a thread 8 consecutive points (three 16-byte loads
// of packed deltas or two per int32 array, six 16-byte stores).  A segment almost always lies inside one tile: the tile's
// derived constants (float64 origin, z scale / offset: the only divisions that do not depend on the point) are computed
// once per segment, the tile extremes are reduced inside the CTA before ONE set of four global atomics per segment
// (per-warp atomics on the 4 words of a tile serialise in L2: 3 000 of them per address cost 150 us at B = 16 x 100 k).
// Segments that straddle a tile boundary (at most B - 1 of them) take a per-point path.
#include "p3p_internal.cuh"

namespace p3p {
namespace {

constexpr int kLasThreads = 256;
constexpr int kPer = 8;                     // consecutive points per thread
constexpr int kSeg = kLasThreads * kPer;    // points per segment

// mm: [B][4] unsigned words under the order-preserving map u(v) = (unsigned)v ^ 0x80000000: min u(X), min u(Y),
// min u(Z), min ~u(Z) (= the maximum of Z); all four shrink under atomicMin and start at 0xFFFFFFFF (one memset).
__device__ __forceinline__ unsigned ukey(int v) { return (unsigned)v ^ 0x80000000u; }
__device__ __forceinline__ int ikey(unsigned u) { return (int)(u ^ 0x80000000u); }

__device__ __forceinline__ int tile_of(const int64_t* __restrict__ offsets, int B, int64_t i) {
    int lo = 0, hi = B;  // offsets[lo] <= i < offsets[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (offsets[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// Where the raw integer coordinates come from: three int32 arrays (las.X / las.Y / las.Z), or the packed transfer format --
// per tile an int32 base and per point three uint16 deltas (6 bytes per point instead of 12: a 56 m tile at the usual
// 1 mm .. 1 cm LAS scale spans < 65536 steps) -- which halves the host -> device copy.
struct LasSource {
    const int32_t* X;
    const int32_t* Y;
    const int32_t* Z;
    const uint16_t* d;     // (total, 3) deltas, or NULL
    const int32_t* base;   // (B, 3)
    __device__ __forceinline__ void tile_base(int b, int& bx, int& by, int& bz) const {
        bx = by = bz = 0;
        if (d) { bx = base[b * 3 + 0]; by = base[b * 3 + 1]; bz = base[b * 3 + 2]; }
    }
    // raw words (deltas, or the coordinates themselves) of the 8 consecutive points from i0 (a multiple of 8) that lie
    // before `end`; the tile's base is added by the caller.  All 8 inside and 16-byte aligned arrays: 16-byte loads.
    __device__ __forceinline__ void get8(int64_t i0, int64_t end, bool vec, int (&x)[kPer], int (&y)[kPer], int (&z)[kPer]) const {
        if (vec && i0 + kPer <= end) {
            if (d) {
                const uint4* q = reinterpret_cast<const uint4*>(d + i0 * 3);
                const uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2);
                const unsigned w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
                for (int e = 0; e < kPer; ++e) {  // halfword 3 e + c of the 24
                    const int h0 = 3 * e, h1 = 3 * e + 1, h2 = 3 * e + 2;
                    x[e] = (int)((w[h0 >> 1] >> (16 * (h0 & 1))) & 0xFFFFu);
                    y[e] = (int)((w[h1 >> 1] >> (16 * (h1 & 1))) & 0xFFFFu);
                    z[e] = (int)((w[h2 >> 1] >> (16 * (h2 & 1))) & 0xFFFFu);
                }
            } else {
                const int4 xa = __ldg(reinterpret_cast<const int4*>(X + i0)), xb = __ldg(reinterpret_cast<const int4*>(X + i0) + 1);
                const int4 ya = __ldg(reinterpret_cast<const int4*>(Y + i0)), yb = __ldg(reinterpret_cast<const int4*>(Y + i0) + 1);
                const int4 za = __ldg(reinterpret_cast<const int4*>(Z + i0)), zb = __ldg(reinterpret_cast<const int4*>(Z + i0) + 1);
                x[0] = xa.x; x[1] = xa.y; x[2] = xa.z; x[3] = xa.w; x[4] = xb.x; x[5] = xb.y; x[6] = xb.z; x[7] = xb.w;
                y[0] = ya.x; y[1] = ya.y; y[2] = ya.z; y[3] = ya.w; y[4] = yb.x; y[5] = yb.y; y[6] = yb.z; y[7] = yb.w;
                z[0] = za.x; z[1] = za.y; z[2] = za.z; z[3] = za.w; z[4] = zb.x; z[5] = zb.y; z[6] = zb.z; z[7] = zb.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < kPer; ++e) {
                x[e] = y[e] = z[e] = 0;
                if (i0 + e < end) {
                    if (d) {
                        const uint16_t* p = d + (i0 + e) * 3;
                        x[e] = (int)p[0]; y[e] = (int)p[1]; z[e] = (int)p[2];
                    } else {
                        x[e] = X[i0 + e]; y[e] = Y[i0 + e]; z[e] = Z[i0 + e];
                    }
                }
            }
        }
    }
    __device__ __forceinline__ bool vector_ok() const {  // 16-byte alignment of the arrays (the ABI asks for it; checked all the same)
        if (d) return (reinterpret_cast<uintptr_t>(d) & 15u) == 0;
        return ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(Z)) & 15u) == 0;
    }
};

// A segment is walked tile by tile (uniform loop; one tile for all but the <= B - 1 segments that hold a tile boundary):
// points of the segment that belong to tile b are those with index in [lo, hi).
__global__ void __launch_bounds__(kLasThreads)
las_minmax_kernel(LasSource src, const int64_t* __restrict__ offsets, int B, int64_t total, unsigned* mm) {
    __shared__ unsigned red[kLasThreads / 32][4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool vec = src.vector_ok();
    const int64_t nseg = (total + kSeg - 1) / kSeg;
    for (int64_t seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
        const int64_t s0 = seg * kSeg, s1 = (s0 + kSeg < total) ? s0 + kSeg : total;
        const int b0 = tile_of(offsets, B, s0), b1 = tile_of(offsets, B, s1 - 1);  // (uniform)
        const int64_t i0 = s0 + (int64_t)tid * kPer;
        int x[kPer], y[kPer], z[kPer];
        src.get8(i0, s1, vec, x, y, z);
        for (int b = b0; b <= b1; ++b) {
            const int64_t lo = offsets[b] > s0 ? offsets[b] : s0, hi = offsets[b + 1] < s1 ? offsets[b + 1] : s1;
            if (hi <= lo) continue;  // (uniform: an empty tile between two others)
            int bx, by, bz;
            src.tile_base(b, bx, by, bz);
            unsigned m[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
#pragma unroll
            for (int e = 0; e < kPer; ++e) {
                if (i0 + e >= lo && i0 + e < hi) {
                    m[0] = min(m[0], ukey(bx + x[e])); m[1] = min(m[1], ukey(by + y[e]));
                    m[2] = min(m[2], ukey(bz + z[e])); m[3] = min(m[3], ~ukey(bz + z[e]));
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) m[k] = __reduce_min_sync(0xffffffffu, m[k]);
            __syncthreads();  // (red of the previous round has been read)
            if (lane == 0) { red[warp][0] = m[0]; red[warp][1] = m[1]; red[warp][2] = m[2]; red[warp][3] = m[3]; }
            __syncthreads();
            if (tid < 4) {
                unsigned v = red[0][tid];
#pragma unroll
                for (int w = 1; w < kLasThreads / 32; ++w) v = min(v, red[w][tid]);
                atomicMin(mm + b * 4 + tid, v);
            }
        }
    }
}

// Everything of a tile that does not depend on the point, in the reference's float64 operation order.
struct TileConsts {
    double sx, sy, sz, ox, oy, oz;
    double left, top, res, inv_res, height;
    double zscale, zoff;
    float w, h, cx, cy;
    int clip, d4, res_pow2;
};

__device__ __forceinline__ TileConsts derive(const p3p_las_tile& t, const unsigned* __restrict__ mm_b, double z_hi) {
    TileConsts c;
    c.sx = t.scale[0]; c.sy = t.scale[1]; c.sz = t.scale[2];
    c.ox = t.offset[0]; c.oy = t.offset[1]; c.oz = t.offset[2];
    c.left = t.left; c.top = t.top;
    if (t.origin_from_min) {  // predictor.py:126: the tile's own minimum is the origin
        c.left = __dadd_rn(__dmul_rn((double)ikey(mm_b[0]), t.scale[0]), t.offset[0]);
        c.top = __dadd_rn(__dmul_rn((double)ikey(mm_b[1]), t.scale[1]), t.offset[1]);
    }
    c.res = t.res; c.height = t.height;
    // x / res == x * (1 / res) bit for bit when res is a power of two (both are the correctly rounded value of the same
    // real number): the usual 0.25 m / pixel takes a multiplication instead of a float64 division per coordinate
    const long long rb = __double_as_longlong(t.res);
    const int ex = (int)((rb >> 52) & 0x7FF);
    c.res_pow2 = ((rb & 0xFFFFFFFFFFFFFll) == 0 && ex > 1 && ex < 2045) ? 1 : 0;
    c.inv_res = c.res_pow2 ? __ddiv_rn(1.0, t.res) : 0.0;
    // MinMaxScaler(feature_range = (0, z_hi)).fit_transform on the tile's z
    const double zmin = __dadd_rn(__dmul_rn((double)ikey(mm_b[2]), t.scale[2]), t.offset[2]);
    const double zmax = __dadd_rn(__dmul_rn((double)ikey(~mm_b[3]), t.scale[2]), t.offset[2]);
    double range = __dsub_rn(zmax, zmin);
    if (range < 10.0 * 2.220446049250313e-16) range = 1.0;  // sklearn _handle_zeros_in_scale
    c.zscale = __ddiv_rn(__dsub_rn(z_hi, 0.0), range);
    c.zoff = __dsub_rn(0.0, __dmul_rn(zmin, c.zscale));
    c.w = (float)t.width; c.h = (float)t.height;
    c.cx = (float)t.center_x; c.cy = (float)t.center_y;
    c.clip = t.clip; c.d4 = t.d4;
    return c;
}

__device__ __forceinline__ void to_pixels(const TileConsts& c, int Xi, int Yi, int Zi, float& fx, float& fy, float& fz) {
    const double x64 = __dadd_rn(__dmul_rn((double)Xi, c.sx), c.ox);
    const double y64 = __dadd_rn(__dmul_rn((double)Yi, c.sy), c.oy);
    const double z64 = __dadd_rn(__dmul_rn((double)Zi, c.sz), c.oz);
    const double dx = __dsub_rn(x64, c.left), dy = __dsub_rn(y64, c.top);
    const double px = c.res_pow2 ? __dmul_rn(dx, c.inv_res) : __ddiv_rn(dx, c.res);
    const double py = __dsub_rn(c.height, c.res_pow2 ? __dmul_rn(dy, c.inv_res) : __ddiv_rn(dy, c.res));
    const double pz = __dadd_rn(__dmul_rn(z64, c.zscale), c.zoff);
    fx = __double2float_rn(px); fy = __double2float_rn(py); fz = __double2float_rn(pz);
    if (c.clip) {  // p3_coco.py:95-96 (np.clip on the float32 array)
        fx = fminf(fmaxf(fx, 0.f), c.w);
        fy = fminf(fmaxf(fy, 0.f), c.h);
    }
    if (c.d4 != P3P_D4_NONE) {  // apply_d4_augmentations_to_lidar (p3_coco.py:114-160), float32 like the reference
        const float ax = fx - c.cx, ay = fy - c.cy;
        float bx = ax, by = ay;
        switch (c.d4) {
            case P3P_D4_R90: bx = ay; by = -ax; break;    // swap, then y = -y
            case P3P_D4_R180: bx = -ax; by = -ay; break;
            case P3P_D4_R270: bx = -ay; by = ax; break;   // swap, then x = -x
            case P3P_D4_V: by = -ay; break;
            case P3P_D4_HVT: bx = -ay; by = -ax; break;   // swap, then both negated
            case P3P_D4_H: bx = -ax; break;
            case P3P_D4_T: bx = ay; by = ax; break;
            default: break;
        }
        fx = bx + c.cx; fy = by + c.cy;
    }
}

__global__ void __launch_bounds__(kLasThreads)
las_pixels_kernel(LasSource src, const int64_t* __restrict__ offsets, int B, int64_t total, const p3p_las_tile* __restrict__ tiles,
                  double z_hi, const unsigned* __restrict__ mm, float* __restrict__ out) {
    __shared__ TileConsts sc;
    const int tid = threadIdx.x;
    const bool vec = src.vector_ok(), vec_out = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    const int64_t nseg = (total + kSeg - 1) / kSeg;
    for (int64_t seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
        const int64_t s0 = seg * kSeg, s1 = (s0 + kSeg < total) ? s0 + kSeg : total;
        const int b0 = tile_of(offsets, B, s0), b1 = tile_of(offsets, B, s1 - 1);  // (uniform)
        const int64_t i0 = s0 + (int64_t)tid * kPer;
        int x[kPer], y[kPer], z[kPer];
        src.get8(i0, s1, vec, x, y, z);
        float f[3 * kPer];
#pragma unroll
        for (int q = 0; q < 3 * kPer; ++q) f[q] = 0.f;
        for (int b = b0; b <= b1; ++b) {
            const int64_t lo = offsets[b] > s0 ? offsets[b] : s0, hi = offsets[b + 1] < s1 ? offsets[b + 1] : s1;
            if (hi <= lo) continue;  // (uniform)
            __syncthreads();  // (sc of the previous round has been read)
            if (tid == 0) sc = derive(tiles[b], mm + b * 4, z_hi);
            __syncthreads();
            const TileConsts c = sc;
            int bx, by, bz;
            src.tile_base(b, bx, by, bz);
#pragma unroll
            for (int e = 0; e < kPer; ++e) {
                if (i0 + e >= lo && i0 + e < hi) to_pixels(c, bx + x[e], by + y[e], bz + z[e], f[3 * e], f[3 * e + 1], f[3 * e + 2]);
            }
        }
        if (vec_out && i0 + kPer <= s1) {
            float4* dst = reinterpret_cast<float4*>(out + i0 * 3);
#pragma unroll
            for (int q = 0; q < 3 * kPer / 4; ++q) dst[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
        } else {
#pragma unroll
            for (int e = 0; e < kPer; ++e) {
                if (i0 + e < s1) { out[(i0 + e) * 3 + 0] = f[3 * e]; out[(i0 + e) * 3 + 1] = f[3 * e + 1]; out[(i0 + e) * 3 + 2] = f[3 * e + 2]; }
            }
        }
    }
}

}  // namespace

int launch_las_to_pixels(const int32_t* X, const int32_t* Y, const int32_t* Z, const uint16_t* deltas, const int32_t* base,
                         const int64_t* offsets, int B, int64_t total, const p3p_las_tile* tiles, double z_hi, int32_t* mm, float* out,
                         cudaStream_t st) {
    if (B <= 0 || total <= 0) return P3P_OK;
    LasSource src;
    src.X = X; src.Y = Y; src.Z = Z; src.d = deltas; src.base = base;
    unsigned* mmu = reinterpret_cast<unsigned*>(mm);
    P3P_CUDA_CHECK(cudaMemsetAsync(mmu, 0xFF, sizeof(unsigned) * 4 * (size_t)B, st));
    const int sms = device_sm_count();
    int64_t grid = (total + kSeg - 1) / kSeg;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    las_minmax_kernel<<<(unsigned)grid, kLasThreads, 0, st>>>(src, offsets, B, total, mmu);
    P3P_CUDA_CHECK(cudaGetLastError());
    las_pixels_kernel<<<(unsigned)grid, kLasThreads, 0, st>>>(src, offsets, B, total, tiles, z_hi, mmu, out);
    P3P_CUDA_CHECK(cudaGetLastError());
    return P3P_OK;
}

}  // namespace p3p
