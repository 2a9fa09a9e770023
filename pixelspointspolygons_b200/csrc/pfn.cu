// pfn.cu -- PillarFeatureNet (decorate + Linear/BN/ReLU/max x2) fused with the BEV scatter (sm_100a).
//
// Replaces `self.voxel_encoder(...)` + `self.middle_encoder(...)` of the reference
// (R:pixelspointspolygons/models/pointpillars/pointpillars_o3d.py:93-95; Open3D-ML PillarFeatureNet,
// PFNLayer, PointPillarsScatter) and, for the fusion encoders, the
// LiDAR half of `torch.cat((x_image, x_lidar), 1)` (R:.../fusion_layers/early_fusion_vit.py:121).
//
// Eval-mode closed form (Appendix A.4), a = gamma / sqrt(var + eps), b = beta - mean * a folded into the weights:
//   h_p   = relu(W0' d_p + b0)                      per kept point p (d_p: 8 decorated channels)
//   h_pad = relu(b0)                                every padded slot (present iff n < M)
//   hmax  = max over the M slots of h
//   o_p   = relu(W1a' h_p + W1b' hmax + b1),  out = max over the M slots of o
//         = relu(max_p(W1a' h_p) + W1b' hmax + b1)  (relu and "+ const" are monotone; the BN scale is already
//                                                    inside W1a', so its sign does not matter)
//
// pfn_tc_kernel (M == 64, C <= 384): per CTA a persistent pipeline
//   4 front-end warps  : one pillar per warp at a time; layer 0 in its affine form
//                        W0' d_p + b0 = Ux x' + Uy y' + Uz z + kappa(pillar)   (x' = x - centre_x, ...)
//                        -> 3 FMA per (point, channel); rows written (tf32 / bf16) straight into the swizzled
//                        K-major B-operand tile in shared memory; hmax via redux.sync.max.f32
//   1 MMA thread       : per pillar pair and 128-channel tile: D[128 ch x 128 pts] = W1a' H^T and
//                        D[128 ch x 16] = W1b' hmax^T, tcgen05.mma, accumulators in TMEM (3 stages)
//   12 epilogue warps  : warp group m owns channel tile m: tcgen05.ld, max over the pillar's 64 columns with
//                        3-input max, + G + b1, relu, then 32-byte NCHW sectors or coalesced NLC rows.
// The decorated (V, M, 8) tensor and every (V, M, *) intermediate of the reference never exist in HBM.
//
// pfn_simt_kernel: exact fp32 FMA, literal 8-channel formulation, any M and C -- the GPU-side cross-check of the
// tensor-core path and the route for configurations the tensor-core kernel does not cover (density ablation).
#include "p3p_internal.cuh"

namespace p3p {

void make_blob_layout(int C, BlobLayout* out) {
    BlobLayout l;
    l.C = C;
    l.Cpad = (C + 127) / 128 * 128;
    l.MT = l.Cpad / 128;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 1023) / 1024 * 1024;
        return o;
    };
    l.off_header = take(64);
    l.off_front = take(10 * 32 * 4);
    l.off_w0 = take(32 * 8 * 4);
    l.off_b0 = take(32 * 4);
    l.off_w1 = take((size_t)l.Cpad * 64 * 4);
    l.off_b1 = take((size_t)l.Cpad * 4);
    l.off_a1 = take((size_t)l.Cpad * 128);  // sized for tf32 rows (128 B); bf16 uses half
    l.off_a2 = take((size_t)l.Cpad * 128);
    l.total_bytes = off;
    *out = l;
}

namespace {

// ------------------------------------------------------------------------------------------------
// weight preparation
// ------------------------------------------------------------------------------------------------
__global__ void pfn_prepare_kernel(p3p_pfn_params p, int precision, char* blob, BlobLayout bl) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    int* header = reinterpret_cast<int*>(blob + bl.off_header);
    float* front = reinterpret_cast<float*>(blob + bl.off_front);
    float* w0 = reinterpret_cast<float*>(blob + bl.off_w0);
    float* b0 = reinterpret_cast<float*>(blob + bl.off_b0);
    float* w1 = reinterpret_cast<float*>(blob + bl.off_w1);
    float* b1 = reinterpret_cast<float*>(blob + bl.off_b1);
    if (tid == 0) {
        header[0] = kBlobMagic; header[1] = precision; header[2] = bl.C; header[3] = bl.Cpad; header[4] = p.center_alias;
    }
    for (int k = tid; k < kC0; k += nthreads) {
        const float a = p.norm0_weight[k] / sqrtf(p.norm0_var[k] + p.eps);
        const float sh = p.norm0_bias[k] - p.norm0_mean[k] * a;
        float w[8];
        for (int i = 0; i < 8; ++i) {
            w[i] = a * p.linear0_weight[k * 8 + i];
            w0[k * 8 + i] = w[i];
        }
        b0[k] = sh;
        // affine form: channels [c0, c1, z, x-mx, y-my, z-mz, x-cx, y-cy]; c0,c1 = (x-cx, y-cy) if alias else (x, y)
        front[0 * 32 + k] = w[0] + w[3] + w[6];          // Ux
        front[1 * 32 + k] = w[1] + w[4] + w[7];          // Uy
        front[2 * 32 + k] = w[2] + w[5];                 // Uz
        front[3 * 32 + k] = p.center_alias ? 0.f : w[0]; // Kcx (multiplies centre_x)
        front[4 * 32 + k] = p.center_alias ? 0.f : w[1]; // Kcy
        front[5 * 32 + k] = w[3];                        // Wmx (multiplies mean_x - centre_x)
        front[6 * 32 + k] = w[4];                        // Wmy
        front[7 * 32 + k] = w[5];                        // Wmz (multiplies mean_z)
        front[8 * 32 + k] = sh;                          // b0
        front[9 * 32 + k] = fmaxf(sh, 0.f);              // h_pad
        if (precision == P3P_PRECISION_TF32) {
            // kind::tf32 reads the upper 19 bits of each fp32 operand (truncation).  Layer 0 is positively homogeneous
            // (relu of an affine map), so scaling all of it by (1 + 2^-12) makes that truncation a round-to-nearest
            // of the unscaled h -- no conversion instruction per value in the kernel.
            const float up = 1.0f + 0x1p-12f;
            for (int r = 0; r < 10; ++r) front[r * 32 + k] *= up;
        }
    }
    const bool tf32 = (precision != P3P_PRECISION_BF16);
    for (int c = tid; c < bl.Cpad; c += nthreads) {
        float a = 0.f, sh = 0.f;
        if (c < bl.C) {
            a = p.norm1_weight[c] / sqrtf(p.norm1_var[c] + p.eps);
            sh = p.norm1_bias[c] - p.norm1_mean[c] * a;
        }
        b1[c] = sh;
        const int t = c >> 7, r = c & 127;
        for (int j = 0; j < 64; ++j) {
            const float v = (c < bl.C) ? a * p.linear1_weight[c * 64 + j] : 0.f;
            w1[c * 64 + j] = v;
            const int k = j & 31;
            char* tile = blob + (j < 32 ? bl.off_a1 : bl.off_a2);
            if (tf32) {
                const size_t o = (size_t)t * 16384 + (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 +
                                 (size_t)(((k >> 2) ^ (r & 7)) * 16) + (size_t)(k & 3) * 4;
                *reinterpret_cast<uint32_t*>(tile + o) = to_tf32(v);
            } else {
                const size_t o = (size_t)t * 8192 + (size_t)(r >> 3) * 512 + (size_t)(r & 7) * 64 +
                                 (size_t)(((k >> 3) ^ ((r >> 1) & 3)) * 16) + (size_t)(k & 7) * 2;
                *reinterpret_cast<unsigned short*>(tile + o) = (unsigned short)(pack_bf16(v, 0.f) & 0xFFFF);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// output addressing shared by both kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_scalar(const PfnArgs& a, int64_t idx, float v) {
    if (a.out_dtype == P3P_DTYPE_F32) {
        static_cast<float*>(a.out)[idx] = v;
    } else {
        static_cast<unsigned short*>(a.out)[idx] = (unsigned short)(pack_bf16(v, 0.f) & 0xFFFF);
    }
}
__device__ __forceinline__ int64_t out_index(const PfnArgs& a, int64_t item, int b, int cell, int c) {
    if (a.item_mode == kItemsCanvas && a.out_layout == P3P_LAYOUT_NCHW)
        return ((int64_t)b * a.c_total + a.c_offset + c) * a.items_per_tile + cell;
    return item * a.bl.C + c;
}

// ------------------------------------------------------------------------------------------------
// exact fp32 kernel (literal formulation)
// ------------------------------------------------------------------------------------------------
constexpr int kSimtThreads = 128;
constexpr int kHStride = 36;  // floats per H row: 16-byte aligned and bank-conflict free for float4 row stores

__global__ void __launch_bounds__(kSimtThreads) pfn_simt_kernel(PfnArgs a) {
    extern __shared__ __align__(16) float smem_f[];
    float* H = smem_f;                                  // [(M + 1)][kHStride]
    float* w0s = H + (size_t)(a.g.M + 1) * kHStride;    // [32][8]
    float* b0s = w0s + 256;                             // [32]
    int* red = reinterpret_cast<int*>(b0s + 32);        // [4][6] fixed-point partial sums
    int* hmax_bits = red + 24;                          // [32]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const float* w0g = reinterpret_cast<const float*>(a.blob + a.bl.off_w0);
    const float* b0g = reinterpret_cast<const float*>(a.blob + a.bl.off_b0);
    const float* w1g = reinterpret_cast<const float*>(a.blob + a.bl.off_w1);
    const float* b1g = reinterpret_cast<const float*>(a.blob + a.bl.off_b1);
    const int alias = reinterpret_cast<const int*>(a.blob + a.bl.off_header)[4];
    for (int i = tid; i < 256; i += kSimtThreads) w0s[i] = w0g[i];
    if (tid < 32) b0s[tid] = b0g[tid];
    __syncthreads();
    const int C = a.bl.C, M = a.g.M;

    for (int64_t item = blockIdx.x; item < a.num_items; item += gridDim.x) {
        const Item it = fetch_item(a, item);
        if (!it.valid) {
            if (a.item_mode == kItemsCanvas)
                for (int c = tid; c < C; c += kSimtThreads) store_scalar(a, out_index(a, item, it.b, it.cell, c), 0.f);
            continue;
        }
        const float4* slot = item_slots(a, it);
        const int n = it.n;
        // cluster mean over the kept points (padded slots are zeros in the reference's sum), on the fixed-point grid
        int sl[3] = {0, 0, 0}, sh[3] = {0, 0, 0};
        for (int r = tid; r < n; r += kSimtThreads) {
            const float4 p = slot[r];
            int lo, hi;
            fix_split(p.x, a.g.fix_scale, lo, hi); sl[0] += lo; sh[0] += hi;
            fix_split(p.y, a.g.fix_scale, lo, hi); sl[1] += lo; sh[1] += hi;
            fix_split(p.z, a.g.fix_scale, lo, hi); sl[2] += lo; sh[2] += hi;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            sl[i] = __reduce_add_sync(0xffffffffu, sl[i]);
            sh[i] = __reduce_add_sync(0xffffffffu, sh[i]);
            if (lane == 0) { red[w * 6 + i] = sl[i]; red[w * 6 + 3 + i] = sh[i]; }
        }
        if (tid < 32) hmax_bits[tid] = 0;
        __syncthreads();
        const float fn = (float)n;
        const float mx = fix_mean(red[0] + red[6] + red[12] + red[18], red[3] + red[9] + red[15] + red[21], a.g.fix_inv, fn);
        const float my = fix_mean(red[1] + red[7] + red[13] + red[19], red[4] + red[10] + red[16] + red[22], a.g.fix_inv, fn);
        const float mz = fix_mean(red[2] + red[8] + red[14] + red[20], red[5] + red[11] + red[17] + red[23], a.g.fix_inv, fn);

        float hm[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) hm[k] = 0.f;
        for (int r = tid; r < n; r += kSimtThreads) {
            const float4 p = slot[r];
            float d[8];
            const float xc = p.x - it.ctr_x, yc = p.y - it.ctr_y;
            d[0] = alias ? xc : p.x; d[1] = alias ? yc : p.y; d[2] = p.z;
            d[3] = p.x - mx; d[4] = p.y - my; d[5] = p.z - mz;
            d[6] = xc; d[7] = yc;
            float* hrow = H + (size_t)r * kHStride;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc = __fmaf_rn(w0s[k * 8 + i], d[i], acc);
                const float h = fmaxf(acc + b0s[k], 0.f);
                hrow[k] = h;
                hm[k] = fmaxf(hm[k], h);
            }
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const float v = warp_max_f32(hm[k]);
            if (lane == 0) atomicMax(&hmax_bits[k], __float_as_int(v));  // h >= 0: int order == float order
        }
        if (n < M && tid < 32) {  // one representative padded slot: relu(BN(0))
            const float hp = fmaxf(b0s[tid], 0.f);
            H[(size_t)n * kHStride + tid] = hp;
            atomicMax(&hmax_bits[tid], __float_as_int(hp));
        }
        __syncthreads();
        const int rows = n + (n < M ? 1 : 0);
        for (int c = tid; c < C; c += kSimtThreads) {
            float wa[32];
            float g = b1g[c];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                wa[k] = w1g[(size_t)c * 64 + k];
                g = __fmaf_rn(w1g[(size_t)c * 64 + 32 + k], __int_as_float(hmax_bits[k]), g);
            }
            float best = -INFINITY;
            for (int r = 0; r < rows; ++r) {
                const float4* hr = reinterpret_cast<const float4*>(H + (size_t)r * kHStride);
                float acc = 0.f;
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 h = hr[k4];
                    acc = __fmaf_rn(wa[k4 * 4 + 0], h.x, acc);
                    acc = __fmaf_rn(wa[k4 * 4 + 1], h.y, acc);
                    acc = __fmaf_rn(wa[k4 * 4 + 2], h.z, acc);
                    acc = __fmaf_rn(wa[k4 * 4 + 3], h.w, acc);
                }
                best = fmaxf(best, acc);
            }
            store_scalar(a, out_index(a, item, it.b, it.cell, c), fmaxf(best + g, 0.f));
        }
        __syncthreads();
    }
}

__global__ void zero_lidar_kernel(PfnArgs a) {
    const int C = a.bl.C;
    const int64_t total = a.num_items * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t idx = i;
        if (a.out_layout == P3P_LAYOUT_NCHW) {
            const int64_t per_tile = (int64_t)C * a.items_per_tile;
            const int64_t b = i / per_tile, rem = i - b * per_tile;
            idx = (b * a.c_total + a.c_offset) * a.items_per_tile + rem;
        }
        store_scalar(a, idx, 0.f);
    }
}

// ------------------------------------------------------------------------------------------------
// tensor-core kernel
// ------------------------------------------------------------------------------------------------
constexpr int kNF = 8;                         // front-end warps (one pillar each at a time)
constexpr int kNS = kNF / 2;                   // B-operand stages: one pillar pair each
constexpr int kEpiWarp0 = 1 + kNF;             // warp 0: TMEM alloc + MMA issue; 1..kNF: front end; then 12 epilogue warps
constexpr int kTcThreads = 32 * (kEpiWarp0 + 12);
constexpr int kUnit = 8;                       // items per epilogue flush (one 32-byte NCHW sector per channel)
constexpr int kPairsPerUnit = kUnit / 2;
constexpr int kTmemStage = 160;  // TMEM columns per channel tile: 128 (pair) + 16 (G), padded

template <bool kTf32>
struct TcCfg {
    static constexpr int RB = kTf32 ? 128 : 64;           // bytes of one K = 32 operand row
    static constexpr int kATile = 128 * RB;               // 128 channels
    static constexpr int kHStage = 128 * RB;              // 2 pillars x 64 rows
    static constexpr int kGStage = 16 * RB;               // N = 16 rows; rows 0,1 = hmax of the pair
    static constexpr uint32_t kLayout = kTf32 ? 2u : 4u;  // SWIZZLE_128B : SWIZZLE_64B
    static constexpr uint32_t kSBO = 8 * RB;
    static constexpr int kKSteps = kTf32 ? 4 : 2;         // UMMA_K = 8 (tf32) / 16 (bf16): 32 bytes per step
    static constexpr int kGroup = kTf32 ? 4 : 8;          // channels per 16-byte operand chunk
    static constexpr int kGroups = 32 / kGroup;
    static constexpr size_t kSmemOperands = (size_t)6 * kATile + (size_t)kNS * (kHStage + kGStage);
    static constexpr size_t kSmemFloats = 10 * 32 + 384 + kNF * 32;
    static constexpr size_t kSmemBytes = 1024 + kSmemOperands + kSmemFloats * 4 + (2 * kNS + 8) * 8 + 16;
};

__device__ __forceinline__ float max32(const float (&v)[32]) {
    float r[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) r[i] = fmax3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    r[10] = fmaxf(v[30], v[31]);
    return fmaxf(fmax3(fmax3(r[0], r[1], r[2]), fmax3(r[3], r[4], r[5]), fmax3(r[6], r[7], r[8])), fmaxf(r[9], r[10]));
}

__device__ __forceinline__ float max64(const float (&v)[64]) {
    float r[22];
#pragma unroll
    for (int i = 0; i < 21; ++i) r[i] = fmax3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    r[21] = v[63];
    float s[8];
#pragma unroll
    for (int i = 0; i < 7; ++i) s[i] = fmax3(r[3 * i], r[3 * i + 1], r[3 * i + 2]);
    s[7] = r[21];
    return fmaxf(fmax3(s[0], s[1], s[2]), fmax3(fmax3(s[3], s[4], s[5]), s[6], s[7]));
}

template <bool kTf32>
__global__ void __launch_bounds__(kTcThreads, 1) pfn_tc_kernel(PfnArgs a) {
    using Cfg = TcCfg<kTf32>;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t raw = smem_u32(smem_dyn);
    unsigned char* base = smem_dyn + ((1024u - (raw & 1023u)) & 1023u);
    unsigned char* sA1 = base;
    unsigned char* sA2 = sA1 + 3 * Cfg::kATile;
    unsigned char* sH = sA2 + 3 * Cfg::kATile;
    unsigned char* sG = sH + kNS * Cfg::kHStage;
    float* sFront = reinterpret_cast<float*>(sG + kNS * Cfg::kGStage);  // [10][32]
    float* sB1 = sFront + 10 * 32;                                      // [384]
    float* sKap = sB1 + 384;                                            // [kNF][32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sKap + kNF * 32);
    uint64_t* h_full = bars;                 // [kNS] front end -> MMA (2 arrivals: the two warps of a pair)
    uint64_t* h_empty = bars + kNS;          // [kNS] MMA -> front end (tcgen05.commit)
    uint64_t* t_full = bars + 2 * kNS;       // [3]   MMA -> epilogue group m (tcgen05.commit)
    uint64_t* t_empty = bars + 2 * kNS + 3;  // [3]   epilogue group m -> MMA (128 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kNS + 6);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int MT = a.bl.MT;
    const int64_t num_units = (a.num_items + kUnit - 1) / kUnit;
    const int64_t my_units = (num_units > blockIdx.x) ? (num_units - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t my_pairs = my_units * kPairsPerUnit;  // pairs this CTA processes, in order p = 0, 1, ...

    // ---- one-time setup --------------------------------------------------------------------------
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 32) {
        for (int i = 0; i < kNS; ++i) { mbar_init(&h_full[i], 2); mbar_init(&h_empty[i], 1); }
        for (int i = 0; i < 3; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 128); }
        fence_mbar_init();
    }
    {
        const uint4* g1 = reinterpret_cast<const uint4*>(a.blob + a.bl.off_a1);
        const uint4* g2 = reinterpret_cast<const uint4*>(a.blob + a.bl.off_a2);
        const int nvec = MT * Cfg::kATile / 16;
        for (int i = tid; i < nvec; i += kTcThreads) {
            reinterpret_cast<uint4*>(sA1)[i] = g1[i];
            reinterpret_cast<uint4*>(sA2)[i] = g2[i];
        }
        for (int i = tid; i < kNS * Cfg::kGStage / 16; i += kTcThreads) reinterpret_cast<uint4*>(sG)[i] = make_uint4(0, 0, 0, 0);
        const float* fg = reinterpret_cast<const float*>(a.blob + a.bl.off_front);
        const float* b1g = reinterpret_cast<const float*>(a.blob + a.bl.off_b1);
        for (int i = tid; i < 10 * 32; i += kTcThreads) sFront[i] = fg[i];
        for (int i = tid; i < 384; i += kTcThreads) sB1[i] = (i < a.bl.Cpad) ? b1g[i] : 0.f;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =========================== MMA issuer ===========================
        // The whole warp walks the loop (warp-uniform control flow keeps descriptors in uniform registers);
        // one elected lane issues the tcgen05 instructions and their commits.
        const bool leader = elect_one();
        const uint32_t idesc_main = make_idesc(kTf32, 128, 128);
        const uint32_t idesc_g = make_idesc(kTf32, 128, 16);
        const uint32_t desc_hi = (Cfg::kSBO >> 4) | (1u << 14) | (Cfg::kLayout << 29);
        const uint32_t a1_lo = (smem_u32(sA1) >> 4) | (1u << 16), a2_lo = (smem_u32(sA2) >> 4) | (1u << 16);
        const uint32_t h_lo = (smem_u32(sH) >> 4) | (1u << 16), g_lo = (smem_u32(sG) >> 4) | (1u << 16);
        int st = 0;
        uint32_t use = 0;
        for (int64_t p = 0; p < my_pairs; ++p) {
            mbar_wait(&h_full[st], use & 1);
            tc_fence_after();
            const uint32_t hs_lo = h_lo + (uint32_t)st * (Cfg::kHStage >> 4), gs_lo = g_lo + (uint32_t)st * (Cfg::kGStage >> 4);
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                if (m < MT) {
                    mbar_wait(&t_empty[m], ((uint32_t)p & 1) ^ 1);
                    tc_fence_after();
                    if (leader) {
                        const uint32_t d_main = tmem_base + m * kTmemStage, d_g = d_main + 128;
#pragma unroll
                        for (int k = 0; k < Cfg::kKSteps; ++k)
                            tc_mma<kTf32>(d_g, ((uint64_t)desc_hi << 32) | (a2_lo + (uint32_t)((m * Cfg::kATile + k * 32) >> 4)),
                                          ((uint64_t)desc_hi << 32) | (gs_lo + (uint32_t)(k * 2)), idesc_g, k > 0);
#pragma unroll
                        for (int k = 0; k < Cfg::kKSteps; ++k)
                            tc_mma<kTf32>(d_main, ((uint64_t)desc_hi << 32) | (a1_lo + (uint32_t)((m * Cfg::kATile + k * 32) >> 4)),
                                          ((uint64_t)desc_hi << 32) | (hs_lo + (uint32_t)(k * 2)), idesc_main, k > 0);
                        tc_commit(&t_full[m]);
                    }
                    __syncwarp();
                }
            }
            if (leader) tc_commit(&h_empty[st]);
            __syncwarp();
            if (++st == kNS) { st = 0; ++use; }
        }
    } else if (warp < kEpiWarp0) {
        // =========================== front end: one pillar per warp ===========================
        const int fw = warp - 1, half = fw & 1, st = fw >> 1;  // this warp fills half `half` of stage `st`
        const float2* Ux2 = reinterpret_cast<const float2*>(sFront + 0 * 32);
        const float2* Uy2 = reinterpret_cast<const float2*>(sFront + 1 * 32);
        const float2* Uz2 = reinterpret_cast<const float2*>(sFront + 2 * 32);
        const float4* Hp4 = reinterpret_cast<const float4*>(sFront + 9 * 32);
        float* kap = sKap + fw * 32;
        const float2* Kp2 = reinterpret_cast<const float2*>(kap);
        const float kcx = sFront[3 * 32 + lane], kcy = sFront[4 * 32 + lane];
        const float wmx = sFront[5 * 32 + lane], wmy = sFront[6 * 32 + lane], wmz = sFront[7 * 32 + lane];
        const float b0l = sFront[8 * 32 + lane];

        // this warp's pairs: p = st, st + kNS, ...; pair p covers items 2p, 2p+1 of the CTA's unit sequence
        auto item_of = [&](int64_t j) -> int64_t {
            const int64_t p = st + j * kNS;
            if (p >= my_pairs) return a.num_items;  // fetch_item -> invalid
            const int64_t u = blockIdx.x + (p / kPairsPerUnit) * (int64_t)gridDim.x;
            return u * kUnit + (p % kPairsPerUnit) * 2 + half;
        };
        const int64_t nseq = (my_pairs > st) ? (my_pairs - st + kNS - 1) / kNS : 0;
        // software pipeline: descriptor two ahead, points one ahead
        Item it_cur = fetch_item(a, item_of(0));
        Item it_nxt = fetch_item(a, item_of(1));
        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
        if (it_cur.valid) {
            const float4* sl = item_slots(a, it_cur);
            if (lane < it_cur.n) c0 = sl[lane];
            if (lane + 32 < it_cur.n) c1 = sl[lane + 32];
        }
        unsigned char* hst = sH + st * Cfg::kHStage;
        unsigned char* gst = sG + st * Cfg::kGStage;
        const int R0 = half * 64 + lane, R1 = R0 + 32;
        for (int64_t j = 0; j < nseq; ++j) {
            const Item it = it_cur;
            const float4 p0 = c0, p1 = c1;
            it_cur = it_nxt;
            c0 = make_float4(0.f, 0.f, 0.f, 0.f); c1 = c0;
            if (it_cur.valid) {
                const float4* sl = item_slots(a, it_cur);
                if (lane < it_cur.n) c0 = sl[lane];
                if (lane + 32 < it_cur.n) c1 = sl[lane + 32];
            }
            it_nxt = fetch_item(a, item_of(j + 2));

            mbar_wait(&h_empty[st], ((uint32_t)j & 1) ^ 1);
            if (it.valid) {
                const int n = it.n;
                float mean3[3];
                {
                    const float c0v[3] = {p0.x, p0.y, p0.z}, c1v[3] = {p1.x, p1.y, p1.z};
#pragma unroll
                    for (int i = 0; i < 3; ++i) {  // padded lanes hold zeros
                        int l0, h0, l1, h1;
                        fix_split(c0v[i], a.g.fix_scale, l0, h0);
                        fix_split(c1v[i], a.g.fix_scale, l1, h1);
                        mean3[i] = fix_mean(__reduce_add_sync(0xffffffffu, l0 + l1), __reduce_add_sync(0xffffffffu, h0 + h1),
                                            a.g.fix_inv, (float)n);
                    }
                }
                const float mpx = mean3[0] - it.ctr_x, mpy = mean3[1] - it.ctr_y, mz = mean3[2];
                float kv = b0l;
                kv = __fmaf_rn(kcx, it.ctr_x, kv);
                kv = __fmaf_rn(kcy, it.ctr_y, kv);
                kv = __fmaf_rn(-wmx, mpx, kv);
                kv = __fmaf_rn(-wmy, mpy, kv);
                kv = __fmaf_rn(-wmz, mz, kv);
                kap[lane] = kv;
                __syncwarp();
                const float2 x0 = make_float2(p0.x - it.ctr_x, p0.x - it.ctr_x), y0 = make_float2(p0.y - it.ctr_y, p0.y - it.ctr_y);
                const float2 z0 = make_float2(p0.z, p0.z);
                const float2 x1 = make_float2(p1.x - it.ctr_x, p1.x - it.ctr_x), y1 = make_float2(p1.y - it.ctr_y, p1.y - it.ctr_y);
                const float2 z1 = make_float2(p1.z, p1.z);
                const bool ok0 = lane < n, ok1 = lane + 32 < n;
#pragma unroll
                for (int g = 0; g < Cfg::kGroups; ++g) {
                    float h0[Cfg::kGroup], h1[Cfg::kGroup], hm[Cfg::kGroup];
#pragma unroll
                    for (int v2 = 0; v2 < Cfg::kGroup / 2; ++v2) {
                        const int i2 = g * (Cfg::kGroup / 2) + v2;
                        const float2 ux = Ux2[i2], uy = Uy2[i2], uz = Uz2[i2], kp = Kp2[i2];
                        const float2 a0 = ffma2(ux, x0, ffma2(uy, y0, ffma2(uz, z0, kp)));
                        const float2 a1 = ffma2(ux, x1, ffma2(uy, y1, ffma2(uz, z1, kp)));
                        h0[v2 * 2 + 0] = fmaxf(a0.x, 0.f); h0[v2 * 2 + 1] = fmaxf(a0.y, 0.f);
                        h1[v2 * 2 + 0] = fmaxf(a1.x, 0.f); h1[v2 * 2 + 1] = fmaxf(a1.y, 0.f);
                    }
                    if (n < 64) {  // padded slots carry relu(BN(0)) (warp-uniform branch)
#pragma unroll
                        for (int v4 = 0; v4 < Cfg::kGroup / 4; ++v4) {
                            const float4 hp = Hp4[g * (Cfg::kGroup / 4) + v4];
                            if (!ok0) { h0[v4 * 4 + 0] = hp.x; h0[v4 * 4 + 1] = hp.y; h0[v4 * 4 + 2] = hp.z; h0[v4 * 4 + 3] = hp.w; }
                            if (!ok1) { h1[v4 * 4 + 0] = hp.x; h1[v4 * 4 + 1] = hp.y; h1[v4 * 4 + 2] = hp.z; h1[v4 * 4 + 3] = hp.w; }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < Cfg::kGroup; ++i) hm[i] = warp_max_f32(fmaxf(h0[i], h1[i]));
                    if constexpr (kTf32) {
                        // operands are already scaled by (1 + 2^-12): the MMA's truncation rounds them to nearest tf32
                        *reinterpret_cast<float4*>(hst + R0 * 128 + ((g ^ (R0 & 7)) * 16)) = make_float4(h0[0], h0[1], h0[2], h0[3]);
                        *reinterpret_cast<float4*>(hst + R1 * 128 + ((g ^ (R1 & 7)) * 16)) = make_float4(h1[0], h1[1], h1[2], h1[3]);
                        if (lane == 0) *reinterpret_cast<float4*>(gst + half * 128 + ((g ^ half) * 16)) = make_float4(hm[0], hm[1], hm[2], hm[3]);
                    } else {
                        const uint4 w0v = make_uint4(pack_bf16(h0[0], h0[1]), pack_bf16(h0[2], h0[3]), pack_bf16(h0[4], h0[5]), pack_bf16(h0[6], h0[7]));
                        const uint4 w1v = make_uint4(pack_bf16(h1[0], h1[1]), pack_bf16(h1[2], h1[3]), pack_bf16(h1[4], h1[5]), pack_bf16(h1[6], h1[7]));
                        const uint4 wmv = make_uint4(pack_bf16(hm[0], hm[1]), pack_bf16(hm[2], hm[3]), pack_bf16(hm[4], hm[5]), pack_bf16(hm[6], hm[7]));
                        *reinterpret_cast<uint4*>(hst + R0 * 64 + ((g ^ ((R0 >> 1) & 3)) * 16)) = w0v;
                        *reinterpret_cast<uint4*>(hst + R1 * 64 + ((g ^ ((R1 >> 1) & 3)) * 16)) = w1v;
                        if (lane == 0) *reinterpret_cast<uint4*>(gst + half * 64 + (g * 16)) = wmv;
                    }
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&h_full[st]);
        }
    } else {
        // =========================== epilogue: warp group m owns channel tile m ===========================
        const int m = (warp - kEpiWarp0) >> 2;
        const int quad = warp & 3;  // TMEM lanes this warp may read: 32 * (warp id % 4)
        const int c = m * 128 + quad * 32 + lane;
        if (m < MT) {
            const float b1c = sB1[c];
            const bool c_ok = c < a.bl.C;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(m * kTmemStage);
            const bool nchw = (a.item_mode == kItemsCanvas) && (a.out_layout == P3P_LAYOUT_NCHW);
            const bool nchw_vec = nchw && (a.items_per_tile % kUnit == 0);
            uint32_t gp = 0;
            for (int64_t u = blockIdx.x; u < num_units; u += gridDim.x) {
                float ob[kUnit];
                const int64_t item0 = u * kUnit;
                unsigned vmask = 0;
#pragma unroll
                for (int i = 0; i < kUnit; ++i) {
                    const int64_t item = item0 + i;
                    bool v = false;
                    if (item < a.num_items) {
                        if (a.item_mode == kItemsCanvas) {
                            v = __ldg(a.ws.cell_desc + item) >= 0;
                        } else {
                            const int b = (int)(item / a.items_per_tile);
                            v = (int)(item - (int64_t)b * a.items_per_tile) < a.ws.num_pil[b];
                        }
                    }
                    vmask |= (v ? 1u : 0u) << i;
                }
#pragma unroll
                for (int q = 0; q < kPairsPerUnit; ++q, ++gp) {
                    mbar_wait(&t_full[m], gp & 1);
                    tc_fence_after();
                    float v[32];
                    tmem_ld32_wait(taddr + 0, v);
                    float mA = max32(v);
                    tmem_ld32_wait(taddr + 32, v);
                    mA = fmaxf(mA, max32(v));
                    tmem_ld32_wait(taddr + 64, v);
                    float mB = max32(v);
                    tmem_ld32_wait(taddr + 96, v);
                    mB = fmaxf(mB, max32(v));
                    float gA, gB;
                    tmem_ld2_wait(taddr + 128, gA, gB);
                    tc_fence_before();
                    mbar_arrive(&t_empty[m]);
                    ob[2 * q + 0] = ((vmask >> (2 * q)) & 1u) ? fmaxf(mA + gA + b1c, 0.f) : 0.f;
                    ob[2 * q + 1] = ((vmask >> (2 * q + 1)) & 1u) ? fmaxf(mB + gB + b1c, 0.f) : 0.f;
                }
                if (!c_ok) continue;
                if (nchw_vec) {
                    const int b = (int)(item0 / a.items_per_tile);
                    const int cell0 = (int)(item0 - (int64_t)b * a.items_per_tile);
                    const int64_t idx = ((int64_t)b * a.c_total + a.c_offset + c) * a.items_per_tile + cell0;
                    if (a.out_dtype == P3P_DTYPE_F32) {
                        float4* dst = reinterpret_cast<float4*>(static_cast<float*>(a.out) + idx);
                        dst[0] = make_float4(ob[0], ob[1], ob[2], ob[3]);
                        dst[1] = make_float4(ob[4], ob[5], ob[6], ob[7]);
                    } else {
                        uint4* dst = reinterpret_cast<uint4*>(static_cast<unsigned short*>(a.out) + idx);
                        dst[0] = make_uint4(pack_bf16(ob[0], ob[1]), pack_bf16(ob[2], ob[3]), pack_bf16(ob[4], ob[5]), pack_bf16(ob[6], ob[7]));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < kUnit; ++i) {
                        const int64_t item = item0 + i;
                        if (item >= a.num_items) break;
                        const bool v = (vmask >> i) & 1u;
                        if (!v && a.item_mode != kItemsCanvas) continue;  // list rows past num_pillars stay untouched
                        const int b = (int)(item / a.items_per_tile);
                        const int cell = (int)(item - (int64_t)b * a.items_per_tile);
                        store_scalar(a, out_index(a, item, b, cell, c), ob[i]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace

int launch_pfn_prepare(const p3p_pfn_params* p, int precision, char* blob, const BlobLayout& bl, cudaStream_t st) {
    pfn_prepare_kernel<<<4, 128, 0, st>>>(*p, precision, blob, bl);
    P3P_CUDA_CHECK(cudaGetLastError());
    return P3P_OK;
}

int launch_pfn_simt(const PfnArgs& a, cudaStream_t st) {
    if (a.num_items <= 0) return P3P_OK;
    const size_t smem = ((size_t)(a.g.M + 1) * kHStride + 256 + 32 + 24 + 32) * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        P3P_CUDA_CHECK(cudaFuncSetAttribute(pfn_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    const int sms = device_sm_count();
    int64_t grid = (int64_t)sms * 8;
    if (grid > a.num_items) grid = a.num_items;
    pfn_simt_kernel<<<(unsigned)grid, kSimtThreads, smem, st>>>(a);
    P3P_CUDA_CHECK(cudaGetLastError());
    return P3P_OK;
}

int launch_zero_lidar(const PfnArgs& a, cudaStream_t st) {
    if (a.num_items <= 0) return P3P_OK;
    zero_lidar_kernel<<<device_sm_count() * 4, 256, 0, st>>>(a);
    P3P_CUDA_CHECK(cudaGetLastError());
    return P3P_OK;
}

int launch_pfn_tc(const PfnArgs& a, int precision, cudaStream_t st) {
    if (a.num_items <= 0) return P3P_OK;
    static bool attr_done = false;
    if (!attr_done) {
        P3P_CUDA_CHECK(cudaFuncSetAttribute(pfn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)TcCfg<true>::kSmemBytes));
        P3P_CUDA_CHECK(cudaFuncSetAttribute(pfn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)TcCfg<false>::kSmemBytes));
        attr_done = true;
    }
    const int64_t units = (a.num_items + kUnit - 1) / kUnit;
    int64_t grid = device_sm_count();
    if (grid > units) grid = units;
    if (precision == P3P_PRECISION_TF32)
        pfn_tc_kernel<true><<<(unsigned)grid, kTcThreads, TcCfg<true>::kSmemBytes, st>>>(a);
    else
        pfn_tc_kernel<false><<<(unsigned)grid, kTcThreads, TcCfg<false>::kSmemBytes, st>>>(a);
    P3P_CUDA_CHECK(cudaGetLastError());
    return P3P_OK;
}

}  // namespace p3p
