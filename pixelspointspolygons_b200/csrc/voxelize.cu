// voxelize.cu -- point -> pillar assignment for a whole jagged batch (sm_100a).
//
// Replaces the per-sample Python loop of Open3D-ML PointPillars.voxelize that the reference calls at
// R:pixelspointspolygons/models/pointpillars/pointpillars_o3d.py:92.
//
// The reference sorts (hash, index) pairs per tile and keeps the first M indices of each run.  Only the
// stable rank of a point inside its run matters, and only ranks < M survive, so no sort is done here:
//
//   rank_kernel<false>  one CTA per chunk of S consecutive points of one tile; every warp owns a contiguous
//                       segment and counts its points per key (hash) in a private shared-memory histogram
//                       (warp match.any aggregation, no atomics); the chunk's per-key totals go to global.
//   rank_kernel<true>   same walk (keys cached in smem); per key the exclusive prefix over the tile's earlier
//                       chunks and over the CTA's earlier warps gives each segment its base rank; a second
//                       walk in index order assigns rank = base + (same-key lanes below me) and writes the
//                       point, if rank < M, to slots[tile][key][rank] = (x, y, z, tile-local index).
//                       Deterministic: no atomics, no dependence on block scheduling.
//   plan_kernel         one CTA per tile: keys in ascending order -> run ordinal (max_voxels cut), cell
//                       coordinates from the rank-0 (lowest index) point, x/y bound filter, final voxel
//                       order, and the canvas owner table (last pillar in voxel order wins a cell).
//   export_kernel       optional: dumps the reference-shaped tensors for the parity tests.
#include "p3p_internal.cuh"

namespace p3p {

namespace {

constexpr int kRankThreads = 256;
constexpr int kRankWarps = kRankThreads / 32;

struct ChunkLoc {
    int b;        // tile, -1 if this CTA has no chunk
    int c;        // chunk index inside the tile
    int nchunks;  // chunks of the tile
    int gstart;   // global index of the tile's chunk 0
    long long p0, p1, tile_start;
};

// Executed by warp 0: map global chunk id -> (tile, local chunk).  Tiles are walked 32 at a time.
__device__ void locate_chunk(int g, const int64_t* __restrict__ offsets, int B, int S, ChunkLoc* out) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    bool found = false;
    for (int t0 = 0; t0 < B && !found; t0 += 32) {
        const int t = t0 + lane;
        long long o0 = 0, o1 = 0;
        if (t < B) { o0 = offsets[t]; o1 = offsets[t + 1]; }
        const long long n = o1 > o0 ? o1 - o0 : 0;
        const int nc = (int)((n + S - 1) / S);
        int incl = nc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int excl = base + incl - nc;
        const bool mine = (t < B) && (g >= excl) && (g < excl + nc);
        const unsigned bal = __ballot_sync(0xffffffffu, mine);
        if (bal) {
            found = true;
            if (mine) {
                out->b = t;
                out->c = g - excl;
                out->nchunks = nc;
                out->gstart = excl;
                out->tile_start = o0;
                out->p0 = o0 + (long long)(g - excl) * S;
                out->p1 = (out->p0 + S < o1) ? out->p0 + S : o1;
            }
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (!found && lane == 0) out->b = -1;
}

template <bool kSecond>
__global__ void __launch_bounds__(kRankThreads)
rank_kernel(const float* __restrict__ pts, int stride, const int64_t* __restrict__ offsets, int B, GridDev g, int S,
            WsPtrs ws, int32_t* __restrict__ point_hash) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = g.num_keys;
    uint16_t* hist = reinterpret_cast<uint16_t*>(smem_raw);                    // [kRankWarps][K]
    uint16_t* keycache = hist + (size_t)kRankWarps * K + (((size_t)kRankWarps * K) & 1);  // [S] (second pass)
    __shared__ ChunkLoc loc;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (w == 0) locate_chunk(blockIdx.x, offsets, B, S, &loc);
    for (int i = tid; i < kRankWarps * K; i += kRankThreads) hist[i] = 0;
    __syncthreads();
    if (loc.b < 0) return;

    const int segS = S / kRankWarps;  // multiple of 32
    const long long seg0 = loc.p0 + (long long)w * segS;
    const long long seg1 = (seg0 + segS < loc.p1) ? seg0 + segS : loc.p1;
    uint16_t* myhist = hist + (size_t)w * K;

    // ---- walk 1: per-warp histogram of this warp's contiguous segment --------------------------
    for (long long base = seg0; base < seg1; base += 32) {
        const long long idx = base + lane;
        int key = -1;
        if (idx < seg1) {
            const float* p = pts + idx * stride;
            key = point_key(g, __ldg(p), __ldg(p + 1), __ldg(p + 2));
            if (!kSecond && point_hash) point_hash[idx] = key;
        }
        const unsigned m = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && lane == (__ffs(m) - 1)) myhist[key] = (uint16_t)(myhist[key] + __popc(m));
        if (kSecond && idx < seg1) keycache[idx - loc.p0] = (key < 0) ? (uint16_t)kInvalidKey : (uint16_t)key;
        __syncwarp();  // the leader lane of a key changes between iterations
    }
    __syncthreads();

    if (!kSecond) {
        uint16_t* dst = ws.chunk_hist + (size_t)blockIdx.x * K;
        for (int k = tid; k < K; k += kRankThreads) {
            unsigned s = 0;
#pragma unroll
            for (int ww = 0; ww < kRankWarps; ++ww) s += hist[(size_t)ww * K + k];
            dst[k] = (uint16_t)s;
        }
        return;
    }

    // ---- base ranks: earlier chunks of the tile, then earlier warps of this CTA ------------------
    const int M = g.M;
    const bool last_chunk = (loc.c == loc.nchunks - 1);
    for (int k = tid; k < K; k += kRankThreads) {
        unsigned run = 0;
        const uint16_t* src = ws.chunk_hist + (size_t)loc.gstart * K + k;
        for (int cc = 0; cc < loc.c; ++cc) run += src[(size_t)cc * K];
#pragma unroll
        for (int ww = 0; ww < kRankWarps; ++ww) {
            const unsigned t = hist[(size_t)ww * K + k];
            hist[(size_t)ww * K + k] = (uint16_t)(run < (unsigned)M ? run : (unsigned)M);  // saturate: rank >= M is dropped
            run += t;
        }
        if (last_chunk) ws.totals[(size_t)loc.b * K + k] = (int)run;
    }
    __syncthreads();

    // ---- walk 2: stable rank in index order, scatter the survivors -------------------------------
    float4* tile_slots = ws.slots + (size_t)loc.b * K * M;
    for (long long base = seg0; base < seg1; base += 32) {
        const long long idx = base + lane;
        int key = -1;
        if (idx < seg1) {
            const int kc = keycache[idx - loc.p0];
            key = (kc == kInvalidKey) ? -1 : kc;
        }
        const unsigned m = __match_any_sync(0xffffffffu, key);
        int basecnt = 0;
        if (key >= 0) basecnt = myhist[key];
        const int rank = basecnt + __popc(m & ((1u << lane) - 1u));
        if (key >= 0 && rank < M) {
            const float* p = pts + idx * stride;
            tile_slots[(size_t)key * M + rank] =
                make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __int_as_float((int)(idx - loc.tile_start)));
        }
        __syncwarp();
        if (key >= 0 && lane == (__ffs(m) - 1)) {
            const int nb = basecnt + __popc(m);
            myhist[key] = (uint16_t)(nb < M ? nb : M);
        }
        __syncwarp();
    }
}

__device__ __forceinline__ int block_exclusive_scan_256(int v, int* warp_tot, int* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();  // protect warp_tot reuse
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int t = warp_tot[i];
        if (i < w) off += t;
        tot += t;
    }
    *total = tot;
    return off + incl - v;
}

__global__ void __launch_bounds__(256)
plan_kernel(const int64_t* __restrict__ offsets, GridDev g, WsPtrs ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = g.num_keys, HW = g.ny * g.nx;
    int* state = reinterpret_cast<int*>(smem_raw);  // [K] packed coords of kept, in-bounds runs, else -1
    int* owner_s = state + K;                       // [HW]
    __shared__ int warp_tot[8];

    const int b = blockIdx.x, tid = threadIdx.x;
    const long long n_b = offsets[b + 1] - offsets[b];
    for (int i = tid; i < HW; i += 256) owner_s[i] = -1;
    if (n_b <= 0) {
        __syncthreads();
        for (int i = tid; i < HW; i += 256) {
            ws.owner[(size_t)b * HW + i] = -1;
            ws.cell_desc[(size_t)b * HW + i] = -1;
        }
        if (tid == 0) ws.num_pil[b] = 0;
        return;
    }
    const int* totals = ws.totals + (size_t)b * K;
    const float4* slots = ws.slots + (size_t)b * K * g.M;
    const int per = (K + 255) / 256;
    const int k0 = tid * per, k1 = (k0 + per < K) ? k0 + per : K;

    int cnt = 0;
    for (int k = k0; k < k1; ++k) cnt += (totals[k] > 0);
    int total_runs;
    int ord = block_exclusive_scan_256(cnt, warp_tot, &total_runs);
    int cnt2 = 0;
    for (int k = k0; k < k1; ++k) {
        int st = -1;
        if (totals[k] > 0) {
            const int r = ord++;
            if (r < g.Vmax) {  // first max_voxels runs in hash order survive (A.1)
                const float4 p = slots[(size_t)k * g.M];  // rank 0 == lowest original index of the run
                int cx, cy, cz;
                point_cell(g, p.x, p.y, p.z, cx, cy, cz);
                if (cy < g.nv[1] && cx < g.nv[0]) st = cx | (cy << 10) | (cz << 20);  // x/y bound filter (A.2)
            }
        }
        state[k] = st;
        cnt2 += (st >= 0);
    }
    int total_pil;
    int ord2 = block_exclusive_scan_256(cnt2, warp_tot, &total_pil);
    for (int k = k0; k < k1; ++k) {
        const int st = state[k];
        if (st < 0) continue;
        const int r = ord2++;
        const size_t pi = (size_t)b * g.Vmax + r;
        ws.pil_key[pi] = k;
        const int t = totals[k];
        ws.pil_n[pi] = t < g.M ? t : g.M;
        ws.pil_coord[pi] = st;
        const int cx = st & 1023, cy = (st >> 10) & 1023;
        atomicMax(&owner_s[cy * g.nx + cx], r);  // scatter collisions: the later row (higher hash) wins (A.5)
    }
    if (tid == 0) ws.num_pil[b] = total_pil;
    __syncthreads();
    for (int i = tid; i < HW; i += 256) {
        const int o = owner_s[i];
        ws.owner[(size_t)b * HW + i] = o;
        int d = -1;
        if (o >= 0) {  // written above by this CTA; visible after the barrier
            const size_t pi = (size_t)b * g.Vmax + o;
            d = ws.pil_key[pi] | (ws.pil_n[pi] << 16);
        }
        ws.cell_desc[(size_t)b * HW + i] = d;
    }
}

__global__ void __launch_bounds__(128)
export_kernel(GridDev g, int B, WsPtrs ws, p3p_voxel_outputs out) {
    const int b = blockIdx.y;
    const int HW = g.ny * g.nx;
    const int np = ws.num_pil[b];
    if (blockIdx.x == 0) {
        if (out.num_pillars && threadIdx.x == 0) out.num_pillars[b] = np;
        if (out.cell_owner)
            for (int i = threadIdx.x; i < HW; i += blockDim.x) out.cell_owner[(size_t)b * HW + i] = ws.owner[(size_t)b * HW + i];
    }
    for (int r = blockIdx.x; r < np; r += gridDim.x) {
        const size_t pi = (size_t)b * g.Vmax + r;
        const int key = ws.pil_key[pi], n = ws.pil_n[pi], pc = ws.pil_coord[pi];
        if (threadIdx.x == 0) {
            if (out.pillar_coords) {
                int* c = out.pillar_coords + pi * 4;
                c[0] = b; c[1] = pc >> 20; c[2] = (pc >> 10) & 1023; c[3] = pc & 1023;
            }
            if (out.pillar_num_points) out.pillar_num_points[pi] = n;
        }
        const float4* slot = ws.slots + ((size_t)b * g.num_keys + key) * g.M;
        for (int s = threadIdx.x; s < g.M; s += blockDim.x) {
            float4 p = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
            if (s < n) p = slot[s];
            if (out.pillar_point_idx) out.pillar_point_idx[pi * g.M + s] = __float_as_int(p.w);
            if (out.pillar_points) {
                float* d = out.pillar_points + (pi * g.M + s) * 3;
                d[0] = p.x; d[1] = p.y; d[2] = p.z;
            }
        }
    }
}

}  // namespace

int launch_voxelize(const float* pts, int stride, const int64_t* offsets, int B, int64_t total, const GridDev& g,
                    const WsLayout& l, const WsPtrs& ws, int32_t* point_hash, cudaStream_t st) {
    (void)total;
    const int K = g.num_keys;
    const size_t hist_elems = (size_t)kRankWarps * K + (((size_t)kRankWarps * K) & 1);
    const size_t smem1 = hist_elems * sizeof(uint16_t);
    const size_t smem2 = smem1 + (size_t)l.chunk_points * sizeof(uint16_t);
    static bool attr_done = false;
    if (!attr_done) {
        P3P_CUDA_CHECK(cudaFuncSetAttribute(rank_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        P3P_CUDA_CHECK(cudaFuncSetAttribute(rank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        P3P_CUDA_CHECK(cudaFuncSetAttribute(plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_done = true;
    }
    if (l.max_chunks > 0) {
        rank_kernel<false><<<l.max_chunks, kRankThreads, smem1, st>>>(pts, stride, offsets, B, g, l.chunk_points, ws, point_hash);
        rank_kernel<true><<<l.max_chunks, kRankThreads, smem2, st>>>(pts, stride, offsets, B, g, l.chunk_points, ws, point_hash);
    }
    const size_t smem3 = (size_t)(K + g.ny * g.nx) * sizeof(int);
    plan_kernel<<<B, 256, smem3, st>>>(offsets, g, ws);
    P3P_CUDA_CHECK(cudaGetLastError());
    return P3P_OK;
}

int launch_export(const GridDev& g, int B, const WsPtrs& ws, const p3p_voxel_outputs* out, cudaStream_t st) {
    dim3 grid((unsigned)(g.Vmax < 256 ? (g.Vmax > 0 ? g.Vmax : 1) : 256), (unsigned)B);
    export_kernel<<<grid, 128, 0, st>>>(g, B, ws, *out);
    P3P_CUDA_CHECK(cudaGetLastError());
    return P3P_OK;
}

}  // namespace p3p
