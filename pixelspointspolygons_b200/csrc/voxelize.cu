// voxelize.cu -- point -> pillar assignment for a whole jagged batch in ONE kernel (sm_100a).
//
// Replaces the per-sample Python loop of Open3D-ML PointPillars.voxelize that the reference calls at
// R:pixelspointspolygons/models/pointpillars/pointpillars_o3d.py:92.
//
// The reference sorts (hash, index) pairs per tile and keeps the first M indices of each run.  Only the stable
// rank of a point inside its run matters, and only ranks < M survive, so nothing is sorted here:
//
//   * a CTA takes a ticket (atomic counter) -> chunk of <= 4096 consecutive points of one tile.  Tickets, not
//     blockIdx, order the chunks, so a CTA only ever waits for CTAs that are already running.
//   * every lane loads its <= 16 points into registers up front (all loads in flight at once; the xyz stream is
//     read from HBM exactly once) and computes the cell hash with the reference's fp32 operation order.
//   * each warp owns a contiguous segment of the chunk and counts it per key in a private shared-memory histogram
//     (match.any aggregation, no atomics); the chunk's per-key counts are published to global memory and a flag
//     is released.
//   * the CTA acquires the flags of the earlier chunks of its tile, sums their counts per key (exclusive prefix
//     over chunks, then over its own warps) and walks its registers again in index order:
//     rank = base + (same-key lanes below me); survivors (rank < M) go to slots[tile][key][rank] =
//     (x, y, z, tile-local index).  Deterministic: no atomics on the data path, no dependence on scheduling.
//   * the last CTA of a tile to finish runs the tile's plan: keys in ascending order -> run ordinal (max_voxels
//     cut), cell coordinates from the rank-0 (lowest index) point, x/y bound filter, final voxel order, and the
//     canvas owner table (last pillar in voxel order wins a cell, Appendix A.5).
//
// export_kernel (optional) dumps the reference-shaped tensors for the parity tests.
#include "p3p_internal.cuh"

namespace p3p {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kIters = kMaxChunkPoints / kThreads;  // points per lane held in registers

struct ChunkLoc {
    int b;        // tile, -1 if this ticket has no chunk
    int c;        // chunk index inside the tile
    int nchunks;  // chunks of the tile
    int gstart;   // global index of the tile's chunk 0
    long long p0, p1, tile_start;
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Executed by warp 0: map ticket -> (tile, local chunk).  Tiles are walked 32 at a time.
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

__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_tot, int* total) {
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
    for (int i = 0; i < kWarps; ++i) {
        const int t = warp_tot[i];
        if (i < w) off += t;
        tot += t;
    }
    *total = tot;
    return off + incl - v;
}

__device__ void write_empty_tile(const GridDev& g, const WsPtrs& ws, int b) {
    const int HW = g.ny * g.nx;
    for (int i = threadIdx.x; i < HW; i += kThreads) {
        ws.owner[(size_t)b * HW + i] = -1;
        ws.cell_desc[(size_t)b * HW + i] = -1;
    }
    if (threadIdx.x == 0) ws.num_pil[b] = 0;
}

// Per-tile plan, run by the CTA that finished the tile's last chunk.  scratch: >= (2 K + HW) ints of shared memory.
__device__ void plan_tile(const GridDev& g, const WsPtrs& ws, int b, int* scratch, int* warp_tot) {
    const int K = g.num_keys, HW = g.ny * g.nx, M = g.M, tid = threadIdx.x;
    int* tot_s = scratch;       // [K] points per key
    int* cell_s = tot_s + K;    // [K] packed cell of the run's first point, -1 if dropped
    int* owner_s = cell_s + K;  // [HW]
    const int* totals = ws.totals + (size_t)b * K;
    const float4* slots = ws.slots + (size_t)b * K * M;
    // one round trip: counts and rank-0 points of every key, all loads independent (written by other CTAs -> L2 loads)
    for (int k = tid; k < K; k += kThreads) {
        const int t = __ldcg(totals + k);
        const float4 p = __ldcg(slots + (size_t)k * M);
        int cell = -1;
        if (t > 0) {
            int cx, cy, cz;
            point_cell(g, p.x, p.y, p.z, cx, cy, cz);
            if (cy < g.nv[1] && cx < g.nv[0]) cell = cx | (cy << 10) | (cz << 20);  // x/y bound filter (A.2)
        }
        tot_s[k] = t;
        cell_s[k] = cell;
    }
    for (int i = tid; i < HW; i += kThreads) owner_s[i] = -1;
    __syncthreads();
    const int per = (K + kThreads - 1) / kThreads;
    const int k0 = tid * per, k1 = (k0 + per < K) ? k0 + per : K;
    int cnt = 0;
    for (int k = k0; k < k1; ++k) cnt += (tot_s[k] > 0);
    int total_runs;
    int ord = block_exclusive_scan(cnt, warp_tot, &total_runs);
    int cnt2 = 0;
    for (int k = k0; k < k1; ++k) {
        if (tot_s[k] > 0) {
            const int r = ord++;
            if (r >= g.Vmax) cell_s[k] = -1;  // only the first max_voxels runs in hash order survive (A.1)
        }
        cnt2 += (cell_s[k] >= 0);
    }
    int total_pil;
    int ord2 = block_exclusive_scan(cnt2, warp_tot, &total_pil);
    for (int k = k0; k < k1; ++k) {
        const int st = cell_s[k];
        if (st < 0) continue;
        const int r = ord2++;
        const size_t pi = (size_t)b * g.Vmax + r;
        const int t = tot_s[k];
        ws.pil_key[pi] = k;
        ws.pil_n[pi] = t < M ? t : M;
        ws.pil_coord[pi] = st;
        const int cx = st & 1023, cy = (st >> 10) & 1023;
        atomicMax(&owner_s[cy * g.nx + cx], r);  // scatter collisions: the later row (higher hash) wins (A.5)
    }
    if (tid == 0) ws.num_pil[b] = total_pil;
    __syncthreads();
    for (int i = tid; i < HW; i += kThreads) {
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

__global__ void __launch_bounds__(kThreads, 3)
voxelize_kernel(const float* __restrict__ pts, int stride, const int64_t* __restrict__ offsets, int B, GridDev g, int S,
                WsPtrs ws, int32_t* __restrict__ point_hash) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = g.num_keys;
    uint16_t* hist = reinterpret_cast<uint16_t*>(smem_raw);  // [kWarps][K]; reused as the plan's scratch
    unsigned* prefix_s = reinterpret_cast<unsigned*>(smem_raw + (((size_t)kWarps * K * sizeof(uint16_t) + 15) / 16) * 16);  // [Kp]
    __shared__ ChunkLoc loc;
    __shared__ int s_ticket, s_last;
    __shared__ int warp_tot[kWarps];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_ticket = (int)atomicAdd(ws.sync, 1u);
    for (int i = tid; i < kWarps * K; i += kThreads) hist[i] = 0;
    __syncthreads();
    const int ticket = s_ticket;
    if (w == 0) locate_chunk(ticket, offsets, B, S, &loc);
    // tiles without points have no chunk: ticket t < B writes the empty plan of tile t
    if (ticket < B && offsets[ticket + 1] <= offsets[ticket]) write_empty_tile(g, ws, ticket);
    __syncthreads();
    if (loc.b < 0) return;

    unsigned* flags = ws.sync + 1;
    unsigned* tile_done = flags + ws.max_chunks;
    const int segS = S / kWarps;  // multiple of 32
    const long long seg0 = loc.p0 + (long long)w * segS;
    const long long seg1 = (seg0 + segS < loc.p1) ? seg0 + segS : loc.p1;
    uint16_t* myhist = hist + (size_t)w * K;

    // ---- load: every lane's points into registers, all loads issued before any use ------------------------------
    float px[kIters], py[kIters], pz[kIters];
    int key[kIters];
#pragma unroll
    for (int j = 0; j < kIters; ++j) {
        const long long idx = seg0 + j * 32 + lane;
        px[j] = 0.f; py[j] = 0.f; pz[j] = 0.f;
        if (j * 32 < segS && idx < seg1) {
            const float* p = pts + idx * stride;
            px[j] = __ldg(p); py[j] = __ldg(p + 1); pz[j] = __ldg(p + 2);
        }
    }
#pragma unroll
    for (int j = 0; j < kIters; ++j) {
        const long long idx = seg0 + j * 32 + lane;
        key[j] = -1;
        if (j * 32 < segS && idx < seg1) {
            key[j] = point_key(g, px[j], py[j], pz[j]);
            if (point_hash) point_hash[idx] = key[j];
        }
    }
    // ---- walk 1: per-warp histogram of this warp's contiguous segment --------------------------------------------
#pragma unroll
    for (int j = 0; j < kIters; ++j) {
        if (j * 32 < segS && seg0 + j * 32 < seg1) {  // warp-uniform
            const unsigned m = __match_any_sync(0xffffffffu, key[j]);
            if (key[j] >= 0 && lane == (__ffs(m) - 1)) myhist[key[j]] = (uint16_t)(myhist[key[j]] + __popc(m));
            __syncwarp();  // the leader lane of a key changes between iterations
        }
    }
    __syncthreads();
    // ---- publish this chunk's per-key counts, then wait for the earlier chunks of the tile ---------------------
    const int Kp = ws.key_stride;  // row stride of chunk_hist: K rounded up to 8 (16-byte rows)
    {
        uint16_t* dst = ws.chunk_hist + (size_t)ticket * Kp;
        for (int k = tid; k < K; k += kThreads) {
            unsigned s = 0;
#pragma unroll
            for (int ww = 0; ww < kWarps; ++ww) s += hist[(size_t)ww * K + k];
            dst[k] = (uint16_t)s;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(flags + ticket, 1u);
        for (int cc = tid; cc < loc.c; cc += kThreads)
            while (ld_acquire(flags + loc.gstart + cc) == 0u) {
            }
        __syncthreads();
    }
    // ---- base ranks: earlier chunks of the tile, then earlier warps of this CTA -----------------------------------
    const int M = g.M;
    const bool last_chunk = (loc.c == loc.nchunks - 1);
    // sum of the earlier chunks' counts, 8 keys per thread with 16-byte L2 loads, 4 rows in flight per batch
    for (int k8 = tid; k8 < Kp / 8; k8 += kThreads) {
        unsigned acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const uint4* src = reinterpret_cast<const uint4*>(ws.chunk_hist + (size_t)loc.gstart * Kp) + k8;
        const size_t row = (size_t)Kp / 8;
        for (int c0 = 0; c0 < loc.c; c0 += 4) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = (c0 + u < loc.c) ? __ldcg(src + (size_t)(c0 + u) * row) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc[0] += v[u].x & 0xFFFFu; acc[1] += v[u].x >> 16; acc[2] += v[u].y & 0xFFFFu; acc[3] += v[u].y >> 16;
                acc[4] += v[u].z & 0xFFFFu; acc[5] += v[u].z >> 16; acc[6] += v[u].w & 0xFFFFu; acc[7] += v[u].w >> 16;
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) prefix_s[k8 * 8 + u] = acc[u];
    }
    __syncthreads();
    for (int k = tid; k < K; k += kThreads) {
        unsigned run = prefix_s[k];
#pragma unroll
        for (int ww = 0; ww < kWarps; ++ww) {
            const unsigned t = hist[(size_t)ww * K + k];
            hist[(size_t)ww * K + k] = (uint16_t)(run < (unsigned)M ? run : (unsigned)M);  // saturate: rank >= M is dropped
            run += t;
        }
        if (last_chunk) ws.totals[(size_t)loc.b * K + k] = (int)run;
    }
    __syncthreads();
    // ---- walk 2: stable rank in index order, scatter the survivors ---------------------------------------------------
    float4* tile_slots = ws.slots + (size_t)loc.b * K * M;
#pragma unroll
    for (int j = 0; j < kIters; ++j) {
        if (j * 32 < segS && seg0 + j * 32 < seg1) {  // warp-uniform
            const int kj = key[j];
            const unsigned m = __match_any_sync(0xffffffffu, kj);
            int basecnt = 0;
            if (kj >= 0) basecnt = myhist[kj];
            const int rank = basecnt + __popc(m & ((1u << lane) - 1u));
            if (kj >= 0 && rank < M) {
                const long long idx = seg0 + j * 32 + lane;
                tile_slots[(size_t)kj * M + rank] = make_float4(px[j], py[j], pz[j], __int_as_float((int)(idx - loc.tile_start)));
            }
            __syncwarp();
            if (kj >= 0 && lane == (__ffs(m) - 1)) {
                const int nb = basecnt + __popc(m);
                myhist[kj] = (uint16_t)(nb < M ? nb : M);
            }
            __syncwarp();
        }
    }
    // ---- the last CTA of the tile to get here plans the tile ------------------------------------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(tile_done + loc.b, 1u) == (unsigned)(loc.nchunks - 1)) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    plan_tile(g, ws, loc.b, reinterpret_cast<int*>(smem_raw), warp_tot);
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

static size_t voxelize_smem_bytes(const GridDev& g) {
    const size_t kp = ((size_t)g.num_keys + 7) / 8 * 8;
    const size_t hist = ((size_t)kWarps * g.num_keys * sizeof(uint16_t) + 15) / 16 * 16 + kp * sizeof(unsigned);
    const size_t plan = (size_t)(2 * g.num_keys + g.ny * g.nx) * sizeof(int);
    return ((hist > plan ? hist : plan) + 15) / 16 * 16;
}

int launch_voxelize(const float* pts, int stride, const int64_t* offsets, int B, int64_t total, const GridDev& g,
                    const WsLayout& l, const WsPtrs& ws, int32_t* point_hash, cudaStream_t st) {
    (void)total;
    const size_t smem = voxelize_smem_bytes(g);
    static bool attr_done = false;
    if (!attr_done) {
        P3P_CUDA_CHECK(cudaFuncSetAttribute(voxelize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    // ticket counter, chunk flags and per-tile completion counters start at zero for every call
    P3P_CUDA_CHECK(cudaMemsetAsync(ws.sync, 0, l.sync_bytes, st));
    voxelize_kernel<<<l.max_chunks, kThreads, smem, st>>>(pts, stride, offsets, B, g, l.chunk_points, ws, point_hash);
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
