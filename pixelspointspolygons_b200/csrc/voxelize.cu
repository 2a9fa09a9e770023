// voxelize.cu -- point -> pillar assignment for a whole jagged batch in ONE kernel (sm_100a).
//
// Replaces the per-sample Python loop of Open3D-ML PointPillars.voxelize that the reference calls at
// R:pixelspointspolygons/models/pointpillars/pointpillars_o3d.py:92.
//
// The reference sorts (hash, index) pairs per tile and keeps the first M indices of each run.  What the rest of the
// path consumes is, per key, the SET of its min(count, M) lowest-index points (the PFN is a max and an exact
// fixed-point mean over that set), so nothing is sorted here:
//
//   * a CTA takes a ticket (atomic counter) -> chunk of <= 4096 consecutive points of one tile.  Tickets, not
//     blockIdx, order the chunks, so a CTA only ever waits for CTAs that are already running.
//   * every lane hashes its <= 16 points (all loads in flight at once) with the reference's fp32 operation order;
//     only the packed (key, rank) word of a point stays in a register.
//   * each warp owns a contiguous segment of the chunk and counts it per key in a private shared-memory histogram,
//     32 consecutive points per step: read the key's count, then one shared-memory atomic per point; same-key lanes
//     of a step receive consecutive counts in arbitrary order.  No retry loops, no match.any.
//   * the chunk's per-key counts are published to global memory and a flag is released; the CTA acquires the flags
//     of the earlier chunks of its tile, sums their counts per key (exclusive prefix over chunks, then over its own
//     warps) and walks its registers again: rank = base + rank-in-segment; survivors (rank < M) re-read their xyz
//     (L1/L2 hit) and go to slots[tile][key][rank].  Inside one step same-key lanes hold their ranks in arbitrary
//     order, which only matters in the step where the key crosses M: that step re-ranks in lane order
//     (match.any, about 1 % of the steps).  The kept set is therefore exactly the reference's; the order inside a
//     pillar is not (export_kernel sorts by index for the parity surface).
//   * the last CTA of a tile to finish runs the tile's plan: keys in ascending order -> run ordinal (max_voxels
//     cut), cell coordinates (decoded from the key; from the lowest-index point when the run holds a point on the
//     x / y max face, i.e. under hash aliasing), x/y bound filter, final voxel order, and the canvas owner table
//     (last pillar in voxel order wins a cell, Appendix A.5).
//
// export_kernel (optional) dumps the reference-shaped tensors for the parity tests.
#include <type_traits>

#include "p3p_internal.cuh"

namespace p3p {

// Optional phase timeline (build with -DP3P_TIMELINE): thread 0 of every CTA stamps %globaltimer at the phase
// boundaries; tools/timeline.py reads them back through p3p_debug_timeline.
#ifdef P3P_TIMELINE
__device__ unsigned long long g_timeline[8192][16];
__device__ __forceinline__ void tl_stamp(int cta, int slot) {
    if (threadIdx.x == 0 && cta < 8192) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_timeline[cta][slot] = t;
    }
}
#define TL(cta, slot) tl_stamp(cta, slot)
extern "C" int p3p_debug_timeline(unsigned long long* host, int ctas) {
    return (int)cudaMemcpyFromSymbol(host, g_timeline, sizeof(unsigned long long) * 16 * (size_t)ctas);
}
#else
#define TL(cta, slot)
#endif

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

// predicated 16-bit shared-memory accesses by 32-bit shared address (no generic-pointer arithmetic, no branches)
__device__ __forceinline__ unsigned lds_u16(uint32_t addr, int pred) {
    unsigned v;
    asm volatile(
        "{\n .reg .pred p;\n .reg .b16 t;\n setp.ne.b32 p, %2, 0;\n mov.b16 t, 0;\n @p ld.shared.u16 t, [%1];\n cvt.u32.u16 %0, t;\n}"
        : "=r"(v)
        : "r"(addr), "r"(pred)
        : "memory");
    return v;
}
__device__ __forceinline__ unsigned lds_u32(uint32_t addr, int pred) {
    unsigned v;
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n mov.b32 %0, 0;\n @p ld.shared.u32 %0, [%1];\n}"
                 : "=r"(v)
                 : "r"(addr), "r"(pred)
                 : "memory");
    return v;
}

// predicated atomic increment of a 16-bit shared-memory counter (through its 32-bit word; counters never overflow into
// their neighbour); returns the counter's previous value
__device__ __forceinline__ unsigned atoms_add_u16(uint32_t addr, int pred) {
    unsigned r;
    const unsigned sh = (addr & 2u) * 8u;
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %3, 0;\n mov.b32 %0, 0;\n @p atom.shared.add.u32 %0, [%1], %2;\n}"
        : "=r"(r)
        : "r"(addr & ~3u), "r"(1u << sh), "r"(pred)
        : "memory");
    return (r >> sh) & 0xFFFFu;
}

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

// Per-tile plan, run by the CTA that finished the tile's last chunk.  Keff: keys that can be non-empty (the regular
// cells only, unless some point of the tile hashed beyond them).  scratch: >= (4 K + HW) ints of shared memory.
__device__ void plan_tile(const GridDev& g, const WsPtrs& ws, int b, int Keff, int* scratch, int* warp_tot) {
    const int K = g.num_keys, HW = g.ny * g.nx, M = g.M, tid = threadIdx.x;
    int* n_s = scratch;         // [K] min(count, M), 0 = empty
    int* cell_s = n_s + K;      // [K] packed cell of the run, -1 if dropped
    int* rkey_s = cell_s + K;   // [K] key of pillar r
    int* rn_s = rkey_s + K;     // [K] n of pillar r
    int* owner_s = rn_s + K;    // [HW]
    const int* totals = ws.totals + (size_t)b * K;
    const float4* slots = ws.slots + (size_t)b * K * M;
    const uint8_t* edge = ws.edge + (size_t)b * K;
    // counts of every key (written by other CTAs -> L2 loads, 8 independent loads per thread and batch); coordinates
    // come from the key itself unless the run holds an edge point, in which case its lowest-index kept point decides
    // (Appendix A.1: hash aliasing)
    for (int k0 = 0; k0 < Keff; k0 += kThreads * 8) {
        int t[8];
        uint8_t e[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k0 + u * kThreads + tid;
            t[u] = 0; e[u] = 0;
            if (k < Keff) { t[u] = __ldcg(totals + k); e[u] = __ldcg(edge + k); }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k0 + u * kThreads + tid;
            if (k >= Keff) continue;
            int cell = -1;
            const int n = t[u] < M ? t[u] : M;
            if (n > 0) {
                int cx, cy, cz;
                if (e[u]) {
                    float4 best = __ldcg(slots + (size_t)k * M);
                    for (int i = 1; i < n; ++i) {
                        const float4 p = __ldcg(slots + (size_t)k * M + i);
                        if (__float_as_int(p.w) < __float_as_int(best.w)) best = p;
                    }
                    point_cell(g, best.x, best.y, best.z, cx, cy, cz);
                } else {
                    cz = k / g.stride2;
                    const int rem = k - cz * g.stride2;
                    cy = rem / g.stride1;
                    cx = rem - cy * g.stride1;
                }
                if (cy < g.nv[1] && cx < g.nv[0]) cell = cx | (cy << 10) | (cz << 20);  // x/y bound filter (A.2)
            }
            n_s[k] = n;
            cell_s[k] = cell;
        }
    }
    for (int i = tid; i < HW; i += kThreads) owner_s[i] = -1;
    __syncthreads();
    const int per = (Keff + kThreads - 1) / kThreads;
    const int k0 = tid * per, k1 = (k0 + per < Keff) ? k0 + per : Keff;
    int cnt = 0;
    for (int k = k0; k < k1; ++k) cnt += (n_s[k] > 0);
    int total_runs;
    int ord = block_exclusive_scan(cnt, warp_tot, &total_runs);
    int cnt2 = 0;
    for (int k = k0; k < k1; ++k) {
        if (n_s[k] > 0) {
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
        const int n = n_s[k];
        ws.pil_key[pi] = k;
        ws.pil_n[pi] = n;
        ws.pil_coord[pi] = st;
        rkey_s[r] = k;
        rn_s[r] = n;
        const int cx = st & 1023, cy = (st >> 10) & 1023;
        atomicMax(&owner_s[cy * g.nx + cx], r);  // scatter collisions: the later row (higher hash) wins (A.5)
    }
    if (tid == 0) ws.num_pil[b] = total_pil;
    __syncthreads();
    for (int i = tid; i < HW; i += kThreads) {
        const int o = owner_s[i];
        ws.owner[(size_t)b * HW + i] = o;
        ws.cell_desc[(size_t)b * HW + i] = (o >= 0) ? (rkey_s[o] | (rn_s[o] << 16)) : -1;
    }
}

// packed per-point word: key (13 bits) | count of the key before the point in the warp segment (10 bits) << 13 |
// position among the same-key lanes of the point's step (5 bits) << 23
constexpr int kKeyBits = 13, kOldBits = 10;
constexpr unsigned kCntMask = (1u << kOldBits) - 1u;
static_assert(kMaxKeys <= (1 << kKeyBits), "key field too narrow");
static_assert(kMaxChunkPoints / kWarps <= (1 << (kOldBits - 1)), "segment count field too narrow");
static_assert(kMaxChunkPoints * 8 < 65536, "8 chunk rows must add up inside 16-bit lanes");

// kFast: packed xyz (stride 3), no per-point hash export, hashes beyond the cells kept (the shipped configuration): the
// per-point branches on those options are resolved at compile time.
template <bool kFast>
__global__ void __launch_bounds__(kThreads, 3)
voxelize_kernel(const float* __restrict__ pts, int stride_arg, const int64_t* __restrict__ offsets, int B, GridDev g, int S,
                WsPtrs ws, int32_t* __restrict__ point_hash, int need_plan, int pdl_from) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = g.num_keys;
    const int Kp = ws.key_stride;  // K rounded up to 8: row stride of chunk_hist (16-byte rows)
    // [kWarps][Kp] (16-byte rows): walk 1: points of the key in the warp's segment; afterwards: points of the key in the earlier warps of the CTA
    uint16_t* hist = reinterpret_cast<uint16_t*>(smem_raw);
    uint16_t* ctot_s = reinterpret_cast<uint16_t*>(smem_raw + (size_t)kWarps * Kp * sizeof(uint16_t));  // [Kp] chunk totals
    unsigned* prefix_s = reinterpret_cast<unsigned*>(ctot_s + Kp);  // [Kp] points of the key in the earlier chunks of the tile
    __shared__ ChunkLoc loc;
    __shared__ int s_ticket, s_last, s_runs[2];
    __shared__ int warp_tot[kWarps];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int stride = kFast ? 3 : stride_arg;
    TL(blockIdx.x, 0);
    if (tid == 0) s_ticket = (int)atomicAdd(ws.sync, 1u);
    {
        uint4* h4 = reinterpret_cast<uint4*>(hist);
        const int n16 = kWarps * Kp * (int)sizeof(uint16_t) / 16;
        for (int i = tid; i < n16; i += kThreads) h4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const int ticket = s_ticket;
    // Let the dependent PFN grid start its prologue on the SMs this grid has left (it waits for this grid's completion
    // with griddepcontrol.wait before it reads anything written here).  Only the last tickets trigger early: the
    // dependent grid launches once every CTA has triggered or exited, i.e. when no CTA of this grid is still waiting
    // for an SM that a waiting PFN CTA could occupy.
    if (ticket >= pdl_from) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (w == 0) locate_chunk(ticket, offsets, B, S, &loc);
    // tiles without points have no chunk: ticket t < B writes the empty plan of tile t
    if (ticket < B && offsets[ticket + 1] <= offsets[ticket]) write_empty_tile(g, ws, ticket);
    __syncthreads();
    if (loc.b < 0) return;
    TL(blockIdx.x, 1);

    unsigned* flags = ws.sync + 1;
    unsigned* tile_done = flags + ws.max_chunks;
    const int segS = S / kWarps;  // multiple of 32
    const long long seg0 = loc.p0 + (long long)w * segS;
    const int seg_n = (int)(((seg0 + segS < loc.p1) ? seg0 + segS : loc.p1) - seg0);  // points of this warp (may be <= 0)
    uint16_t* myhist = hist + (size_t)w * Kp;
    uint8_t* edge = ws.edge + (size_t)loc.b * K;
    const float* seg_pts = pts + seg0 * stride;

    // ---- hash + walk 1, software pipelined: the loads of the second half are in flight while the first half is
    //      counted.  Walk 1 = per-warp histogram of this warp's contiguous segment, 32 consecutive points per step.
    int pk[kIters];
    // bit 0: some key may lie outside the regular cells or be aliased (a point on a max face: z == z_max, y == y_max, ...)
    // bit 1: some point sits on the x / y max face (its run's coordinates are not decodable from the key)
    int hi_key = 0;
    constexpr int kBatch = kIters / 2;
    float px[kBatch], py[kBatch], pz[kBatch];
    // `full` (a std::bool_constant) resolves the bounds tests of a full 512-point segment at compile time: straight-line
    // code, loads at immediate offsets from one lane pointer (every chunk but the last of a tile).
    const float* lane_pts = seg_pts + (size_t)lane * stride;
    auto load_batch = [&](int j0, auto full) {
        constexpr bool kFull = decltype(full)::value;
#pragma unroll
        for (int jj = 0; jj < kBatch; ++jj) {
            const int i = (j0 + jj) * 32 + lane;
            if (!kFull) { px[jj] = 0.f; py[jj] = 0.f; pz[jj] = 0.f; }
            if (kFull || i < seg_n) {
                const float* p = lane_pts + (size_t)((j0 + jj) * 32) * stride;
                px[jj] = __ldg(p); py[jj] = __ldg(p + 1); pz[jj] = __ldg(p + 2);
            }
        }
    };
    auto hash_batch = [&](int j0, auto full) {
        constexpr bool kFull = decltype(full)::value;
        unsigned edge_bits = 0;  // points on the x / y max face (rare: one vote per batch, flags set on a slow path)
        int kmax = -1;
#pragma unroll
        for (int jj = 0; jj < kBatch; ++jj) {
            const int j = j0 + jj, i = j * 32 + lane;
            pk[j] = -1;
            if (kFull || i < seg_n) {
                bool on_edge;
                const int k = point_key<!kFast>(g, px[jj], py[jj], pz[jj], on_edge);  // (kFast: the launcher checked the overflow flag)
                pk[j] = k;
                edge_bits |= on_edge ? (1u << j) : 0u;
                kmax = k > kmax ? k : kmax;
                if (!kFast && point_hash) point_hash[seg0 + i] = k;
            }
        }
        hi_key |= ((kmax >= g.num_cells) ? 1 : 0) | (edge_bits ? 3 : 0);
        if (__any_sync(0xffffffffu, edge_bits != 0)) {
#pragma unroll
            for (int jj = 0; jj < kBatch; ++jj)
                if ((edge_bits >> (j0 + jj)) & 1u) edge[pk[j0 + jj]] = 1;  // idempotent flag, read by the tile's plan
        }
    };
    const uint32_t myhist_sa = smem_u32(myhist);
    auto count_batch = [&](int j0, auto full) {
        constexpr bool kFull = decltype(full)::value;
#pragma unroll
        for (int jj = 0; jj < kBatch; ++jj) {
            const int j = j0 + jj;
            if (kFull || j * 32 < seg_n) {  // warp-uniform
                const int k = pk[j];
                const int act = k >= 0 ? 1 : 0;
                const uint32_t sa = myhist_sa + 2u * (unsigned)(act ? k : 0);
                // count of the key before this step, then one atomic per point: same-key lanes of the step receive
                // consecutive counts in arbitrary order (shared-memory operations of a warp execute in program order)
                const unsigned before = lds_u16(sa, act);
                const unsigned old = atoms_add_u16(sa, act);
                if (act) pk[j] = k | (int)(old << kKeyBits) | (int)((old - before) << (kKeyBits + kOldBits));
            }
        }
    };
    auto walk1 = [&](auto full) {
        load_batch(0, full);
        hash_batch(0, full);
        load_batch(kBatch, full);
        TL(blockIdx.x, 2);
        count_batch(0, full);
        hash_batch(kBatch, full);
        count_batch(kBatch, full);
    };
    const bool full_seg = (seg_n == kIters * 32);  // warp-uniform: all 16 steps of the warp hold 32 points
    if (full_seg) walk1(std::true_type{}); else walk1(std::false_type{});
    hi_key = (__syncthreads_or(hi_key & 1) ? 1 : 0) | (__syncthreads_or(hi_key & 2) ? 2 : 0);  // (the intrinsic ORs predicates)
    TL(blockIdx.x, 3);
    if (tid < 2) s_runs[tid] = 0;
    // ---- publish this chunk's per-key counts (full rows: zeros beyond the keys in use) ------------------------------
    const int Kreg = (g.num_cells + 7) / 8 * 8 < Kp ? (g.num_cells + 7) / 8 * 8 : Kp;  // regular cells, 16-byte granular
    {
        // 8 keys per thread: 16-byte rows, counts added inside their 16-bit lanes (a chunk holds <= 4096 points);
        // keys nobody hashed to stay zero, so the published rows are complete
        uint4* dst = reinterpret_cast<uint4*>(ws.chunk_hist + (size_t)ticket * Kp);
        for (int k8 = tid; k8 < Kp / 8; k8 += kThreads) {
            uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int ww = 0; ww < kWarps; ++ww) {
                uint4* cell = reinterpret_cast<uint4*>(hist + (size_t)ww * Kp) + k8;
                const uint4 t = *cell;
                *cell = acc;  // points of the keys in the earlier warps
                acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
            }
            reinterpret_cast<uint4*>(ctot_s)[k8] = acc;
            dst[k8] = acc;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(flags + ticket, 1u | ((unsigned)hi_key << 1));  // bit 0: published
        TL(blockIdx.x, 4);
        // wait for the earlier chunks of the tile; learn whether any of them holds keys beyond the regular cells
        int hi_before = 0;
        for (int cc = tid; cc < loc.c; cc += kThreads) {
            unsigned f;
            while ((f = ld_acquire(flags + loc.gstart + cc)) == 0u) {
            }
            hi_before |= (int)(f >> 1);
        }
        hi_key |= (__syncthreads_or(hi_before & 1) ? 1 : 0) | (__syncthreads_or(hi_before & 2) ? 2 : 0);
    }
    TL(blockIdx.x, 5);
    // ---- prefix over the earlier chunks: 8 keys per thread with 16-byte L2 loads, 8 rows in flight per batch -----------
    const int M = g.M;
    const bool last_chunk = (loc.c == loc.nchunks - 1);
    const int Kuse = (hi_key & 1) ? Kp : Kreg;  // keys that can be non-empty in this tile so far
    int open_keys = 0;  // keys of this chunk that still have room (base < M): if none, nothing of the chunk is kept
    for (int k8 = tid; k8 < Kp / 8; k8 += kThreads) {
        unsigned acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (k8 * 8 < Kuse) {
            const uint4* src = reinterpret_cast<const uint4*>(ws.chunk_hist + (size_t)loc.gstart * Kp) + k8;
            const size_t row = (size_t)Kp / 8;
            const uint4* rp = src;
            auto add_rows = [&](const uint4 (&v)[8]) {
                // 8 rows of 16-bit counts (<= 4096 each) add up inside their 16-bit lanes without a carry
                uint4 sum = v[0];
#pragma unroll
                for (int u = 1; u < 8; ++u) { sum.x += v[u].x; sum.y += v[u].y; sum.z += v[u].z; sum.w += v[u].w; }
                acc[0] += sum.x & 0xFFFFu; acc[1] += sum.x >> 16; acc[2] += sum.y & 0xFFFFu; acc[3] += sum.y >> 16;
                acc[4] += sum.z & 0xFFFFu; acc[5] += sum.z >> 16; acc[6] += sum.w & 0xFFFFu; acc[7] += sum.w >> 16;
            };
            int c0 = 0;
            for (; c0 + 8 <= loc.c; c0 += 8, rp += 8 * row) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldcg(rp + u * row);
                add_rows(v);
            }
            if (c0 < loc.c) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = (c0 + u < loc.c) ? __ldcg(rp + u * row) : make_uint4(0, 0, 0, 0);
                add_rows(v);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k8 * 8 + u;
            const unsigned own = ctot_s[k];
            prefix_s[k] = acc[u];
            open_keys |= (own > 0 && acc[u] < (unsigned)M) ? 1 : 0;
            if (last_chunk) {
                const unsigned tot = acc[u] + own;
                if (k < K) ws.totals[(size_t)loc.b * K + k] = (int)tot;
                ctot_s[k] = (uint16_t)(tot < (unsigned)M ? tot : (unsigned)M);  // from here on: min(count, M) of the tile
            }
        }
    }
    open_keys = __syncthreads_or(open_keys);
    TL(blockIdx.x, 6);
    // ---- the tile's last chunk knows every count: when no run can be cut, filtered or aliased, the canvas owner table
    //      follows from the keys directly and the tile needs no plan ------------------------------
    if (last_chunk) {
        int direct = 0, hi_ok = 0;
        if (!need_plan && !(hi_key & 2)) {
            int r = 0, h = 0;
            for (int k = tid; k < Kuse && k < K; k += kThreads) {
                const int nz = ctot_s[k] > 0 ? 1 : 0;
                if (k < g.num_cells) r += nz; else h += nz;
            }
            r = __reduce_add_sync(0xffffffffu, r);
            h = __reduce_add_sync(0xffffffffu, h);
            if (lane == 0) { atomicAdd(&s_runs[0], r); atomicAdd(&s_runs[1], h); }
            __syncthreads();
            const int R = s_runs[0], H = s_runs[1];
            // runs are numbered in key order and the first max_voxels survive: regular cells all survive iff R <= Vmax;
            // the runs beyond them survive all (R + H <= Vmax) or not at all (R == Vmax); anything else needs the plan
            hi_ok = (R + H <= g.Vmax) ? 1 : 0;
            direct = (R <= g.Vmax) && (H == 0 || hi_ok || R == g.Vmax);
        }
        if (direct) {
            const int HW = g.ny * g.nx;
            const int top = hi_ok ? g.ext[2] : g.ext[2] - 1;  // highest z layer whose runs survive
            for (int cell = tid; cell < HW; cell += kThreads) {
                const int cy = cell / g.nx, cx = cell - cy * g.nx;
                int d = -1;
                if (cx < g.nv[0] && cy < g.nv[1]) {
                    for (int cz = top; cz >= 0; --cz) {  // the last pillar in voxel order (highest key) owns the cell
                        const int k = cx + cy * g.stride1 + cz * g.stride2;
                        const int n = (k < Kuse && k < K) ? (int)ctot_s[k] : 0;
                        if (n > 0) { d = k | (n << 16); break; }
                    }
                }
                ws.cell_desc[(size_t)loc.b * HW + cell] = d;
            }
        }
        if (tid == 0) ws.tile_hi[loc.b] = (hi_key & 1) | (direct << 1);
    }
    // ---- walk 2: rank = earlier chunks + earlier warps + rank in segment; scatter the survivors ---------------------
    auto walk2 = [&](auto full) {
        constexpr bool kFull = decltype(full)::value;
        // pass A: ranks (registers + shared memory only); pk[j] becomes key | rank << 13 | group position << 24 (-1: no key).
        // Same-key lanes of a step own their ranks in arbitrary order, which only matters in the one step where the key
        // crosses M: those steps are found with ONE warp vote for the whole chunk and re-ranked in lane (= index) order.
        const uint32_t prefix_sa = smem_u32(prefix_s);
        constexpr int kRankBits = 11;  // provisional rank field (clamped): M + 31 < 2^11
        unsigned cross = 0;
#pragma unroll
        for (int j = 0; j < kIters; ++j) {
            if (kFull || j * 32 < seg_n) {  // warp-uniform
                const int kj = pk[j] < 0 ? -1 : (pk[j] & ((1 << kKeyBits) - 1));
                const int old = (pk[j] >> kKeyBits) & (int)kCntMask, rnd = (pk[j] >> (kKeyBits + kOldBits)) & 31;
                const int act = kj >= 0 ? 1 : 0;
                const unsigned pre = act ? lds_u32(prefix_sa + 4u * (unsigned)kj, act) : (unsigned)M;
                const unsigned wbase = lds_u16(myhist_sa + 2u * (unsigned)(act ? kj : 0), act);
                unsigned rank = pre < (unsigned)M ? pre + wbase + (unsigned)old : (unsigned)M;
                if ((rank >= (unsigned)M) && (rank < (unsigned)(M + rnd))) cross |= 1u << j;
                rank = rank < (1u << kRankBits) - 1u ? rank : (1u << kRankBits) - 1u;
                pk[j] = act ? (kj | (int)(rank << kKeyBits) | (rnd << (kKeyBits + kRankBits))) : -1;
            } else {
                pk[j] = -1;
            }
        }
        cross = __reduce_or_sync(0xffffffffu, cross);
        if (cross) {  // warp-uniform; about 1 % of the steps are flagged
#pragma unroll
            for (int j = 0; j < kIters; ++j) {
                if (cross & (1u << j)) {
                    const int kj = pk[j] < 0 ? -1 : (pk[j] & ((1 << kKeyBits) - 1));
                    unsigned rank = (unsigned)(pk[j] >> kKeyBits) & ((1u << kRankBits) - 1u);
                    const unsigned rnd = (unsigned)(pk[j] >> (kKeyBits + kRankBits)) & 31u;
                    const unsigned m = __match_any_sync(0xffffffffu, kj);
                    // lanes of a run whose first count lies below M (the others keep a rank >= M)
                    if (kj >= 0 && rank - rnd < (unsigned)M) {
                        rank = rank - rnd + (unsigned)__popc(m & ((1u << lane) - 1u));
                        pk[j] = kj | (int)(rank << kKeyBits);
                    }
                }
            }
        }
        TL(blockIdx.x, 7);
        // pass B: survivors re-read their xyz (L1 / L2 hits), all loads of a batch in flight before the first store
        float4* tile_slots = ws.slots + (size_t)loc.b * K * M;
#pragma unroll
        for (int j0 = 0; j0 < kIters; j0 += kBatch) {
#pragma unroll
            for (int jj = 0; jj < kBatch; ++jj) {
                // survivors: a key and a (provisional or re-ranked) rank below M
                if (pk[j0 + jj] >= 0 && (((unsigned)pk[j0 + jj] >> kKeyBits) & ((1u << kRankBits) - 1u)) >= (unsigned)M) pk[j0 + jj] = -1;
                if (pk[j0 + jj] >= 0) {
                    const float* p = lane_pts + (size_t)((j0 + jj) * 32) * stride;
                    px[jj] = __ldg(p); py[jj] = __ldg(p + 1); pz[jj] = __ldg(p + 2);
                }
            }
#pragma unroll
            for (int jj = 0; jj < kBatch; ++jj) {
                const int j = j0 + jj, i = j * 32 + lane;
                if (pk[j] >= 0) {
                    const int kj = pk[j] & ((1 << kKeyBits) - 1), rank = (pk[j] >> kKeyBits) & ((1 << kRankBits) - 1);
                    tile_slots[(size_t)kj * M + rank] =
                        make_float4(px[jj], py[jj], pz[jj], __int_as_float((int)(seg0 + i - loc.tile_start)));
                }
            }
        }
    };
    if (open_keys) {
        if (full_seg) walk2(std::true_type{}); else walk2(std::false_type{});
    }
    // ---- the last CTA of the tile to get here plans the tile ------------------------------------------------------------
    TL(blockIdx.x, 8);
    __threadfence();
    __syncthreads();
    TL(blockIdx.x, 9);
    if (tid == 0) s_last = (atomicAdd(tile_done + loc.b, 1u) == (unsigned)(loc.nchunks - 1)) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    TL(blockIdx.x, 10);
    const int tile_hi = __ldcg(ws.tile_hi + loc.b);
    if (tile_hi & 2) return;  // owner table already written by the tile's last chunk
    plan_tile(g, ws, loc.b, (tile_hi & 1) ? K : (g.num_cells < K ? g.num_cells : K), reinterpret_cast<int*>(smem_raw), warp_tot);
    __syncthreads();
    TL(blockIdx.x, 11);
}

// Parity surface (p3p_voxel_outputs): the reference lists a pillar's points by ascending index; slots are unordered.
__global__ void __launch_bounds__(128)
export_kernel(GridDev g, int B, WsPtrs ws, p3p_voxel_outputs out) {
    __shared__ int idx_s[1024];
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
        __syncthreads();
        for (int s = threadIdx.x; s < n; s += blockDim.x) idx_s[s] = __float_as_int(slot[s].w);
        __syncthreads();
        for (int s = threadIdx.x; s < g.M; s += blockDim.x) {
            int pos = s;  // padding rows stay where they are
            float4 p = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
            if (s < n) {
                p = slot[s];
                const int me = idx_s[s];
                pos = 0;
                for (int i = 0; i < n; ++i) pos += (idx_s[i] < me) ? 1 : 0;  // indices are distinct
            }
            if (out.pillar_point_idx) out.pillar_point_idx[pi * g.M + pos] = __float_as_int(p.w);
            if (out.pillar_points) {
                float* d = out.pillar_points + (pi * g.M + pos) * 3;
                d[0] = p.x; d[1] = p.y; d[2] = p.z;
            }
        }
    }
}

}  // namespace

static size_t voxelize_smem_bytes(const GridDev& g) {
    const size_t kp = ((size_t)g.num_keys + 7) / 8 * 8;
    const size_t hist = (size_t)kWarps * kp * sizeof(uint16_t) + kp * (sizeof(uint16_t) + sizeof(unsigned));
    const size_t plan = (size_t)(4 * g.num_keys + g.ny * g.nx) * sizeof(int);
    return ((hist > plan ? hist : plan) + 15) / 16 * 16;
}

int launch_voxelize(const float* pts, int stride, const int64_t* offsets, int B, int64_t total, const GridDev& g,
                    const WsLayout& l, const WsPtrs& ws, int32_t* point_hash, int need_plan, cudaStream_t st) {
    (void)total;
    const size_t smem = voxelize_smem_bytes(g);
    static bool attr_done = false;
    if (!attr_done) {
        P3P_CUDA_CHECK(cudaFuncSetAttribute(voxelize_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        P3P_CUDA_CHECK(cudaFuncSetAttribute(voxelize_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    // ticket counter, chunk flags and per-tile completion counters start at zero for every call
    P3P_CUDA_CHECK(cudaMemsetAsync(ws.sync, 0, l.sync_bytes, st));
    const int pdl_from = l.max_chunks - device_sm_count();  // tickets of the last (partial) wave
    if (stride == 3 && point_hash == nullptr && !(g.flags & kFlagDropOverflow))
        voxelize_kernel<true><<<l.max_chunks, kThreads, smem, st>>>(pts, stride, offsets, B, g, l.chunk_points, ws, point_hash, need_plan, pdl_from);
    else
        voxelize_kernel<false><<<l.max_chunks, kThreads, smem, st>>>(pts, stride, offsets, B, g, l.chunk_points, ws, point_hash, need_plan, pdl_from);
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
