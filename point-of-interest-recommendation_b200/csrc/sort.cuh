// sort.cuh -- the integer side of the hot path: stable radix sort of (row index, occurrence id)
// pairs, exclusive scan, and segment (= sorted-unique) construction.  This replaces Theano's
// `Unique(False,False,False)` (GRU.py:329-331, GRU_Spatial.py:149-153, GeoIE.py:147-153) and,
// because the sort keeps every occurrence, also gives the duplicate lists the sparse update
// needs in order to sum gradients in a fixed order.  All results are bit-exact integers.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// exclusive scan (uint32), multi-level
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_block(const uint32_t* in, uint32_t* out, uint32_t* block_sums, int64_t n) {
    __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)tid * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + i;
        v[i] = idx < n ? in[idx] : 0u;
        sum += v[i];
    }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t t = lane < SCAN_THREADS / 32 ? warp_tot[lane] : 0u;
        uint32_t ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t s = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += s;
        }
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = ti - t;
        if (lane == SCAN_THREADS / 32 - 1) block_sums[blockIdx.x] = ti;
    }
    __syncthreads();
    uint32_t excl = warp_tot[wid] + inc - sum;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + i;
        if (idx < n) out[idx] = excl;
        excl += v[i];
    }
}

__global__ void k_scan_add(uint32_t* out, const uint32_t* block_off, int64_t n) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) out[idx] += block_off[idx / SCAN_TILE];
}

// up to SCAN_SINGLE_MAX elements in ONE launch.  Warp w owns the contiguous run [w * rpw * 32, (w + 1) * rpw * 32) and walks
// it in rows of 32 (coalesced): pass 1 adds the rows up to the warp's total, the 32 warp totals are scanned in the block,
// pass 2 walks the rows again (L1 / L2 hits) writing exclusive prefixes with a running carry.  out may alias in.
constexpr int SCAN_SINGLE_MAX = 65536;
__global__ void __launch_bounds__(1024)
k_scan_single(const uint32_t* in, uint32_t* out, int n, uint32_t* total_out) {
    __shared__ uint32_t warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int rpw = (n + 1023) / 1024;                     // rows of 32 per warp
    const int b0 = wid * rpw * 32;
    uint32_t sum = 0;
    for (int r = 0; r < rpw; ++r) { const int i = b0 + r * 32 + lane; if (i < n) sum += in[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) warp_tot[wid] = sum;
    __syncthreads();
    if (wid == 0) {
        uint32_t t = warp_tot[lane], ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t s2 = __shfl_up_sync(0xffffffffu, ti, o); if (lane >= o) ti += s2; }
        warp_tot[lane] = ti - t;
        if (lane == 31 && total_out) *total_out = ti;
    }
    __syncthreads();
    uint32_t carry = warp_tot[wid];
    for (int r = 0; r < rpw; ++r) {
        const int i = b0 + r * 32 + lane;
        const uint32_t v = i < n ? in[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (i < n) out[i] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// out may alias in.  total_dev (optional) receives the sum of all elements.
static int exclusive_scan_u32(poi_engine* e, const uint32_t* in, uint32_t* out, int64_t n,
                              uint32_t* total_dev) {
    if (n <= 0) {
        if (total_dev) POI_CK(e, cudaMemsetAsync(total_dev, 0, 4, e->stream));
        return 0;
    }
    if (n <= SCAN_SINGLE_MAX) {
        POI_LAUNCH(e, k_scan_single, 1, 1024, 0, in, out, (int)n, total_dev);
        return 0;
    }
    int64_t nb = poi_cdiv(n, SCAN_TILE);
    uint32_t* bs = nullptr;
    POI_TRY(arena_get(e, (size_t)nb + 1, &bs));
    POI_LAUNCH(e, k_scan_block, (unsigned)nb, SCAN_THREADS, 0, in, out, bs, n);
    if (nb > 1) {
        POI_TRY(exclusive_scan_u32(e, bs, bs, nb, total_dev));
        POI_LAUNCH(e, k_scan_add, (unsigned)poi_cdiv(n, 256), 256, 0, out, bs, n);
    } else if (total_dev) {
        POI_CK(e, cudaMemcpyAsync(total_dev, bs, 4, cudaMemcpyDeviceToDevice, e->stream));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// small path: one CTA, bitonic sort of 64-bit (key<<32 | occurrence) composites in shared memory.
// The occurrence id in the low word makes the order total, hence identical to a stable sort.
// ---------------------------------------------------------------------------------------------
constexpr int SMALL_SORT_MAX = 4096;

__global__ void __launch_bounds__(1024)
k_small_sort(const uint32_t* keys_in, int n, int np2, uint32_t* keys_out, uint32_t* vals_out) {
    extern __shared__ unsigned long long s_comp[];
    for (int i = threadIdx.x; i < np2; i += blockDim.x)
        s_comp[i] = i < n ? (((unsigned long long)keys_in[i] << 32) | (unsigned)i) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    bool asc = (i & k) == 0;
                    unsigned long long a = s_comp[i], b = s_comp[ixj];
                    if ((a > b) == asc) { s_comp[i] = b; s_comp[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        unsigned long long c = s_comp[i];
        keys_out[i] = (uint32_t)(c >> 32);
        vals_out[i] = (uint32_t)(c & 0xffffffffu);
    }
}

// The same sort followed by the segment arrays, in the same launch (lists <= SMALL_SORT_MAX keys: every index list of a
// one-by-one call, the reference's own mode -- 4 launches per list became 1).  Thread t owns the np2 / blockDim consecutive
// sorted entries [t * per, (t + 1) * per): head flags, a block scan of the per-thread head counts, then the writes.
__global__ void __launch_bounds__(1024)
k_small_sort_seg(const uint32_t* keys_in, int n, int np2, uint32_t* keys_out, uint32_t* vals_out,
                 uint32_t* seg_start, uint32_t* uniq, uint32_t* n_unique, uint32_t* seg_of_occ) {
    extern __shared__ unsigned long long s_comp[];
    __shared__ uint32_t warp_tot[32];
    for (int i = threadIdx.x; i < np2; i += blockDim.x)
        s_comp[i] = i < n ? (((unsigned long long)keys_in[i] << 32) | (unsigned)i) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    bool asc = (i & k) == 0;
                    unsigned long long a = s_comp[i], b = s_comp[ixj];
                    if ((a > b) == asc) { s_comp[i] = b; s_comp[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (np2 + (int)blockDim.x - 1) / (int)blockDim.x;
    const int i0 = tid * per, i1 = min(n, i0 + per);
    uint32_t heads = 0;
    for (int i = i0; i < i1; ++i) {
        const uint32_t k = (uint32_t)(s_comp[i] >> 32);
        heads += (i == 0 || k != (uint32_t)(s_comp[i - 1] >> 32)) ? 1u : 0u;
    }
    uint32_t inc = heads;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    uint32_t ex = inc - heads;
    for (int ww = 0; ww < wid; ++ww) ex += warp_tot[ww];
    for (int i = i0; i < i1; ++i) {
        const unsigned long long c = s_comp[i];
        const uint32_t k = (uint32_t)(c >> 32), v = (uint32_t)(c & 0xffffffffu);
        const bool head = i == 0 || k != (uint32_t)(s_comp[i - 1] >> 32);
        const uint32_t sid = head ? ex : ex - 1u;
        keys_out[i] = k; vals_out[i] = v;
        if (head) { seg_start[sid] = (uint32_t)i; uniq[sid] = k; ++ex; }
        if (seg_of_occ) seg_of_occ[v] = sid;
        if (i == n - 1) { *n_unique = sid + 1u; seg_start[sid + 1u] = (uint32_t)n; }
    }
}

// ---------------------------------------------------------------------------------------------
// large path: LSD radix sort, 8 bits per pass, stable
// ---------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS)
k_radix_hist(const uint32_t* keys, int64_t n, int shift, uint32_t* hist, int nblocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        int64_t idx = base + j * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// Warp w owns the RS_ITEMS consecutive 32-key chunks (w*RS_ITEMS + j); ranks are assigned in
// index order inside a chunk (match_any + popc of lower lanes), chunk after chunk inside a warp,
// warp after warp inside the CTA, CTA after CTA via the scanned histogram -> stable.
__global__ void __launch_bounds__(RS_THREADS)
k_radix_scatter(const uint32_t* keys_in, const uint32_t* vals_in, int64_t n, int shift,
                const uint32_t* offs, int nblocks, uint32_t* keys_out, uint32_t* vals_out) {
    __shared__ uint32_t wcount[RS_THREADS / 32][256];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < (RS_THREADS / 32) * 256; i += RS_THREADS) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * RS_ITEMS * 32;
    uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        int64_t idx = base + j * 32 + lane;
        bool ok = idx < n;
        key[j] = ok ? keys_in[idx] : 0u;
        val[j] = ok ? (vals_in ? vals_in[idx] : (uint32_t)idx) : 0u;
        uint32_t dig = ok ? ((key[j] >> shift) & 255u) : 0xffffffffu;
        uint32_t peers = __match_any_sync(0xffffffffu, dig);
        uint32_t before = 0;
        if (ok) before = wcount[w][dig];
        __syncwarp();
        if (ok && (peers & lt_mask) == 0) wcount[w][dig] = before + __popc(peers);
        __syncwarp();
        rank[j] = before + __popc(peers & lt_mask);
    }
    __syncthreads();
    {   // thread == digit: turn per-warp counts into per-warp bases (global offset included)
        uint32_t run = offs[(size_t)tid * nblocks + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < RS_THREADS / 32; ++ww) {
            uint32_t c = wcount[ww][tid];
            wcount[ww][tid] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        int64_t idx = base + j * 32 + lane;
        if (idx < n) {
            uint32_t dig = (key[j] >> shift) & 255u;
            uint32_t pos = wcount[w][dig] + rank[j];
            keys_out[pos] = key[j];
            vals_out[pos] = val[j];
        }
    }
}

struct SegList;
static int fused_sort_launch(poi_engine* e, const uint32_t* keys_in, int64_t n, uint32_t bound, uint32_t* k0, uint32_t* v0, SegList* seg);

// Sort (key, occurrence-id) pairs by key, stable.  keys < bound.  Outputs live in the arena.
static int sort_pairs(poi_engine* e, const uint32_t* keys_in, int64_t n, uint32_t bound,
                      uint32_t** keys_sorted, uint32_t** vals_sorted) {
    uint32_t *k0 = nullptr, *v0 = nullptr;
    POI_TRY(arena_get(e, (size_t)std::max<int64_t>(n, 1), &k0));
    POI_TRY(arena_get(e, (size_t)std::max<int64_t>(n, 1), &v0));
    *keys_sorted = k0; *vals_sorted = v0;
    if (n <= 0) return 0;
    if (n <= SMALL_SORT_MAX) {
        int np2 = 32;
        while (np2 < n) np2 <<= 1;
        int threads = std::min(1024, std::max(32, np2 / 2));
        POI_LAUNCH(e, k_small_sort, 1, threads, (size_t)np2 * 8, keys_in, (int)n, np2, k0, v0);
        return 0;
    }
    if (e->fused_sort) return fused_sort_launch(e, keys_in, n, bound, k0, v0, nullptr);
    int bits = 1;
    while (bits < 32 && (1ull << bits) < (unsigned long long)bound) ++bits;
    int passes = (bits + 7) / 8;
    uint32_t *k1 = nullptr, *v1 = nullptr, *hist = nullptr;
    int nblocks = (int)poi_cdiv(n, RS_TILE);
    if (passes > 1) {
        POI_TRY(arena_get(e, (size_t)n, &k1));
        POI_TRY(arena_get(e, (size_t)n, &v1));
    }
    POI_TRY(arena_get(e, (size_t)256 * nblocks, &hist));
    // ping-pong so that the final pass lands in (k0, v0)
    const uint32_t* ck = keys_in; const uint32_t* cv = nullptr;
    for (int p = 0; p < passes; ++p) {
        bool to0 = ((passes - 1 - p) % 2) == 0;
        uint32_t* ok = to0 ? k0 : k1; uint32_t* ov = to0 ? v0 : v1;
        POI_LAUNCH(e, k_radix_hist, nblocks, RS_THREADS, 0, ck, n, p * 8, hist, nblocks);
        POI_TRY(exclusive_scan_u32(e, hist, hist, (int64_t)256 * nblocks, nullptr));
        POI_LAUNCH(e, k_radix_scatter, nblocks, RS_THREADS, 0, ck, cv, n, p * 8, hist, nblocks, ok, ov);
        ck = ok; cv = ov;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// segments = runs of equal keys in the sorted order
// ---------------------------------------------------------------------------------------------
struct SegList {
    uint32_t* keys = nullptr;       // sorted keys              [n]
    uint32_t* vals = nullptr;       // occurrence ids, ascending inside a segment [n]
    uint32_t* seg_start = nullptr;  // [n_unique + 1]
    uint32_t* uniq = nullptr;       // sorted unique keys       [n_unique]
    uint32_t* n_unique = nullptr;   // device scalar
    uint32_t* seg_of_occ = nullptr; // optional inverse map occurrence -> segment [n]
    int64_t n = 0;
};

__global__ void k_seg_flags(const uint32_t* keys, int64_t n, uint32_t* flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

__global__ void k_seg_write(const uint32_t* keys, const uint32_t* vals, const uint32_t* excl, int64_t n,
                            uint32_t* seg_start, uint32_t* uniq, uint32_t* n_unique,
                            uint32_t* seg_of_occ) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool head = (i == 0 || keys[i] != keys[i - 1]);
    uint32_t sid = head ? excl[i] : excl[i] - 1u;
    if (head) { seg_start[sid] = (uint32_t)i; uniq[sid] = keys[i]; }
    if (seg_of_occ) seg_of_occ[vals[i]] = sid;
    if (i == n - 1) { uint32_t nu = sid + 1u; *n_unique = nu; seg_start[nu] = (uint32_t)n; }
}

// Segment heads and ranks in two kernels instead of flags -> 3-launch scan -> write: k_seg_rank_block scans the head flags
// (computed from the sorted keys on the fly) inside each SCAN_TILE block and leaves the block totals; after the block
// totals are scanned (one launch), k_seg_write2 adds the block offset while it writes the segment arrays.
__global__ void __launch_bounds__(SCAN_THREADS)
k_seg_rank_block(const uint32_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ excl, uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)tid * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
    uint32_t prev = base > 0 && base - 1 < n ? keys[base - 1] : 0u;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const int64_t idx = base + i;
        uint32_t k = idx < n ? keys[idx] : 0u;
        v[i] = idx < n && (idx == 0 || k != prev) ? 1u : 0u;
        prev = k;
        sum += v[i];
    }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t t = lane < SCAN_THREADS / 32 ? warp_tot[lane] : 0u, ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t s2 = __shfl_up_sync(0xffffffffu, ti, o); if (lane >= o) ti += s2; }
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = ti - t;
        if (lane == SCAN_THREADS / 32 - 1) block_sums[blockIdx.x] = ti;
    }
    __syncthreads();
    uint32_t ex = warp_tot[wid] + inc - sum;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { const int64_t idx = base + i; if (idx < n) excl[idx] = ex; ex += v[i]; }
}

__global__ void k_seg_write2(const uint32_t* keys, const uint32_t* vals, const uint32_t* excl, const uint32_t* block_off, int64_t n,
                             uint32_t* seg_start, uint32_t* uniq, uint32_t* n_unique, uint32_t* seg_of_occ) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool head = (i == 0 || keys[i] != keys[i - 1]);
    const uint32_t ex = excl[i] + block_off[i / SCAN_TILE];
    const uint32_t sid = head ? ex : ex - 1u;
    if (head) { seg_start[sid] = (uint32_t)i; uniq[sid] = keys[i]; }
    if (seg_of_occ) seg_of_occ[vals[i]] = sid;
    if (i == n - 1) { uint32_t nu = sid + 1u; *n_unique = nu; seg_start[nu] = (uint32_t)n; }
}

// ---------------------------------------------------------------------------------------------
// Fused path (n > SMALL_SORT_MAX): ONE persistent launch does every radix pass and, optionally, the segment arrays.
// G <= #SMs CTAs of 1024 threads, all resident, each owning a contiguous chunk of the key list; the phases are separated by
// a grid barrier (counter + generation in engine-owned global memory, self-resetting, so the launch can be replayed):
//   per pass:  chunk histogram (warp-private counters, match_any aggregation) -> ghist[b][digit]   | barrier
//              digit base = sum of all smaller digits + the same digit in earlier chunks; stable ranks; scatter | barrier
//   segments:  head flags of the chunk counted -> gheads[b] | barrier | ranks = heads in earlier chunks + block scan; write.
// Replaces 5 launches per pass + 3 for the segments (18 launches / 0.19 ms for 2^20 keys below 2^20) by one.
// ---------------------------------------------------------------------------------------------
constexpr int FS_THREADS = 1024;
constexpr int FS_WARPS = FS_THREADS / 32;
constexpr int FS_ITEMS = 8;
constexpr int FS_TILE = FS_THREADS * FS_ITEMS;

struct FusedSortArgs {
    const uint32_t* keys_in; int64_t n; int passes; int64_t chunk;
    uint32_t *k0, *v0, *k1, *v1;      // the last pass lands in (k0, v0)
    uint32_t* ghist;                  // [G][256]
    uint32_t* bar;                    // [0] arrivals, [1] generation
    int want_seg;
    uint32_t *seg_start, *uniq, *n_unique, *seg_of_occ, *gheads;   // gheads [G]
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void grid_barrier(uint32_t* bar, uint32_t& gen, int G) {
    __syncthreads();
    if (threadIdx.x == 0) {
        if (G > 1) {
            __threadfence();
            const uint32_t old = atomicAdd(bar, 1u);
            if (old == (uint32_t)G - 1u) {
                atomicExch(bar, 0u);
                __threadfence();
                atomicAdd(bar + 1, 1u);
            } else {
                while (ld_acquire_u32(bar + 1) == gen) { }
            }
            __threadfence();
        }
        ++gen;
    }
    __syncthreads();
}

// head flags of one tile (warp w owns entries [t0 + 256 w, t0 + 256 w + 256), FS_ITEMS rows of 32): ballots per row, per-(warp,
// row) counts exclusive-scanned over the CTA in entry order.  Returns the tile's number of heads; bal[] / s_scan stay valid.
__device__ __forceinline__ uint32_t seg_tile_scan(const uint32_t* ck, int64_t t0, int64_t c0, int64_t c1, uint32_t first_prev,
                                                  uint32_t (&key)[FS_ITEMS], uint32_t (&bal)[FS_ITEMS], uint32_t* s_scan, uint32_t* s_wtot) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t base = t0 + (int64_t)w * FS_ITEMS * 32;
    uint32_t prev0[FS_ITEMS];
#pragma unroll
    for (int j = 0; j < FS_ITEMS; ++j) {
        const int64_t idx = base + j * 32 + lane;
        key[j] = idx < c1 ? __ldcg(ck + idx) : 0u;
        prev0[j] = 0u;
        if (lane == 0 && idx < c1 && idx > c0) prev0[j] = __ldcg(ck + idx - 1);
    }
#pragma unroll
    for (int j = 0; j < FS_ITEMS; ++j) {
        const int64_t idx = base + j * 32 + lane;
        uint32_t pk = __shfl_up_sync(0xffffffffu, key[j], 1);
        if (lane == 0) pk = idx == c0 ? first_prev : prev0[j];
        const bool head = idx < c1 && (idx == 0 || key[j] != pk);
        bal[j] = __ballot_sync(0xffffffffu, head);
    }
    if (lane < FS_ITEMS) {
        uint32_t c = 0;
#pragma unroll
        for (int j = 0; j < FS_ITEMS; ++j) if (lane == j) c = __popc(bal[j]);
        s_scan[w * FS_ITEMS + lane] = c;
    }
    __syncthreads();
    uint32_t v = 0, inc = 0;
    if (tid < FS_WARPS * FS_ITEMS) {               // 256 counts, one per thread of warps 0..7
        v = s_scan[tid]; inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_wtot[w] = inc;
    }
    __syncthreads();
    uint32_t total = 0;
#pragma unroll
    for (int ww = 0; ww < FS_WARPS * FS_ITEMS / 32; ++ww) total += s_wtot[ww];
    if (tid < FS_WARPS * FS_ITEMS) {
        uint32_t wb = 0;
        for (int ww = 0; ww < w; ++ww) wb += s_wtot[ww];
        s_scan[tid] = wb + inc - v;
    }
    __syncthreads();
    return total;
}

template <bool SINGLE>
__global__ void __launch_bounds__(FS_THREADS, 1)
k_sort_seg_fused(FusedSortArgs a) {
    __shared__ uint32_t wcount[FS_WARPS][256];
    __shared__ uint32_t dbase[256];
    __shared__ uint32_t part[2][4][256];
    __shared__ uint32_t wtot[FS_WARPS];
    __shared__ uint32_t s_scan[FS_WARPS * FS_ITEMS];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    const int64_t c0 = (int64_t)b * a.chunk, c1 = min(a.n, c0 + a.chunk);
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t gen = 0;
    if (tid == 0) gen = ld_acquire_u32(a.bar + 1);       // before this CTA's first arrival: nobody can have advanced it yet

    const uint32_t* ck = a.keys_in; const uint32_t* cv = nullptr;
    for (int p = 0; p < a.passes; ++p) {
        const bool to0 = ((a.passes - 1 - p) % 2) == 0;
        uint32_t* ok = to0 ? a.k0 : a.k1; uint32_t* ov = to0 ? a.v0 : a.v1;
        const int shift = p * 8;
        uint32_t key[FS_ITEMS], val[FS_ITEMS], rank[FS_ITEMS];
        for (int i = tid; i < FS_WARPS * 256; i += FS_THREADS) (&wcount[0][0])[i] = 0;
        __syncthreads();
        // stable ranks inside the tile: index order inside a 32-key row (match_any + popc of the lower lanes), row after row
        // inside a warp (per-warp digit counters); the counters double as the tile's histogram
        auto rank_tile = [&](int64_t t0) {
            const int64_t base = t0 + (int64_t)w * FS_ITEMS * 32;
#pragma unroll
            for (int j = 0; j < FS_ITEMS; ++j) {
                const int64_t idx = base + j * 32 + lane;
                key[j] = idx < c1 ? __ldcg(ck + idx) : 0u;
                val[j] = idx < c1 ? (cv ? __ldcg(cv + idx) : (uint32_t)idx) : 0u;
            }
#pragma unroll
            for (int j = 0; j < FS_ITEMS; ++j) {
                const int64_t idx = base + j * 32 + lane;
                const bool okk = idx < c1;
                const uint32_t dig = okk ? ((key[j] >> shift) & 255u) : 0xffffffffu;
                const uint32_t peers = __match_any_sync(0xffffffffu, dig);
                uint32_t before = 0;
                if (okk) before = wcount[w][dig];
                __syncwarp();
                if (okk && (peers & lt_mask) == 0) wcount[w][dig] = before + __popc(peers);
                __syncwarp();
                rank[j] = before + __popc(peers & lt_mask);
            }
        };
        // per-warp counts -> per-warp bases (running digit base included), then the scatter out of the registers
        auto scatter_tile = [&](int64_t t0) {
            if (tid < 256) {
                uint32_t run = dbase[tid];
#pragma unroll 8
                for (int ww = 0; ww < FS_WARPS; ++ww) { const uint32_t c = wcount[ww][tid]; wcount[ww][tid] = run; run += c; }
                dbase[tid] = run;
            }
            __syncthreads();
            const int64_t base = t0 + (int64_t)w * FS_ITEMS * 32;
#pragma unroll
            for (int j = 0; j < FS_ITEMS; ++j) {
                const int64_t idx = base + j * 32 + lane;
                if (idx < c1) {
                    const uint32_t pos = wcount[w][(key[j] >> shift) & 255u] + rank[j];
                    ok[pos] = key[j]; ov[pos] = val[j];
                }
            }
        };
        // ---- chunk histogram ----
        if (SINGLE) {
            rank_tile(c0);
        } else {
            for (int64_t r0 = c0 + (int64_t)w * 32; r0 < c1; r0 += 4 * FS_THREADS) {
                uint32_t dg[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t idx = r0 + (int64_t)u * FS_THREADS + lane;
                    dg[u] = idx < c1 ? ((__ldcg(ck + idx) >> shift) & 255u) : 0xffffffffu;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t peers = __match_any_sync(0xffffffffu, dg[u]);
                    if (dg[u] != 0xffffffffu && (peers & lt_mask) == 0) wcount[w][dg[u]] += __popc(peers);
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t h = 0;
#pragma unroll
            for (int ww = 0; ww < FS_WARPS; ++ww) h += wcount[ww][tid];
            a.ghist[(size_t)b * 256 + tid] = h;
        }
        grid_barrier(a.bar, gen, G);
        // ---- digit bases: all smaller digits of every chunk + this digit in the earlier chunks ----
        {
            const int d = tid & 255, q = tid >> 8;
            uint32_t pre = 0, tot = 0;
            for (int bb0 = q; bb0 < G; bb0 += 32) {
                uint32_t x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int bb = bb0 + 4 * u; x[u] = bb < G ? __ldcg(a.ghist + (size_t)bb * 256 + d) : 0u; }
#pragma unroll
                for (int u = 0; u < 8; ++u) { tot += x[u]; if (bb0 + 4 * u < b) pre += x[u]; }
            }
            part[0][q][d] = pre; part[1][q][d] = tot;
        }
        __syncthreads();
        uint32_t pre = 0, tot = 0;
        if (tid < 256) {
            pre = part[0][0][tid] + part[0][1][tid] + part[0][2][tid] + part[0][3][tid];
            tot = part[1][0][tid] + part[1][1][tid] + part[1][2][tid] + part[1][3][tid];
            uint32_t inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (lane == 31) wtot[w] = inc;
            tot = inc - tot;                              // exclusive inside the warp
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t wb = 0;
            for (int ww = 0; ww < w; ++ww) wb += wtot[ww];
            dbase[tid] = wb + tot + pre;
        }
        __syncthreads();
        // ---- scatter ----
        if (SINGLE) {
            scatter_tile(c0);
        } else {
            for (int64_t t0 = c0; t0 < c1; t0 += FS_TILE) {
                for (int i = tid; i < FS_WARPS * 256; i += FS_THREADS) (&wcount[0][0])[i] = 0;
                __syncthreads();
                rank_tile(t0);
                __syncthreads();
                scatter_tile(t0);
                __syncthreads();
            }
        }
        grid_barrier(a.bar, gen, G);
        ck = ok; cv = ov;
    }
    if (!a.want_seg) return;
    // ---- segments over the sorted list (ck, cv) ----
    const uint32_t first_prev = c0 > 0 && c0 < a.n ? __ldcg(ck + c0 - 1) : 0u;
    uint32_t key[FS_ITEMS], bal[FS_ITEMS];
    {
        uint32_t heads = 0;
        for (int64_t t0 = c0; t0 < c1; t0 += FS_TILE) heads += seg_tile_scan(ck, t0, c0, c1, first_prev, key, bal, s_scan, wtot);
        if (tid == 0) a.gheads[b] = heads;
    }
    grid_barrier(a.bar, gen, G);
    if (tid < 32) {
        uint32_t t = 0;
        for (int bb0 = lane; bb0 < b; bb0 += 128) {
            uint32_t x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = bb0 + 32 * u < b ? __ldcg(a.gheads + bb0 + 32 * u) : 0u;
            t += x[0] + x[1] + x[2] + x[3];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) s_carry = t;
    }
    __syncthreads();
    uint32_t carry = s_carry;
    for (int64_t t0 = c0; t0 < c1; t0 += FS_TILE) {
        uint32_t tile_heads = 0;
        if (!SINGLE) tile_heads = seg_tile_scan(ck, t0, c0, c1, first_prev, key, bal, s_scan, wtot);   // SINGLE: still in place
        const int64_t base = t0 + (int64_t)w * FS_ITEMS * 32;
        uint32_t vv[FS_ITEMS];
        if (a.seg_of_occ) {
#pragma unroll
            for (int j = 0; j < FS_ITEMS; ++j) { const int64_t idx = base + j * 32 + lane; vv[j] = idx < c1 ? __ldcg(cv + idx) : 0u; }
        }
#pragma unroll
        for (int j = 0; j < FS_ITEMS; ++j) {
            const int64_t idx = base + j * 32 + lane;
            if (idx < c1) {
                const bool head = (bal[j] >> lane) & 1u;
                const uint32_t ex = carry + s_scan[w * FS_ITEMS + j] + __popc(bal[j] & lt_mask);
                const uint32_t sid = head ? ex : ex - 1u;
                if (head) { a.seg_start[sid] = (uint32_t)idx; a.uniq[sid] = key[j]; }
                if (a.seg_of_occ) a.seg_of_occ[vv[j]] = sid;
                if (idx == a.n - 1) { *a.n_unique = sid + 1u; a.seg_start[sid + 1u] = (uint32_t)a.n; }
            }
        }
        carry += tile_heads;
        __syncthreads();
    }
}

// one launch: sorted (keys, vals) into (k0, v0); segment arrays when `seg` is given
static int fused_sort_launch(poi_engine* e, const uint32_t* keys_in, int64_t n, uint32_t bound, uint32_t* k0, uint32_t* v0, SegList* seg) {
    int bits = 1;
    while (bits < 32 && (1ull << bits) < (unsigned long long)bound) ++bits;
    FusedSortArgs a; memset(&a, 0, sizeof(a));
    a.keys_in = keys_in; a.n = n; a.passes = (bits + 7) / 8; a.k0 = k0; a.v0 = v0;
    a.chunk = (int64_t)FS_TILE * poi_cdiv(n, (int64_t)FS_TILE * e->num_sms);
    const int G = (int)poi_cdiv(n, a.chunk);
    if (a.passes > 1) { POI_TRY(arena_get(e, (size_t)n, &a.k1)); POI_TRY(arena_get(e, (size_t)n, &a.v1)); }
    POI_TRY(arena_get(e, (size_t)G * 256, &a.ghist));
    POI_TRY(arena_get(e, (size_t)G, &a.gheads));
    a.bar = e->grid_bar;
    if (seg) { a.want_seg = 1; a.seg_start = seg->seg_start; a.uniq = seg->uniq; a.n_unique = seg->n_unique; a.seg_of_occ = seg->seg_of_occ; }
    if (a.chunk == FS_TILE) POI_LAUNCH(e, k_sort_seg_fused<true>, G, FS_THREADS, 0, a);
    else POI_LAUNCH(e, k_sort_seg_fused<false>, G, FS_THREADS, 0, a);
    return 0;
}

static int build_segments(poi_engine* e, const uint32_t* keys_dev, int64_t n, uint32_t bound,
                          bool want_inverse, SegList* out) {
    out->n = n;
    POI_CAT(e, CAT_INDEX, 0, 0);
    size_t nn = (size_t)std::max<int64_t>(n, 1);
    const bool fused = e->fused_sort && n > SMALL_SORT_MAX;
    const bool small_fused = e->fused_sort && n > 0 && n <= SMALL_SORT_MAX;
    if (fused || small_fused) {
        POI_TRY(arena_get(e, nn, &out->keys));
        POI_TRY(arena_get(e, nn, &out->vals));
    } else {
        POI_TRY(sort_pairs(e, keys_dev, n, bound, &out->keys, &out->vals));
    }
    uint32_t* excl = nullptr;
    if (!fused && !small_fused) POI_TRY(arena_get(e, nn, &excl));
    POI_TRY(arena_get(e, nn + 1, &out->seg_start));
    POI_TRY(arena_get(e, nn, &out->uniq));
    POI_TRY(arena_get(e, 4, &out->n_unique));
    out->seg_of_occ = nullptr;
    if (want_inverse) POI_TRY(arena_get(e, nn, &out->seg_of_occ));
    if (n <= 0) { POI_CK(e, cudaMemsetAsync(out->n_unique, 0, 4, e->stream)); return 0; }
    if (fused) return fused_sort_launch(e, keys_dev, n, bound, out->keys, out->vals, out);
    if (small_fused) {
        int np2 = 32;
        while (np2 < n) np2 <<= 1;
        const int threads = std::min(1024, std::max(32, np2 / 2));
        POI_LAUNCH(e, k_small_sort_seg, 1, threads, (size_t)np2 * 8, keys_dev, (int)n, np2, out->keys, out->vals, out->seg_start,
                   out->uniq, out->n_unique, out->seg_of_occ);
        return 0;
    }
    unsigned g = (unsigned)poi_cdiv(n, 256);
    const int64_t nb = poi_cdiv(n, SCAN_TILE);
    uint32_t* bs = nullptr;
    POI_TRY(arena_get(e, (size_t)nb + 1, &bs));
    POI_LAUNCH(e, k_seg_rank_block, (unsigned)nb, SCAN_THREADS, 0, out->keys, n, excl, bs);
    POI_TRY(exclusive_scan_u32(e, bs, bs, nb, nullptr));
    POI_LAUNCH(e, k_seg_write2, g, 256, 0, out->keys, out->vals, excl, bs, n, out->seg_start, out->uniq,
               out->n_unique, out->seg_of_occ);
    return 0;
}
