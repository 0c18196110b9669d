// rows.cuh -- the HBM-bound row kernels: embedding gather, sparse row SGD (segment gather-reduce),
// streaming sum of squares.  All accesses are 128-bit and coalesced along the row.
#pragma once
#include "common.cuh"
#include "sort.cuh"

// ---------------------------------------------------------------------------------------------
// gather: out[i,:] = table[idx[i],:]           (AdvancedSubtensor1; GRU.py:327 etc.)
// LPR lanes cooperate on one row (LPR*16 B per access); each lane group keeps UNR rows in flight.
// ---------------------------------------------------------------------------------------------
template <int LPR, int UNR>
__global__ void __launch_bounds__(256)
k_gather_rows(const float* __restrict__ table, int dim4, const int32_t* __restrict__ idx,
              int64_t n_idx, float* __restrict__ out) {
    const int lane = threadIdx.x % LPR;
    const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    const int64_t n_groups = (int64_t)gridDim.x * blockDim.x / LPR;
    const float4* tab4 = reinterpret_cast<const float4*>(table);
    float4* out4 = reinterpret_cast<float4*>(out);
    for (int64_t r0 = group * UNR; r0 < n_idx; r0 += n_groups * UNR) {
        int64_t src[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) src[u] = (r0 + u < n_idx) ? (int64_t)idx[r0 + u] : -1;
        for (int c = lane; c < dim4; c += LPR) {
            float4 v[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u)
                if (src[u] >= 0) v[u] = __ldg(tab4 + src[u] * dim4 + c);
#pragma unroll
            for (int u = 0; u < UNR; ++u)
                if (src[u] >= 0) __stcs(out4 + (r0 + u) * dim4 + c, v[u]);
        }
    }
}

static int launch_gather_rows(poi_engine* e, const float* table, int dim, const int32_t* idx,
                              int64_t n_idx, float* out) {
    if (n_idx <= 0) return 0;
    const int dim4 = dim / 4;
    const int UNR = 4;
    POI_CAT(e, CAT_GATHER, 0, 2.0 * (double)n_idx * dim * 4 + 4.0 * (double)n_idx);
    int lpr = dim4 <= 8 ? 8 : (dim4 <= 16 ? 16 : 32);
    int64_t groups_needed = poi_cdiv(n_idx, UNR);
    int64_t threads_needed = groups_needed * lpr;
    unsigned grid = (unsigned)std::min<int64_t>(poi_cdiv(threads_needed, 256), (int64_t)e->num_sms * 16);
    grid = std::max(grid, 1u);
    if (lpr == 8)       POI_LAUNCH(e, (k_gather_rows<8, 4>), grid, 256, 0, table, dim4, idx, n_idx, out);
    else if (lpr == 16) POI_LAUNCH(e, (k_gather_rows<16, 4>), grid, 256, 0, table, dim4, idx, n_idx, out);
    else                POI_LAUNCH(e, (k_gather_rows<32, 4>), grid, 256, 0, table, dim4, idx, n_idx, out);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// sum of squares, fp64 accumulation, fixed-order two-stage reduction   (model.l2.eval())
// ---------------------------------------------------------------------------------------------
constexpr int SUMSQ_BLOCKS_PER_SM = 8;

__global__ void __launch_bounds__(256)
k_sumsq_partial(const float* __restrict__ x, int64_t n, double* __restrict__ part) {
    __shared__ double sh[8];
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n / 4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    double acc = 0.0;
    for (int64_t i = tid; i < n4; i += nth) {
        float4 v = __ldcs(x4 + i);
        float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        acc += (double)s;
    }
    if (tid == 0) for (int64_t i = n4 * 4; i < n; ++i) acc += (double)x[i] * (double)x[i];
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        part[blockIdx.x] = t;
    }
}

__global__ void k_sum_partials_d(const double* __restrict__ part, int n, double* out, int out_stride, int ncols) {
    // one thread per column, sequential over partial rows: fixed order, deterministic
    int c = threadIdx.x;
    if (c < ncols) {
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += part[(size_t)i * ncols + c];
        out[c * out_stride] = t;
    }
}

static int launch_sumsq(poi_engine* e, const float* x, int64_t n, double* out_dev) {
    int blocks = (int)std::min<int64_t>(std::max<int64_t>(poi_cdiv(n / 4 + 1, 256), 1), (int64_t)e->num_sms * SUMSQ_BLOCKS_PER_SM);
    double* part = nullptr;
    POI_TRY(arena_get(e, (size_t)blocks, &part));
    POI_CAT(e, CAT_REDUCE, 0, (double)n * 4);
    POI_LAUNCH(e, k_sumsq_partial, blocks, 256, 0, x, n, part);
    POI_LAUNCH(e, k_sum_partials_d, 1, 32, 0, part, blocks, out_dev, 1, 1);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// sparse row SGD by segment gather-reduce
//
//   table[r] <- table[r] - alpha * ( sum_{occ in seg(r)} grad(occ) + lambda * count(r) * table[r] )
//
// = `T.set_subtensor(uiq_x, uiq_x - lr * T.grad(cost, self.lt)[uiq_pqs])` (GRU.py:372-373,
// GRU_Spatial.py:212-215): the dense Theano gradient sums every duplicate occurrence, and the L2
// term of the cost counts every gathered row (padding included, GRU.py:365).  Here the dense
// (n_item+1) x d gradient is never materialised: each unique row walks its own occurrence list
// (ascending occurrence id = fixed summation order, no atomics, bit-reproducible).
//
// The per-occurrence gradient is described by up to two (row pointer, scale) pairs.
// ---------------------------------------------------------------------------------------------
enum RowSrcMode { SRC_DENSE_GRADS = 0, SRC_GRU_LT = 1, SRC_GRU_DI = 2, SRC_NONE = 3 };

struct RowSrc {
    int mode;
    // SRC_DENSE_GRADS: grad rows [n x dim]
    const float* grads;
    // SRC_GRU_*: DX [T*B x ldx] (cols [0,d) -> lt, [d,2d) -> di), Hc [T*B x d], e [T*B]
    const float* DX; int ldx; const float* Hc; const float* ev;
    int B; int T; int64_t LB;   // LB = lmax * B  (p occurrences [0,LB), q occurrences [LB,2LB))
    int dim;
    // optional per-occurrence L2 multiplicities (multi-GPU owner side: each received gradient row stands
    // for `weights[occ]` gathered occurrences on the sending rank); NULL = 1 each
    const float* weights;
    // emit mode (multi-GPU sender side): instead of updating the table, write the duplicate-summed
    // gradient of segment sg to emit_rows[(emit_by_key ? key : sg) * dim] and its count to emit_cnt[.]
    float* emit_rows; float* emit_cnt; int emit_by_key;
    // segments of ONE occurrence are left alone (their row was already updated in place by the kernel that produced the
    // gradients: prme_k.cuh / geoie_k.cuh); only rows that occur several times in the batch are summed here
    int skip_single;
};

__device__ __forceinline__ void row_finish(const RowSrc& src, float* table, int dim4, uint32_t key, uint32_t sg,
                                           int c, float4 a, float cntf, float alpha, float lambda) {
    if (src.emit_rows) {
        size_t slot = src.emit_by_key ? key : sg;
        st4(src.emit_rows + (slot * dim4 + c) * 4, a);
        if (c == 0) src.emit_cnt[slot] = cntf;
    } else {
        float* row = table + ((size_t)key * dim4 + c) * 4;
        const float lc = lambda * cntf;
        float4 r = ld4(row);
        r.x -= alpha * (a.x + lc * r.x); r.y -= alpha * (a.y + lc * r.y);
        r.z -= alpha * (a.z + lc * r.z); r.w -= alpha * (a.w + lc * r.w);
        st4(row, r);
    }
}

struct OccPair { const float* p1; const float* p2; float s2; };

__device__ __forceinline__ OccPair occ_begin(const RowSrc& s, uint32_t occ) {
    OccPair o; o.p1 = nullptr; o.p2 = nullptr; o.s2 = 0.f;
    if (s.mode == SRC_DENSE_GRADS) {
        if (s.grads) o.p1 = s.grads + (size_t)occ * s.dim;
    } else if (s.mode == SRC_GRU_LT) {
        bool isq = occ >= (uint64_t)s.LB;
        int64_t m = isq ? (int64_t)occ - s.LB : (int64_t)occ;     // m = t*B + b
        int64_t TB = (int64_t)s.T * s.B;
        if (!isq && m < TB) o.p1 = s.DX + (size_t)m * s.ldx;      // d cost / d x_t, t < T
        if (m >= s.B && m < TB + s.B) {                           // 1 <= t <= T: pair j = t-1
            int64_t mm = m - s.B;
            float ev = s.ev[mm];
            if (ev != 0.f) { o.p2 = s.Hc + (size_t)mm * s.dim; o.s2 = isq ? -ev : ev; }
        }
    } else if (s.mode == SRC_GRU_DI) {
        int64_t TB = (int64_t)s.T * s.B;
        if ((int64_t)occ < TB) o.p1 = s.DX + (size_t)occ * s.ldx + s.dim;
    }
    return o;
}

constexpr int ROW_LONG_THRESH = 96;

// NCH float4 chunks per lane: dim <= 128*NCH.  A warp takes up to 32 consecutive segments at a time: lane l reads the bounds, key and
// first occurrence of segment l (coalesced, one round trip for all 32), the warp then walks the segments with the metadata
// broadcast by shuffles, and the table row is requested BEFORE the gradient rows are summed -- per segment the warp waits for
// one memory round trip (gradient rows and table row in flight together) instead of four dependent ones.
template <int NCH>
__global__ void __launch_bounds__(256, NCH == 1 ? 6 : (NCH == 2 ? 5 : (NCH == 4 ? 3 : 2)))
k_rows_update_warp(SegList seg, float* __restrict__ table, int dim4, float alpha, float lambda,
                   RowSrc src, int long_thresh, uint32_t* long_list, uint32_t* long_count) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nu = *seg.n_unique;
    const bool emit = src.emit_rows != nullptr;
    int gs = 32;                                   // segments per warp and round: as many as leave every warp >= 4 rounds (balance)
    while (gs > 1 && nu < nwarps * gs * 4) gs >>= 1;
    for (int64_t sg0 = warp * gs; sg0 < nu; sg0 += nwarps * gs) {
        const int64_t my = sg0 + lane;
        uint32_t s0l = 0, s1l = 0, keyl = 0, v0l = 0;
        if (lane < gs && my < nu) { s0l = seg.seg_start[my]; s1l = seg.seg_start[my + 1]; keyl = seg.uniq[my]; v0l = seg.vals[s0l]; }
        const int nseg = (int)min((int64_t)gs, nu - sg0);
        for (int j = 0; j < nseg; ++j) {
            const uint32_t s0 = __shfl_sync(0xffffffffu, s0l, j), s1 = __shfl_sync(0xffffffffu, s1l, j);
            const uint32_t key = __shfl_sync(0xffffffffu, keyl, j), v0 = __shfl_sync(0xffffffffu, v0l, j);
            const uint32_t cnt = s1 - s0;
            const int64_t sg = sg0 + j;
            if (src.skip_single && cnt == 1u) continue;
            if ((int)cnt > long_thresh) {
                if (lane == 0) { uint32_t pos = atomicAdd(long_count, 1u); long_list[pos] = (uint32_t)sg; }
                continue;
            }
            float4 trow[NCH];
            if (!emit) {
#pragma unroll
                for (int k = 0; k < NCH; ++k) { const int c = lane + 32 * k; if (c < dim4) trow[k] = ld4(table + ((size_t)key * dim4 + c) * 4); }
            }
            float4 acc[NCH];
#pragma unroll
            for (int k = 0; k < NCH; ++k) acc[k] = f4zero();
            float cntf = (float)cnt;
            if (src.weights) { cntf = 0.f; for (uint32_t i = s0; i < s1; ++i) cntf += src.weights[seg.vals[i]]; }
            for (uint32_t i = s0; i < s1; ++i) {
                OccPair o = occ_begin(src, i == s0 ? v0 : seg.vals[i]);
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    int c = lane + 32 * k;
                    if (c < dim4) {
                        if (o.p1) acc[k] = f4add(acc[k], ld4(o.p1 + 4 * c));
                        if (o.p2) acc[k] = f4fma(o.s2, ld4(o.p2 + 4 * c), acc[k]);
                    }
                }
            }
            if (emit) {
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    int c = lane + 32 * k;
                    if (c < dim4) row_finish(src, table, dim4, key, (uint32_t)sg, c, acc[k], cntf, alpha, lambda);
                }
            } else {
                const float lc = lambda * cntf;
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    const int c = lane + 32 * k;
                    if (c < dim4) {
                        float4 r = trow[k];
                        r.x -= alpha * (acc[k].x + lc * r.x); r.y -= alpha * (acc[k].y + lc * r.y);
                        r.z -= alpha * (acc[k].z + lc * r.z); r.w -= alpha * (acc[k].w + lc * r.w);
                        st4(table + ((size_t)key * dim4 + c) * 4, r);
                    }
                }
            }
        }
    }
}

// long segments (hot rows: the pad row, popular POIs, every distance interval): split into chunks of
// ROW_CHUNK occurrences; one CTA reduces one chunk to a partial row (warp w takes occurrences
// w, w+8, ... of the chunk, the 8 warp rows are then added in warp order), a second kernel adds the
// chunk partials of a segment in chunk order and applies the SGD step.  Fixed order -> deterministic.
constexpr int ROW_CHUNK = 256;

__global__ void k_long_offsets(SegList seg, const uint32_t* __restrict__ long_list,
                               const uint32_t* __restrict__ long_count, uint32_t* __restrict__ chunk_off) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t nl = *long_count;
    uint32_t run = 0;
    for (uint32_t li = 0; li < nl; ++li) {
        chunk_off[li] = run;
        uint32_t sg = long_list[li];
        run += (seg.seg_start[sg + 1] - seg.seg_start[sg] + ROW_CHUNK - 1) / ROW_CHUNK;
    }
    chunk_off[nl] = run;
}

// Small tables (the 201 distance-interval rows of Distance2Pre: every row is a hot row): all segments go through the chunked
// path, so the list of "long" segments is simply 0 .. n_unique-1 and the chunk offsets are one block scan (n_unique <= 1024)
// -- no pass of the per-segment kernel (measured 50 us for 201 rows) and no serial offset loop (15 us).
__global__ void __launch_bounds__(1024)
k_all_long_offsets(SegList seg, uint32_t* __restrict__ long_list, uint32_t* __restrict__ long_count, uint32_t* __restrict__ chunk_off) {
    __shared__ uint32_t wtot[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t nu = *seg.n_unique;
    uint32_t c = 0;
    if ((uint32_t)tid < nu) c = (seg.seg_start[tid + 1] - seg.seg_start[tid] + ROW_CHUNK - 1) / ROW_CHUNK;
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wtot[wid] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < wid; ++w) base += wtot[w];
    if ((uint32_t)tid < nu) { long_list[tid] = (uint32_t)tid; chunk_off[tid] = base + inc - c; }
    if ((uint32_t)tid + 1 == nu) chunk_off[nu] = base + inc;
    if (tid == 0) { *long_count = nu; if (nu == 0) chunk_off[0] = 0; }
}

__global__ void __launch_bounds__(256)
k_long_partial(SegList seg, int dim4, RowSrc src, const uint32_t* __restrict__ long_list,
               const uint32_t* __restrict__ long_count, const uint32_t* __restrict__ chunk_off,
               float4* __restrict__ partial, float* __restrict__ partial_w) {
    extern __shared__ float4 s_part[];          // [8][dim4]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t nl = *long_count;
    const uint32_t total = chunk_off[nl];
    for (uint32_t ch = blockIdx.x; ch < total; ch += gridDim.x) {
        // locate the long entry that owns global chunk ch (binary search over chunk_off)
        uint32_t lo = 0, hi = nl;
        while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (chunk_off[mid] <= ch) lo = mid; else hi = mid; }
        const uint32_t sg = long_list[lo];
        const uint32_t s0 = seg.seg_start[sg] + (ch - chunk_off[lo]) * ROW_CHUNK;
        const uint32_t s1 = min(s0 + ROW_CHUNK, seg.seg_start[sg + 1]);
        if (w == 0) {                               // L2 multiplicity of this chunk
            float wsum = 0.f;
            for (uint32_t i = s0 + lane; i < s1; i += 32) wsum += src.weights ? src.weights[seg.vals[i]] : 1.f;
            wsum = warp_sum(wsum);
            if (lane == 0) partial_w[ch] = wsum;
        }
        for (int c0 = 0; c0 < dim4; c0 += 32) {
            int c = c0 + lane;
            float4 acc = f4zero();
            for (uint32_t i = s0 + w; i < s1; i += 8) {
                OccPair o = occ_begin(src, seg.vals[i]);
                if (c < dim4) {
                    if (o.p1) acc = f4add(acc, ld4(o.p1 + 4 * c));
                    if (o.p2) acc = f4fma(o.s2, ld4(o.p2 + 4 * c), acc);
                }
            }
            if (c < dim4) s_part[w * dim4 + c] = acc;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < dim4; c += blockDim.x) {
            float4 a = s_part[c];
#pragma unroll
            for (int ww = 1; ww < 8; ++ww) a = f4add(a, s_part[ww * dim4 + c]);
            partial[(size_t)ch * dim4 + c] = a;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128)
k_long_final(SegList seg, float* __restrict__ table, int dim4, float alpha, float lambda, RowSrc src,
             const uint32_t* __restrict__ long_list, const uint32_t* __restrict__ long_count,
             const uint32_t* __restrict__ chunk_off, const float4* __restrict__ partial,
             const float* __restrict__ partial_w) {
    const uint32_t nl = *long_count;
    for (uint32_t li = blockIdx.x; li < nl; li += gridDim.x) {
        const uint32_t sg = long_list[li];
        const uint32_t c0 = chunk_off[li], c1 = chunk_off[li + 1];
        float cntf = 0.f;
        for (uint32_t ch = c0; ch < c1; ++ch) cntf += partial_w[ch];
        for (int c = threadIdx.x; c < dim4; c += blockDim.x) {
            float4 a = f4zero();
            for (uint32_t ch = c0; ch < c1; ++ch) a = f4add(a, partial[(size_t)ch * dim4 + c]);
            row_finish(src, table, dim4, seg.uniq[sg], sg, c, a, cntf, alpha, lambda);
        }
    }
}

static int launch_rows_update(poi_engine* e, const SegList& seg, float* table, int dim,
                              float alpha, float lambda, const RowSrc& src, int long_thresh,
                              double algo_bytes = 0.0, int64_t table_rows = 0) {
    if (seg.n <= 0) return 0;
    POI_CAT(e, CAT_ROWS, 0, algo_bytes);
    const int dim4 = dim / 4;
    uint32_t *long_list = nullptr, *long_count = nullptr;
    POI_TRY(arena_get(e, (size_t)seg.n, &long_list));
    POI_TRY(arena_get(e, 4, &long_count));
    if (table_rows > 0 && table_rows <= 1024 && seg.n > 8 * table_rows && !src.skip_single) {
        // small table, many occurrences per row: the chunked path for every segment (k_all_long_offsets)
        const size_t max_chunks = (size_t)seg.n / ROW_CHUNK + (size_t)table_rows + 2;
        uint32_t* chunk_off = nullptr; float4* partial = nullptr; float* partial_w = nullptr;
        POI_TRY(arena_get(e, (size_t)table_rows + 2, &chunk_off));
        POI_TRY(arena_get(e, max_chunks * dim4, &partial));
        POI_TRY(arena_get(e, max_chunks, &partial_w));
        POI_LAUNCH(e, k_all_long_offsets, 1, 1024, 0, seg, long_list, long_count, chunk_off);
        const size_t smem = (size_t)8 * dim4 * sizeof(float4);
        unsigned lgrid = (unsigned)std::min<int64_t>((int64_t)max_chunks, (int64_t)e->num_sms * 8);
        POI_LAUNCH(e, k_long_partial, lgrid, 256, smem, seg, dim4, src, long_list, long_count, chunk_off, partial, partial_w);
        unsigned fgrid = (unsigned)std::min<int64_t>(table_rows, (int64_t)e->num_sms * 8);
        POI_LAUNCH(e, k_long_final, fgrid, 128, 0, seg, table, dim4, alpha, lambda, src, long_list, long_count, chunk_off, partial, partial_w);
        return 0;
    }
    POI_CK(e, cudaMemsetAsync(long_count, 0, 4, e->stream));
    // one wave of resident CTAs (a grid-stride loop over more CTAs than fit would run its tail at a fraction of the occupancy)
    static int occ[4] = {0, 0, 0, 0};
    const int oi = dim4 <= 32 ? 0 : (dim4 <= 64 ? 1 : (dim4 <= 128 ? 2 : 3));
    if (!occ[oi]) {
        int o = 0;
        if (oi == 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_rows_update_warp<1>, 256, 0);
        else if (oi == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_rows_update_warp<2>, 256, 0);
        else if (oi == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_rows_update_warp<4>, 256, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_rows_update_warp<8>, 256, 0);
        occ[oi] = std::max(o, 1);
    }
    int64_t warps = std::min<int64_t>(seg.n, (int64_t)e->num_sms * occ[oi] * 8);
    unsigned grid = (unsigned)std::max<int64_t>(poi_cdiv(warps * 32, 256), 1);
    if (dim4 <= 32)       POI_LAUNCH(e, (k_rows_update_warp<1>), grid, 256, 0, seg, table, dim4, alpha, lambda, src, long_thresh, long_list, long_count);
    else if (dim4 <= 64)  POI_LAUNCH(e, (k_rows_update_warp<2>), grid, 256, 0, seg, table, dim4, alpha, lambda, src, long_thresh, long_list, long_count);
    else if (dim4 <= 128) POI_LAUNCH(e, (k_rows_update_warp<4>), grid, 256, 0, seg, table, dim4, alpha, lambda, src, long_thresh, long_list, long_count);
    else if (dim4 <= 256) POI_LAUNCH(e, (k_rows_update_warp<8>), grid, 256, 0, seg, table, dim4, alpha, lambda, src, long_thresh, long_list, long_count);
    else POI_FAIL(e, "row dim %d too large (max 1024)", dim);
    // long segments: chunk partials, then ordered final sum (none can exist when the whole list is shorter than the
    // threshold -- the one-by-one calls of the reference's semantics -- so the three launches are skipped)
    if (seg.n <= long_thresh) return 0;
    const size_t max_long = (size_t)seg.n / (size_t)std::max(long_thresh, 1) + 2;
    const size_t max_chunks = (size_t)seg.n / ROW_CHUNK + max_long + 2;
    uint32_t* chunk_off = nullptr; float4* partial = nullptr; float* partial_w = nullptr;
    POI_TRY(arena_get(e, max_long + 1, &chunk_off));
    POI_TRY(arena_get(e, max_chunks * dim4, &partial));
    POI_TRY(arena_get(e, max_chunks, &partial_w));
    POI_LAUNCH(e, k_long_offsets, 1, 32, 0, seg, long_list, long_count, chunk_off);
    size_t smem = (size_t)8 * dim4 * sizeof(float4);
    unsigned lgrid = (unsigned)std::min<int64_t>((int64_t)max_chunks, (int64_t)e->num_sms * 8);
    POI_LAUNCH(e, k_long_partial, lgrid, 256, smem, seg, dim4, src, long_list, long_count, chunk_off, partial, partial_w);
    unsigned fgrid = (unsigned)std::min<int64_t>((int64_t)max_long, (int64_t)e->num_sms * 8);
    POI_LAUNCH(e, k_long_final, fgrid, 128, 0, seg, table, dim4, alpha, lambda, src, long_list, long_count, chunk_off, partial, partial_w);
    return 0;
}
