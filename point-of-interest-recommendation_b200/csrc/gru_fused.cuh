// gru_fused.cuh -- the GRU forward recurrence as ONE persistent tcgen05 kernel.
//
// The reference's `theano.scan` runs the cell once per time step (GRU.py:345-360,
// GRU_Spatial.py:170-197); the per-step path of this engine does the same with two GEMM launches
// per step.  Users are independent inside a step and coupled across steps only through their own
// h, so a CTA can own 128 users for the WHOLE sequence: h_t never leaves the SM between steps.
//
//   warps 0-3  epilogue, thread = one user row (= TMEM lane).  Accumulator -> gate math -> next A
//              operand (r*h, then h_t) written straight into shared memory in the UMMA layout
//              (128B-swizzled K-major, hi/lo split for 3xTF32).  z and (1-z)*h_prev are stashed in the
//              TMEM columns they came from, so epilogue 2 reloads nothing.  Z/R/C/H go to global
//              memory (saved for BPTT) through a per-warp staging tile so that stores are coalesced.
//   warps 4-5  Wh producers: stream the Wh tiles (L2-resident, identical every step) through a ring
//   warps 6-7  AX producers: cp.async the hoisted input projection rows of the coming chunks into a
//              ring (coalesced, no registers), completion signalled on mbarriers
//   warp  8    one thread issues tcgen05.mma:  D1z|D1r = h . Wh[0:2]^T,  D2 = (r*h) . Wh[2]^T
//
// Per step the chain is GEMM1 -> epilogue1 -> GEMM2 -> epilogue2; everything else is prefetch.
// H must be a multiple of 32, <= 128.
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

namespace fused {
using namespace tc;

constexpr int FM = 128;               // users per CTA = UMMA M
constexpr int F_THREADS = 384;      // 8 epilogue warps, 2 AX producer warps, MMA warp, Wh bulk-copy warp
constexpr int A_KB_BYTES = FM * 128;  // one 32-float k-block of the A tile
constexpr int AX_TILE = FM * 64;      // 128 rows x 16 floats
constexpr int AX_STAGES = 2;         // per column half
constexpr int OUT_STG = 32 * 64;      // per-warp output staging: 32 rows x 16 floats

// -DPOI_FUSED_TRACE (tools/fused_trace.py builds a separate library): CTA 0 records SM clock stamps of the
// recurrence's hand-offs, 16 slots per time step, read back with poi_debug_fused_trace
#ifdef POI_FUSED_TRACE
__device__ long long g_trace[2][512 * 16];
#define FTR(dir, step, slot) do { if (blockIdx.x == 0) g_trace[dir][(step) * 16 + (slot)] = clock64(); } while (0)
#define FTR_ADD(dir, step, slot, v) do { if (blockIdx.x == 0) g_trace[dir][(step) * 16 + (slot)] += (v); } while (0)
#define FTR_NOW() clock64()
#else
#define FTR(dir, step, slot) do { } while (0)
#define FTR_ADD(dir, step, slot, v) do { } while (0)
#define FTR_NOW() 0ll
#endif

__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 x;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(a));
    return x;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// 64-byte rows (16 floats): XOR the 16-byte chunk index with bits 1..2 of the row -> lane=row accesses
// and 4-lanes-per-row accesses are both bank-conflict free
__device__ __forceinline__ uint32_t sw64(int row, int c) { return row * 64 + ((c ^ ((row >> 1) & 3)) << 4); }

template <bool SPLIT3>
__device__ __forceinline__ void a_store16(uint32_t a_hi, uint32_t a_lo, int row, int col, const float (&v)[16]) {
    const int kb = col >> 5, cc0 = (col & 31) >> 2;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t off = kb * A_KB_BYTES + row * 128 + (((cc0 + q) ^ (row & 7)) << 4);
        float4 x = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        if (SPLIT3) { float4 hi, lo; split4(x, hi, lo); sts4(a_hi + off, hi); sts4(a_lo + off, lo); }
        else sts4(a_hi + off, x);
    }
}
// exact fp32 values of the A tile (hi + lo is exact by construction)
template <bool SPLIT3>
__device__ __forceinline__ void a_load16(uint32_t a_hi, uint32_t a_lo, int row, int col, float (&v)[16]) {
    const int kb = col >> 5, cc0 = (col & 31) >> 2;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t off = kb * A_KB_BYTES + row * 128 + (((cc0 + q) ^ (row & 7)) << 4);
        float4 x = lds4(a_hi + off);
        if (SPLIT3) { float4 l = lds4(a_lo + off); x.x += l.x; x.y += l.y; x.z += l.z; x.w += l.w; }
        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
}
// one warp: 32 rows x 16 floats from registers (lane = row) to global with 4 lanes per row
__device__ __forceinline__ void warp_store_chunk(uint32_t stg, int lane, const float (&v)[16], float* gbase, int ld,
                                                 int rows_valid) {
#pragma unroll
    for (int q = 0; q < 4; ++q) sts4(stg + sw64(lane, q), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    __syncwarp();
    const int qq = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rr = (lane >> 2) + 8 * i;
        float4 x = lds4(stg + sw64(rr, qq));
        if (rr < rows_valid) *reinterpret_cast<float4*>(gbase + (size_t)rr * ld + 4 * qq) = x;
    }
    __syncwarp();
}

template <bool SPLIT3>
__device__ __forceinline__ void mma_kblock(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi, uint32_t w_lo,
                                           uint32_t idesc, bool first) {
    const uint64_t dA = make_sdesc(a_hi), dW = make_sdesc(w_hi);
    uint32_t acc = first ? 0u : 1u;
    if (SPLIT3) {
        const uint64_t dAl = make_sdesc(a_lo), dWl = make_sdesc(w_lo);
#pragma unroll
        for (int k = 0; k < 4; ++k) { umma_tf32(d_tmem, dAl + 2 * k, dW + 2 * k, idesc, acc); acc = 1u; }
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, dA + 2 * k, dWl + 2 * k, idesc, 1u);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { umma_tf32(d_tmem, dA + 2 * k, dW + 2 * k, idesc, acc); acc = 1u; }
}

// Wh staged ONCE per call in the exact shared-memory image the UMMA wants (tile = one gate x one 32-float
// k-block: H rows x 128 B, 128B-swizzled, hi then lo for 3xTF32), so that inside the recurrence a tile is one
// bulk async copy (TMA, cp.async.bulk) issued by a single thread instead of 64 threads converting it
template <bool SPLIT3>
__global__ void k_stage_wh(const float* __restrict__ wh, int H, uint8_t* __restrict__ image) {
    const int KB = H >> 5;
    const int64_t n = (int64_t)3 * KB * H * 8;
    const uint32_t w_tile = H * 128, w_stage = w_tile * (SPLIT3 ? 2 : 1);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int t12 = (int)(idx / (H * 8)), rem = (int)(idx % (H * 8)), row = rem >> 3, c = rem & 7;
        const int gate = t12 / KB, kb = t12 % KB;
        float4 v = *reinterpret_cast<const float4*>(wh + ((size_t)gate * H + row) * H + kb * 32 + c * 4);
        uint8_t* dst = image + (size_t)t12 * w_stage + row * 128 + ((c ^ (row & 7)) << 4);
        if (SPLIT3) { float4 hi, lo; split4(v, hi, lo); *reinterpret_cast<float4*>(dst) = hi; *reinterpret_cast<float4*>(dst + w_tile) = lo; }
        else *reinterpret_cast<float4*>(dst) = v;
    }
}

// tensor-core modes only: the gate non-linearities with the fast exp / divide (error ~1e-6, far inside the
// 1e-4 parity bar; the fp32 FMA mode keeps expf/tanhf)
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

template <bool SPLIT3>
__global__ void __launch_bounds__(F_THREADS, 1)
k_gru_fwd_fused(const float* __restrict__ AX, const uint8_t* __restrict__ wimg, float* __restrict__ Hs, float* __restrict__ Z,
                float* __restrict__ R, float* __restrict__ C, int B, int T, int H) {
    constexpr int WST = 2;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t w_full[WST], w_empty[WST], ax_full[2][AX_STAGES], ax_empty[2][AX_STAGES], a_ready, d1_full, d2_full;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KB = H >> 5, NCH = H >> 4, HCH = NCH >> 1;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_hi = sbase, a_lo = sbase + KB * A_KB_BYTES;
    const uint32_t w_tile = H * 128, w_stage = w_tile * (SPLIT3 ? 2 : 1);
    const uint32_t w_base = sbase + KB * A_KB_BYTES * (SPLIT3 ? 2 : 1);
    const uint32_t ax_base = w_base + WST * w_stage;           // [half][stage] tiles of AX_TILE bytes
    const int m0 = blockIdx.x * FM;
    uint32_t ncols = 32; while (ncols < (uint32_t)(3 * H)) ncols <<= 1;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < WST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int s = 0; s < AX_STAGES; ++s) { mbar_init(&ax_full[h][s], 32); mbar_init(&ax_empty[h][s], 128); }
        mbar_init(&a_ready, 256); mbar_init(&d1_full, 1); mbar_init(&d2_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc_tf32(FM, H);

    if (warp < 8) {
        // ================================ epilogue (8 warps) ================================
        // warp w: TMEM lane quadrant q = w & 3 (rows 32q..32q+31), column half hf = w >> 2
        const int q = warp & 3, hf = warp >> 2;
        const int row = q * 32 + lane;                         // TMEM lane == row of the tile
        const bool ok = m0 + row < B;
        const uint32_t tl = (uint32_t)(q * 32) << 16;
        const int rows_valid = min(32, B - (m0 + q * 32));
        const int k_beg = hf * HCH, k_end = k_beg + HCH;
        if (hf == 0 || true) {   // h_{-1} = 0 -> this warp's half of the A tile
            float zero[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) zero[i] = 0.f;
            for (int k = k_beg; k < k_end; ++k) a_store16<SPLIT3>(a_hi, a_lo, row, 16 * k, zero);
            fence_async_smem();
            mbar_arrive(&a_ready);
        }
        int64_t axi = 0;                                       // position in this half's AX ring
        // wait for the next AX tile, read this thread's row, return this warp's 2 KB slice of the tile (free to
        // be reused as output staging once the whole warp has read) and the barrier to release it on
        auto ax_take = [&](float (&v)[16], uint32_t& slice, uint64_t*& rel) {
            const int s = (int)(axi % AX_STAGES);
            const long long tw0 = FTR_NOW();
            mbar_wait(&ax_full[hf][s], (uint32_t)(axi / AX_STAGES) & 1);
            if (warp == 0 && lane == 0) FTR_ADD(0, (int)(axi / (3 * HCH)), 12, FTR_NOW() - tw0);
            slice = ax_base + (hf * AX_STAGES + s) * AX_TILE + q * OUT_STG;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float4 x = lds4(slice + sw64(lane, c));
                v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
            }
            __syncwarp();
            rel = &ax_empty[hf][s];
            ++axi;
        };
        for (int j = 0; j < T; ++j) {
            const size_t wrow = (size_t)j * B + m0 + q * 32;   // first global row of this warp's quadrant at step j
            // ---- epilogue 1: z, r, r*h ; stash z and (1-z)*h ----
            mbar_wait(&d1_full, j & 1);
            tc_fence_after();
            if (lane == 0 && (warp == 0 || warp == 7)) FTR(0, j, warp == 0 ? 4 : 8);
            for (int k = k_beg; k < k_end; ++k) {
                const int c0 = 16 * k;
                float a[16], b[16], hv[16], dz[16], dr[16];
                uint32_t sl_z, sl_r; uint64_t *rel_z, *rel_r;
                ax_take(a, sl_z, rel_z); ax_take(b, sl_r, rel_r);
                tmem_ld16(tmem + tl + (uint32_t)c0, dz);
                tmem_ld16(tmem + tl + (uint32_t)(H + c0), dr);
                a_load16<SPLIT3>(a_hi, a_lo, row, c0, hv);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float zz = sigmoid_fast(dz[i] + a[i]);
                    const float rr = sigmoid_fast(dr[i] + b[i]);
                    const float h_ = ok ? hv[i] : 0.f;
                    dz[i] = zz; dr[i] = rr;
                    a[i] = rr * h_;                 // r*h -> next A operand
                    b[i] = (1.f - zz) * h_;         // (1-z)*h_prev, used by epilogue 2
                }
                tmem_st16(tmem + tl + (uint32_t)c0, dz);              // stash z
                tmem_st16(tmem + tl + (uint32_t)(H + c0), b);         // stash (1-z)*h_prev
                a_store16<SPLIT3>(a_hi, a_lo, row, c0, a);
                warp_store_chunk(sl_z, lane, dz, Z + wrow * H + c0, H, rows_valid);
                warp_store_chunk(sl_r, lane, dr, R + wrow * H + c0, H, rows_valid);
                mbar_arrive(rel_z); mbar_arrive(rel_r);
            }
            fence_async_smem();
            tc_fence_before();
            if (lane == 0 && (warp == 0 || warp == 7)) FTR(0, j, warp == 0 ? 5 : 9);
            mbar_arrive(&a_ready);                 // this warp's part of the r*h tile is in place
            // ---- epilogue 2: c, h_t ----
            mbar_wait(&d2_full, j & 1);
            tc_fence_after();
            if (lane == 0 && (warp == 0 || warp == 7)) FTR(0, j, warp == 0 ? 6 : 10);
            for (int k = k_beg; k < k_end; ++k) {
                const int c0 = 16 * k;
                float a[16], dc[16], zz[16], u[16];
                uint32_t sl; uint64_t* rel;
                ax_take(a, sl, rel);
                tmem_ld16(tmem + tl + (uint32_t)(2 * H + c0), dc);
                tmem_ld16(tmem + tl + (uint32_t)c0, zz);
                tmem_ld16(tmem + tl + (uint32_t)(H + c0), u);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float cc = tanh_fast(dc[i] + a[i]);
                    dc[i] = cc;
                    u[i] = ok ? u[i] + zz[i] * cc : 0.f;   // h_t = (1-z) h_prev + z c
                }
                a_store16<SPLIT3>(a_hi, a_lo, row, c0, u);
                warp_store_chunk(sl, lane, dc, C + wrow * H + c0, H, rows_valid);
                warp_store_chunk(sl, lane, u, Hs + (wrow + B) * H + c0, H, rows_valid);
                mbar_arrive(rel);
            }
            fence_async_smem();
            tc_fence_before();
            if (lane == 0 && (warp == 0 || warp == 7)) FTR(0, j, warp == 0 ? 7 : 11);
            mbar_arrive(&a_ready);                 // this warp's part of the h_t tile is in place
        }
    } else if (warp < 10) {
        // ================================ AX producers (one warp per column half, cp.async) ================================
        const int hf = warp - 8;
        const int per_step = 3 * HCH;
        const int64_t n_tiles = (int64_t)T * per_step;
        for (int64_t ai = 0; ai < n_tiles; ++ai) {
            const int s = (int)(ai % AX_STAGES);
            if (ai >= AX_STAGES) mbar_wait(&ax_empty[hf][s], (uint32_t)((ai / AX_STAGES) - 1) & 1);
            const int j = (int)(ai / per_step), ti = (int)(ai % per_step);
            const int gate = ti < 2 * HCH ? (ti & 1) : 2;
            const int k = hf * HCH + (ti < 2 * HCH ? (ti >> 1) : ti - 2 * HCH);
            const float* src = AX + ((size_t)j * B + m0) * 3 * H + gate * H + 16 * k;
            const uint32_t dst = ax_base + (hf * AX_STAGES + s) * AX_TILE;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                int f = lane + i * 32, rw = f >> 2, c = f & 3;
                if (m0 + rw < B)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + sw64(rw, c)), "l"(src + (size_t)rw * 3 * H + 4 * c) : "memory");
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&ax_full[hf][s])) : "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp == 10) {
        if (lane == 0) {
            // ================================ MMA issuer ================================
            int64_t ws = 0; uint32_t pa = 0;
            for (int j = 0; j < T; ++j) {
                mbar_wait(&a_ready, pa & 1); ++pa;     // h_{j-1} tile staged (and the stashes of step j-1 consumed)
                tc_fence_after();
                FTR(0, j, 0);
                for (int half = 0; half < 2; ++half) {
                    for (int kb = 0; kb < KB; ++kb, ++ws) {
                        const int s = (int)(ws % WST);
                        const long long tw0 = FTR_NOW();
                        mbar_wait(&w_full[s], (uint32_t)(ws / WST) & 1);
                        FTR_ADD(0, j, 13, FTR_NOW() - tw0);
                        tc_fence_after();
                        const uint32_t sW = w_base + s * w_stage;
                        mma_kblock<SPLIT3>(tmem + half * H, a_hi + kb * A_KB_BYTES, a_lo + kb * A_KB_BYTES, sW, sW + w_tile, idesc, kb == 0);
                        umma_commit(&w_empty[s]);
                    }
                }
                umma_commit(&d1_full);
                FTR(0, j, 1);
                mbar_wait(&a_ready, pa & 1); ++pa;     // r*h tile staged
                tc_fence_after();
                FTR(0, j, 2);
                for (int kb = 0; kb < KB; ++kb, ++ws) {
                    const int s = (int)(ws % WST);
                    const long long tw0 = FTR_NOW();
                    mbar_wait(&w_full[s], (uint32_t)(ws / WST) & 1);
                    FTR_ADD(0, j, 14, FTR_NOW() - tw0);
                    tc_fence_after();
                    const uint32_t sW = w_base + s * w_stage;
                    mma_kblock<SPLIT3>(tmem + 2 * H, a_hi + kb * A_KB_BYTES, a_lo + kb * A_KB_BYTES, sW, sW + w_tile, idesc, kb == 0);
                    umma_commit(&w_empty[s]);
                }
                umma_commit(&d2_full);
                FTR(0, j, 3);
            }
        }
    } else if (lane == 0) {
        // ================================ Wh tiles: one bulk async copy (TMA) per tile ================================
        const int64_t n_tiles = (int64_t)T * 3 * KB;
        for (int64_t ws = 0; ws < n_tiles; ++ws) {
            const int s = (int)(ws % WST);
            if (ws >= WST) mbar_wait(&w_empty[s], (uint32_t)((ws / WST) - 1) & 1);
            const uint32_t bar = smem_u32(&w_full[s]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(w_stage) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(w_base + s * w_stage), "l"(wimg + (size_t)(ws % (3 * KB)) * w_stage), "r"(w_stage), "r"(bar) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

// ---------------------------------------------------------------------------------------------
// Backward recurrence (BPTT through the cell) as one persistent kernel, same roles.
// Per step j = T-1 .. 0, for the 128 users of the CTA (dh carried in TMEM, never in global memory):
//   P : dht = dh + DHl_j ; da_c = dht z (1-c^2) ; da_z = dht (c-h') z(1-z) ; keep = dht (1-z)
//       A <- da_c                                         GEMM_M  : m   = da_c . Wh[2]
//   M1: A <- da_z                                         GEMM_DH1: dhn = da_z . Wh[0]
//   M2: da_r = m h' r(1-r) ; keep += m r      (overlaps GEMM_DH1)
//   M3: A <- da_r                                         GEMM_DH2: dhn += da_r . Wh[1]
//   D : dh = keep + dhn  -> P of step j-1
// TMEM columns: [0,H) D_m (then the da_r stash), [H,2H) D_dh, [2H,3H) keep, [3H,4H) da_z.
// Inputs (DHl, Z, C, H_prev, R) stream through cp.async rings in 16-column tiles, DA_z/DA_r/DA_c are
// stored through the consumed ring slices (coalesced), the transposed Wh tiles are bulk-copied (TMA) from
// an image staged once per call.
// ---------------------------------------------------------------------------------------------
template <bool SPLIT3>
__global__ void k_stage_wh_bwd(const float* __restrict__ wh, int H, uint8_t* __restrict__ image) {
    const int KB = H >> 5;
    const int64_t n = (int64_t)3 * KB * H * 8;
    const uint32_t w_tile = H * 128, w_stage = w_tile * (SPLIT3 ? 2 : 1);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int t12 = (int)(idx / (H * 8)), rem = (int)(idx % (H * 8)), row = rem >> 3, c = rem & 7;
        const int g = t12 / KB, kb = t12 % KB;
        const int gsel = g == 0 ? 2 : g - 1;             // tile order: Wh[2] (c), Wh[0] (z), Wh[1] (r)
        // B operand row n, k-major: element (n, k) = Wh[gsel][k][n]
        const float* src = wh + ((size_t)gsel * H + kb * 32 + c * 4) * H + row;
        float4 v = make_float4(src[0], src[H], src[2 * (size_t)H], src[3 * (size_t)H]);
        uint8_t* dst = image + (size_t)t12 * w_stage + row * 128 + ((c ^ (row & 7)) << 4);
        if (SPLIT3) { float4 hi, lo; split4(v, hi, lo); *reinterpret_cast<float4*>(dst) = hi; *reinterpret_cast<float4*>(dst + w_tile) = lo; }
        else *reinterpret_cast<float4*>(dst) = v;
    }
}

template <bool SPLIT3>
__global__ void __launch_bounds__(F_THREADS, 1)
k_gru_bwd_fused(const float* __restrict__ DHl, const float* __restrict__ Z, const float* __restrict__ R,
                const float* __restrict__ C, const float* __restrict__ Hs, const uint8_t* __restrict__ wimg,
                float* __restrict__ DA, int B, int T, int H) {
    constexpr int WST = 2;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t w_full[WST], w_empty[WST], in_full[2][AX_STAGES], in_empty[2][AX_STAGES], a_ready, dm_full, dh1_done, ddh_full;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KB = H >> 5, NCH = H >> 4, HCH = NCH >> 1;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_hi = sbase, a_lo = sbase + KB * A_KB_BYTES;
    const uint32_t w_tile = H * 128, w_stage = w_tile * (SPLIT3 ? 2 : 1);
    const uint32_t w_base = sbase + KB * A_KB_BYTES * (SPLIT3 ? 2 : 1);
    const uint32_t in_base = w_base + WST * w_stage;
    const int m0 = blockIdx.x * FM;
    uint32_t ncols = 32; while (ncols < (uint32_t)(4 * H)) ncols <<= 1;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < WST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int s = 0; s < AX_STAGES; ++s) { mbar_init(&in_full[h][s], 32); mbar_init(&in_empty[h][s], 128); }
        mbar_init(&a_ready, 256); mbar_init(&dm_full, 1); mbar_init(&dh1_done, 1); mbar_init(&ddh_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc_tf32(FM, H);
    const uint32_t T_M = 0, T_DH = (uint32_t)H, T_KEEP = (uint32_t)(2 * H), T_DAZ = (uint32_t)(3 * H);

    if (warp < 8) {
        // ================================ epilogue (8 warps) ================================
        const int q = warp & 3, hf = warp >> 2;
        const int row = q * 32 + lane;
        const bool ok = m0 + row < B;
        const uint32_t tl = (uint32_t)(q * 32) << 16;
        const int rows_valid = min(32, B - (m0 + q * 32));
        const int k_beg = hf * HCH, k_end = k_beg + HCH;
        int64_t ini = 0;
        auto in_take = [&](float (&v)[16], uint32_t& slice, uint64_t*& rel) {
            const int s = (int)(ini % AX_STAGES);
            const long long tw0 = FTR_NOW();
            mbar_wait(&in_full[hf][s], (uint32_t)(ini / AX_STAGES) & 1);
            if (warp == 0 && lane == 0) FTR_ADD(1, (int)(ini / (6 * HCH)), 13, FTR_NOW() - tw0);
            slice = in_base + (hf * AX_STAGES + s) * AX_TILE + q * OUT_STG;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float4 x = lds4(slice + sw64(lane, c));
                v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
            }
            __syncwarp();
            rel = &in_empty[hf][s];
            ++ini;
        };
        for (int j = T - 1; j >= 0; --j) {
            const int it = T - 1 - j;                           // iteration number (barrier phases)
            const size_t wrow = (size_t)j * B + m0 + q * 32;
            float* DAw = DA + wrow * 3 * H;
            // ---- D + P: dh = keep + dhn (0 for the last step) ; gate derivatives ; A <- da_c ----
            if (it > 0) { mbar_wait(&ddh_full, (it - 1) & 1); tc_fence_after(); }
            if (lane == 0 && warp == 0) FTR(1, it, 6);
            for (int k = k_beg; k < k_end; ++k) {
                const int c0 = 16 * k;
                float dl[16], zz[16], cc[16], hp[16], dh[16], kp[16];
                uint32_t s0, s1, s2, s3; uint64_t *r0, *r1, *r2, *r3;
                // the ring has 2 stages: release the first two tiles as soon as they are in registers, keep the
                // slices of the last two as store staging
                in_take(dl, s0, r0); mbar_arrive(r0);
                in_take(zz, s1, r1); mbar_arrive(r1);
                in_take(cc, s2, r2); in_take(hp, s3, r3);
                if (it > 0) {
                    tmem_ld16(tmem + tl + T_DH + (uint32_t)c0, dh);
                    tmem_ld16(tmem + tl + T_KEEP + (uint32_t)c0, kp);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) { dh[i] = 0.f; kp[i] = 0.f; }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float dht = ok ? dh[i] + kp[i] + dl[i] : 0.f;
                    const float z_ = zz[i], c_ = cc[i];
                    const float dac = dht * z_ * (1.f - c_ * c_);
                    const float daz = dht * (c_ - hp[i]) * z_ * (1.f - z_);
                    kp[i] = ok ? dht * (1.f - z_) : 0.f;        // keep
                    dl[i] = ok ? dac : 0.f;                     // da_c
                    dh[i] = ok ? daz : 0.f;                     // da_z
                }
                tmem_st16(tmem + tl + T_KEEP + (uint32_t)c0, kp);
                tmem_st16(tmem + tl + T_DAZ + (uint32_t)c0, dh);
                a_store16<SPLIT3>(a_hi, a_lo, row, c0, dl);
                warp_store_chunk(s2, lane, dh, DAw + c0, 3 * H, rows_valid);            // DA_z
                warp_store_chunk(s3, lane, dl, DAw + 2 * H + c0, 3 * H, rows_valid);    // DA_c
                mbar_arrive(r2); mbar_arrive(r3);
            }
            fence_async_smem();
            tc_fence_before();
            if (lane == 0 && warp == 0) FTR(1, it, 7);
            mbar_arrive(&a_ready);                              // A = da_c
            // ---- M1: A <- da_z (after GEMM_M has read da_c) ----
            mbar_wait(&dm_full, it & 1);
            tc_fence_after();
            if (lane == 0 && warp == 0) FTR(1, it, 8);
            if (j > 0) {
                for (int k = k_beg; k < k_end; ++k) {
                    float dz[16];
                    tmem_ld16(tmem + tl + T_DAZ + (uint32_t)(16 * k), dz);
                    a_store16<SPLIT3>(a_hi, a_lo, row, 16 * k, dz);
                }
                fence_async_smem();
                tc_fence_before();
                if (lane == 0 && warp == 0) FTR(1, it, 9);
                mbar_arrive(&a_ready);                          // A = da_z
            }
            // ---- M2: da_r, keep += m r  (runs while GEMM_DH1 executes) ----
            for (int k = k_beg; k < k_end; ++k) {
                const int c0 = 16 * k;
                float rr[16], hp[16], mm[16], kp[16];
                uint32_t s0, s1; uint64_t *r0, *r1;
                in_take(rr, s0, r0); in_take(hp, s1, r1);
                tmem_ld16(tmem + tl + T_M + (uint32_t)c0, mm);
                tmem_ld16(tmem + tl + T_KEEP + (uint32_t)c0, kp);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float m_ = ok ? mm[i] : 0.f, r_ = rr[i];
                    kp[i] = kp[i] + m_ * r_;
                    mm[i] = ok ? m_ * hp[i] * r_ * (1.f - r_) : 0.f;       // da_r
                }
                tmem_st16(tmem + tl + T_KEEP + (uint32_t)c0, kp);
                tmem_st16(tmem + tl + T_M + (uint32_t)c0, mm);              // stash da_r over m
                warp_store_chunk(s0, lane, mm, DAw + H + c0, 3 * H, rows_valid);        // DA_r
                mbar_arrive(r0); mbar_arrive(r1);
            }
            if (lane == 0 && warp == 0) FTR(1, it, 10);
            // ---- M3: A <- da_r (after GEMM_DH1 has read da_z) ----
            if (j > 0) {
                mbar_wait(&dh1_done, it & 1);
                tc_fence_after();
                if (lane == 0 && warp == 0) FTR(1, it, 11);
                for (int k = k_beg; k < k_end; ++k) {
                    float dr[16];
                    tmem_ld16(tmem + tl + T_M + (uint32_t)(16 * k), dr);
                    a_store16<SPLIT3>(a_hi, a_lo, row, 16 * k, dr);
                }
                fence_async_smem();
                tc_fence_before();
                if (lane == 0 && warp == 0) FTR(1, it, 12);
                mbar_arrive(&a_ready);                          // A = da_r
            }
        }
    } else if (warp < 10) {
        // ================================ input producers (one warp per column half, cp.async) ================================
        const int hf = warp - 8;
        const int per_step = 6 * HCH;
        const int64_t n_tiles = (int64_t)T * per_step;
        for (int64_t ai = 0; ai < n_tiles; ++ai) {
            const int s = (int)(ai % AX_STAGES);
            if (ai >= AX_STAGES) mbar_wait(&in_empty[hf][s], (uint32_t)((ai / AX_STAGES) - 1) & 1);
            const int j = T - 1 - (int)(ai / per_step), ti = (int)(ai % per_step);
            int kind, kk;                                       // 0 DHl, 1 Z, 2 C, 3 H_prev, 4 R
            if (ti < 4 * HCH) { kind = ti & 3; kk = ti >> 2; }
            else { const int t2 = ti - 4 * HCH; kind = (t2 & 1) ? 3 : 4; kk = t2 >> 1; }
            const float* base = kind == 0 ? DHl : kind == 1 ? Z : kind == 2 ? C : kind == 3 ? Hs : R;
            const float* src = base + ((size_t)j * B + m0) * H + 16 * (hf * HCH + kk);
            const uint32_t dst = in_base + (hf * AX_STAGES + s) * AX_TILE;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                int f = lane + i * 32, rw = f >> 2, c = f & 3;
                if (m0 + rw < B)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + sw64(rw, c)), "l"(src + (size_t)rw * H + 4 * c) : "memory");
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&in_full[hf][s])) : "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp == 10) {
        if (lane == 0) {
            // ================================ MMA issuer ================================
            int64_t ws = 0; uint32_t pa = 0;
            auto gemm = [&](uint32_t dcol, bool fresh) {
                for (int kb = 0; kb < KB; ++kb, ++ws) {
                    const int s = (int)(ws % WST);
                    const long long tw0 = FTR_NOW();
                    mbar_wait(&w_full[s], (uint32_t)(ws / WST) & 1);
                    FTR_ADD(1, (int)(ws / (3 * KB)), 14, FTR_NOW() - tw0);
                    tc_fence_after();
                    const uint32_t sW = w_base + s * w_stage;
                    mma_kblock<SPLIT3>(tmem + dcol, a_hi + kb * A_KB_BYTES, a_lo + kb * A_KB_BYTES, sW, sW + w_tile, idesc, fresh && kb == 0);
                    umma_commit(&w_empty[s]);
                }
            };
            for (int j = T - 1; j >= 0; --j) {
                mbar_wait(&a_ready, pa & 1); ++pa; tc_fence_after();    // A = da_c
                FTR(1, T - 1 - j, 0);
                gemm(T_M, true);
                umma_commit(&dm_full);
                FTR(1, T - 1 - j, 1);
                if (j > 0) {
                    mbar_wait(&a_ready, pa & 1); ++pa; tc_fence_after();    // A = da_z
                    FTR(1, T - 1 - j, 2);
                    gemm(T_DH, true);
                    umma_commit(&dh1_done);
                    FTR(1, T - 1 - j, 3);
                    mbar_wait(&a_ready, pa & 1); ++pa; tc_fence_after();    // A = da_r
                    FTR(1, T - 1 - j, 4);
                    gemm(T_DH, false);
                    umma_commit(&ddh_full);
                    FTR(1, T - 1 - j, 5);
                }
            }
        }
    } else if (lane == 0) {
        // ================================ transposed Wh tiles: one bulk async copy (TMA) each ================================
        const int64_t n_tiles = (int64_t)(T - 1) * 3 * KB + KB;         // the last step (j = 0) needs only Wh[2]
        for (int64_t ws = 0; ws < n_tiles; ++ws) {
            const int s = (int)(ws % WST);
            if (ws >= WST) mbar_wait(&w_empty[s], (uint32_t)((ws / WST) - 1) & 1);
            const uint32_t bar = smem_u32(&w_full[s]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(w_stage) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(w_base + s * w_stage), "l"(wimg + (size_t)(ws % (3 * KB)) * w_stage), "r"(w_stage), "r"(bar) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

template <bool SPLIT3>
static int launch_bwd_inst(poi_engine* e, const float* DHl, const float* Z, const float* R, const float* C, const float* Hs,
                           const float* wh, float* DA, int B, int T, int H) {
    const int KB = H / 32;
    const size_t w_stage = (size_t)H * 128 * (SPLIT3 ? 2 : 1);
    uint8_t* wimg = nullptr;
    POI_TRY(arena_get(e, (size_t)3 * KB * w_stage, &wimg));
    POI_CAT(e, CAT_ELTWISE, 0, 0);
    POI_LAUNCH(e, (k_stage_wh_bwd<SPLIT3>), 48, 256, 0, wh, H, wimg);
    size_t smem = (size_t)KB * A_KB_BYTES * (SPLIT3 ? 2 : 1) + 2 * w_stage + (size_t)2 * AX_STAGES * AX_TILE + 1024;
    POI_CK(e, cudaFuncSetAttribute(k_gru_bwd_fused<SPLIT3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    POI_CAT(e, CAT_GEMM, 2.0 * (double)B * T * 3 * H * H, 0);
    POI_LAUNCH(e, (k_gru_bwd_fused<SPLIT3>), (unsigned)poi_cdiv(B, FM), F_THREADS, smem, DHl, Z, R, C, Hs, wimg, DA, B, T, H);
    return 0;
}

static int launch_gru_bwd_fused(poi_engine* e, const float* DHl, const float* Z, const float* R, const float* C,
                                const float* Hs, const float* wh, float* DA, int B, int T, int H, bool split3) {
    if (split3) return launch_bwd_inst<true>(e, DHl, Z, R, C, Hs, wh, DA, B, T, H);
    return launch_bwd_inst<false>(e, DHl, Z, R, C, Hs, wh, DA, B, T, H);
}

static inline bool fwd_supported(int H) { return H % 32 == 0 && H >= 32 && H <= 128; }

// RH = R * H_prev (the r-gated state the candidate GEMM consumed): the fused kernel keeps it on chip, the
// weight-gradient stage wants it in memory
__global__ void k_mul_rh(const float* __restrict__ R, const float* __restrict__ Hprev, float* __restrict__ RH, int64_t n4) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 r = reinterpret_cast<const float4*>(R)[i], h = reinterpret_cast<const float4*>(Hprev)[i];
    reinterpret_cast<float4*>(RH)[i] = make_float4(r.x * h.x, r.y * h.y, r.z * h.z, r.w * h.w);
}

template <bool SPLIT3>
static int launch_fwd_inst(poi_engine* e, const float* AX, const float* wh, float* Hs, float* Z, float* R, float* C,
                           int B, int T, int H) {
    const int KB = H / 32;
    const size_t w_stage = (size_t)H * 128 * (SPLIT3 ? 2 : 1);
    uint8_t* wimg = nullptr;
    POI_TRY(arena_get(e, (size_t)3 * KB * w_stage, &wimg));
    POI_CAT(e, CAT_ELTWISE, 0, 0);
    POI_LAUNCH(e, (k_stage_wh<SPLIT3>), 48, 256, 0, wh, H, wimg);
    size_t smem = (size_t)KB * A_KB_BYTES * (SPLIT3 ? 2 : 1) + 2 * w_stage + (size_t)2 * AX_STAGES * AX_TILE + 1024;
    POI_CK(e, cudaFuncSetAttribute(k_gru_fwd_fused<SPLIT3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    POI_CAT(e, CAT_GEMM, 2.0 * (double)B * T * 3 * H * H, 0);
    POI_LAUNCH(e, (k_gru_fwd_fused<SPLIT3>), (unsigned)poi_cdiv(B, FM), F_THREADS, smem, AX, wimg, Hs, Z, R, C, B, T, H);
    return 0;
}

static int launch_gru_fwd_fused(poi_engine* e, const float* AX, const float* wh, float* Hs, float* Z, float* R,
                                float* C, float* RH, int B, int T, int H, bool split3) {
    if (split3) POI_TRY(launch_fwd_inst<true>(e, AX, wh, Hs, Z, R, C, B, T, H));
    else POI_TRY(launch_fwd_inst<false>(e, AX, wh, Hs, Z, R, C, B, T, H));
    const int64_t n4 = (int64_t)T * B * H / 4;
    POI_CAT(e, CAT_ELTWISE, 0, 3.0 * (double)n4 * 16);
    POI_LAUNCH(e, k_mul_rh, (unsigned)poi_cdiv(n4, 256), 256, 0, R, Hs, RH, n4);
    return 0;
}

}  // namespace fused
