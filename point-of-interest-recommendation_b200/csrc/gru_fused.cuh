// gru_fused.cuh -- the GRU recurrence as ONE persistent tcgen05 kernel per direction.
//
// The reference's `theano.scan` runs the cell once per time step (GRU.py:345-360,
// GRU_Spatial.py:170-197); the first version of this engine did the same with two GEMM launches
// per step.  Users are independent inside a step and across steps only through their own h, so a
// CTA can own 128 users for the WHOLE sequence: h_t never leaves the SM between steps.
//
//   warps 0-3  epilogue: thread = one user row.  TMEM accumulator -> gate math -> Z/R/C/H to global
//              (saved for BPTT) and the next A operand (r*h, then h_t) straight into shared memory
//              in the UMMA layout (128B-swizzled K-major, hi/lo split for 3xTF32)
//   warps 4-7  producers: stream the Wh tiles (L2-resident, identical every step) through a ring
//   warp  8    one thread issues tcgen05.mma:  D1z|D1r = h . Wh[0:2]^T,  D2 = (r*h) . Wh[2]^T
//
// Per step the chain is GEMM1 -> epilogue1 -> GEMM2 -> epilogue2; nothing else is on it: the input
// projection AX (hoisted GEMM) and Wh arrive by prefetch.  H must be a multiple of 32, <= 128.
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

namespace fused {
using namespace tc;

constexpr int FM = 128;               // users per CTA = UMMA M
constexpr int F_THREADS = 288;
constexpr int A_KB_BYTES = FM * 128;  // one 32-float k-block of the A tile

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
__device__ __forceinline__ void ld16(const float* p, bool ok, float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float4 x = ok ? *reinterpret_cast<const float4*>(p + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
}
__device__ __forceinline__ void st16(float* p, const float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(p + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// issue the MMAs of one k-block: A tile (hi/lo) x W stage (hi/lo) -> D
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

template <bool SPLIT3>
__global__ void __launch_bounds__(F_THREADS, 1)
k_gru_fwd_fused(const float* __restrict__ AX, const float* __restrict__ wh, float* Hs, float* Z, float* R,
                float* C, float* RH, int B, int T, int H) {
    constexpr int STAGES = SPLIT3 ? 3 : 4;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t w_full[STAGES], w_empty[STAGES], a_ready, d1_full, d2_full;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KB = H >> 5;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_hi = sbase, a_lo = sbase + KB * A_KB_BYTES;
    const uint32_t w_base = sbase + KB * A_KB_BYTES * (SPLIT3 ? 2 : 1);
    const uint32_t w_tile = H * 128, w_stage = w_tile * (SPLIT3 ? 2 : 1);
    const int m0 = blockIdx.x * FM;
    uint32_t ncols = 32; while (ncols < (uint32_t)(3 * H)) ncols <<= 1;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&w_full[s], 128); mbar_init(&w_empty[s], 1); }
        mbar_init(&a_ready, 128); mbar_init(&d1_full, 1); mbar_init(&d2_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc_tf32(FM, H);

    if (warp < 4) {
        // ================================ epilogue ================================
        const int row = tid;                       // TMEM lane == row of the tile
        const int64_t m = (int64_t)m0 + row;
        const bool ok = m < B;
        const uint32_t tl = (uint32_t)(warp * 32) << 16;
        {   // h_{-1} = 0 -> A tile
            float zero[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) zero[i] = 0.f;
            for (int c0 = 0; c0 < H; c0 += 16) a_store16<SPLIT3>(a_hi, a_lo, row, c0, zero);
            fence_async_smem();
            mbar_arrive(&a_ready);
        }
        for (int j = 0; j < T; ++j) {
            const size_t rb = (size_t)j * B + (ok ? m : 0);
            const float* ax = AX + rb * 3 * H;
            const float* hp = Hs + rb * H;
            float* hn = Hs + (rb + B) * H;
            float *zj = Z + rb * H, *rj = R + rb * H, *cj = C + rb * H, *rhj = RH + rb * H;
            // ---- epilogue 1: z, r, r*h ----
            float axz[16], axr[16], hv[16];
            ld16(ax, ok, axz); ld16(ax + H, ok, axr); ld16(hp, ok, hv);          // in flight while GEMM1 runs
            mbar_wait(&d1_full, j & 1);
            tc_fence_after();
            for (int c0 = 0; c0 < H; c0 += 16) {
                float nz[16], nr[16], nh[16];
                const bool more = c0 + 16 < H;
                if (more) { ld16(ax + c0 + 16, ok, nz); ld16(ax + H + c0 + 16, ok, nr); ld16(hp + c0 + 16, ok, nh); }
                float dz[16], dr[16], zz[16], rr[16], rh[16];
                tmem_ld16(tmem + tl + (uint32_t)c0, dz);
                tmem_ld16(tmem + tl + (uint32_t)(H + c0), dr);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    zz[i] = sigmoidf_(dz[i] + axz[i]);
                    rr[i] = sigmoidf_(dr[i] + axr[i]);
                    rh[i] = ok ? rr[i] * hv[i] : 0.f;
                }
                if (ok) { st16(zj + c0, zz); st16(rj + c0, rr); st16(rhj + c0, rh); }
                a_store16<SPLIT3>(a_hi, a_lo, row, c0, rh);
                if (more) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) { axz[i] = nz[i]; axr[i] = nr[i]; hv[i] = nh[i]; }
                }
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(&a_ready);                 // r*h tile ready, D1 drained
            // ---- epilogue 2: c, h_t ----
            float axc[16], zv[16];
            ld16(ax + 2 * H, ok, axc); ld16(zj, ok, zv); ld16(hp, ok, hv);
            mbar_wait(&d2_full, j & 1);
            tc_fence_after();
            for (int c0 = 0; c0 < H; c0 += 16) {
                float nc[16], nz[16], nh[16];
                const bool more = c0 + 16 < H;
                if (more) { ld16(ax + 2 * H + c0 + 16, ok, nc); ld16(zj + c0 + 16, ok, nz); ld16(hp + c0 + 16, ok, nh); }
                float dc[16], cc[16], hh[16];
                tmem_ld16(tmem + tl + (uint32_t)(2 * H + c0), dc);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    cc[i] = tanhf(dc[i] + axc[i]);
                    hh[i] = ok ? (1.f - zv[i]) * hv[i] + zv[i] * cc[i] : 0.f;
                }
                if (ok) { st16(cj + c0, cc); st16(hn + c0, hh); }
                a_store16<SPLIT3>(a_hi, a_lo, row, c0, hh);
                if (more) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) { axc[i] = nc[i]; zv[i] = nz[i]; hv[i] = nh[i]; }
                }
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(&a_ready);                 // h_t tile ready, D2 drained
        }
    } else if (warp < 8) {
        // ================================ Wh producers ================================
        const int ptid = tid - 128;
        const int LW = H >> 4;                     // 16-byte chunks per thread per tile (H*8/128)
        const int64_t n_tiles = (int64_t)T * 3 * KB;
        float4 r0[8], r1[8];
        auto gload = [&](int64_t ws, float4 (&rg)[8]) {
            const int t12 = (int)(ws % (3 * KB)), gate = t12 / KB, kb = t12 % KB;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < LW) {
                    int f = ptid + i * 128, rw = f >> 3, c = f & 7;
                    rg[i] = __ldg(reinterpret_cast<const float4*>(wh + ((size_t)gate * H + rw) * H + kb * 32 + c * 4));
                }
            }
        };
        auto stage_in = [&](int64_t ws, const float4 (&rg)[8]) {
            const int s = (int)(ws % STAGES);
            if (ws >= STAGES) mbar_wait(&w_empty[s], (uint32_t)((ws / STAGES) - 1) & 1);
            const uint32_t sW = w_base + s * w_stage, sWl = sW + w_tile;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < LW) {
                    int f = ptid + i * 128, rw = f >> 3, c = f & 7;
                    uint32_t off = rw * 128 + ((c ^ (rw & 7)) << 4);
                    if (SPLIT3) { float4 hi, lo; split4(rg[i], hi, lo); sts4(sW + off, hi); sts4(sWl + off, lo); }
                    else sts4(sW + off, rg[i]);
                }
            }
            fence_async_smem();
            mbar_arrive(&w_full[s]);
        };
        if (n_tiles > 0) gload(0, r0);
        for (int64_t ws = 0; ws < n_tiles; ws += 2) {
            if (ws + 1 < n_tiles) gload(ws + 1, r1);
            stage_in(ws, r0);
            if (ws + 1 < n_tiles) {
                if (ws + 2 < n_tiles) gload(ws + 2, r0);
                stage_in(ws + 1, r1);
            }
        }
    } else if (lane == 0) {
        // ================================ MMA issuer ================================
        int64_t ws = 0; uint32_t pa = 0;
        for (int j = 0; j < T; ++j) {
            mbar_wait(&a_ready, pa & 1); ++pa;     // h_{j-1} tile staged (and D2 of the previous step drained)
            tc_fence_after();
            for (int half = 0; half < 2; ++half) {
                for (int kb = 0; kb < KB; ++kb, ++ws) {
                    const int s = (int)(ws % STAGES);
                    mbar_wait(&w_full[s], (uint32_t)(ws / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t sW = w_base + s * w_stage;
                    mma_kblock<SPLIT3>(tmem + half * H, a_hi + kb * A_KB_BYTES, a_lo + kb * A_KB_BYTES, sW, sW + w_tile, idesc, kb == 0);
                    umma_commit(&w_empty[s]);
                }
            }
            umma_commit(&d1_full);
            mbar_wait(&a_ready, pa & 1); ++pa;     // r*h tile staged
            tc_fence_after();
            for (int kb = 0; kb < KB; ++kb, ++ws) {
                const int s = (int)(ws % STAGES);
                mbar_wait(&w_full[s], (uint32_t)(ws / STAGES) & 1);
                tc_fence_after();
                const uint32_t sW = w_base + s * w_stage;
                mma_kblock<SPLIT3>(tmem + 2 * H, a_hi + kb * A_KB_BYTES, a_lo + kb * A_KB_BYTES, sW, sW + w_tile, idesc, kb == 0);
                umma_commit(&w_empty[s]);
            }
            umma_commit(&d2_full);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

static inline bool fwd_supported(int H) { return H % 32 == 0 && H >= 32 && H <= 128; }

template <bool SPLIT3>
static int launch_fwd_inst(poi_engine* e, const float* AX, const float* wh, float* Hs, float* Z, float* R, float* C,
                           float* RH, int B, int T, int H) {
    constexpr int STAGES = SPLIT3 ? 3 : 4;
    const int KB = H / 32;
    size_t smem = (size_t)KB * A_KB_BYTES * (SPLIT3 ? 2 : 1) + (size_t)STAGES * H * 128 * (SPLIT3 ? 2 : 1) + 1024;
    POI_CK(e, cudaFuncSetAttribute(k_gru_fwd_fused<SPLIT3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // flops the tensor pipe is asked for: per step 2*B*H*3H (x3 products in 3xTF32 are not counted twice)
    POI_CAT(e, CAT_GEMM, 2.0 * (double)B * T * 3 * H * H, 0);
    POI_LAUNCH(e, (k_gru_fwd_fused<SPLIT3>), (unsigned)poi_cdiv(B, FM), F_THREADS, smem, AX, wh, Hs, Z, R, C, RH, B, T, H);
    return 0;
}

static int launch_gru_fwd_fused(poi_engine* e, const float* AX, const float* wh, float* Hs, float* Z, float* R,
                                float* C, float* RH, int B, int T, int H, bool split3) {
    if (split3) return launch_fwd_inst<true>(e, AX, wh, Hs, Z, R, C, RH, B, T, H);
    return launch_fwd_inst<false>(e, AX, wh, Hs, Z, R, C, RH, B, T, H);
}

}  // namespace fused
