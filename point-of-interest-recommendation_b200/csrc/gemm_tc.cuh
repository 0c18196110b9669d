// gemm_tc.cuh -- tcgen05 (5th-gen tensor core) GEMM:  C[m,n] = sum_k A[m*lda+k] * W[n*ldw+k]
//
// fp32 operands in global memory, TF32 tensor-core products, fp32 accumulation in TMEM.
//   mode 1 (SPLIT3): error-compensated 3xTF32.  Each fp32 x is split while it is staged into shared
//                    memory into hi = rn_tf32(x) and lo = rn_tf32(x - hi), and
//                    D += A_lo.W_hi + A_hi.W_lo + A_hi.W_hi  -> ~2^-21 relative error, fp32-faithful.
//   mode 2         : single-pass TF32 (the tensor core truncates the fp32 words itself).
//
// One CTA = one 128 x BN output tile: all 8 warps stage 128x32 / BNx32 fp32 tiles (global -> regs ->
// split -> 128B-swizzled K-major shared memory, the canonical UMMA layout), one elected thread issues
// tcgen05.mma (M=128, N=BN, K=8 per instruction), tcgen05.commit releases the stage through an
// mbarrier; the epilogue reads the accumulator with tcgen05.ld (32 lanes x 16 columns per warp) and
// calls the same fused epilogue functors as the FMA path (gemm_simt.cuh / gru.cuh).
#pragma once
#include "common.cuh"

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100):
//  [0,14) start>>4 | [16,30) LBO>>4 (=1, unused when swizzled) | [32,46) SBO>>4 (1024 B between 8-row
//  groups) | [46,48) version=1 | [61,64) layout type 2 = SWIZZLE_128B.  Tile base 1024-B aligned.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor for kind::tf32: c=F32 (bit 4), a=b=TF32 (2 at bits 7, 10), K-major both,
// N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void sts4(uint32_t saddr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// round-to-nearest TF32 (low 13 mantissa bits zero -> exactly what the tensor core reads) with two
// integer ops; a carry out of the mantissa bumps the exponent, which is the correct rounding
__device__ __forceinline__ float tf32_rn(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// hi = rn_tf32(x); lo = x - hi (exact in fp32; the tensor core truncates it to TF32: second-order error)
__device__ __forceinline__ void split4(float4 v, float4& hi, float4& lo) {
    hi = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
    lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
}

// Epilogue functors may split their global READS from the math: `Pre pre(m, n) const` issues the loads of one 4-column
// group, `operator()(m, n, v, pre)` consumes them.  The per-tile GEMM kernel then requests the inputs of four rows at a time
// -- the first four before it waits for the accumulator -- instead of paying one memory round trip per row after it
// (the per-time-step recurrence GEMMs of H > 128 spent a third of their life there).
template <class T, class = void> struct epi_has_pre { static constexpr bool value = false; };
template <class T> struct epi_has_pre<T, decltype((void)sizeof(typename T::Pre))> { static constexpr bool value = true; };

// thread-block-cluster helpers of the split-K pair (k_gemm_tn_tc<..., CSPLIT = true>)
__device__ __forceinline__ uint32_t tcx_mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void tcx_st_cluster4(uint32_t caddr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(caddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void tcx_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}

// rows per request batch; two batches are kept in registers (the inputs of batch j + 1 are in flight while batch j is consumed)
template <class T, bool = epi_has_pre<T>::value> struct epi_pre_rows { static constexpr int value = 4; };
template <class T> struct epi_pre_rows<T, true> { static constexpr int value = sizeof(typename T::Pre) <= 32 ? 4 : 2; };

constexpr int BM = 128;
constexpr int BK = 32;            // 32 fp32 = 128 B = one swizzle row
constexpr int PRODUCERS = 256;    // warps 0..7: stage operands, later run the epilogue
constexpr int THREADS = PRODUCERS + 32;   // + warp 8: one elected thread issues the MMAs

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int BN, bool SPLIT3>
struct Cfg {
    static constexpr int STAGES = SPLIT3 ? 3 : 4;
    static constexpr int A_BYTES = BM * BK * 4;
    static constexpr int W_BYTES = BN * BK * 4;
    static constexpr int STAGE_BYTES = (A_BYTES + W_BYTES) * (SPLIT3 ? 2 : 1);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
    static constexpr int LA = BM * 8 / PRODUCERS;     // 16-byte chunks per thread, A tile
    static constexpr int LW = BN * 8 / PRODUCERS;     // W tile
};

// Warp-specialised: producers run ahead through a STAGES-deep ring (full/empty mbarriers, no block-wide
// barrier in the main loop) with the next k-block's global loads already in flight in registers;
// blockIdx.z selects a K range [z*k_per_split, ...) (split-K for the weight-gradient GEMMs).
//
// CSPLIT (few tiles, long K: the per-time-step recurrence GEMMs at a few hundred users per GPU): the two K halves of a tile
// run as a cluster of two CTAs (cluster dims (1, 1, 2), rank = blockIdx.z) on two SMs -- the length of the dependent MMA
// chain, which is what such a launch costs, halves.  Rank 1 pushes its accumulator into rank 0's shared memory
// (st.shared::cluster, conflict-free slots), a cluster barrier orders it, rank 0 adds it (first half + second half: fixed
// order) and runs the epilogue.  Requires a functor with Pre.
template <int BN, bool SPLIT3, class Epi, bool CSPLIT = false>
__global__ void __launch_bounds__(THREADS, 1)
k_gemm_tn_tc(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
             int M, int N, int K, int k_per_split, Epi epi) {
    using C = Cfg<BN, SPLIT3>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mbar_full[C::STAGES];
    __shared__ uint64_t mbar_empty[C::STAGES];
    __shared__ uint64_t mbar_done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    const int KB = kend > kbeg ? (kend - kbeg + BK - 1) / BK : 0;
    // CSPLIT: [BM x BN] fp32 landing zone of rank 1's accumulator, behind the pipeline stages (BN = 64: 32 KB; measured against
    // reusing the idle stages behind a second cluster barrier: the extra barrier costs 0.5 us per launch)
    const uint32_t part_base = sbase + C::STAGES * C::STAGE_BYTES;
    static_assert(!CSPLIT || BN == 64, "cluster split-K is instantiated for 64-wide tiles only");
    const bool second_half = CSPLIT && blockIdx.z == 1;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&mbar_full[s], PRODUCERS); mbar_init(&mbar_empty[s], 1); }
        mbar_init(&mbar_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp < 8) {
        // ------------------------------ producers ------------------------------
        float4 ra[4][C::LA], rw[4][C::LW];       // 4 register sets: loads run 3 k-blocks ahead of the staging
        // per-thread constants: chunk f = tid + i*256 -> row = (tid>>3) + 32 i, 16-byte chunk c = tid & 7, so the
        // swizzle term (row & 7) and the chunk are fixed per thread; only +i*32 rows / +k0 move
        const int srow = tid >> 3, sc = tid & 7;
        const uint32_t soff = srow * 128 + ((sc ^ (srow & 7)) << 4);
        const float* pA = A + (size_t)(m0 + srow) * lda + sc * 4;
        const float* pW = W + (size_t)(n0 + srow) * ldw + sc * 4;
        const size_t strA = (size_t)32 * lda, strW = (size_t)32 * ldw;
        uint32_t okA = 0, okW = 0;
#pragma unroll
        for (int i = 0; i < C::LA; ++i) okA |= (m0 + srow + 32 * i < M ? 1u : 0u) << i;
#pragma unroll
        for (int i = 0; i < C::LW; ++i) okW |= (n0 + srow + 32 * i < N ? 1u : 0u) << i;
        auto gload = [&](int kb, float4 (&ra_)[C::LA], float4 (&rw_)[C::LW]) {
            const int k0 = kbeg + kb * BK;
            const bool kin = k0 + sc * 4 < kend;
#pragma unroll
            for (int i = 0; i < C::LA; ++i)
                ra_[i] = (kin && ((okA >> i) & 1u)) ? *reinterpret_cast<const float4*>(pA + i * strA + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < C::LW; ++i)
                rw_[i] = (kin && ((okW >> i) & 1u)) ? *reinterpret_cast<const float4*>(pW + i * strW + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        auto stage_in = [&](int kb, const float4 (&ra_)[C::LA], const float4 (&rw_)[C::LW]) {
            const int s = kb % C::STAGES;
            if (kb >= C::STAGES) mbar_wait(&mbar_empty[s], ((kb / C::STAGES) - 1) & 1);   // MMAs reading this stage are done
            const uint32_t sA = sbase + s * C::STAGE_BYTES + soff, sW = sA + C::A_BYTES;
            const uint32_t sAl = sW + C::W_BYTES, sWl = sAl + C::A_BYTES;
#pragma unroll
            for (int i = 0; i < C::LA; ++i) {
                if (SPLIT3) { float4 hi, lo; split4(ra_[i], hi, lo); sts4(sA + i * 4096, hi); sts4(sAl + i * 4096, lo); }
                else sts4(sA + i * 4096, ra_[i]);
            }
#pragma unroll
            for (int i = 0; i < C::LW; ++i) {
                if (SPLIT3) { float4 hi, lo; split4(rw_[i], hi, lo); sts4(sW + i * 4096, hi); sts4(sWl + i * 4096, lo); }
                else sts4(sW + i * 4096, rw_[i]);
            }
            fence_async_smem();           // generic-proxy writes -> visible to the tensor-core (async) proxy
            mbar_arrive(&mbar_full[s]);
        };
#pragma unroll
        for (int u = 0; u < 3; ++u) if (u < KB) gload(u, ra[u], rw[u]);
        for (int kb = 0; kb < KB; kb += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (kb + u < KB) {
                    if (kb + u + 3 < KB) gload(kb + u + 3, ra[(u + 3) & 3], rw[(u + 3) & 3]);
                    stage_in(kb + u, ra[u], rw[u]);
                }
            }
        }
    } else if (lane == 0) {
        // ------------------------------ MMA issuer ------------------------------
        constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % C::STAGES;
            mbar_wait(&mbar_full[s], (kb / C::STAGES) & 1);
            tc_fence_after();
            const uint32_t sA = sbase + s * C::STAGE_BYTES, sW = sA + C::A_BYTES;
            const uint32_t sAl = sW + C::W_BYTES, sWl = sAl + C::A_BYTES;
            const uint64_t dA = make_sdesc(sA), dW = make_sdesc(sW);
            uint32_t acc = kb > 0 ? 1u : 0u;
            if (SPLIT3) {
                const uint64_t dAl = make_sdesc(sAl), dWl = make_sdesc(sWl);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) { umma_tf32(tmem, dAl + 2 * k, dW + 2 * k, idesc, acc); acc = 1u; }
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) umma_tf32(tmem, dA + 2 * k, dWl + 2 * k, idesc, 1u);
            }
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) { umma_tf32(tmem, dA + 2 * k, dW + 2 * k, idesc, acc); acc = 1u; }
            umma_commit(&mbar_empty[s]);
            if (kb == KB - 1) umma_commit(&mbar_done);
        }
    }

    // ---- epilogue (warps 0..7): TMEM -> registers -> per-warp smem transpose -> fused functor ----
    // tcgen05.ld hands each thread one accumulator ROW (lane = row); calling the functor like that makes
    // every global access touch 32 different lines.  Each warp therefore bounces its 32 rows x 32 columns
    // through a private 4 KB swizzled staging tile (the pipeline stages are free by now) and calls the
    // functor with 8 lanes per row: 4 rows x 128 contiguous bytes per instruction.
    if (warp < 8) {
        const uint32_t stg = sbase + warp * 4096;
        const int rbase = m0 + (warp & 3) * 32;
        const int cbeg = (warp >> 2) * (BN / 2);
        const int rr0 = lane >> 3, qq = lane & 7;
        if constexpr (epi_has_pre<Epi>::value) {
            constexpr int SR = epi_pre_rows<Epi>::value, NB = 8 / SR;
            typename Epi::Pre pf[2][SR];
            auto request = [&](typename Epi::Pre (&set)[SR], int c0, int j) {
                const int n = n0 + c0 + 4 * qq;
#pragma unroll
                for (int i = 0; i < SR; ++i) { const int m = rbase + rr0 + 4 * (SR * j + i); if (m < M && n < N) set[i] = epi.pre(m, n); }
            };
            if (!second_half) request(pf[0], cbeg, 0);
            if (KB > 0) { mbar_wait(&mbar_done, 0); tc_fence_after(); }
            if constexpr (CSPLIT) {
                if (second_half) {
                    const uint32_t dst = tcx_mapa(part_base, 0);
                    int ci = 0;
#pragma unroll 1
                    for (int c0 = cbeg; c0 < cbeg + BN / 2; c0 += 32, ++ci) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float v[16];
                            if (KB > 0) tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(c0 + 16 * h), v);
                            else {
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] = 0.f;
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                tcx_st_cluster4(dst + (uint32_t)(((((warp * (BN / 64) + ci) * 2 + h) * 4 + q) * 32 + lane) * 16),
                                                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
                        }
                    }
                }
                tcx_cluster_sync();
            }
            int ci = 0;
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + BN / 2 && !second_half; c0 += 32, ++ci) {
                float v[16];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (KB > 0) tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(c0 + 16 * h), v);
                    else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = 0.f;
                    }
                    if constexpr (CSPLIT) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float4 pp;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pp.x), "=f"(pp.y), "=f"(pp.z), "=f"(pp.w)
                                         : "r"(part_base + (uint32_t)(((((warp * (BN / 64) + ci) * 2 + h) * 4 + q) * 32 + lane) * 16)));
                            v[4 * q] += pp.x; v[4 * q + 1] += pp.y; v[4 * q + 2] += pp.z; v[4 * q + 3] += pp.w;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        sts4(stg + lane * 128 + (((4 * h + q) ^ (lane & 7)) << 4), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
                }
                __syncwarp();
                const int n = n0 + c0 + 4 * qq;
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    if (j + 1 < NB) request(pf[(j + 1) & 1], c0, j + 1);
                    else if (c0 + 32 < cbeg + BN / 2) request(pf[0], c0 + 32, 0);
#pragma unroll
                    for (int i = 0; i < SR; ++i) {
                        const int rr = rr0 + 4 * (SR * j + i);
                        float4 x;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                                     : "r"(stg + rr * 128 + ((qq ^ (rr & 7)) << 4)));
                        const int m = rbase + rr;
                        if (m < M && n < N) { float q4[4] = {x.x, x.y, x.z, x.w}; epi(m, n, q4, pf[j & 1][i]); }
                    }
                }
                __syncwarp();
            }
        } else {
        if (KB > 0) { mbar_wait(&mbar_done, 0); tc_fence_after(); }
#pragma unroll 1
        for (int c0 = cbeg; c0 < cbeg + BN / 2; c0 += 32) {
            float v[16];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (KB > 0) tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(c0 + 16 * h), v);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    sts4(stg + lane * 128 + (((4 * h + q) ^ (lane & 7)) << 4), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = rr0 + 4 * i;
                float4 x;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                             : "r"(stg + rr * 128 + ((qq ^ (rr & 7)) << 4)));
                const int m = rbase + rr, n = n0 + c0 + 4 * qq;
                if (m < M && n < N) { float q4[4] = {x.x, x.y, x.z, x.w}; epi(m, n, q4); }
            }
            __syncwarp();
        }
        }
    }
    if constexpr (CSPLIT) { if (warp >= 8) tcx_cluster_sync(); }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, BN);
}

// ---------------------------------------------------------------------------------------------
// Persistent variant for the large GEMMs: grid = #SMs, every CTA walks tiles t = blockIdx.x, +gridDim.x, ...
// with TWO TMEM accumulators, so the epilogue of tile i (4 dedicated warps) overlaps the main loop of
// tile i+1 (8 producer warps + the MMA thread never stop).  Tile order keeps the n-blocks of one m-block
// adjacent (the A tile is re-read from L2, not HBM).
// ---------------------------------------------------------------------------------------------
constexpr int P_THREADS = PRODUCERS + 32 + 128;     // 8 producer warps, MMA warp, 4 epilogue warps

template <int BN, bool SPLIT3, class Epi>
__global__ void __launch_bounds__(P_THREADS, 1)
k_gemm_tn_tc_persist(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                     int M, int N, int K, int k_per_split, int tiles_m, int tiles_n, int splits, Epi epi) {
    using C = Cfg<BN, SPLIT3>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mbar_full[C::STAGES], mbar_empty[C::STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stg_base = sbase + C::STAGES * C::STAGE_BYTES;      // 4 x 4 KB epilogue staging
    const int total = tiles_m * tiles_n * splits;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&mbar_full[s], PRODUCERS); mbar_init(&mbar_empty[s], 1); }
        mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
        mbar_init(&acc_empty[0], 128); mbar_init(&acc_empty[1], 128);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    auto decode = [&](int t, int& m0, int& n0, int& kbeg, int& kend) {
        const int z = t / (tiles_m * tiles_n), r = t % (tiles_m * tiles_n);
        m0 = (r / tiles_n) * BM; n0 = (r % tiles_n) * BN;
        kbeg = z * k_per_split; kend = min(K, kbeg + k_per_split);
    };

    if (warp < 8) {
        // ------------------------------ producers ------------------------------
        // The (tile, k-block) sequence of this CTA is walked as ONE stream: the global loads run two k-blocks ahead of the
        // staging ACROSS tile boundaries (measured on the per-tile version: the first two k-blocks of every tile waited a
        // full memory round trip with the tensor pipe drained -- a fifth of the producers' time at K = 256).
        float4 ra[3][C::LA], rw[3][C::LW];
        const int srow = tid >> 3, sc = tid & 7;
        const uint32_t soff = srow * 128 + ((sc ^ (srow & 7)) << 4);
        const size_t strA = (size_t)32 * lda, strW = (size_t)32 * ldw;
        int64_t gkb = 0;
        // load cursor
        int lt = blockIdx.x, lkb = 0, lKB = 0, lkbeg = 0, lkend = 0;
        const float* pA = A; const float* pW = W;
        uint32_t okA = 0, okW = 0;
        auto open_tile = [&]() {
            if (lt >= total) return;
            int m0, n0; decode(lt, m0, n0, lkbeg, lkend);
            lKB = (lkend - lkbeg + BK - 1) / BK; lkb = 0;
            pA = A + (size_t)(m0 + srow) * lda + sc * 4;
            pW = W + (size_t)(n0 + srow) * ldw + sc * 4;
            okA = 0; okW = 0;
#pragma unroll
            for (int i = 0; i < C::LA; ++i) okA |= (m0 + srow + 32 * i < M ? 1u : 0u) << i;
#pragma unroll
            for (int i = 0; i < C::LW; ++i) okW |= (n0 + srow + 32 * i < N ? 1u : 0u) << i;
        };
        open_tile();
        auto load_next = [&](float4 (&ra_)[C::LA], float4 (&rw_)[C::LW]) -> bool {
            if (lt >= total) return false;
            const int k0 = lkbeg + lkb * BK;
            const bool kin = k0 + sc * 4 < lkend;
#pragma unroll
            for (int i = 0; i < C::LA; ++i)
                ra_[i] = (kin && ((okA >> i) & 1u)) ? *reinterpret_cast<const float4*>(pA + i * strA + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < C::LW; ++i)
                rw_[i] = (kin && ((okW >> i) & 1u)) ? *reinterpret_cast<const float4*>(pW + i * strW + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (++lkb == lKB) { lt += gridDim.x; open_tile(); }
            return true;
        };
        auto stage_in = [&](const float4 (&ra_)[C::LA], const float4 (&rw_)[C::LW]) {
            const int s = (int)(gkb % C::STAGES);
            if (gkb >= C::STAGES) mbar_wait(&mbar_empty[s], (uint32_t)((gkb / C::STAGES) - 1) & 1);
            const uint32_t sA = sbase + s * C::STAGE_BYTES + soff, sW = sA + C::A_BYTES;
            const uint32_t sAl = sW + C::W_BYTES, sWl = sAl + C::A_BYTES;
#pragma unroll
            for (int i = 0; i < C::LA; ++i) {
                if (SPLIT3) { float4 hi, lo; split4(ra_[i], hi, lo); sts4(sA + i * 4096, hi); sts4(sAl + i * 4096, lo); }
                else sts4(sA + i * 4096, ra_[i]);
            }
#pragma unroll
            for (int i = 0; i < C::LW; ++i) {
                if (SPLIT3) { float4 hi, lo; split4(rw_[i], hi, lo); sts4(sW + i * 4096, hi); sts4(sWl + i * 4096, lo); }
                else sts4(sW + i * 4096, rw_[i]);
            }
            fence_async_smem();
            mbar_arrive(&mbar_full[s]);
            ++gkb;
        };
        bool v0 = load_next(ra[0], rw[0]), v1 = load_next(ra[1], rw[1]), v2 = false;
        for (;;) {
            v2 = load_next(ra[2], rw[2]); if (!v0) break; stage_in(ra[0], rw[0]);
            v0 = load_next(ra[0], rw[0]); if (!v1) break; stage_in(ra[1], rw[1]);
            v1 = load_next(ra[1], rw[1]); if (!v2) break; stage_in(ra[2], rw[2]);
        }
    } else if (warp == 8) {
        if (lane == 0) {
            // ------------------------------ MMA issuer ------------------------------
            constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
            int64_t gkb = 0; int it = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
                int m0, n0, kbeg, kend; decode(t, m0, n0, kbeg, kend);
                const int KB = (kend - kbeg + BK - 1) / BK;
                const int buf = it & 1;
                if (it >= 2) { mbar_wait(&acc_empty[buf], (uint32_t)((it >> 1) - 1) & 1); tc_fence_after(); }
                const uint32_t d_tmem = tmem + buf * BN;
                for (int kb = 0; kb < KB; ++kb, ++gkb) {
                    const int s = (int)(gkb % C::STAGES);
                    mbar_wait(&mbar_full[s], (uint32_t)(gkb / C::STAGES) & 1);
                    tc_fence_after();
                    const uint32_t sA = sbase + s * C::STAGE_BYTES, sW = sA + C::A_BYTES;
                    const uint32_t sAl = sW + C::W_BYTES, sWl = sAl + C::A_BYTES;
                    const uint64_t dA = make_sdesc(sA), dW = make_sdesc(sW);
                    uint32_t acc = kb > 0 ? 1u : 0u;
                    if (SPLIT3) {
                        const uint64_t dAl = make_sdesc(sAl), dWl = make_sdesc(sWl);
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) { umma_tf32(d_tmem, dAl + 2 * k, dW + 2 * k, idesc, acc); acc = 1u; }
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) umma_tf32(d_tmem, dA + 2 * k, dWl + 2 * k, idesc, 1u);
                    }
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) { umma_tf32(d_tmem, dA + 2 * k, dW + 2 * k, idesc, acc); acc = 1u; }
                    umma_commit(&mbar_empty[s]);
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else {
        // ------------------------------ epilogue (warps 9..12; TMEM lane quadrant = warp % 4) ------------------------------
        const int q = warp & 3;
        const uint32_t stg = stg_base + q * 4096;
        const int rr0 = lane >> 3, qq = lane & 7;
        int it = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
            int m0, n0, kbeg, kend; decode(t, m0, n0, kbeg, kend);
            const int buf = it & 1;
            const int rbase = m0 + q * 32;
            if constexpr (epi_has_pre<Epi>::value) {
                // the functor's inputs of the first rows are requested while the MMAs of this tile are still running
                constexpr int SR = epi_pre_rows<Epi>::value, NB = 8 / SR;
                typename Epi::Pre pf[2][SR];
                auto request = [&](typename Epi::Pre (&set)[SR], int c0, int j) {
                    const int n = n0 + c0 + 4 * qq;
#pragma unroll
                    for (int i = 0; i < SR; ++i) { const int m = rbase + rr0 + 4 * (SR * j + i); if (m < M && n < N) set[i] = epi.pre(m, n); }
                };
                request(pf[0], 0, 0);
                mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    float v[16];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0 + 16 * h), v);
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd)
                            sts4(stg + lane * 128 + (((4 * h + qd) ^ (lane & 7)) << 4), make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]));
                    }
                    __syncwarp();
                    const int n = n0 + c0 + 4 * qq;
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        if (j + 1 < NB) request(pf[(j + 1) & 1], c0, j + 1);
                        else if (c0 + 32 < BN) request(pf[0], c0 + 32, 0);
#pragma unroll
                        for (int i = 0; i < SR; ++i) {
                            const int rr = rr0 + 4 * (SR * j + i);
                            float4 x;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                                         : "r"(stg + rr * 128 + ((qq ^ (rr & 7)) << 4)));
                            const int m = rbase + rr;
                            if (m < M && n < N) { float q4[4] = {x.x, x.y, x.z, x.w}; epi(m, n, q4, pf[j & 1][i]); }
                        }
                    }
                    __syncwarp();
                }
            } else {
            mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                float v[16];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0 + 16 * h), v);
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd)
                        sts4(stg + lane * 128 + (((4 * h + qd) ^ (lane & 7)) << 4), make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]));
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = rr0 + 4 * i;
                    float4 x;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                                 : "r"(stg + rr * 128 + ((qq ^ (rr & 7)) << 4)));
                    const int m = rbase + rr, n = n0 + c0 + 4 * qq;
                    if (m < M && n < N) { float q4[4] = {x.x, x.y, x.z, x.w}; epi(m, n, q4); }
                }
                __syncwarp();
            }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);          // accumulator drained: the MMA thread may overwrite it
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 2 * BN);
}

// ---------------------------------------------------------------------------------------------
// Weight-gradient GEMM without transposes:  C[i,j] = sum_m A[m,i] * Bm[m,j]  (reduction over the rows).
// Both operands are MN-major for the UMMA.  For 32-bit (tf32) MN-major operands the only shared-memory layout the
// tensor core accepts is SWIZZLE_128B_BASE32B (layout type 1; cute::UMMA::Layout_MN_SW128_32B_Atom): atoms of
// 4 k-rows x 128 B (32 columns) in which the 32-byte unit index is XORed with (k-row & 3).  A stage holds 32 rows m
// (the K dimension) x 128 columns, stored as [MN block of 32 columns][group of 4 rows][4 rows x 128 B]:
// SBO = 512 B (next 4 rows), LBO = 4 KB (next 32 columns); one UMMA (K = 8) consumes two row groups = 1 KB.
// Global rows are read as they lie (512 contiguous bytes per warp), no transposed copies.
// blockIdx.z = split over m; partials reduced in split order by k_reduce_update / k_reduce_only.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_sdesc_mn(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)(4096 >> 4) << 16;      // LBO: next MN block (32 columns)
    d |= (uint64_t)(512 >> 4) << 32;       // SBO: next group of 4 k-rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                // SWIZZLE_128B_BASE32B
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int M, int N) {
    return make_idesc_tf32(M, N) | (1u << 15) | (1u << 16);     // a_major = b_major = MN
}

template <bool SPLIT3>
__global__ void __launch_bounds__(THREADS, 1)
k_gemm_atb_tc(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
              int64_t Mrows, int N1, int N2, int64_t m_per_split, float* __restrict__ part,
              float* __restrict__ colsum_part) {
    // colsum_part != NULL: the CTAs of the first column tile (blockIdx.x == 0) also emit the column sums of A over
    // their m range, colsum_part[blockIdx.z][i] (the bias gradients: no separate pass over A)
    constexpr int BN = 128;
    constexpr int STAGES = SPLIT3 ? 3 : 4;
    __shared__ float4 cs_sh[8][32];
    constexpr int T_BYTES = BK * 128 * 4;                       // one operand tile: 32 rows x 128 columns
    constexpr int STAGE_BYTES = 2 * T_BYTES * (SPLIT3 ? 2 : 1);
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mbar_full[STAGES], mbar_empty[STAGES], mbar_done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int64_t mbeg = (int64_t)blockIdx.z * m_per_split;
    const int64_t mend = mbeg + m_per_split < Mrows ? mbeg + m_per_split : Mrows;
    const int KB = mend > mbeg ? (int)((mend - mbeg + BK - 1) / BK) : 0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&mbar_full[s], PRODUCERS); mbar_init(&mbar_empty[s], 1); }
        mbar_init(&mbar_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp < 8) {
        float4 ra[3][4], rb[3][4];
        // chunk f = tid + 256 i : k-row kr = (tid >> 5) + 8 i, float4 column c = tid & 31 (fixed per thread)
        const int kr0 = tid >> 5, cq = tid & 31;
        // k-row kr = kr0 + 8 i: row group kr >> 2 = (kr0 >> 2) + 2 i, row in group kr0 & 3; 32-byte unit ((cq & 7) >> 1) ^ (kr0 & 3)
        const uint32_t soff = (cq >> 3) * 4096 + (kr0 >> 2) * 512 + (kr0 & 3) * 128 +
                              (((((cq & 7) >> 1) ^ (kr0 & 3)) << 5) | ((cq & 1) << 4));       // + i * 1024
        const bool okA = i0 + 4 * cq < N1, okB = j0 + 4 * cq < N2;
        const bool do_cs = colsum_part != nullptr && blockIdx.x == 0;
        float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* pA = A + (size_t)kr0 * lda + i0 + 4 * cq;
        const float* pB = Bm + (size_t)kr0 * ldb + j0 + 4 * cq;
        auto gload = [&](int kb, float4 (&ra_)[4], float4 (&rb_)[4]) {
            const int64_t mk = mbeg + (int64_t)kb * BK;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t m = mk + kr0 + 8 * i;
                ra_[i] = (okA && m < mend) ? *reinterpret_cast<const float4*>(pA + (size_t)(mk + 8 * i) * lda) : make_float4(0.f, 0.f, 0.f, 0.f);
                rb_[i] = (okB && m < mend) ? *reinterpret_cast<const float4*>(pB + (size_t)(mk + 8 * i) * ldb) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto stage_in = [&](int kb, const float4 (&ra_)[4], const float4 (&rb_)[4]) {
            const int s = kb % STAGES;
            if (kb >= STAGES) mbar_wait(&mbar_empty[s], ((kb / STAGES) - 1) & 1);
            const uint32_t sA = sbase + s * STAGE_BYTES + soff, sB = sA + T_BYTES;
            const uint32_t sAl = sB + T_BYTES, sBl = sAl + T_BYTES;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cs.x += ra_[i].x; cs.y += ra_[i].y; cs.z += ra_[i].z; cs.w += ra_[i].w;      // rows kr0, kr0+8, ... in order
                if (SPLIT3) {
                    float4 hi, lo;
                    split4(ra_[i], hi, lo); sts4(sA + i * 1024, hi); sts4(sAl + i * 1024, lo);
                    split4(rb_[i], hi, lo); sts4(sB + i * 1024, hi); sts4(sBl + i * 1024, lo);
                } else { sts4(sA + i * 1024, ra_[i]); sts4(sB + i * 1024, rb_[i]); }
            }
            fence_async_smem();
            mbar_arrive(&mbar_full[s]);
        };
#pragma unroll
        for (int u = 0; u < 2; ++u) if (u < KB) gload(u, ra[u], rb[u]);
        for (int kb = 0; kb < KB; kb += 3) {
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                if (kb + u < KB) {
                    if (kb + u + 2 < KB) gload(kb + u + 2, ra[(u + 2) % 3], rb[(u + 2) % 3]);
                    stage_in(kb + u, ra[u], rb[u]);
                }
            }
        }
        if (do_cs) {        // fixed-order sum over the 8 row phases
            cs_sh[kr0][cq] = cs;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (kr0 == 0 && okA) {
                float4 t = cs_sh[0][cq];
#pragma unroll
                for (int r = 1; r < 8; ++r) { const float4 u = cs_sh[r][cq]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
                *reinterpret_cast<float4*>(colsum_part + (size_t)blockIdx.z * N1 + i0 + 4 * cq) = t;
            }
        }
    } else if (lane == 0) {
        constexpr uint32_t idesc = make_idesc_tf32_mn(BM, BN);
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % STAGES;
            mbar_wait(&mbar_full[s], (kb / STAGES) & 1);
            tc_fence_after();
            const uint32_t sA = sbase + s * STAGE_BYTES, sB = sA + T_BYTES, sAl = sB + T_BYTES, sBl = sAl + T_BYTES;
            const uint64_t dA = make_sdesc_mn(sA), dB = make_sdesc_mn(sB);
            uint32_t acc = kb > 0 ? 1u : 0u;
            // one UMMA (K = 8) consumes two groups of 4 k-rows: advance the start address by 1 KB (64 units)
            if (SPLIT3) {
                const uint64_t dAl = make_sdesc_mn(sAl), dBl = make_sdesc_mn(sBl);
#pragma unroll
                for (int k = 0; k < 4; ++k) { umma_tf32(tmem, dAl + 64 * k, dB + 64 * k, idesc, acc); acc = 1u; }
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(tmem, dA + 64 * k, dBl + 64 * k, idesc, 1u);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { umma_tf32(tmem, dA + 64 * k, dB + 64 * k, idesc, acc); acc = 1u; }
            umma_commit(&mbar_empty[s]);
            if (kb == KB - 1) umma_commit(&mbar_done);
        }
    }
    if (warp < 8) {
        if (KB > 0) { mbar_wait(&mbar_done, 0); tc_fence_after(); }
        const uint32_t stg = sbase + warp * 4096;
        const int rbase = i0 + (warp & 3) * 32;
        const int cbeg = (warp >> 2) * (BN / 2);
        const int rr0 = lane >> 3, qq = lane & 7;
        float* P = part + (size_t)blockIdx.z * N1 * N2;
#pragma unroll 1
        for (int c0 = cbeg; c0 < cbeg + BN / 2; c0 += 32) {
            float v[16];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (KB > 0) tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(c0 + 16 * h), v);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    sts4(stg + lane * 128 + (((4 * h + q) ^ (lane & 7)) << 4), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = rr0 + 4 * i;
                float4 x;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                             : "r"(stg + rr * 128 + ((qq ^ (rr & 7)) << 4)));
                const int m = rbase + rr, n = j0 + c0 + 4 * qq;
                if (m < N1 && n < N2) *reinterpret_cast<float4*>(P + (size_t)m * N2 + n) = x;
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, BN);
}

}  // namespace tc

static inline bool tc_gemm_supported(int64_t M, int N, int K, int lda, int ldw) {
    return K >= 32 && (K % 4) == 0 && (lda % 4) == 0 && (ldw % 4) == 0 && M >= 32 && N >= 16 && M <= 0x7fffffffLL;
}

template <int BN, bool SPLIT3, class Epi>
static int launch_tc_inst(poi_engine* e, const float* A, int lda, const float* W, int ldw, int64_t M, int N, int K,
                          const Epi& epi, int splits = 1, int k_per_split = 0) {
    using C = tc::Cfg<BN, SPLIT3>;
    static bool attr_set = false;
    if (!attr_set) {
        POI_CK(e, cudaFuncSetAttribute(tc::k_gemm_tn_tc<BN, SPLIT3, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    if (splits <= 1) { splits = 1; k_per_split = K; }
    dim3 grid((unsigned)poi_cdiv(N, BN), (unsigned)poi_cdiv(M, tc::BM), (unsigned)splits);
    POI_LAUNCH(e, (tc::k_gemm_tn_tc<BN, SPLIT3, Epi>), grid, tc::THREADS, C::SMEM_BYTES, A, lda, W, ldw, (int)M, N, K,
               k_per_split, epi);
    return 0;
}

// the two K halves of every tile as a cluster of two CTAs (see k_gemm_tn_tc, CSPLIT)
template <int BN, bool SPLIT3, class Epi>
static int launch_tc_csplit(poi_engine* e, const float* A, int lda, const float* W, int ldw, int64_t M, int N, int K, const Epi& epi) {
    using C = tc::Cfg<BN, SPLIT3>;
    const int smem = C::SMEM_BYTES + tc::BM * BN * 4;
    static bool attr_set = false;
    if (!attr_set) {
        POI_CK(e, cudaFuncSetAttribute(tc::k_gemm_tn_tc<BN, SPLIT3, Epi, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    const int k_per_split = (int)poi_cdiv(poi_cdiv(K, 2), tc::BK) * tc::BK;
    ProfRec* pr = e->kprof ? prof_begin(e) : nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)poi_cdiv(N, BN), (unsigned)poi_cdiv(M, tc::BM), 2);
    cfg.blockDim = dim3(tc::THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = e->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 2;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t st = cudaLaunchKernelEx(&cfg, tc::k_gemm_tn_tc<BN, SPLIT3, Epi, true>, A, lda, W, ldw, (int)M, N, K, k_per_split, epi);
    if (pr) cudaEventRecord(pr->b, e->stream);
    e->cur_flops = 0.0; e->cur_bytes = 0.0;
    e->launches++;
    if (st == cudaSuccess) st = cudaPeekAtLastError();
    if (st != cudaSuccess) POI_FAIL(e, "launch k_gemm_tn_tc (cluster split-K) failed: %s", cudaGetErrorString(st));
    return 0;
}

template <int BN, bool SPLIT3, class Epi>
static int launch_tc_persist(poi_engine* e, const float* A, int lda, const float* W, int ldw, int64_t M, int N, int K,
                             const Epi& epi, int splits, int k_per_split) {
    using C = tc::Cfg<BN, SPLIT3>;
    const int smem = C::STAGES * C::STAGE_BYTES + 4 * 4096 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        POI_CK(e, cudaFuncSetAttribute(tc::k_gemm_tn_tc_persist<BN, SPLIT3, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    if (splits <= 1) { splits = 1; k_per_split = K; }
    const int tiles_m = (int)poi_cdiv(M, tc::BM), tiles_n = (int)poi_cdiv(N, BN);
    const int64_t total = (int64_t)tiles_m * tiles_n * splits;
    unsigned grid = (unsigned)std::min<int64_t>(total, e->num_sms);
    POI_LAUNCH(e, (tc::k_gemm_tn_tc_persist<BN, SPLIT3, Epi>), grid, tc::P_THREADS, smem, A, lda, W, ldw, (int)M, N, K,
               k_per_split, tiles_m, tiles_n, splits, epi);
    return 0;
}

template <class Epi>
static int launch_gemm_tn_tc(poi_engine* e, const float* A, int lda, const float* W, int ldw, int64_t M, int N, int K,
                             const Epi& epi, bool split3) {
    // achieved = algorithmic flops (3xTF32 issues three MMAs per product; counted once)
    POI_CAT(e, CAT_GEMM, 2.0 * (double)M * N * K, 0);
    const int64_t tiles128 = poi_cdiv(M, tc::BM) * poi_cdiv(N, 128);
    if (e->persistent_gemm && tiles128 >= 2 * e->num_sms) {       // many tiles: persistent CTAs with overlapped epilogue
        if (split3) return launch_tc_persist<128, true>(e, A, lda, W, ldw, M, N, K, epi, 1, 0);
        return launch_tc_persist<128, false>(e, A, lda, W, ldw, M, N, K, epi, 1, 0);
    }
    // fewer tiles than two waves of 128-wide ones: take the width with fewer waves (a 64-wide tile costs ~0.8 of a 128-wide
    // one: same MMA instruction count, half the W staging and epilogue)
    const int64_t tiles64 = poi_cdiv(M, tc::BM) * poi_cdiv(N, 64);
    const double cost128 = (double)poi_cdiv(tiles128, e->num_sms), cost64 = 0.8 * (double)poi_cdiv(tiles64, e->num_sms);
    const bool wide = cost128 <= cost64;
    if constexpr (tc::epi_has_pre<Epi>::value) {
        // one partial wave and a long reduction: what the launch costs is the dependent MMA chain of one tile -> halve it
        // (64-wide tiles only: measured at 1024 x 1024 x 512, 128-wide tiles + split were 6 us SLOWER than 64-wide unsplit)
        if (e->gemm_csplit && split3 && K >= 256 && 2 * tiles64 <= e->num_sms)
            return launch_tc_csplit<64, true>(e, A, lda, W, ldw, M, N, K, epi);
    }
    if (split3) {
        if (wide) return launch_tc_inst<128, true>(e, A, lda, W, ldw, M, N, K, epi);
        return launch_tc_inst<64, true>(e, A, lda, W, ldw, M, N, K, epi);
    }
    if (wide) return launch_tc_inst<128, false>(e, A, lda, W, ldw, M, N, K, epi);
    return launch_tc_inst<64, false>(e, A, lda, W, ldw, M, N, K, epi);
}

// ---------------------------------------------------------------------------------------------
// weight gradients on the tensor cores:  C[i,j] = sum_m A[m,i] * Bm[m,j]
// Both operands are transposed once ([N x Mp], Mp = pad4(M), K-major for the UMMA) and the long
// reduction over m = (t, b) is split over blockIdx.z; partials are reduced in split order by the
// same k_reduce_update / k_reduce_only kernels as the FMA path.
// ---------------------------------------------------------------------------------------------
// out[c*ldo + r] = in[r*ld_in + c]; 64x64 tiles, 128-bit global accesses on both sides; Cc, ld_in, ldo % 4 == 0
__global__ void __launch_bounds__(256)
k_transpose_ld(const float* __restrict__ in, int ld_in, int64_t R, int Cc, float* __restrict__ out, int64_t ldo) {
    __shared__ float tile[64][65];
    const int64_t r0 = (int64_t)blockIdx.y * 64; const int c0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;          // 16 float4 columns x 16 rows per pass
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t r = r0 + ty + 16 * i; int c = c0 + tx * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < R && c < Cc) v = *reinterpret_cast<const float4*>(in + (size_t)r * ld_in + c);
        tile[ty + 16 * i][tx * 4 + 0] = v.x; tile[ty + 16 * i][tx * 4 + 1] = v.y;
        tile[ty + 16 * i][tx * 4 + 2] = v.z; tile[ty + 16 * i][tx * 4 + 3] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int c = c0 + ty + 16 * i; int64_t r = r0 + tx * 4;
        if (c < Cc && r < ldo)
            *reinterpret_cast<float4*>(out + (size_t)c * ldo + r) =
                make_float4(tile[tx * 4 + 0][ty + 16 * i], tile[tx * 4 + 1][ty + 16 * i],
                            tile[tx * 4 + 2][ty + 16 * i], tile[tx * 4 + 3][ty + 16 * i]);
    }
}

static int launch_transpose_ld(poi_engine* e, const float* in, int ld_in, int64_t R, int Cc, float* out, int64_t ldo) {
    dim3 grid((unsigned)poi_cdiv(Cc, 64), (unsigned)poi_cdiv(ldo, 64));
    POI_CAT(e, CAT_WGRAD, 0, 2.0 * (double)R * Cc * 4);
    POI_LAUNCH(e, k_transpose_ld, grid, 256, 0, in, ld_in, R, Cc, out, ldo);
    return 0;
}

struct EpiPartial {       // split-K partial tile: part[blockIdx.z][m][n]
    float* part; int N1, N2;
    __device__ __forceinline__ void operator()(int m, int n, const float (&v)[4]) const {
        float* c = part + ((size_t)blockIdx.z * N1 + m) * N2 + n;
        if (n + 3 < N2) st4(c, make_float4(v[0], v[1], v[2], v[3]));
        else for (int i = 0; i < 4 && n + i < N2; ++i) c[i] = v[i];
    }
};

// Split count of the tensor-core weight-gradient contractions.  The tensor core adds into its fp32 accumulator
// with truncation, not round-to-nearest, so the error of one accumulation chain grows linearly with its length
// (measured: 5291-row chains, 126976 x 384 x 256 -> 1.8e-4 of a typical entry); partials are summed in fp32 with
// round-to-nearest by k_reduce_update.  Chains are therefore capped at ATB_MAX_CHAIN rows (more, shorter splits
// than one wave needs) as long as the partial buffer stays small.
constexpr int64_t ATB_MAX_CHAIN = 2048;
constexpr int64_t ATB_MAX_PARTIAL_BYTES = (int64_t)6 << 30;     // C5 at 8192 users: 2 M rows -> ~1000 partial tiles per output tile
static inline int atb_tc_splits(int num_sms, int tiles, int64_t kblocks, int N1, int N2) {
    int64_t splits = std::max<int64_t>(1, std::min<int64_t>(num_sms / std::max(tiles, 1), kblocks));
    const int64_t want = poi_cdiv(kblocks * tc::BK, ATB_MAX_CHAIN);
    const int64_t cap = std::max<int64_t>(1, ATB_MAX_PARTIAL_BYTES / ((int64_t)N1 * N2 * 4));
    splits = std::max(splits, std::min(want, cap));
    // fill the last wave: the launch lasts ceil(tiles * splits / #SMs) CTA lifetimes either way -- more, shorter chains make
    // every one of them shorter (c2, Ui gradient: 6 tiles x 62 splits = 2.5 waves of 64 k-blocks -> 74 splits = 3.0 waves of 54)
    const int64_t waves = poi_cdiv((int64_t)tiles * splits, num_sms);
    const int64_t full = waves * num_sms / std::max(tiles, 1);
    if (full > splits && full <= cap) splits = full;
    return (int)std::min(splits, kblocks);
}

// At [N1 x Mp], Bt [N2 x Mp] already transposed (leading dimension Mp, zero padded); fills plan
static int launch_gemm_atb_tc(poi_engine* e, const float* At, const float* Bt, int64_t Mp, int N1, int N2,
                              bool split3, AtbPlan* plan) {
    const int tiles = (int)(poi_cdiv(N1, tc::BM) * poi_cdiv(N2, 128));
    int64_t kblocks = poi_cdiv(Mp, tc::BK);
    int splits = atb_tc_splits(e->num_sms, tiles, kblocks, N1, N2);
    int64_t kps = poi_cdiv(kblocks, splits) * tc::BK;
    splits = (int)poi_cdiv(Mp, kps);
    plan->splits = splits; plan->m_per_split = kps; plan->N1 = N1; plan->N2 = N2;
    POI_TRY(arena_get(e, (size_t)splits * N1 * N2, &plan->part));
    POI_CAT(e, CAT_WGRAD, 2.0 * (double)Mp * N1 * N2, 0);
    EpiPartial epi{plan->part, N1, N2};
    if (split3) return launch_tc_inst<128, true>(e, At, (int)Mp, Bt, (int)Mp, N1, N2, (int)Mp, epi, splits, (int)kps);
    return launch_tc_inst<128, false>(e, At, (int)Mp, Bt, (int)Mp, N1, N2, (int)Mp, epi, splits, (int)kps);
}

// A [M x lda] (columns i < N1 used), Bm [M x ldb] (columns j < N2 used), as they lie in memory; N1, N2 % 4 == 0,
// lda, ldb % 4 == 0, 16-byte aligned bases.  colsum != NULL: also the column sums of A (plan with N1 = 1, N2 = N1).
static int launch_gemm_atb_tc_mn(poi_engine* e, const float* A, int lda, const float* Bm, int ldb, int64_t M, int N1, int N2,
                                 bool split3, AtbPlan* plan, AtbPlan* colsum = nullptr) {
    const int tiles = (int)(poi_cdiv(N1, tc::BM) * poi_cdiv(N2, 128));
    const int64_t kblocks = poi_cdiv(M, tc::BK);
    int splits = atb_tc_splits(e->num_sms, tiles, kblocks, N1, N2);
    const int64_t mps = poi_cdiv(kblocks, splits) * tc::BK;
    splits = (int)poi_cdiv(M, mps);
    plan->splits = splits; plan->m_per_split = mps; plan->N1 = N1; plan->N2 = N2;
    POI_TRY(arena_get(e, (size_t)splits * N1 * N2, &plan->part));
    float* cs_part = nullptr;
    if (colsum) {
        colsum->splits = splits; colsum->m_per_split = mps; colsum->N1 = 1; colsum->N2 = N1;
        POI_TRY(arena_get(e, (size_t)splits * N1, &colsum->part));
        cs_part = colsum->part;
    }
    POI_CAT(e, CAT_WGRAD, 2.0 * (double)M * N1 * N2, 0);
    dim3 grid((unsigned)poi_cdiv(N2, 128), (unsigned)poi_cdiv(N1, tc::BM), (unsigned)splits);
    if (split3) {
        constexpr int smem = 3 * (2 * tc::BK * 128 * 4 * 2) + 1024;
        static bool set3 = false;
        if (!set3) { POI_CK(e, cudaFuncSetAttribute(tc::k_gemm_atb_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set3 = true; }
        POI_LAUNCH(e, (tc::k_gemm_atb_tc<true>), grid, tc::THREADS, smem, A, lda, Bm, ldb, M, N1, N2, mps, plan->part, cs_part);
    } else {
        constexpr int smem = 4 * (2 * tc::BK * 128 * 4) + 1024;
        static bool set1 = false;
        if (!set1) { POI_CK(e, cudaFuncSetAttribute(tc::k_gemm_atb_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set1 = true; }
        POI_LAUNCH(e, (tc::k_gemm_atb_tc<false>), grid, tc::THREADS, smem, A, lda, Bm, ldb, M, N1, N2, mps, plan->part, cs_part);
    }
    return 0;
}
