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
//
// Cluster split (CL = 1, 2 or 4 CTAs per 128 users): with B / 128 CTAs most SMs idle (B = 4096 -> 32 of 148) and
// the per-step chain is bound by the epilogue passes of one SM.  A thread-block cluster of CL CTAs shares the
// 128 users; CTA `cr` owns the gate columns [cr*H/CL, (cr+1)*H/CL): it multiplies against its rows of Wh only
// (N = H/CL per gate) and runs the gate math for its columns.  Its slice of the next A operand (r*h, h_t) is a
// whole number of 32-column k-blocks, i.e. one CONTIGUOUS 16 KB-per-k-block piece of the UMMA-layout A tile: the
// epilogue warps write it into the CTA's own tile (st.shared, as in the single-CTA case) and the MMA thread then
// pushes it into the same place of every peer's tile with one bulk async copy per peer and hi/lo half
// (cp.async.bulk shared::cta -> shared::cluster, distributed shared memory), completing bytes on the PEER's
// a_ready transaction barrier -- no per-thread remote stores (measured: 32-lane st.shared::cluster with a 128 B
// lane stride costs ~370 cycles per instruction).  The accumulator-ready barriers are armed by tcgen05.commit
// multicast to the whole cluster (count CL): a slice is overwritten, and sent, only when every CTA's MMA has
// finished reading the previous operand.
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
// issue-only variants: several TMEM accesses in flight, one wait for all of them
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),
          "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16_issue(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
          "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 64-byte rows (16 floats): XOR the 16-byte chunk index with bits 1..2 of the row -> lane=row accesses
// and 4-lanes-per-row accesses are both bank-conflict free
__device__ __forceinline__ uint32_t sw64(int row, int c) { return row * 64 + ((c ^ ((row >> 1) & 3)) << 4); }


// ---- thread-block-cluster primitives (distributed shared memory) ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
// accumulator-ready signal to the same barrier of every CTA in the cluster
template <int CL>
__device__ __forceinline__ void umma_commit_cl(uint64_t* bar) {
    if (CL == 1) umma_commit(bar);
    else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                      ::"r"(smem_u32(bar)), "h"((uint16_t)((1u << CL) - 1)) : "memory");
}
// this CTA's k-blocks of the next A operand -> the same offsets of the hi half of every peer CTA's A tile; bytes complete
// on the peer's a_ready[this rank] barrier.  3xTF32: the slice travels UNSPLIT (the fp32 values from `send`, half the
// bytes of hi + lo -- the DSMEM transfer is what the GEMMs wait for) and the receiver splits it in place
// (convert_incoming); 1xTF32: the hi tile itself is the slice.  One thread.
template <bool SPLIT3, int CL>
__device__ __forceinline__ void send_slice(uint32_t a_hi, uint32_t send, uint32_t slice_off, uint32_t slice_bytes,
                                           uint64_t* a_ready, int cr) {
    const uint32_t bar = smem_u32(&a_ready[cr]);
    const uint32_t src = SPLIT3 ? send : a_hi + slice_off;
#pragma unroll
    for (int i = 1; i < CL; ++i) {
        const uint32_t r = (uint32_t)((cr + i) % CL);
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(mapa_u32(a_hi + slice_off, r)), "r"(src), "r"(slice_bytes), "r"(mapa_u32(bar, r)) : "memory");
    }
}
// MMA thread, once per operand hand-over: the local slice is staged (s_done); arm one transaction barrier per peer
// and push the local slice to the peers.  The GEMM that follows starts on the LOCAL k-blocks at once and waits for
// each peer's k-blocks only when it reaches them, so the transfer overlaps the MMAs.
template <bool SPLIT3, int CL>
__device__ __forceinline__ void begin_exchange(uint64_t* s_done, uint32_t& ps, uint64_t* a_ready,
                                               uint32_t a_hi, uint32_t send, uint32_t slice_off, uint32_t slice_bytes, int cr,
                                               bool exchange) {
    mbar_wait(s_done, ps & 1); ++ps;
    if (CL > 1 && exchange) {
#pragma unroll
        for (int o = 0; o < CL; ++o)
            if (o != cr)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&a_ready[o])), "r"(slice_bytes) : "memory");
        send_slice<SPLIT3, CL>(a_hi, send, slice_off, slice_bytes, a_ready, cr);
    }
    tc_fence_after();
}
// Epilogue threads, after they have staged their own slice: as each peer's unsplit slice lands in the hi half of the A
// tile (arrival order: owner cr-1, cr-2, ...), split it in place into hi / lo and tell the MMA thread (conv_done[owner]).
// Thread (row, hf) converts columns [16 hf, 16 hf + 16) of its row in every k-block of the owner.
template <bool SPLIT3, int CL>
__device__ __forceinline__ void convert_incoming(uint64_t* a_ready, uint64_t* conv_done, uint32_t& pc, uint32_t a_hi, uint32_t a_lo,
                                                 int cr, int KBc, int row, int hf) {
    if (CL == 1 || !SPLIT3) return;
#pragma unroll
    for (int i = 1; i < CL; ++i) {
        const int o = (cr - i + CL) % CL;
        mbar_wait_cl(&a_ready[o], pc & 1);
        for (int kbl = 0; kbl < KBc; ++kbl) {
            const uint32_t base = (uint32_t)(o * KBc + kbl) * A_KB_BYTES + row * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t off = base + (((hf * 4 + q) ^ (row & 7)) << 4);
                float4 hi, lo;
                split4(lds4(a_hi + off), hi, lo);
                sts4(a_hi + off, hi); sts4(a_lo + off, lo);
            }
        }
        fence_async_smem();
        mbar_arrive(&conv_done[o]);
    }
    ++pc;
}

// send != 0: also the unsplit values into the send buffer (k-block kb - kb0 of the slice this CTA owns)
template <bool SPLIT3>
__device__ __forceinline__ void a_store16(uint32_t a_hi, uint32_t a_lo, int row, int col, const float (&v)[16],
                                          uint32_t send = 0u, int kb0 = 0) {
    const int kb = col >> 5, cc0 = (col & 31) >> 2;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t in_kb = row * 128 + (((cc0 + q) ^ (row & 7)) << 4);
        const uint32_t off = kb * A_KB_BYTES + in_kb;
        float4 x = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        if (SPLIT3) {
            float4 hi, lo; split4(x, hi, lo); sts4(a_hi + off, hi); sts4(a_lo + off, lo);
            if (send) sts4(send + (kb - kb0) * A_KB_BYTES + in_kb, x);
        } else sts4(a_hi + off, x);
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

// ---- TS mode: the A operand lives in TENSOR MEMORY (lane = row, column = k; hi at tA, lo at tA + H) ----
// Measured (tools/umma_rate.cu, B200): one tcgen05.mma.kind::tf32 M = 128, K = 8 takes 105 cycles with A in shared
// memory and 63.5 with A in tensor memory, for every N <= 128 -- at the N = 32..64 of a cluster-split recurrence the
// instruction count, not N, sets the GEMM time, so the clustered 3xTF32 kernels keep A in TMEM: the epilogue threads
// (thread = row = TMEM lane) write their slice with tcgen05.st, the peers' slices are split out of the landing buffer
// into TMEM by the same threads.
constexpr uint32_t TS_A_COL = 256;          // first TMEM column of the A operand (accumulators and stashes stay below)

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// one 32-column k-block of A (TMEM columns ta_hi.., ta_lo..) against a W tile (hi, lo) in shared memory
__device__ __forceinline__ void mma_kblock_ts(uint32_t d_tmem, uint32_t ta_hi, uint32_t ta_lo, uint32_t w_hi, uint32_t w_lo,
                                              uint32_t idesc, bool first) {
    const uint64_t dW = make_sdesc(w_hi), dWl = make_sdesc(w_lo);
    uint32_t acc = first ? 0u : 1u;
#pragma unroll
    for (int k = 0; k < 4; ++k) { umma_tf32_ts(d_tmem, ta_lo + 8 * k, dW + 2 * k, idesc, acc); acc = 1u; }
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_tf32_ts(d_tmem, ta_hi + 8 * k, dWl + 2 * k, idesc, 1u);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_tf32_ts(d_tmem, ta_hi + 8 * k, dW + 2 * k, idesc, 1u);
}
// 16 values of this thread's row (columns col..col+15 of the operand): hi / lo into TMEM (stores ISSUED, not awaited), the
// unsplit values into the send buffer (k-block kb - kb0 of the slice this CTA owns).  tl = this warp's TMEM lane offset.
__device__ __forceinline__ void a_put16_ts(uint32_t tmem, uint32_t tl, int H, int row, int col, const float (&v)[16],
                                           uint32_t send, int kb0) {
    float hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { hi[i] = tf32_rn(v[i]); lo[i] = v[i] - hi[i]; }
    tmem_st16_issue(tmem + tl + TS_A_COL + (uint32_t)col, hi);          // the caller waits (tmem_wait_st) before it signals
    tmem_st16_issue(tmem + tl + TS_A_COL + (uint32_t)(H + col), lo);
    if (send) {
        const int kb = col >> 5, cc0 = (col & 31) >> 2;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            sts4(send + (kb - kb0) * A_KB_BYTES + row * 128 + (((cc0 + q) ^ (row & 7)) << 4),
                 make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    }
}
// the unsplit values of this thread's own slice, as a_put16_ts left them in the send buffer
__device__ __forceinline__ void send_load16(uint32_t send, int kb0, int row, int col, float (&v)[16]) {
    const int kb = col >> 5, cc0 = (col & 31) >> 2;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float4 x = lds4(send + (kb - kb0) * A_KB_BYTES + row * 128 + (((cc0 + q) ^ (row & 7)) << 4));
        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
}
// TS mode, PUSH variant of the exchange: once the 8 epilogue warps have staged the own slice (unsplit) in `send`, the same
// 256 threads copy it into every peer's landing buffer with COALESCED distributed-shared-memory stores (a warp writes 512
// contiguous bytes per instruction; the first version of this kernel stored with a 128-byte lane stride and paid ~370
// cycles per instruction) and arrive (release.cluster) on the peer's a_ready[this rank].  No copy engine, no async
// proxy: the receiver reads the landing buffer with ordinary loads.  et = 0..255, the thread's index among the epilogue
// threads.
// MEASURED (c2, CL = 4): the push makes an exchange + GEMM phase ~10.6 k cycles against ~6.9 k with the bulk copies
// (step 2.78 ms vs 2.49 ms) -- distributed-shared-memory STORES are slow even when coalesced; kept for reference, off.
constexpr int POI_FUSED_PUSH = 0;           // 1: coalesced st.shared::cluster push; 0: bulk copies issued by the MMA thread
__device__ __forceinline__ void stc4(uint32_t caddr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(caddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int CL>
__device__ __forceinline__ void push_slice_ts(uint32_t send, uint32_t landing_slice, uint32_t slice_bytes, uint64_t* a_ready, int cr, int et) {
    asm volatile("bar.sync 1, 256;" ::: "memory");                       // the whole slice is in `send`
    uint32_t dst[CL], bar[CL];
#pragma unroll
    for (int i = 1; i < CL; ++i) {
        const uint32_t r = (uint32_t)((cr + i) % CL);
        dst[i] = mapa_u32(landing_slice, r); bar[i] = mapa_u32(smem_u32(&a_ready[cr]), r);
    }
    for (uint32_t off = (uint32_t)et * 16u; off < slice_bytes; off += 4096u) {
        const float4 v = lds4(send + off);
#pragma unroll
        for (int i = 1; i < CL; ++i) stc4(dst[i] + off, v);
    }
#pragma unroll
    for (int i = 1; i < CL; ++i)
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar[i]) : "memory");
}

// TS variant of convert_incoming: the peers' unsplit slices go from the landing buffer (the hi half of the smem A
// tile) into TMEM as hi / lo
template <int CL>
__device__ __forceinline__ void convert_incoming_ts(uint64_t* a_ready, uint64_t* conv_done, uint32_t& pc, uint32_t landing,
                                                    uint32_t tmem, uint32_t tl, int H, int cr, int KBc, int row, int hf) {
#pragma unroll
    for (int i = 1; i < CL; ++i) {
        const int o = (cr - i + CL) % CL;
        mbar_wait_cl(&a_ready[o], pc & 1);
        for (int kbl = 0; kbl < KBc; ++kbl) {
            const int kb = o * KBc + kbl;
            const uint32_t base = (uint32_t)kb * A_KB_BYTES + row * 128;
            float hi[16], lo[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 x = lds4(landing + base + (((hf * 4 + q) ^ (row & 7)) << 4));
                float4 h4, l4; split4(x, h4, l4);
                hi[4 * q] = h4.x; hi[4 * q + 1] = h4.y; hi[4 * q + 2] = h4.z; hi[4 * q + 3] = h4.w;
                lo[4 * q] = l4.x; lo[4 * q + 1] = l4.y; lo[4 * q + 2] = l4.z; lo[4 * q + 3] = l4.w;
            }
            const uint32_t col = (uint32_t)(kb * 32 + hf * 16);
            tmem_st16_issue(tmem + tl + TS_A_COL + col, hi);
            tmem_st16_issue(tmem + tl + TS_A_COL + (uint32_t)H + col, lo);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&conv_done[o]);
    }
    ++pc;
}

// Wh staged ONCE per call in the exact shared-memory image the UMMA wants (tile = one gate x one 32-float
// k-block: H rows x 128 B, 128B-swizzled, hi then lo for 3xTF32), so that inside the recurrence a tile is one
// bulk async copy (TMA, cp.async.bulk) issued by a single thread instead of 64 threads converting it
template <bool SPLIT3>
__global__ void k_stage_wh(const float* __restrict__ wh, int H, int CL, int merge, uint8_t* __restrict__ image) {
    // image = [cluster rank][tile t] of Hc = H / CL rows (the rows of Wh the rank owns) x 128 B, tiles in the order
    // the rank's MMA thread consumes them: GEMM1 k-blocks in arrival order (owner rank, rank-1, ...), z and r tile
    // of each; then GEMM2 (c gate) over the same k-block order.  merge: the z and r tiles of a k-block form ONE tile of
    // 2 Hc rows (hi: z rows, r rows; lo: z rows, r rows) so that GEMM1 is a single N = 2 Hc instruction per k-step.
    const int KB = H >> 5, Hc = H / CL, KBc = KB / CL;
    const int64_t n = (int64_t)3 * KB * H * 8;
    const uint32_t w_tile = Hc * 128, w_stage = w_tile * (SPLIT3 ? 2 : 1);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx & 7), rl = (int)((idx >> 3) % Hc);
        const int t = (int)((idx / (8 * Hc)) % (3 * KB)), rank = (int)(idx / ((int64_t)8 * Hc * 3 * KB));
        int gate, pos;
        if (t < 2 * KB) { pos = t >> 1; gate = t & 1; } else { pos = t - 2 * KB; gate = 2; }
        const int owner = (rank - pos / KBc + CL) % CL, kb = owner * KBc + pos % KBc;
        float4 v = *reinterpret_cast<const float4*>(wh + ((size_t)gate * H + rank * Hc + rl) * H + kb * 32 + c * 4);
        const uint32_t in_tile = rl * 128 + ((c ^ (rl & 7)) << 4);
        uint8_t* rank_base = image + (size_t)rank * 3 * KB * w_stage;
        uint8_t *dst_hi, *dst_lo;
        if (merge && SPLIT3) {
            uint8_t* tb = gate < 2 ? rank_base + (size_t)pos * 2 * w_stage : rank_base + (size_t)KB * 2 * w_stage + (size_t)pos * w_stage;
            dst_hi = tb + (gate == 1 ? w_tile : 0) + in_tile;
            dst_lo = tb + (gate < 2 ? 2 * w_tile : w_tile) + (gate == 1 ? w_tile : 0) + in_tile;
        } else {
            dst_hi = rank_base + (size_t)t * w_stage + in_tile;
            dst_lo = dst_hi + w_tile;
        }
        if (SPLIT3) { float4 hi, lo; split4(v, hi, lo); *reinterpret_cast<float4*>(dst_hi) = hi; *reinterpret_cast<float4*>(dst_lo) = lo; }
        else *reinterpret_cast<float4*>(dst_hi) = v;
    }
}

// tensor-core modes only: the gate non-linearities with the fast exp / divide (error ~1e-6, far inside the
// 1e-4 parity bar; the fp32 FMA mode keeps expf/tanhf)
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

template <bool SPLIT3, int CL>
__global__ void __launch_bounds__(F_THREADS, 1)
k_gru_fwd_fused(const float* __restrict__ AX, const uint8_t* __restrict__ wimg, float* __restrict__ Hs, float* __restrict__ Z,
                float* __restrict__ R, float* __restrict__ C, int B, int T, int H) {
    constexpr int WST = CL == 4 ? 4 : 2;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t w_full[4], w_empty[4], ax_full[2][4], ax_empty[2][4], s_done, a_ready[CL], conv_done[CL], d1_full, d2_full;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cr = CL == 1 ? 0 : (int)cluster_ctarank();      // this CTA's slice of the gate columns
    const int Hc = H / CL;                                      // columns of each gate owned by this CTA
    const int KB = H >> 5, NCH = Hc >> 4, HCH = NCH >> 1;
    const int col0 = cr * Hc;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_hi = sbase, a_lo = sbase + KB * A_KB_BYTES;
    const uint32_t w_tile = Hc * 128, w_stage = w_tile * (SPLIT3 ? 2 : 1);
    const uint32_t w_base = sbase + KB * A_KB_BYTES * (SPLIT3 ? 2 : 1);
    const int KBc = Hc >> 5;                                    // k-blocks of the operand this CTA owns
    // unsplit copy of the own slice, the source of the bulk copies to the peers (3xTF32 clusters only)
    const uint32_t send = (CL > 1 && SPLIT3) ? w_base + WST * w_stage + 2 * AX_STAGES * AX_TILE : 0u;
    // TS mode (A operand in tensor memory) frees the lo half of the smem A tile: the AX ring moves there (4 stages when
    // it fits) and the W ring takes the old AX region too, in slots of 2 * w_stage -- room for the merged z|r tiles
    const bool merge = SPLIT3 && CL > 1 && KB * A_KB_BYTES >= 4 * AX_TILE;
    const int AXS = merge && KB * A_KB_BYTES >= 8 * AX_TILE ? 4 : AX_STAGES;
    const uint32_t ax_base = merge ? a_lo : w_base + WST * w_stage;           // [half][stage] tiles of AX_TILE bytes
    const uint32_t w_slot = merge ? 2 * w_stage : w_stage;
    const int NWS = merge ? min(4, (int)((WST * w_stage + 2 * AX_STAGES * AX_TILE) / w_slot)) : WST;
    const int m0 = (blockIdx.x / CL) * FM;
    const uint8_t* wimg_c = wimg + (size_t)cr * 3 * KB * w_stage;
    constexpr bool TS = SPLIT3 && CL > 1;                       // A operand in tensor memory (see TS mode above)
    uint32_t ncols = 32; while (ncols < (uint32_t)(3 * Hc)) ncols <<= 1;
    if (TS) ncols = 512;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < 4; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int s = 0; s < 4; ++s) { mbar_init(&ax_full[h][s], 32); mbar_init(&ax_empty[h][s], 128); }
        mbar_init(&s_done, 256); mbar_init(&d1_full, CL);
#pragma unroll
        for (int o = 0; o < CL; ++o) { mbar_init(&a_ready[o], (SPLIT3 && CL > 1 && POI_FUSED_PUSH) ? 256 : 1); mbar_init(&conv_done[o], 256); }
        mbar_init(&d2_full, CL);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                             // every CTA's barriers exist before any remote arrive
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc_tf32(FM, Hc), idesc2 = make_idesc_tf32(FM, 2 * Hc);

    if (warp < 8) {
        // ================================ epilogue (8 warps) ================================
        // warp w: TMEM lane quadrant q = w & 3 (rows 32q..32q+31), column half hf = w >> 2 of this CTA's columns
        const int q = warp & 3, hf = warp >> 2;
        const int row = q * 32 + lane;                         // TMEM lane == row of the tile
        const bool ok = m0 + row < B;
        const uint32_t tl = (uint32_t)(q * 32) << 16;
        const int rows_valid = min(32, B - (m0 + q * 32));
        const int k_beg = hf * HCH, k_end = k_beg + HCH;
        {   // h_{-1} = 0: the WHOLE local A tile (every CTA zeroes its own copy, nothing to exchange)
            float zero[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) zero[i] = 0.f;
            for (int k = hf * (H >> 5); k < (hf + 1) * (H >> 5); ++k) {
                if (TS) a_put16_ts(tmem, tl, H, row, 16 * k, zero, 0u, 0);
                else a_store16<SPLIT3>(a_hi, a_lo, row, 16 * k, zero);
            }
            if (TS) for (int k = k_beg; k < k_end; ++k) a_put16_ts(tmem, tl, H, row, col0 + 16 * k, zero, send, cr * KBc);   // own h_{-1} slice, unsplit
            if (TS) tmem_wait_st();
            fence_async_smem(); tc_fence_before(); mbar_arrive(&s_done);
        }
        int64_t axi = 0;                                       // position in this half's AX ring
        uint32_t pc = 0;                                       // operand hand-overs converted so far (barrier phase)
        // wait for the next AX tile, read this thread's row, return this warp's 2 KB slice of the tile (free to
        // be reused as output staging once the whole warp has read) and the barrier to release it on
        auto ax_take = [&](float (&v)[16], uint32_t& slice, uint64_t*& rel) {
            const int s = (int)(axi % AXS);
            const long long tw0 = FTR_NOW();
            mbar_wait(&ax_full[hf][s], (uint32_t)(axi / AXS) & 1);
            if (warp == 0 && lane == 0) FTR_ADD(0, (int)(axi / (3 * HCH)), 12, FTR_NOW() - tw0);
            slice = ax_base + (hf * AXS + s) * AX_TILE + q * OUT_STG;
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
            if (CL == 1) mbar_wait(&d1_full, j & 1); else mbar_wait_cl(&d1_full, j & 1);   // EVERY CTA's GEMM1 has read its A tile
            tc_fence_after();
            if (lane == 0 && (warp == 0 || warp == 7)) FTR(0, j, warp == 0 ? 4 : 8);
            for (int k = k_beg; k < k_end; ++k) {
                const int lc0 = 16 * k, c0 = col0 + lc0;       // local (TMEM) / global (A tile, memory) column
                float a[16], b[16], hv[16], dz[16], dr[16];
                uint32_t sl_z, sl_r; uint64_t *rel_z, *rel_r;
                tmem_ld16_issue(tmem + tl + (uint32_t)lc0, dz);
                tmem_ld16_issue(tmem + tl + (uint32_t)(Hc + lc0), dr);
                ax_take(a, sl_z, rel_z); ax_take(b, sl_r, rel_r);
                if (TS) send_load16(send, cr * KBc, row, c0, hv); else a_load16<SPLIT3>(a_hi, a_lo, row, c0, hv);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float zz = sigmoid_fast(dz[i] + a[i]);
                    const float rr = sigmoid_fast(dr[i] + b[i]);
                    const float h_ = ok ? hv[i] : 0.f;
                    dz[i] = zz; dr[i] = rr;
                    a[i] = rr * h_;                 // r*h -> next A operand
                    b[i] = (1.f - zz) * h_;         // (1-z)*h_prev, used by epilogue 2
                }
                tmem_st16_issue(tmem + tl + (uint32_t)lc0, dz);             // stash z
                tmem_st16_issue(tmem + tl + (uint32_t)(Hc + lc0), b);       // stash (1-z)*h_prev
                if (TS) a_put16_ts(tmem, tl, H, row, c0, a, send, cr * KBc); else a_store16<SPLIT3>(a_hi, a_lo, row, c0, a, send, cr * KBc);
                tmem_wait_st();
                if (k == k_end - 1) {       // the operand slice is complete: hand it over BEFORE the stores to global memory
                    if (lane == 0 && (warp == 0 || warp == 7)) FTR(0, j, warp == 0 ? 5 : 9);
                    fence_async_smem(); tc_fence_before(); mbar_arrive(&s_done);
                }
                warp_store_chunk(sl_z, lane, dz, Z + wrow * H + c0, H, rows_valid);
                warp_store_chunk(sl_r, lane, dr, R + wrow * H + c0, H, rows_valid);
                mbar_arrive(rel_z); mbar_arrive(rel_r);
            }
            if (TS && POI_FUSED_PUSH) push_slice_ts<CL>(send, a_hi + (uint32_t)cr * KBc * A_KB_BYTES, (uint32_t)KBc * A_KB_BYTES, a_ready, cr, (int)threadIdx.x);
            if (TS) convert_incoming_ts<CL>(a_ready, conv_done, pc, a_hi, tmem, tl, H, cr, KBc, row, hf);
            else convert_incoming<SPLIT3, CL>(a_ready, conv_done, pc, a_hi, a_lo, cr, KBc, row, hf);     // the peers' r*h slices
            // ---- epilogue 2: c, h_t ----
            if (CL == 1) mbar_wait(&d2_full, j & 1); else mbar_wait_cl(&d2_full, j & 1);
            tc_fence_after();
            if (lane == 0 && (warp == 0 || warp == 7)) FTR(0, j, warp == 0 ? 6 : 10);
            for (int k = k_beg; k < k_end; ++k) {
                const int lc0 = 16 * k, c0 = col0 + lc0;
                float a[16], dc[16], zz[16], u[16];
                uint32_t sl; uint64_t* rel;
                tmem_ld16_issue(tmem + tl + (uint32_t)(2 * Hc + lc0), dc);
                tmem_ld16_issue(tmem + tl + (uint32_t)lc0, zz);
                tmem_ld16_issue(tmem + tl + (uint32_t)(Hc + lc0), u);
                ax_take(a, sl, rel);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float cc = tanh_fast(dc[i] + a[i]);
                    dc[i] = cc;
                    u[i] = ok ? u[i] + zz[i] * cc : 0.f;   // h_t = (1-z) h_prev + z c
                }
                if (TS) { a_put16_ts(tmem, tl, H, row, c0, u, send, cr * KBc); tmem_wait_st(); }
                else a_store16<SPLIT3>(a_hi, a_lo, row, c0, u, send, cr * KBc);
                if (k == k_end - 1) {       // the h_t slice is complete: hand it over before the stores to global memory
                    if (lane == 0 && (warp == 0 || warp == 7)) FTR(0, j, warp == 0 ? 7 : 11);
                    fence_async_smem(); tc_fence_before(); mbar_arrive(&s_done);
                }
                warp_store_chunk(sl, lane, dc, C + wrow * H + c0, H, rows_valid);
                warp_store_chunk(sl, lane, u, Hs + (wrow + B) * H + c0, H, rows_valid);
                mbar_arrive(rel);
            }
            if (j + 1 < T) {                                                                               // the peers' h_t slices
                if (TS && POI_FUSED_PUSH) push_slice_ts<CL>(send, a_hi + (uint32_t)cr * KBc * A_KB_BYTES, (uint32_t)KBc * A_KB_BYTES, a_ready, cr, (int)threadIdx.x);
                if (TS) convert_incoming_ts<CL>(a_ready, conv_done, pc, a_hi, tmem, tl, H, cr, KBc, row, hf);
                else convert_incoming<SPLIT3, CL>(a_ready, conv_done, pc, a_hi, a_lo, cr, KBc, row, hf);
            }
        }
    } else if (warp < 10) {
        // ================================ AX producers (one warp per column half, cp.async) ================================
        const int hf = warp - 8;
        const int per_step = 3 * HCH;
        const int64_t n_tiles = (int64_t)T * per_step;
        for (int64_t ai = 0; ai < n_tiles; ++ai) {
            const int s = (int)(ai % AXS);
            if (ai >= AXS) mbar_wait(&ax_empty[hf][s], (uint32_t)((ai / AXS) - 1) & 1);
            const int j = (int)(ai / per_step), ti = (int)(ai % per_step);
            const int gate = ti < 2 * HCH ? (ti & 1) : 2;
            const int k = hf * HCH + (ti < 2 * HCH ? (ti >> 1) : ti - 2 * HCH);
            const float* src = AX + ((size_t)j * B + m0) * 3 * H + gate * H + col0 + 16 * k;
            const uint32_t dst = ax_base + (hf * AXS + s) * AX_TILE;
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
            int64_t ws = 0; uint32_t pa = 0, ps = 0;
            const uint32_t slice_off = (uint32_t)cr * KBc * A_KB_BYTES, slice_bytes = (uint32_t)KBc * A_KB_BYTES;
            // one GEMM over K = H in ARRIVAL order: own k-blocks first, then owner cr-1, cr-2, ... (the order the
            // peers send in); `tiles_per_kb` W tiles per k-block, accumulators dcol0 + t * Hc.  The Wh image of this
            // rank is staged in exactly this order (k_stage_wh).
            // nmul = 2: the tile holds the z and the r rows (merged GEMM1 tile, N = 2 Hc)
            auto gemm = [&](uint32_t dcol0, int tiles_per_kb, int nmul, bool exchanged, int j, int slot) {
                for (int i = 0; i < CL; ++i) {
                    const int o = (cr - i + CL) % CL;
                    if (CL > 1 && i > 0 && exchanged) {
                        if (SPLIT3) mbar_wait(&conv_done[o], pa & 1); else mbar_wait_cl(&a_ready[o], pa & 1);
                        tc_fence_after();
                    }
                    for (int kbl = 0; kbl < KBc; ++kbl) {
                        const int kb = o * KBc + kbl;
                        for (int t = 0; t < tiles_per_kb; ++t, ++ws) {
                            const int s = (int)(ws % NWS);
                            const long long tw0 = FTR_NOW();
                            mbar_wait(&w_full[s], (uint32_t)(ws / NWS) & 1);
                            FTR_ADD(0, j, slot, FTR_NOW() - tw0);
                            tc_fence_after();
                            const uint32_t sW = w_base + s * w_slot;
                            if (TS) mma_kblock_ts(tmem + dcol0 + t * Hc, tmem + TS_A_COL + kb * 32, tmem + TS_A_COL + H + kb * 32, sW, sW + nmul * w_tile,
                                                  nmul == 2 ? idesc2 : idesc, i == 0 && kbl == 0);
                            else mma_kblock<SPLIT3>(tmem + dcol0 + t * Hc, a_hi + kb * A_KB_BYTES, a_lo + kb * A_KB_BYTES, sW, sW + w_tile, idesc,
                                                    i == 0 && kbl == 0);
                            umma_commit(&w_empty[s]);
                        }
                    }
                }
                if (CL > 1 && exchanged) ++pa;
            };
            for (int j = 0; j < T; ++j) {
                // h_{j-1}: own slice staged (j = 0: the zero tile, nothing to exchange)
                begin_exchange<SPLIT3, CL>(&s_done, ps, a_ready, a_hi, send, slice_off, slice_bytes, cr, j > 0 && !(TS && POI_FUSED_PUSH));
                FTR(0, j, 0);
                if (merge) gemm(0u, 1, 2, j > 0, j, 13);        // D1z | D1r in one N = 2 Hc instruction per k-step
                else gemm(0u, 2, 1, j > 0, j, 13);              // D1z, D1r
                umma_commit_cl<CL>(&d1_full);
                FTR(0, j, 1);
                begin_exchange<SPLIT3, CL>(&s_done, ps, a_ready, a_hi, send, slice_off, slice_bytes, cr, !(TS && POI_FUSED_PUSH));   // r*h
                FTR(0, j, 2);
                gemm((uint32_t)(2 * Hc), 1, 1, true, j, 14);    // D2
                umma_commit_cl<CL>(&d2_full);
                FTR(0, j, 3);
            }
        }
    } else if (lane == 0) {
        // ================================ Wh tiles: one bulk async copy (TMA) per tile ================================
        const int per_step = merge ? 2 * KB : 3 * KB;           // merged: KB z|r tiles of 2 * w_stage, then KB c tiles of w_stage
        const int64_t n_tiles = (int64_t)T * per_step;
        for (int64_t ws = 0; ws < n_tiles; ++ws) {
            const int s = (int)(ws % NWS);
            if (ws >= NWS) mbar_wait(&w_empty[s], (uint32_t)((ws / NWS) - 1) & 1);
            const int ti = (int)(ws % per_step);
            uint32_t bytes = w_stage; size_t src = (size_t)ti * w_stage;
            if (merge) {
                if (ti < KB) { bytes = 2 * w_stage; src = (size_t)ti * 2 * w_stage; }
                else src = (size_t)KB * 2 * w_stage + (size_t)(ti - KB) * w_stage;
            }
            const uint32_t bar = smem_u32(&w_full[s]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(w_base + s * w_slot), "l"(wimg_c + src), "r"(bytes), "r"(bar) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();         // no CTA leaves while a peer may still write into its A tile / barriers
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
__global__ void k_stage_wh_bwd(const float* __restrict__ wh, int H, int CL, uint8_t* __restrict__ image) {
    // image = [cluster rank][g * KB + pos] tiles of Hc = H / CL rows (the OUTPUT columns n the rank owns) x 128 B;
    // g: Wh[2] (c), Wh[0] (z), Wh[1] (r); pos: k-blocks in arrival order (owner rank, rank-1, ...)
    const int KB = H >> 5, Hc = H / CL, KBc = KB / CL;
    const int64_t n = (int64_t)3 * KB * H * 8;
    const uint32_t w_tile = Hc * 128, w_stage = w_tile * (SPLIT3 ? 2 : 1);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx & 7), rl = (int)((idx >> 3) % Hc);
        const int t = (int)((idx / (8 * Hc)) % (3 * KB)), rank = (int)(idx / ((int64_t)8 * Hc * 3 * KB));
        const int g = t / KB, pos = t % KB;
        const int gsel = g == 0 ? 2 : g - 1;
        const int owner = (rank - pos / KBc + CL) % CL, kb = owner * KBc + pos % KBc;
        // B operand row n = rank*Hc + rl, k-major: element (n, k) = Wh[gsel][k][n]
        const float* src = wh + ((size_t)gsel * H + kb * 32 + c * 4) * H + rank * Hc + rl;
        float4 v = make_float4(src[0], src[H], src[2 * (size_t)H], src[3 * (size_t)H]);
        uint8_t* dst = image + ((size_t)rank * 3 * KB + t) * w_stage + rl * 128 + ((c ^ (rl & 7)) << 4);
        if (SPLIT3) { float4 hi, lo; split4(v, hi, lo); *reinterpret_cast<float4*>(dst) = hi; *reinterpret_cast<float4*>(dst + w_tile) = lo; }
        else *reinterpret_cast<float4*>(dst) = v;
    }
}

template <bool SPLIT3, int CL>
__global__ void __launch_bounds__(F_THREADS, 1)
k_gru_bwd_fused(const float* __restrict__ DHl, const float* __restrict__ Z, const float* __restrict__ R,
                const float* __restrict__ C, const float* __restrict__ Hs, const uint8_t* __restrict__ wimg,
                float* __restrict__ DA, int B, int T, int H) {
    constexpr int WST = CL == 4 ? 4 : 2;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t w_full[WST], w_empty[WST], in_full[2][4], in_empty[2][4], s_done, a_ready[CL], conv_done[CL], dm_full, dh1_done, ddh_full;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cr = CL == 1 ? 0 : (int)cluster_ctarank();
    const int Hc = H / CL;
    const int KB = H >> 5, NCH = Hc >> 4, HCH = NCH >> 1;
    const int col0 = cr * Hc;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_hi = sbase, a_lo = sbase + KB * A_KB_BYTES;
    const uint32_t w_tile = Hc * 128, w_stage = w_tile * (SPLIT3 ? 2 : 1);
    const uint32_t w_base = sbase + KB * A_KB_BYTES * (SPLIT3 ? 2 : 1);
    const int KBc = Hc >> 5;
    const uint32_t send = (CL > 1 && SPLIT3) ? w_base + WST * w_stage + 2 * AX_STAGES * AX_TILE : 0u;
    // input ring: with the A operand in tensor memory the lo half of the smem A tile is free -- a 4-stage ring there
    // holds the whole input set of the P phase one step ahead (2 stages: the epilogue waited ~2 k cycles per step)
    const int AXS = (SPLIT3 && CL > 1 && KB * A_KB_BYTES >= 8 * AX_TILE) ? 4 : AX_STAGES;
    const uint32_t in_base = AXS == 4 ? a_lo : w_base + WST * w_stage;
    const int m0 = (blockIdx.x / CL) * FM;
    const uint8_t* wimg_c = wimg + (size_t)cr * 3 * KB * w_stage;
    constexpr bool TS = SPLIT3 && CL > 1;
    uint32_t ncols = 32; while (ncols < (uint32_t)(4 * Hc)) ncols <<= 1;
    if (TS) ncols = 512;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < WST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int s = 0; s < 4; ++s) { mbar_init(&in_full[h][s], 32); mbar_init(&in_empty[h][s], 128); }
        mbar_init(&s_done, 256); mbar_init(&dm_full, CL);
#pragma unroll
        for (int o = 0; o < CL; ++o) { mbar_init(&a_ready[o], (SPLIT3 && CL > 1 && POI_FUSED_PUSH) ? 256 : 1); mbar_init(&conv_done[o], 256); }
        mbar_init(&dh1_done, CL); mbar_init(&ddh_full, CL);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc_tf32(FM, Hc);
    const uint32_t T_M = 0, T_DH = (uint32_t)Hc, T_KEEP = (uint32_t)(2 * Hc), T_DAZ = (uint32_t)(3 * Hc);

    if (warp < 8) {
        // ================================ epilogue (8 warps) ================================
        const int q = warp & 3, hf = warp >> 2;
        const int row = q * 32 + lane;
        const bool ok = m0 + row < B;
        const uint32_t tl = (uint32_t)(q * 32) << 16;
        const int rows_valid = min(32, B - (m0 + q * 32));
        const int k_beg = hf * HCH, k_end = k_beg + HCH;
        auto wait_acc = [&](uint64_t* bar, uint32_t par) { if (CL == 1) mbar_wait(bar, par); else mbar_wait_cl(bar, par); };
        int64_t ini = 0;
        uint32_t pc = 0;
        auto in_take = [&](float (&v)[16], uint32_t& slice, uint64_t*& rel) {
            const int s = (int)(ini % AXS);
            const long long tw0 = FTR_NOW();
            mbar_wait(&in_full[hf][s], (uint32_t)(ini / AXS) & 1);
            if (warp == 0 && lane == 0) FTR_ADD(1, (int)(ini / (6 * HCH)), 13, FTR_NOW() - tw0);
            slice = in_base + (hf * AXS + s) * AX_TILE + q * OUT_STG;
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
            // (every CTA's GEMM_DH2 of the previous iteration has read its A tile)
            if (it > 0) { wait_acc(&ddh_full, (it - 1) & 1); tc_fence_after(); }
            if (lane == 0 && warp == 0) FTR(1, it, 6);
            for (int k = k_beg; k < k_end; ++k) {
                const int lc0 = 16 * k, c0 = col0 + lc0;
                float dl[16], zz[16], cc[16], hp[16], dh[16], kp[16];
                uint32_t s0, s1, s2, s3; uint64_t *r0, *r1, *r2, *r3;
                // the ring has 2 stages: release the first two tiles as soon as they are in registers, keep the
                // slices of the last two as store staging
                if (it > 0) {
                    tmem_ld16_issue(tmem + tl + T_DH + (uint32_t)lc0, dh);
                    tmem_ld16_issue(tmem + tl + T_KEEP + (uint32_t)lc0, kp);
                }
                in_take(dl, s0, r0); mbar_arrive(r0);
                in_take(zz, s1, r1); mbar_arrive(r1);
                in_take(cc, s2, r2); in_take(hp, s3, r3);
                if (it > 0) {
                    tmem_wait_ld();
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
                tmem_st16_issue(tmem + tl + T_KEEP + (uint32_t)lc0, kp);
                tmem_st16_issue(tmem + tl + T_DAZ + (uint32_t)lc0, dh);
                if (TS) a_put16_ts(tmem, tl, H, row, c0, dl, send, cr * KBc); else a_store16<SPLIT3>(a_hi, a_lo, row, c0, dl, send, cr * KBc);
                tmem_wait_st();
                if (k == k_end - 1) {       // A = da_c is complete: hand it over before the stores to global memory
                    if (lane == 0 && warp == 0) FTR(1, it, 7);
                    fence_async_smem(); tc_fence_before(); mbar_arrive(&s_done);
                }
                warp_store_chunk(s2, lane, dh, DAw + c0, 3 * H, rows_valid);            // DA_z
                warp_store_chunk(s3, lane, dl, DAw + 2 * H + c0, 3 * H, rows_valid);    // DA_c
                mbar_arrive(r2); mbar_arrive(r3);
            }
            if (TS && POI_FUSED_PUSH) push_slice_ts<CL>(send, a_hi + (uint32_t)cr * KBc * A_KB_BYTES, (uint32_t)KBc * A_KB_BYTES, a_ready, cr, (int)threadIdx.x);
            if (TS) convert_incoming_ts<CL>(a_ready, conv_done, pc, a_hi, tmem, tl, H, cr, KBc, row, hf);
            else convert_incoming<SPLIT3, CL>(a_ready, conv_done, pc, a_hi, a_lo, cr, KBc, row, hf);
            // ---- M1: A <- da_z (after every CTA's GEMM_M has read da_c) ----
            wait_acc(&dm_full, it & 1);
            tc_fence_after();
            if (lane == 0 && warp == 0) FTR(1, it, 8);
            if (j > 0) {
                for (int k = k_beg; k < k_end; ++k) {
                    float dz[16];
                    tmem_ld16(tmem + tl + T_DAZ + (uint32_t)(16 * k), dz);
                    if (TS) a_put16_ts(tmem, tl, H, row, col0 + 16 * k, dz, send, cr * KBc);
                    else a_store16<SPLIT3>(a_hi, a_lo, row, col0 + 16 * k, dz, send, cr * KBc);
                }
                if (TS) tmem_wait_st();
                if (lane == 0 && warp == 0) FTR(1, it, 9);
                fence_async_smem(); tc_fence_before(); mbar_arrive(&s_done);                  // A = da_z
                if (TS && POI_FUSED_PUSH) push_slice_ts<CL>(send, a_hi + (uint32_t)cr * KBc * A_KB_BYTES, (uint32_t)KBc * A_KB_BYTES, a_ready, cr, (int)threadIdx.x);
                if (TS) convert_incoming_ts<CL>(a_ready, conv_done, pc, a_hi, tmem, tl, H, cr, KBc, row, hf);
                else convert_incoming<SPLIT3, CL>(a_ready, conv_done, pc, a_hi, a_lo, cr, KBc, row, hf);
            }
            // ---- M2: da_r, keep += m r  (runs while GEMM_DH1 executes) ----
            for (int k = k_beg; k < k_end; ++k) {
                const int lc0 = 16 * k, c0 = col0 + lc0;
                float rr[16], hp[16], mm[16], kp[16];
                uint32_t s0, s1; uint64_t *r0, *r1;
                tmem_ld16_issue(tmem + tl + T_M + (uint32_t)lc0, mm);
                tmem_ld16_issue(tmem + tl + T_KEEP + (uint32_t)lc0, kp);
                in_take(rr, s0, r0); in_take(hp, s1, r1);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float m_ = ok ? mm[i] : 0.f, r_ = rr[i];
                    kp[i] = kp[i] + m_ * r_;
                    mm[i] = ok ? m_ * hp[i] * r_ * (1.f - r_) : 0.f;       // da_r
                }
                tmem_st16_issue(tmem + tl + T_KEEP + (uint32_t)lc0, kp);
                tmem_st16_issue(tmem + tl + T_M + (uint32_t)lc0, mm);       // stash da_r over m
                tmem_wait_st();
                warp_store_chunk(s0, lane, mm, DAw + H + c0, 3 * H, rows_valid);        // DA_r
                mbar_arrive(r0); mbar_arrive(r1);
            }
            if (lane == 0 && warp == 0) FTR(1, it, 10);
            // ---- M3: A <- da_r (after every CTA's GEMM_DH1 has read da_z) ----
            if (j > 0) {
                wait_acc(&dh1_done, it & 1);
                tc_fence_after();
                if (lane == 0 && warp == 0) FTR(1, it, 11);
                for (int k = k_beg; k < k_end; ++k) {
                    float dr[16];
                    tmem_ld16(tmem + tl + T_M + (uint32_t)(16 * k), dr);
                    if (TS) a_put16_ts(tmem, tl, H, row, col0 + 16 * k, dr, send, cr * KBc);
                    else a_store16<SPLIT3>(a_hi, a_lo, row, col0 + 16 * k, dr, send, cr * KBc);
                }
                if (TS) tmem_wait_st();
                if (lane == 0 && warp == 0) FTR(1, it, 12);
                fence_async_smem(); tc_fence_before(); mbar_arrive(&s_done);                  // A = da_r
                if (TS && POI_FUSED_PUSH) push_slice_ts<CL>(send, a_hi + (uint32_t)cr * KBc * A_KB_BYTES, (uint32_t)KBc * A_KB_BYTES, a_ready, cr, (int)threadIdx.x);
                if (TS) convert_incoming_ts<CL>(a_ready, conv_done, pc, a_hi, tmem, tl, H, cr, KBc, row, hf);
                else convert_incoming<SPLIT3, CL>(a_ready, conv_done, pc, a_hi, a_lo, cr, KBc, row, hf);
            }
        }
    } else if (warp < 10) {
        // ================================ input producers (one warp per column half, cp.async) ================================
        const int hf = warp - 8;
        const int per_step = 6 * HCH;
        const int64_t n_tiles = (int64_t)T * per_step;
        for (int64_t ai = 0; ai < n_tiles; ++ai) {
            const int s = (int)(ai % AXS);
            if (ai >= AXS) mbar_wait(&in_empty[hf][s], (uint32_t)((ai / AXS) - 1) & 1);
            const int j = T - 1 - (int)(ai / per_step), ti = (int)(ai % per_step);
            int kind, kk;                                       // 0 DHl, 1 Z, 2 C, 3 H_prev, 4 R
            if (ti < 4 * HCH) { kind = ti & 3; kk = ti >> 2; }
            else { const int t2 = ti - 4 * HCH; kind = (t2 & 1) ? 3 : 4; kk = t2 >> 1; }
            const float* base = kind == 0 ? DHl : kind == 1 ? Z : kind == 2 ? C : kind == 3 ? Hs : R;
            const float* src = base + ((size_t)j * B + m0) * H + col0 + 16 * (hf * HCH + kk);
            const uint32_t dst = in_base + (hf * AXS + s) * AX_TILE;
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
            int64_t ws = 0; uint32_t pa = 0, ps = 0;
            const uint32_t slice_off = (uint32_t)cr * KBc * A_KB_BYTES, slice_bytes = (uint32_t)KBc * A_KB_BYTES;
            // operand hand-over + one GEMM over K = H in arrival order (own k-blocks, then owner cr-1, cr-2, ...)
            auto gemm = [&](uint32_t dcol, bool fresh, int it, int slot) {
                begin_exchange<SPLIT3, CL>(&s_done, ps, a_ready, a_hi, send, slice_off, slice_bytes, cr, !(TS && POI_FUSED_PUSH));
                FTR(1, it, slot);
                for (int i = 0; i < CL; ++i) {
                    const int o = (cr - i + CL) % CL;
                    if (CL > 1 && i > 0) {
                        if (SPLIT3) mbar_wait(&conv_done[o], pa & 1); else mbar_wait_cl(&a_ready[o], pa & 1);
                        tc_fence_after();
                    }
                    for (int kbl = 0; kbl < KBc; ++kbl, ++ws) {
                        const int kb = o * KBc + kbl;
                        const int s = (int)(ws % WST);
                        const long long tw0 = FTR_NOW();
                        mbar_wait(&w_full[s], (uint32_t)(ws / WST) & 1);
                        FTR_ADD(1, (int)(ws / (3 * KB)), 14, FTR_NOW() - tw0);
                        tc_fence_after();
                        const uint32_t sW = w_base + s * w_stage;
                        if (TS) mma_kblock_ts(tmem + dcol, tmem + TS_A_COL + kb * 32, tmem + TS_A_COL + H + kb * 32, sW, sW + w_tile, idesc,
                                              fresh && i == 0 && kbl == 0);
                        else mma_kblock<SPLIT3>(tmem + dcol, a_hi + kb * A_KB_BYTES, a_lo + kb * A_KB_BYTES, sW, sW + w_tile, idesc,
                                                fresh && i == 0 && kbl == 0);
                        umma_commit(&w_empty[s]);
                    }
                }
                if (CL > 1) ++pa;
            };
            for (int j = T - 1; j >= 0; --j) {
                gemm(T_M, true, T - 1 - j, 0);                  // A = da_c
                umma_commit_cl<CL>(&dm_full);
                FTR(1, T - 1 - j, 1);
                if (j > 0) {
                    gemm(T_DH, true, T - 1 - j, 2);             // A = da_z
                    umma_commit_cl<CL>(&dh1_done);
                    FTR(1, T - 1 - j, 3);
                    gemm(T_DH, false, T - 1 - j, 4);            // A = da_r
                    umma_commit_cl<CL>(&ddh_full);
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
                         ::"r"(w_base + s * w_stage), "l"(wimg_c + (size_t)(ws % (3 * KB)) * w_stage), "r"(w_stage), "r"(bar) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

// cluster launch of the fused kernels (same bookkeeping as POI_LAUNCH)
template <class... KArgs, class... Args>
static int launch_clustered(poi_engine* e, const char* name, void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem,
                            int cl, Args... args) {
    ProfRec* pr = e->kprof ? prof_begin(e) : nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = e->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t st = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
    if (pr) cudaEventRecord(pr->b, e->stream);
    e->cur_flops = 0.0; e->cur_bytes = 0.0;
    e->launches++;
    if (st == cudaSuccess) st = cudaPeekAtLastError();
    if (st != cudaSuccess) POI_FAIL(e, "launch %s failed: %s", name, cudaGetErrorString(st));
    return 0;
}

// CTAs per 128-user group: as many as divide the gate columns into slices of >= 32 while the whole grid is still
// co-resident (one CTA per SM); e->fused_cluster forces 1, 2 or 4
static inline int pick_cluster(poi_engine* e, int B, int H) {
    const int groups = (int)poi_cdiv(B, FM);
    int cl = 1;
    for (int c = 2; c <= 4; c *= 2)
        if (H % (32 * c) == 0 && groups * c <= e->num_sms) cl = c;
    if (e->fused_cluster == 1 || e->fused_cluster == 2 || e->fused_cluster == 4) {
        cl = e->fused_cluster;
        while (cl > 1 && H % (32 * cl) != 0) cl >>= 1;
    }
    return cl;
}
static inline size_t fused_smem(int H, int cl, bool split3) {
    const int KB = H / 32, wst = cl == 4 ? 4 : 2;
    const size_t w_stage = (size_t)(H / cl) * 128 * (split3 ? 2 : 1);
    const size_t send = (cl > 1 && split3) ? (size_t)(KB / cl) * A_KB_BYTES : 0;       // unsplit copy of the own slice
    return (size_t)KB * A_KB_BYTES * (split3 ? 2 : 1) + wst * w_stage + (size_t)2 * AX_STAGES * AX_TILE + send + 1024;
}

template <bool SPLIT3, int CL>
static int launch_bwd_cl(poi_engine* e, const float* DHl, const float* Z, const float* R, const float* C, const float* Hs,
                         const uint8_t* wimg, float* DA, int B, int T, int H) {
    const size_t smem = fused_smem(H, CL, SPLIT3);
    POI_CK(e, cudaFuncSetAttribute(k_gru_bwd_fused<SPLIT3, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    POI_CAT(e, CAT_RECUR_BWD, 2.0 * (double)B * T * 3 * H * H, 0);
    const unsigned grid = (unsigned)poi_cdiv(B, FM) * CL;
    if (CL == 1) { POI_LAUNCH(e, (k_gru_bwd_fused<SPLIT3, 1>), grid, F_THREADS, smem, DHl, Z, R, C, Hs, wimg, DA, B, T, H); return 0; }
    return launch_clustered(e, "k_gru_bwd_fused", k_gru_bwd_fused<SPLIT3, CL>, grid, F_THREADS, smem, CL, DHl, Z, R, C, Hs, wimg, DA, B, T, H);
}

template <bool SPLIT3>
static int launch_bwd_inst(poi_engine* e, const float* DHl, const float* Z, const float* R, const float* C, const float* Hs,
                           const float* wh, float* DA, int B, int T, int H) {
    const int KB = H / 32;
    const int cl = pick_cluster(e, B, H);
    const size_t w_stage = (size_t)(H / cl) * 128 * (SPLIT3 ? 2 : 1);
    uint8_t* wimg = nullptr;
    POI_TRY(arena_get(e, (size_t)cl * 3 * KB * w_stage, &wimg));
    POI_CAT(e, CAT_ELTWISE, 0, 0);
    POI_LAUNCH(e, (k_stage_wh_bwd<SPLIT3>), 48, 256, 0, wh, H, cl, wimg);
    if (cl == 4) return launch_bwd_cl<SPLIT3, 4>(e, DHl, Z, R, C, Hs, wimg, DA, B, T, H);
    if (cl == 2) return launch_bwd_cl<SPLIT3, 2>(e, DHl, Z, R, C, Hs, wimg, DA, B, T, H);
    return launch_bwd_cl<SPLIT3, 1>(e, DHl, Z, R, C, Hs, wimg, DA, B, T, H);
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

template <bool SPLIT3, int CL>
static int launch_fwd_cl(poi_engine* e, const float* AX, const uint8_t* wimg, float* Hs, float* Z, float* R, float* C,
                         int B, int T, int H) {
    const size_t smem = fused_smem(H, CL, SPLIT3);
    POI_CK(e, cudaFuncSetAttribute(k_gru_fwd_fused<SPLIT3, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    POI_CAT(e, CAT_RECUR_FWD, 2.0 * (double)B * T * 3 * H * H, 0);
    const unsigned grid = (unsigned)poi_cdiv(B, FM) * CL;
    if (CL == 1) { POI_LAUNCH(e, (k_gru_fwd_fused<SPLIT3, 1>), grid, F_THREADS, smem, AX, wimg, Hs, Z, R, C, B, T, H); return 0; }
    return launch_clustered(e, "k_gru_fwd_fused", k_gru_fwd_fused<SPLIT3, CL>, grid, F_THREADS, smem, CL, AX, wimg, Hs, Z, R, C, B, T, H);
}

template <bool SPLIT3>
static int launch_fwd_inst(poi_engine* e, const float* AX, const float* wh, float* Hs, float* Z, float* R, float* C,
                           int B, int T, int H) {
    const int KB = H / 32;
    const int cl = pick_cluster(e, B, H);
    const size_t w_stage = (size_t)(H / cl) * 128 * (SPLIT3 ? 2 : 1);
    uint8_t* wimg = nullptr;
    POI_TRY(arena_get(e, (size_t)cl * 3 * KB * w_stage, &wimg));
    POI_CAT(e, CAT_ELTWISE, 0, 0);
    const int merge = (SPLIT3 && cl > 1 && (size_t)KB * A_KB_BYTES >= (size_t)4 * AX_TILE) ? 1 : 0;    // as in k_gru_fwd_fused
    POI_LAUNCH(e, (k_stage_wh<SPLIT3>), 48, 256, 0, wh, H, cl, merge, wimg);
    if (cl == 4) return launch_fwd_cl<SPLIT3, 4>(e, AX, wimg, Hs, Z, R, C, B, T, H);
    if (cl == 2) return launch_fwd_cl<SPLIT3, 2>(e, AX, wimg, Hs, Z, R, C, B, T, H);
    return launch_fwd_cl<SPLIT3, 1>(e, AX, wimg, Hs, Z, R, C, B, T, H);
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
