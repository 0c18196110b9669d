// gru_small.cuh -- the GRU recurrence for TINY batches (B <= 8), i.e. the reference's one-by-one mode
// (`theano.scan` over one user's sequence, GRU.py:345-360, GRU_Spatial.py:170-197).
//
// With one user the recurrent products are matrix-VECTOR products: a 128-row tensor-core tile would carry one live
// row and cost the same ~13 us per time step as a full tile.  Here one CTA owns one user for the whole sequence and
// keeps Wh (3 x H x H fp32 <= 192 KB) resident in shared memory; a step is two GEMVs in plain fp32 FMA (exact fp32,
// fixed summation order), the gate math, and three block barriers -- about a microsecond.  Forward stores Z, R, C, H
// and r*h for BPTT exactly like the fused tensor-core kernels; backward carries dh in shared memory.
//   forward : thread j owns output column j;  Wh is staged TRANSPOSED ([g][k][j]) so that a warp reads consecutive
//             addresses for a fixed k (conflict-free) and h[k] is a broadcast
//   backward: m = da_c . Wh[2], dhn = da_z . Wh[0] + da_r . Wh[1] contract over the OUTPUT index, so the natural
//             layout [g][k][n] is already conflict-free for thread n
#pragma once
#include "common.cuh"

namespace small {

constexpr int SMALL_B_MAX = 8;
constexpr int S_THREADS = 256;

static inline bool supported(int B, int H) { return B >= 1 && B <= SMALL_B_MAX && H >= 4 && H <= 128 && H % 4 == 0; }
static inline size_t smem_bytes(int H) { return ((size_t)3 * H * H + 8 * H) * sizeof(float); }

__global__ void __launch_bounds__(S_THREADS, 1)
k_gru_fwd_small(const float* __restrict__ AX, const float* __restrict__ wh, float* __restrict__ Hs, float* __restrict__ Z,
                float* __restrict__ R, float* __restrict__ C, float* __restrict__ RH, int B, int T, int H) {
    extern __shared__ float sm[];
    float* wT = sm;                         // [3][H(k)][H(j)]
    float* h = wT + (size_t)3 * H * H;      // [H]
    float* rh = h + H;                      // [H]
    float* zs = rh + H;                     // [H]
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < 3 * H * H; i += S_THREADS) {
        const int g = i / (H * H), rem = i % (H * H), j = rem / H, k = rem % H;      // coalesced global read of wh[g][j][k]
        wT[((size_t)g * H + k) * H + j] = wh[i];
    }
    for (int i = tid; i < H; i += S_THREADS) h[i] = 0.f;                             // h_{-1} = 0 (GRU.py:63)
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        const size_t row = (size_t)t * B + b;
        const float* ax = AX + row * 3 * H;
        // z | r: thread j < 2H, gate g = j / H
        float acc = 0.f;
        const int g = tid / H, jj = tid - g * H;
        if (tid < 2 * H) {
            const float* w = wT + (size_t)g * H * H + jj;
#pragma unroll 4
            for (int k = 0; k < H; ++k) acc = fmaf(w[(size_t)k * H], h[k], acc);
            const float v = 1.f / (1.f + expf(-(acc + ax[tid])));
            if (g == 0) { zs[jj] = v; Z[row * H + jj] = v; }
            else { const float x = v * h[jj]; rh[jj] = x; R[row * H + jj] = v; RH[row * H + jj] = x; }
        }
        __syncthreads();
        float hn = 0.f;
        if (tid < H) {
            const float* w = wT + (size_t)2 * H * H + tid;
            float a2 = 0.f;
#pragma unroll 4
            for (int k = 0; k < H; ++k) a2 = fmaf(w[(size_t)k * H], rh[k], a2);
            const float c = tanhf(a2 + ax[2 * H + tid]);
            const float z = zs[tid];
            hn = (1.f - z) * h[tid] + z * c;                 // GRU.py:351
            C[row * H + tid] = c;
            Hs[(row + B) * H + tid] = hn;
        }
        __syncthreads();                                     // every thread has read h before it is replaced
        if (tid < H) h[tid] = hn;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(S_THREADS, 1)
k_gru_bwd_small(const float* __restrict__ DHl, const float* __restrict__ Z, const float* __restrict__ R,
                const float* __restrict__ C, const float* __restrict__ Hs, const float* __restrict__ wh,
                float* __restrict__ DA, int B, int T, int H) {
    extern __shared__ float sm[];
    float* w = sm;                          // [3][H(k)][H(n)] as in memory
    float* dac = w + (size_t)3 * H * H;     // [H]
    float* daz = dac + H;
    float* dar = daz + H;
    const int b = blockIdx.x, n = threadIdx.x;
    for (int i = n; i < 3 * H * H; i += S_THREADS) w[i] = wh[i];
    __syncthreads();
    float dh = 0.f;                          // d loss / d h_j carried from step j+1 (thread n owns column n)
    for (int j = T - 1; j >= 0; --j) {
        const size_t row = (size_t)j * B + b;
        float keep = 0.f, hp = 0.f, r = 0.f;
        if (n < H) {
            const float z = Z[row * H + n], c = C[row * H + n];
            hp = Hs[row * H + n]; r = R[row * H + n];
            const float dht = dh + DHl[row * H + n];
            const float a_c = dht * z * (1.f - c * c);
            const float a_z = dht * (c - hp) * z * (1.f - z);
            keep = dht * (1.f - z);
            dac[n] = a_c; daz[n] = a_z;
            DA[row * 3 * H + n] = a_z; DA[row * 3 * H + 2 * H + n] = a_c;
        }
        __syncthreads();
        if (n < H) {
            float m = 0.f;
            const float* w2 = w + (size_t)2 * H * H + n;
#pragma unroll 4
            for (int k = 0; k < H; ++k) m = fmaf(dac[k], w2[(size_t)k * H], m);
            const float a_r = m * hp * r * (1.f - r);
            keep += m * r;
            dar[n] = a_r;
            DA[row * 3 * H + H + n] = a_r;
        }
        __syncthreads();
        if (n < H && j > 0) {
            float s = 0.f;
            const float* w0 = w + n; const float* w1 = w + (size_t)H * H + n;
#pragma unroll 4
            for (int k = 0; k < H; ++k) { s = fmaf(daz[k], w0[(size_t)k * H], s); s = fmaf(dar[k], w1[(size_t)k * H], s); }
            dh = keep + s;
        }
        __syncthreads();                                     // dac / daz / dar are rewritten by the next step
    }
}

static int launch_fwd(poi_engine* e, const float* AX, const float* wh, float* Hs, float* Z, float* R, float* C, float* RH,
                      int B, int T, int H) {
    const size_t smem = smem_bytes(H);
    POI_CK(e, cudaFuncSetAttribute(k_gru_fwd_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    POI_CAT(e, CAT_RECUR_FWD, 2.0 * (double)B * T * 3 * H * H, 0);
    POI_LAUNCH(e, k_gru_fwd_small, (unsigned)B, S_THREADS, smem, AX, wh, Hs, Z, R, C, RH, B, T, H);
    return 0;
}
static int launch_bwd(poi_engine* e, const float* DHl, const float* Z, const float* R, const float* C, const float* Hs,
                      const float* wh, float* DA, int B, int T, int H) {
    const size_t smem = smem_bytes(H);
    POI_CK(e, cudaFuncSetAttribute(k_gru_bwd_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    POI_CAT(e, CAT_RECUR_BWD, 2.0 * (double)B * T * 3 * H * H, 0);
    POI_LAUNCH(e, k_gru_bwd_small, (unsigned)B, S_THREADS, smem, DHl, Z, R, C, Hs, wh, DA, B, T, H);
    return 0;
}

}  // namespace small
