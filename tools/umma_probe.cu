// umma_probe.cu -- like umma_probe.cu, but (1) the accumulator is pre-filled by a K-major MMA so a dropped MMA is
// visible, (2) also probes MN-major B (A = K-major identity) and the no-swizzle MN-major layout.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I point-of-interest-recommendation_b200/csrc tools/umma_probe.cu -o tools/umma_probe
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "sort.cuh"
#include "rows.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
using namespace tc;

// which: 0 = data operand is A (B identity), 1 = data operand is B (A identity)
__global__ void k_probe(uint32_t lbo, uint32_t sbo, uint32_t mn, uint32_t ltype, int which, int pass, int prefill, float* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    float* sm = (float*)(smem_raw + (sbase - smem_u32(smem_raw)));
    const int AF = 65536 / 4;                       // data region: 64 KB
    for (int f = tid; f < AF; f += blockDim.x) {
        int chunk = f >> 2;
        sm[f] = pass == 0 ? (float)(chunk & 1023) : (pass == 1 ? (float)(f & 3) : (float)(chunk >> 10) + 1.f);
    }
    float* sb = sm + AF;                            // identity region: 128 rows x 128 B, K-major SW128
    for (int f = tid; f < 128 * 32; f += blockDim.x) sb[f] = 0.f;
    __syncthreads();
    if (tid < 8) { int j = tid, k = tid; sb[j * 32 + (((k >> 2) ^ (j & 7)) << 2) + (k & 3)] = 1.f; }
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_async_smem();
    if (warp == 0) tmem_alloc(&tmem_s, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_s;
    if (tid == 0) {
        uint64_t dI = make_sdesc(sbase + 65536);
        if (prefill) {      // D = data(K-major view) x identity: something non-trivial everywhere in columns 0..7
            umma_tf32(tmem, make_sdesc(sbase), dI, make_idesc_tf32(128, 128), 0u);
        }
        uint64_t dD;
        if (mn) dD = (uint64_t)((sbase >> 4) & 0x3fffu) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)ltype << 61);
        else dD = make_sdesc(sbase);
        uint32_t idesc = make_idesc_tf32(128, 128);
        if (mn) idesc |= which == 0 ? (1u << 15) : (1u << 16);
        if (pass != 3) {
            if (which == 0) umma_tf32(tmem, dD, dI, idesc, 0u);
            else umma_tf32(tmem, dI, dD, idesc, 0u);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (warp < 4) {
        for (int c = 0; c < 128; c += 16) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
            for (int k = 0; k < 16; ++k) out[tid * 128 + c + k] = v[k];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

int main() {
    float* d; cudaMalloc(&d, 128 * 128 * 4);
    static float h[3][128 * 128];
    const int smem = 65536 + 16384 + 1024;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    struct { uint32_t lbo, sbo, mn, ltype; int which, prefill; } cfgs[] = {
        {16, 1024, 0, 2, 0, 0},                               // K-major SW128 (known good)
        {4096, 512, 1, 1, 0, 1}, {4096, 512, 1, 1, 1, 1},     // MN-major SWIZZLE_128B_BASE32B, data = A / B
        {512, 4096, 1, 1, 0, 1},
        {4096, 1024, 1, 2, 0, 1}};                            // MN-major plain SW128: reads as zeros for tf32
    const int rows[] = {0, 1, 2, 3, 4, 7, 8, 9, 31, 32, 33, 64, 127};
    for (auto& c : cfgs) {
        for (int pass = 0; pass < 3; ++pass) {
            cudaMemset(d, 0, sizeof(h[0]));
            k_probe<<<1, 128, smem>>>(c.lbo, c.sbo, c.mn, c.ltype, c.which, pass, c.prefill, d);
            cudaError_t er = cudaDeviceSynchronize();
            if (er != cudaSuccess) { printf("cfg lbo=%u sbo=%u mn=%u: CUDA error %s\n", c.lbo, c.sbo, c.mn, cudaGetErrorString(er)); return 1; }
            cudaMemcpy(h[pass], d, sizeof(h[0]), cudaMemcpyDeviceToHost);
        }
        printf("== lbo=%u sbo=%u mn=%u ltype=%u data=%c prefill=%d : byte offset of data(i,k) read by the MMA (-16384.. = untouched/none)\n",
               c.lbo, c.sbo, c.mn, c.ltype, c.which ? 'B' : 'A', c.prefill);
        for (int r : rows) {
            printf("i=%3d:", r);
            for (int k = 0; k < 8; ++k) {
                // which==0: D[i][k] = A(i,k) at lane i col k.  which==1: D[k][j] = B(j,k) at lane k col j=i
                int idx = c.which == 0 ? r * 128 + k : k * 128 + r;
                int hi = (int)h[2][idx] - 1;
                printf(" %6d", (hi * 1024 + (int)h[0][idx]) * 16 + (int)h[1][idx] * 4);
            }
            printf("\n");
        }
    }
    return 0;
}
