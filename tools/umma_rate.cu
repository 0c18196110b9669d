// umma_rate.cu -- how long does one tcgen05.mma.kind::tf32 (M = 128, K = 8) take as a function of N, with the A operand
// in shared memory (SS) and in tensor memory (TS)?  One CTA, one issuing thread, R back-to-back instructions, SM clock
// from first issue to the commit's mbarrier completion.  Also checks the TS-mode A layout (lane = row, column = k).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I point-of-interest-recommendation_b200/csrc tools/umma_rate.cu -o tools/umma_rate
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "sort.cuh"
#include "rows.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
using namespace tc;

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// mode 0: SS, same A k-slice every instruction; 1: SS, walking the 4 k-slices of a 128 B row; 2: TS
__global__ void k_rate(int N, int R, int mode, long long* cycles, float* dout) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    float* sm = (float*)(smem_raw + (sbase - smem_u32(smem_raw)));
    // A tile: 128 rows x 128 B (K-major SW128) at sbase; B tile: 256 rows x 128 B at sbase + 16384
    for (int f = tid; f < 128 * 32; f += blockDim.x) sm[f] = 0.f;
    for (int f = tid; f < 256 * 32; f += blockDim.x) sm[4096 + f] = 0.f;
    __syncthreads();
    // B[j][k] = j * 8 + k for k < 8 (first k-slice), swizzled: 16-byte chunk (k >> 2) ^ (j & 7)
    for (int j = tid; j < 256; j += blockDim.x)
        for (int k = 0; k < 8; ++k) sm[4096 + j * 32 + (((k >> 2) ^ (j & 7)) << 2) + (k & 3)] = (float)(j * 8 + k);
    // A[i][k] = (k == i % 8)
    for (int i = tid; i < 128; i += blockDim.x) { int k = i & 7; sm[i * 32 + (((k >> 2) ^ (i & 7)) << 2) + (k & 3)] = 1.f; }
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_async_smem();
    if (warp == 0) tmem_alloc(&tmem_s, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_s;
    const uint32_t tA = tmem + 256;                     // TS mode: A at columns 256..263, lane = row, column = k
    if (warp < 4) {
        const int row = warp * 32 + lane;
        for (int k = 0; k < 8; ++k) tmem_st1(tA + ((uint32_t)(warp * 32) << 16) + k, (k == (row & 7)) ? 1.f : 0.f);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(128, N);
        const uint64_t dA = make_sdesc(sbase), dB = make_sdesc(sbase + 16384);
        const long long t0 = clock64();
        for (int r = 0; r < R; ++r) {
            if (mode == 2) umma_tf32_ts(tmem, tA, dB, idesc, r > 0);
            else umma_tf32(tmem, dA + (mode == 1 ? 2 * (r & 3) : 0), dB, idesc, r > 0);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        cycles[0] = clock64() - t0;
    }
    __syncthreads();
    tc_fence_after();
    if (warp < 4) {                                      // D[i][j], first 16 columns
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
        for (int k = 0; k < 16; ++k) dout[tid * 16 + k] = v[k];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long* dc; float* dd;
    cudaMalloc(&dc, 8); cudaMalloc(&dd, 128 * 16 * 4);
    const int smem = 16384 + 32768 + 1024;
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    static float h[128 * 16];
    const char* names[3] = {"SS same k-slice", "SS walking k-slices", "TS (A in TMEM)"};
    for (int mode = 0; mode < 3; ++mode) {
        for (int N : {16, 32, 64, 128, 256}) {
            long long c[2] = {0, 0};
            for (int rep = 0; rep < 2; ++rep) {
                const int R = rep == 0 ? 64 : 256;
                k_rate<<<1, 128, smem>>>(N, R, mode, dc, dd);
                cudaError_t er = cudaDeviceSynchronize();
                if (er != cudaSuccess) { printf("mode %d N %d: CUDA error %s\n", mode, N, cudaGetErrorString(er)); return 1; }
                cudaMemcpy(&c[rep], dc, 8, cudaMemcpyDeviceToHost);
            }
            cudaMemcpy(h, dd, sizeof(h), cudaMemcpyDeviceToHost);
            // one instruction computes D[i][j] = B[j][i % 8] = j * 8 + i % 8; R accumulated copies of it (mode 1 adds zeros for r & 3 != 0)
            const float per = h[5 * 16 + 3];            // i = 5, j = 3 -> 3 * 8 + 5 = 29 per contributing instruction
            printf("%-22s N=%3d  %7.1f cycles per UMMA (R=256: %lld, R=64: %lld)   D[5][3] = %.0f (29 per contributing instruction)\n",
                   names[mode], N, (double)(c[1] - c[0]) / 192.0, c[1], c[0], per);
        }
    }
    return 0;
}
