// eval.cuh -- scoring + top-K: users . items^T (+ wd * prob) fused with a
// streaming top-K so that the I-wide score rows are only ever materialised one item chunk at a time.
// Replaces compute_sub_all_scores (GRU.py:93-96, GRU_Spatial.py:117-125) + np.argpartition/argsort
// (Valuate.py:91-100,133-146).
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"
#include "sampling.cuh"

struct EpiScore {       // S[b, i] = acc + wd * prob[b, i0 + i]
    float* S; int lds; const float* prob; int64_t n_item; int64_t i0; float wd; int N;
    __device__ __forceinline__ void operator()(int m, int n, const float (&v)[4]) const {
        for (int k = 0; k < 4 && n + k < N; ++k) {
            float s = v[k];
            if (prob) s += wd * prob[(size_t)m * n_item + i0 + n + k];
            S[(size_t)m * lds + n + k] = s;
        }
    }
};

// Distance2Pre scoring without the U x I `prob` / `ulptai` matrices (GRU_Spatial.py:77-78,117-125,
// Load_Data_by_length.py:183-235): prob[b, i] = sts[b, iv] * [iv < dist_num] with iv = the distance interval between user
// b's last training POI and item i, computed here from the coordinates (fp64 haversine of cal_dis, the same device
// function the negative-distance binning uses) instead of being read from a precomputed matrix.
struct EpiScoreGeo {
    float* S; int lds; const float* sts; int nD; const double* ucoord; const double* icoord; double dd; int dist_num;
    int64_t i0; float wd; int N;
    __device__ __forceinline__ void operator()(int m, int n, const float (&v)[4]) const {
        const double ulat = ucoord[2 * (size_t)m], ulon = ucoord[2 * (size_t)m + 1];
        for (int k = 0; k < 4 && n + k < N; ++k) {
            const size_t it = (size_t)(i0 + n + k);
            const int iv = haversine_interval(ulat, ulon, icoord[2 * it], icoord[2 * it + 1], dd, dist_num);
            float s = v[k];
            if (iv < dist_num) s += wd * sts[(size_t)m * nD + iv];
            S[(size_t)m * lds + n + k] = s;
        }
    }
};

// one CTA per user row: merge the running top-K (vals/idx, descending) with a chunk of scores.
// K rounds of block arg-max; ties resolve to the smaller item index (deterministic).
__global__ void __launch_bounds__(256)
k_topk_merge(const float* __restrict__ S, int lds, int chunk, int64_t i0, int K,
             float* __restrict__ best_val, int32_t* __restrict__ best_idx, int first) {
    extern __shared__ float s_sc[];                         // chunk + K candidates
    __shared__ float rv[8]; __shared__ int ri[8];
    int32_t* s_id = reinterpret_cast<int32_t*>(s_sc + chunk + K);
    const int b = blockIdx.x, tid = threadIdx.x;
    const int total = chunk + (first ? 0 : K);
    for (int i = tid; i < chunk; i += blockDim.x) s_sc[i] = S[(size_t)b * lds + i];
    if (!first) for (int i = tid; i < K; i += blockDim.x) { s_sc[chunk + i] = best_val[(size_t)b * K + i]; s_id[i] = best_idx[(size_t)b * K + i]; }
    __syncthreads();
    for (int r = 0; r < K; ++r) {
        float bv = -INFINITY; int bi = 0x7fffffff; int bpos = -1;
        for (int i = tid; i < total; i += blockDim.x) {
            float v = s_sc[i];
            int id = i < chunk ? (int)(i0 + i) : s_id[i - chunk];
            if (v > bv || (v == bv && id < bi)) { bv = v; bi = id; bpos = i; }
        }
        // warp then block arg-max (value desc, index asc)
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            int op = __shfl_xor_sync(0xffffffffu, bpos, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bpos = op; }
        }
        __shared__ int rp[8];
        if ((tid & 31) == 0) { rv[tid >> 5] = bv; ri[tid >> 5] = bi; rp[tid >> 5] = bpos; }
        __syncthreads();
        if (tid == 0) {
            float fv = rv[0]; int fi = ri[0], fp = rp[0];
            for (int w = 1; w < 8; ++w) if (rv[w] > fv || (rv[w] == fv && ri[w] < fi)) { fv = rv[w]; fi = ri[w]; fp = rp[w]; }
            best_val[(size_t)b * K + r] = fv; best_idx[(size_t)b * K + r] = fi;
            if (fp >= 0) s_sc[fp] = -INFINITY;
        }
        __syncthreads();
    }
}
