// gemm_simt.cuh -- fp32 FMA GEMMs with fused epilogues (gemm mode 0: bit-faithful fp32 path).
//
//   gemm_tn : C[m,n] = sum_k A[m*lda + k] * W[n*ldw + k]        ("TN": both operands K-major)
//             -- T.dot(ui[:2], x), T.dot(wh[:2], h), T.dot(wh[2], r*h), T.dot(vs, h)
//                (GRU.py:346-350, GRU_Spatial.py:173-180) batched over users / time
//   gemm_atb: C[i,j] = sum_m A[m*lda + i] * Bm[m*ldb + j]       (weight gradients, split over m)
//
// The epilogue functor receives 4 consecutive output columns of one row.
#pragma once
#include "common.cuh"

template <int BM, int BN, int BK, int TM, int TN, class Epi>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
k_gemm_tn(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
          int M, int N, int K, Epi epi) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int CM = TM / 4, CN = TN / 4;
    constexpr int SA = BM + 4, SW = BN + 4;
    constexpr int KQ = BK / 4;
    constexpr int LA = (BM * KQ + NT - 1) / NT, LW = (BN * KQ + NT - 1) / NT;
    __shared__ __align__(16) float As[2][BK][SA];
    __shared__ __align__(16) float Ws[2][BK][SW];
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[LA], rw[LW];
    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            int f = tid + i * NT;
            int row = f / KQ, kq = f % KQ;
            float4 v = f4zero();
            if (f < BM * KQ && m0 + row < M && k0 + kq * 4 < K)
                v = ld4(A + (size_t)(m0 + row) * lda + k0 + kq * 4);
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < LW; ++i) {
            int f = tid + i * NT;
            int row = f / KQ, kq = f % KQ;
            float4 v = f4zero();
            if (f < BN * KQ && n0 + row < N && k0 + kq * 4 < K)
                v = __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + row) * ldw + k0 + kq * 4));
            rw[i] = v;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            int f = tid + i * NT;
            if (f < BM * KQ) {
                int row = f / KQ, kq = f % KQ;
                As[buf][kq * 4 + 0][row] = ra[i].x; As[buf][kq * 4 + 1][row] = ra[i].y;
                As[buf][kq * 4 + 2][row] = ra[i].z; As[buf][kq * 4 + 3][row] = ra[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < LW; ++i) {
            int f = tid + i * NT;
            if (f < BN * KQ) {
                int row = f / KQ, kq = f % KQ;
                Ws[buf][kq * 4 + 0][row] = rw[i].x; Ws[buf][kq * 4 + 1][row] = rw[i].y;
                Ws[buf][kq * 4 + 2][row] = rw[i].z; Ws[buf][kq * 4 + 3][row] = rw[i].w;
            }
        }
    };

    const int KT = (K + BK - 1) / BK;
    gload(0);
    sstore(0);
    __syncthreads();
    int cur = 0;
    for (int kt = 0; kt < KT; ++kt) {
        if (kt + 1 < KT) gload((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int c = 0; c < CM; ++c) {
                float4 v = *reinterpret_cast<const float4*>(&As[cur][k][c * (BM / CM) + ty * 4]);
                a[c * 4 + 0] = v.x; a[c * 4 + 1] = v.y; a[c * 4 + 2] = v.z; a[c * 4 + 3] = v.w;
            }
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                float4 v = *reinterpret_cast<const float4*>(&Ws[cur][k][c * (BN / CN) + tx * 4]);
                b[c * 4 + 0] = v.x; b[c * 4 + 1] = v.y; b[c * 4 + 2] = v.z; b[c * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < KT) sstore(cur ^ 1);
        __syncthreads();
        cur ^= 1;
    }

#pragma unroll
    for (int ci = 0; ci < CM; ++ci)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            int m = m0 + ci * (BM / CM) + ty * 4 + ii;
            if (m >= M) continue;
#pragma unroll
            for (int cj = 0; cj < CN; ++cj) {
                int n = n0 + cj * (BN / CN) + tx * 4;
                if (n >= N) continue;
                float v[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) v[jj] = acc[ci * 4 + ii][cj * 4 + jj];
                epi(m, n, v);
            }
        }
}

// tile choice: big tiles for the hoisted (all-timestep) GEMMs, small tiles when M = one timestep
template <class Epi>
static int launch_gemm_tn(poi_engine* e, const float* A, int lda, const float* W, int ldw,
                          int64_t M, int N, int K, const Epi& epi) {
    if (M <= 0 || N <= 0) return 0;
    if (M > 0x7fffffffLL) POI_FAIL(e, "gemm M too large");
    POI_CAT(e, CAT_GEMM, 2.0 * (double)M * N * K, 0);
    int64_t big_ctas = poi_cdiv(M, 128) * poi_cdiv(N, 128);
    if (big_ctas >= 2 * e->num_sms) {
        dim3 grid((unsigned)poi_cdiv(N, 128), (unsigned)poi_cdiv(M, 128));
        POI_LAUNCH(e, (k_gemm_tn<128, 128, 16, 8, 8, Epi>), grid, 256, 0, A, lda, W, ldw, (int)M, N, K, epi);
    } else {
        dim3 grid((unsigned)poi_cdiv(N, 64), (unsigned)poi_cdiv(M, 64));
        POI_LAUNCH(e, (k_gemm_tn<64, 64, 16, 4, 4, Epi>), grid, 256, 0, A, lda, W, ldw, (int)M, N, K, epi);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// C[i,j] = sum_m A[m, i] * Bm[m, j], split over m; partials [splits][N1][N2] reduced later in
// split order (deterministic).
// ---------------------------------------------------------------------------------------------
template <int BI, int BJ, int BK, int TI, int TJ>
__global__ void __launch_bounds__((BI / TI) * (BJ / TJ))
k_gemm_atb(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
           int64_t M, int N1, int N2, int64_t m_per_split, float* __restrict__ part) {
    constexpr int NT = (BI / TI) * (BJ / TJ);
    constexpr int IQ = BI / 4, JQ = BJ / 4;
    constexpr int LA = (BK * IQ + NT - 1) / NT, LB = (BK * JQ + NT - 1) / NT;
    __shared__ __align__(16) float As[2][BK][BI];
    __shared__ __align__(16) float Bs[2][BK][BJ];
    const int tid = threadIdx.x;
    const int tx = tid % (BJ / TJ), ty = tid / (BJ / TJ);
    const int i0 = blockIdx.y * BI, j0 = blockIdx.x * BJ;
    const int64_t mb = (int64_t)blockIdx.z * m_per_split;
    const int64_t me = mb + m_per_split < M ? mb + m_per_split : M;

    float acc[TI][TJ];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[i][j] = 0.f;

    float4 ra[LA], rb[LB];
    auto gload = [&](int64_t mk) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            int f = tid + i * NT;
            int k = f / IQ, iq = f % IQ;
            float4 v = f4zero();
            if (f < BK * IQ && mk + k < me && i0 + iq * 4 < N1)
                v = ld4(A + (size_t)(mk + k) * lda + i0 + iq * 4);
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            int f = tid + i * NT;
            int k = f / JQ, jq = f % JQ;
            float4 v = f4zero();
            if (f < BK * JQ && mk + k < me && j0 + jq * 4 < N2)
                v = ld4(Bm + (size_t)(mk + k) * ldb + j0 + jq * 4);
            rb[i] = v;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            int f = tid + i * NT;
            if (f < BK * IQ) *reinterpret_cast<float4*>(&As[buf][f / IQ][(f % IQ) * 4]) = ra[i];
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            int f = tid + i * NT;
            if (f < BK * JQ) *reinterpret_cast<float4*>(&Bs[buf][f / JQ][(f % JQ) * 4]) = rb[i];
        }
    };

    const int64_t KT = (me - mb + BK - 1) / BK;
    if (KT > 0) { gload(mb); sstore(0); }
    __syncthreads();
    int cur = 0;
    for (int64_t kt = 0; kt < KT; ++kt) {
        if (kt + 1 < KT) gload(mb + (kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TI], b[TJ];
#pragma unroll
            for (int c = 0; c < TI / 4; ++c) {
                float4 v = *reinterpret_cast<const float4*>(&As[cur][k][c * (BI / (TI / 4)) + ty * 4]);
                a[c * 4 + 0] = v.x; a[c * 4 + 1] = v.y; a[c * 4 + 2] = v.z; a[c * 4 + 3] = v.w;
            }
#pragma unroll
            for (int c = 0; c < TJ / 4; ++c) {
                float4 v = *reinterpret_cast<const float4*>(&Bs[cur][k][c * (BJ / (TJ / 4)) + tx * 4]);
                b[c * 4 + 0] = v.x; b[c * 4 + 1] = v.y; b[c * 4 + 2] = v.z; b[c * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < KT) sstore(cur ^ 1);
        __syncthreads();
        cur ^= 1;
    }
    float* P = part + (size_t)blockIdx.z * N1 * N2;
#pragma unroll
    for (int ci = 0; ci < TI / 4; ++ci)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            int i = i0 + ci * (BI / (TI / 4)) + ty * 4 + ii;
            if (i >= N1) continue;
#pragma unroll
            for (int cj = 0; cj < TJ / 4; ++cj) {
                int j = j0 + cj * (BJ / (TJ / 4)) + tx * 4;
                if (j >= N2) continue;
                st4(P + (size_t)i * N2 + j, make_float4(acc[ci * 4 + ii][cj * 4 + 0], acc[ci * 4 + ii][cj * 4 + 1],
                                                         acc[ci * 4 + ii][cj * 4 + 2], acc[ci * 4 + ii][cj * 4 + 3]));
            }
        }
}

struct AtbPlan { int splits; int64_t m_per_split; float* part; int N1, N2; };

// N1, N2 multiples of 4 (pad columns are the caller's business)
static int launch_gemm_atb(poi_engine* e, const float* A, int lda, const float* Bm, int ldb,
                           int64_t M, int N1, int N2, AtbPlan* plan) {
    int tiles = (int)(poi_cdiv(N1, 64) * poi_cdiv(N2, 64));
    int splits = (int)std::max<int64_t>(1, std::min<int64_t>(poi_cdiv(2 * e->num_sms, tiles), poi_cdiv(std::max<int64_t>(M, 1), 256)));
    int64_t mps = poi_align_up((size_t)poi_cdiv(std::max<int64_t>(M, 1), splits), 16);
    splits = (int)std::max<int64_t>(1, poi_cdiv(std::max<int64_t>(M, 1), mps));
    plan->splits = splits; plan->m_per_split = mps; plan->N1 = N1; plan->N2 = N2;
    POI_TRY(arena_get(e, (size_t)splits * N1 * N2, &plan->part));
    POI_CAT(e, CAT_WGRAD, 2.0 * (double)M * N1 * N2, 0);
    dim3 grid((unsigned)poi_cdiv(N2, 64), (unsigned)poi_cdiv(N1, 64), (unsigned)splits);
    POI_LAUNCH(e, (k_gemm_atb<64, 64, 16, 4, 4>), grid, 256, 0, A, lda, Bm, ldb, M, N1, N2, mps, plan->part);
    return 0;
}

// theta[i, j] (ld = ldt) <- theta - alpha * (gscale * sum_splits part[.][i][j] + lambda * theta),
// i < n1_true, j < n2_true.  The dense SGD step `par - lr * gra` (GRU.py:371, GRU_Spatial.py:211)
// fused with the split reduction.
__global__ void k_reduce_update(const float* __restrict__ part, int splits, int N1, int N2,
                                float* __restrict__ theta, int ldt, int n1_true, int n2_true,
                                float alpha, float lambda) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n1_true * n2_true) return;
    int i = (int)(idx / n2_true), j = (int)(idx % n2_true);
    float g = 0.f;
    for (int s = 0; s < splits; ++s) g += part[((size_t)s * N1 + i) * N2 + j];
    float th = theta[(size_t)i * ldt + j];
    theta[(size_t)i * ldt + j] = th - alpha * (g + lambda * th);
}

// multi-GPU: only sum the split partials into a gradient buffer (the SGD step happens after the all-reduce)
__global__ void k_reduce_only(const float* __restrict__ part, int splits, int N1, int N2,
                              float* __restrict__ out, int ldo, int n1_true, int n2_true) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n1_true * n2_true) return;
    int i = (int)(idx / n2_true), j = (int)(idx % n2_true);
    float g = 0.f;
    for (int s = 0; s < splits; ++s) g += part[((size_t)s * N1 + i) * N2 + j];
    out[(size_t)i * ldo + j] = g;
}

// theta or gradient buffer, depending on `grad_out`
static int launch_reduce_to(poi_engine* e, const AtbPlan& p, float* theta, int ldt, int n1_true, int n2_true,
                            float alpha, float lambda, float* grad_out);

static int launch_reduce_update(poi_engine* e, const AtbPlan& p, float* theta, int ldt,
                                int n1_true, int n2_true, float alpha, float lambda) {
    int64_t n = (int64_t)n1_true * n2_true;
    POI_CAT(e, CAT_WGRAD, 0, 0);
    POI_LAUNCH(e, k_reduce_update, (unsigned)poi_cdiv(n, 256), 256, 0, p.part, p.splits, p.N1, p.N2,
               theta, ldt, n1_true, n2_true, alpha, lambda);
    return 0;
}

// column sums of A [M x N] (ld = lda): partial[split][N], fixed row partition
__global__ void __launch_bounds__(256)
k_colsum_partial(const float* __restrict__ A, int lda, int64_t M, int N, int64_t m_per_split,
                 float* __restrict__ part) {
    __shared__ float4 sh[16][16];
    const int cq = threadIdx.x & 15, rl = threadIdx.x >> 4;       // 16 float4 columns x 16 row lanes
    const int c = (blockIdx.x * 16 + cq) * 4;
    const int64_t mb = (int64_t)blockIdx.y * m_per_split;
    const int64_t me = mb + m_per_split < M ? mb + m_per_split : M;
    float4 acc = f4zero();
    if (c < N)
        for (int64_t m = mb + rl; m < me; m += 16) acc = f4add(acc, ld4(A + (size_t)m * lda + c));
    sh[rl][cq] = acc;
    __syncthreads();
    if (rl == 0 && c < N) {
        float4 t = sh[0][cq];
#pragma unroll
        for (int r = 1; r < 16; ++r) t = f4add(t, sh[r][cq]);
        st4(part + (size_t)blockIdx.y * N + c, t);
    }
}

static int launch_colsum(poi_engine* e, const float* A, int lda, int64_t M, int N, AtbPlan* plan) {
    int colblocks = (int)poi_cdiv(N, 64);
    int splits = (int)std::max<int64_t>(1, std::min<int64_t>(poi_cdiv(2 * e->num_sms, colblocks), poi_cdiv(std::max<int64_t>(M, 1), 64)));
    int64_t mps = poi_cdiv(std::max<int64_t>(M, 1), splits);
    plan->splits = splits; plan->m_per_split = mps; plan->N1 = 1; plan->N2 = N;
    POI_TRY(arena_get(e, (size_t)splits * N, &plan->part));
    dim3 grid((unsigned)colblocks, (unsigned)splits);
    POI_CAT(e, CAT_WGRAD, 0, (double)M * N * 4);
    POI_LAUNCH(e, k_colsum_partial, grid, 256, 0, A, lda, M, N, mps, plan->part);
    return 0;
}

// All split reductions of a step in ONE launch (the six k_reduce_update launches of a Distance2Pre step cost 14 us each,
// most of it launch latency and an un-batched chain of loads).  Entry t owns the blocks [block0[t], block0[t + 1]); every
// output element sums its partials in split order (fixed), then either takes the SGD step or lands in the gradient buffer.
constexpr int REDUCE_MAX = 8;
struct ReduceSet {
    const float* part[REDUCE_MAX]; float* dst[REDUCE_MAX];
    int splits[REDUCE_MAX], N1[REDUCE_MAX], N2[REDUCE_MAX], ld[REDUCE_MAX], n1[REDUCE_MAX], n2[REDUCE_MAX];
    unsigned block0[REDUCE_MAX + 1];
    int n, update;
};
__global__ void __launch_bounds__(256)
k_reduce_multi(ReduceSet rs, float alpha, float lambda) {
    int t = 0;
    while (t + 1 < rs.n && blockIdx.x >= rs.block0[t + 1]) ++t;
    const int64_t idx = (int64_t)(blockIdx.x - rs.block0[t]) * blockDim.x + threadIdx.x;
    const int n2 = rs.n2[t];
    if (idx >= (int64_t)rs.n1[t] * n2) return;
    const int i = (int)(idx / n2), j = (int)(idx % n2);
    const float* p = rs.part[t] + (size_t)i * rs.N2[t] + j;
    const size_t stride = (size_t)rs.N1[t] * rs.N2[t];
    const int splits = rs.splits[t];
    float g = 0.f;
    int s = 0;
    for (; s + 8 <= splits; s += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(p + (size_t)(s + u) * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) g += v[u];
    }
    for (; s < splits; ++s) g += __ldg(p + (size_t)s * stride);
    float* d = rs.dst[t] + (size_t)i * rs.ld[t] + j;
    if (rs.update) { const float th = *d; *d = th - alpha * (g + lambda * th); }
    else *d = g;
}
static inline void reduce_set_add(ReduceSet& rs, const AtbPlan& p, float* theta, int ldt, int n1_true, int n2_true, float* grad_out) {
    const int t = rs.n++;
    rs.part[t] = p.part; rs.splits[t] = p.splits; rs.N1[t] = p.N1; rs.N2[t] = p.N2;
    rs.dst[t] = grad_out ? grad_out : theta; rs.ld[t] = ldt; rs.n1[t] = n1_true; rs.n2[t] = n2_true;
    rs.update = grad_out ? 0 : 1;
    rs.block0[t + 1] = rs.block0[t] + (unsigned)poi_cdiv((int64_t)n1_true * n2_true, 256);
}
static int launch_reduce_set(poi_engine* e, const ReduceSet& rs, float alpha, float lambda) {
    if (rs.n == 0) return 0;
    POI_CAT(e, CAT_WGRAD, 0, 0);
    POI_LAUNCH(e, k_reduce_multi, rs.block0[rs.n], 256, 0, rs, alpha, lambda);
    return 0;
}

static int launch_reduce_to(poi_engine* e, const AtbPlan& p, float* theta, int ldt, int n1_true, int n2_true,
                            float alpha, float lambda, float* grad_out) {
    if (!grad_out) return launch_reduce_update(e, p, theta, ldt, n1_true, n2_true, alpha, lambda);
    int64_t n = (int64_t)n1_true * n2_true;
    POI_CAT(e, CAT_WGRAD, 0, 0);
    POI_LAUNCH(e, k_reduce_only, (unsigned)poi_cdiv(n, 256), 256, 0, p.part, p.splits, p.N1, p.N2, grad_out, ldt,
               n1_true, n2_true);
    return 0;
}
