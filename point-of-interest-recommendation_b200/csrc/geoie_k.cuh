// geoie_k.cuh -- GeoIE with K negatives per target, one mini-batch step over a batch of users (BASELINE.json C4:
// "GeoIE ... neg=100"; throughput mode, EXTENSION semantics -- the reference trains one user per call with one negative,
// GeoIE.py:129-194; the mini-batch rule is the reference's own Bpr one, BPR.py:351-397; oracle: geoie_train_batch_k).
//
// Per user (n = L - 1 targets, target i = position i + 1, history j <= i), candidates c = {p_{i+1}, q_{i+1,1..K}}:
//     s(c) = sum_{j<=i} (g[p_j] . h[c]) a dist(p_j, c)^b / (i + 1)          (GeoIE.py:155-159; the t.z term cancels)
//     loss = sum_i sum_k log sigmoid(s(p) - s(q_k)),  cost = -loss + lambda/2 (L2 of every gathered g, h, z row)
// With eps_c = d cost / d s(c) (= sigma(-(s_p - s_qk)) for a negative, -sum_k of those for the positive) and
// w_jc = a dist^b / (i + 1):   d cost / d h[c] = eps_c sum_j w_jc g_j,   d cost / d g_j += eps_c w_jc h[c].
//
// One CTA per user, 8 warps.  G (the user's n history rows of g) sits in shared memory with a padded row stride, and
// column-wise in registers (thread t owns column t of every G row and of the dG accumulator).  Candidates stream
// through in tiles of 16: a warp scores two candidates per pass -- LANE j computes the whole dot g_j . h[c] (and the
// haversine weight for ITS history POI), so no cross-lane reduction is needed per dot -- then all 256 threads turn the
// tile's coefficients into d h[c] (written straight back) and dG (registers).  The distances are recomputed from an fp32
// coordinate table (sin^2 form of the loader's haversine; the reference precomputes n x n matrices on the host).
// A row of g / h that occurs ONCE in the batch is updated in place by the thread block that read it (read once + written
// once = the algorithmic traffic); occurrences of rows that occur several times emit their gradient row and are summed in
// fixed order by rows.cuh (skip_single).  The z rows receive no loss gradient (t.z cancels): their update is the L2 decay
// z -= alpha lambda cnt z, one streaming read-modify-write per unique row by rows.cuh.  No atomics; same bits on a re-run.
#pragma once
#include "common.cuh"
#include "rows.cuh"

constexpr int GEO_TILE = 16;          // candidates per tile (2 per warp)
constexpr int GEO_MAXN = 32;          // history length: one lane per history position

struct GeoBatch {
    const int32_t* P;                 // [Bu x L]       row of g (and of coords_g) for every position
    const int32_t* Ph;                // [Bu x L]       row of h / z (and of coords_h) for every position; == P on one GPU
    const int32_t* Q;                 // [Bu x L x K]   rows of h / z (and of coords_h) of the negatives (position 0 unused)
    const float4* coords_g;           // (lat, lon, cos(lat), 0) in degrees, indexed like P
    const float4* coords_h;           // the same, indexed like Ph / Q; == coords_g on one GPU
    int Bu, L, K;
    // (multi-GPU: the tables are compact copies of the rows the batch touches and P / Ph / Q hold SLOTS, mf_mg.cuh)
};

// keys of the h / z occurrences: o = (b n + i)(K + 1) + c, c = 0 the positive p_{i+1}, c >= 1 the negatives; then the
// g occurrences o_g = b n + j
__global__ void k_geoie_keys(GeoBatch gb, uint32_t* __restrict__ keys_h, uint32_t* __restrict__ keys_g) {
    const int n = gb.L - 1, C = gb.K + 1;
    const int64_t tot = (int64_t)gb.Bu * n * C;
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o < tot) {
        const int c = (int)(o % C); const int64_t bi = o / C; const int i = (int)(bi % n); const int b = (int)(bi / n);
        keys_h[o] = (uint32_t)(c == 0 ? gb.Ph[(size_t)b * gb.L + i + 1] : gb.Q[((size_t)b * gb.L + i + 1) * gb.K + c - 1]);
    }
    if (o < (int64_t)gb.Bu * n) { const int j = (int)(o % n), b = (int)(o / n); keys_g[o] = (uint32_t)gb.P[(size_t)b * gb.L + j]; }
}

// sin x: odd polynomial to x^9 for |x| < 0.5 (relative error < 2e-9: the half-angles of POIs a few km -- or a few thousand
// km -- apart), the library routine beyond.  The hardware approximation (__sinf, absolute error 2^-21) is useless here: the
// angles are ~1e-3 rad and it is their RELATIVE error that sets the distance.
__device__ __forceinline__ float geo_sin(float x) {
    if (fabsf(x) >= 0.5f) return sinf(x);
    const float x2 = x * x;
    return x * (1.f + x2 * (-1.f / 6.f + x2 * (1.f / 120.f + x2 * (-1.f / 5040.f + x2 * (1.f / 362880.f)))));
}
// asin x for x in [0, 1]: odd series to x^9 below 0.25 (relative error < 1e-7), library routine beyond
__device__ __forceinline__ float geo_asin(float x) {
    if (x >= 0.25f) return asinf(x);
    const float x2 = x * x;
    return x * (1.f + x2 * (1.f / 6.f + x2 * (3.f / 40.f + x2 * (15.f / 336.f + x2 * (105.f / 3456.f)))));
}
__device__ __forceinline__ float geo_dist_km(float lat1, float lon1, float coslat1, float lat2, float lon2, float coslat2) {
    // Load_Data_GeoIE.py:28-42: 12742 asin(sqrt(c)), c = sin^2(dlat/2) + cos(lat1) cos(lat2) sin^2(dlon/2) -- the form the
    // loader's docstring states is equivalent to its (1 - cos)/2 expression, without the cancellation in float32
    const float p = 0.017453292519943295f;
    const float sa = geo_sin((lat1 - lat2) * p * 0.5f), sb = geo_sin((lon1 - lon2) * p * 0.5f);
    const float c = sa * sa + coslat1 * coslat2 * sb * sb;
    return 12742.0f * geo_asin(sqrtf(fminf(c, 1.0f)));
}

// dots of both candidates of this warp against history row `lane`; HS = padded row stride of Gs
__device__ __forceinline__ void geo_dots2(const float* __restrict__ Gs, int HS, int H, int lane, const float* __restrict__ h0,
                                          const float* __restrict__ h1, float& d0, float& d1) {
    const float4* g4 = reinterpret_cast<const float4*>(Gs + (size_t)lane * HS);
    const float4* a4 = reinterpret_cast<const float4*>(h0); const float4* b4 = reinterpret_cast<const float4*>(h1);
    // four independent partial sums per candidate (one per float4 component): eight FMA chains in flight per lane -- with
    // one or two CTAs per SM the FMA latency, not the issue rate, bounded the single-accumulator version (ncu: "wait" stalls)
    float4 x0 = f4zero(), x1 = f4zero();
#pragma unroll 4
    for (int k = 0; k < (H >> 2); ++k) {
        const float4 g = g4[k], a = a4[k], b = b4[k];
        x0.x = fmaf(g.x, a.x, x0.x); x0.y = fmaf(g.y, a.y, x0.y); x0.z = fmaf(g.z, a.z, x0.z); x0.w = fmaf(g.w, a.w, x0.w);
        x1.x = fmaf(g.x, b.x, x1.x); x1.y = fmaf(g.y, b.y, x1.y); x1.z = fmaf(g.z, b.z, x1.z); x1.w = fmaf(g.w, b.w, x1.w);
    }
    d0 = (x0.x + x0.y) + (x0.z + x0.w); d1 = (x1.x + x1.y) + (x1.z + x1.w);
}

// lane j: weight pieces for candidate at (clat, clon): pw = d^b, lg = ln d (0 when d = 0, Theano's switch in the gradient
// of pow).  Masked lanes (j > i) return 0.
__device__ __forceinline__ void geo_weight(bool on, float hlat, float hlon, float hcos, float4 cc, float b, float& pw, float& pwlog) {
    pw = 0.f; pwlog = 0.f;
    if (!on) return;
    const float dkm = geo_dist_km(hlat, hlon, hcos, cc.x, cc.y, cc.z);
    if (dkm == 0.f) { pw = b > 0.f ? 0.f : (b == 0.f ? 1.f : INFINITY); return; }      // 0 ** b as powf gives it; log part 0
    // d ** b = 2 ** (b log2 d) on the special-function unit: absolute error 2^-22 on log2 d, i.e. ~1e-7 relative on d ** b
    const float l2 = __log2f(dkm);
    pw = exp2f(b * l2);
    pwlog = pw * l2 * 0.6931471805599453f;
}

template <int NCOL>          // columns per thread: H <= 256 * NCOL
__global__ void __launch_bounds__(256, NCOL == 1 ? 2 : 1)
k_geoie_batch_k(float* __restrict__ g, float* __restrict__ h, const double* __restrict__ ab, int H, GeoBatch gb,
                const uint8_t* __restrict__ single_h, const uint8_t* __restrict__ single_g,
                float alpha, float lambda, float* __restrict__ GH, float* __restrict__ GG,
                double* __restrict__ part /* [grid][3]: loss, d/da, d/db */) {
    extern __shared__ __align__(16) float geo_sm[];
    const int n = gb.L - 1, K = gb.K, C = K + 1, HS = H + 4;
    float* Gs = geo_sm;                                   // [n][HS]
    float* Ht2 = Gs + (size_t)GEO_MAXN * HS;              // [2][GEO_TILE][H] candidate rows: the tile in use + the one streaming in (cp.async)
    float* Hp = Ht2 + (size_t)2 * GEO_TILE * H;           // [H]            the positive's row
    float* coef = Hp + H;                                 // [GEO_TILE][32] eps_c w_jc
    float* wP = coef + GEO_TILE * 32;                     // [32]           w_jp of the positive
    float* sc = wP + 32;                                  // [16] scalars: 0 s_p, 1 A_p, 2 B_p, 3 E ; [8..15] per-warp E partials
    __shared__ int32_t sq[128];                           // the target's K negatives (row ids), staged once per target
    __shared__ int32_t sx[GEO_TILE];                      // row ids of the tile's candidates
    __shared__ uint8_t sfl[GEO_TILE];                     // 1 = the row occurs once in the batch (update in place)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float a = (float)ab[0], b = (float)ab[1];
    double loss_acc = 0.0, ga_acc = 0.0, gb_acc = 0.0;    // meaningful in lane 0 of each warp
    __shared__ double sred[8][3];

    for (int u = blockIdx.x; u < gb.Bu; u += gridDim.x) {
        const int32_t* Pu = gb.P + (size_t)u * gb.L;
        // ---- the user's history rows of g: shared memory (row-major, padded) + this thread's columns in registers ----
        for (int f = tid; f < GEO_MAXN * (H >> 2); f += 256) {          // rows j >= n: zeros (their lanes are masked, not NaN)
            const int j = f / (H >> 2), c4 = f - j * (H >> 2);
            *reinterpret_cast<float4*>(Gs + (size_t)j * HS + 4 * c4) = j < n ? ld4(g + (size_t)Pu[j] * H + 4 * c4) : f4zero();
        }
        float hlat = 0.f, hlon = 0.f, hcos = 1.f;
        const int32_t* Phu = gb.Ph + (size_t)u * gb.L;
        if (lane < n) { const float4 cc = gb.coords_g[Pu[lane]]; hlat = cc.x; hlon = cc.y; hcos = cc.z; }
        __syncthreads();
        float Gcol[NCOL][GEO_MAXN], dG[NCOL][GEO_MAXN];
#pragma unroll
        for (int q = 0; q < NCOL; ++q) {
            const int col = tid + 256 * q;
#pragma unroll
            for (int j = 0; j < GEO_MAXN; ++j) { Gcol[q][j] = (col < H && j < n) ? Gs[(size_t)j * HS + col] : 0.f; dG[q][j] = 0.f; }
        }
        for (int i = 0; i < n; ++i) {
            const float inv = 1.0f / (float)(i + 1);
            const bool on = lane <= i;
            const size_t occ0 = ((size_t)u * n + i) * C;                 // occurrence id of the positive of this target
            const int32_t* Qi = gb.Q + ((size_t)u * gb.L + i + 1) * K;
            for (int k = tid; k < K; k += 256) sq[k] = Qi[k];
            // this warp's rows of tile `t0` -> buffer `bf`, asynchronously (16 B per lane and request)
            auto prefetch = [&](int t0, int bf) {
                float* hb = Ht2 + ((size_t)bf * GEO_TILE + 2 * warp) * H;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int k = t0 + 2 * warp + r;
                    if (k < K) {
                        const float* src = h + (size_t)Qi[k] * H;
                        for (int c4 = lane; c4 < (H >> 2); c4 += 32)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(hb + (size_t)r * H + 4 * c4)), "l"(src + 4 * c4) : "memory");
                    } else {
                        for (int c4 = lane; c4 < (H >> 2); c4 += 32) *reinterpret_cast<float4*>(hb + (size_t)r * H + 4 * c4) = f4zero();
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            prefetch(0, 0);
            // ---- the positive: warp 0 scores it; its row stays in Hp until the target's negatives are done ----
            if (warp == 0) {
                const int32_t x = Phu[i + 1];
                for (int c4 = lane; c4 < (H >> 2); c4 += 32) *reinterpret_cast<float4*>(Hp + 4 * c4) = ld4(h + (size_t)x * H + 4 * c4);
                __syncwarp();
                float d0, d1;
                geo_dots2(Gs, HS, H, lane, Hp, Hp, d0, d1);
                float pw, pwl;
                geo_weight(on, hlat, hlon, hcos, gb.coords_h[x], b, pw, pwl);
                const float w = a * pw * inv;
                wP[lane] = w;
                const float s = warp_sum(on ? d0 * w : 0.f), A = warp_sum(on ? d0 * pw * inv : 0.f), Bv = warp_sum(on ? d0 * a * pwl * inv : 0.f);
                if (lane == 0) { sc[0] = s; sc[1] = A; sc[2] = Bv; }
            }
            __syncthreads();
            const float s_p = sc[0];
            float E_w = 0.f;                                              // this warp's sum of eps over its negatives
            // ---- negatives, GEO_TILE per pass ----
            for (int t0 = 0, bf = 0; t0 < K; t0 += GEO_TILE, bf ^= 1) {
                const int k0 = t0 + 2 * warp, k1 = k0 + 1;               // this warp's two candidates (negative numbers)
                const bool v0 = k0 < K, v1 = k1 < K;
                const int32_t x0 = v0 ? sq[k0] : 0, x1 = v1 ? sq[k1] : 0;
                float* Ht = Ht2 + (size_t)bf * GEO_TILE * H;
                float* h0 = Ht + (size_t)(2 * warp) * H; float* h1 = h0 + H;
                asm volatile("cp.async.wait_group 0;" ::: "memory");       // this warp's two rows of the tile have landed
                __syncwarp();
                if (t0 + GEO_TILE < K) prefetch(t0 + GEO_TILE, bf ^ 1);     // next tile streams in behind the compute
                float d0, d1;
                geo_dots2(Gs, HS, H, lane, h0, h1, d0, d1);
                float pw0, pwl0, pw1, pwl1;
                geo_weight(on && v0, hlat, hlon, hcos, gb.coords_h[x0], b, pw0, pwl0);
                geo_weight(on && v1, hlat, hlon, hcos, gb.coords_h[x1], b, pw1, pwl1);
                const float w0 = a * pw0 * inv, w1 = a * pw1 * inv;
                const float s0 = warp_sum(d0 * w0), s1 = warp_sum(d1 * w1);          // masked lanes have w = 0
                const float A0 = warp_sum(d0 * pw0 * inv), A1 = warp_sum(d1 * pw1 * inv);
                const float B0 = warp_sum(d0 * a * pwl0 * inv), B1 = warp_sum(d1 * a * pwl1 * inv);
                const float xk0 = s_p - s0, xk1 = s_p - s1;
                const float e0 = v0 ? sigmoidf_(-xk0) : 0.f, e1 = v1 ? sigmoidf_(-xk1) : 0.f;   // d cost / d s(q_k)
                coef[(2 * warp) * 32 + lane] = e0 * w0; coef[(2 * warp + 1) * 32 + lane] = e1 * w1;
                if (lane == 0) {
                    sx[2 * warp] = x0; sx[2 * warp + 1] = x1;
                    sfl[2 * warp] = v0 ? single_h[occ0 + 1 + k0] : 0; sfl[2 * warp + 1] = v1 ? single_h[occ0 + 1 + k1] : 0;
                }
                E_w += e0 + e1;
                if (lane == 0) {
                    if (v0) loss_acc += (double)logsigmoidf_(xk0);
                    if (v1) loss_acc += (double)logsigmoidf_(xk1);
                    ga_acc += (double)(e0 * A0 + e1 * A1); gb_acc += (double)(e0 * B0 + e1 * B1);
                }
                __syncthreads();
                // ---- all threads: column t of d h[c] for the tile's candidates, and the dG accumulators ----
                const int nc = min(GEO_TILE, K - t0);
#pragma unroll
                for (int q = 0; q < NCOL; ++q) {
                    const int col = tid + 256 * q;
                    if (col < H) {
                        for (int cw = 0; cw < nc; ++cw) {
                            const float hv = Ht[(size_t)cw * H + col];
                            const float4* cf4 = reinterpret_cast<const float4*>(coef + cw * 32);
                            float4 dh4 = f4zero();                       // four independent chains for d h[c]
#pragma unroll
                            for (int j4 = 0; j4 < GEO_MAXN / 4; ++j4) {
                                if (4 * j4 > i) break;                   // rows j > i carry a zero coefficient
                                const float4 cf = cf4[j4];
                                dG[q][4 * j4 + 0] = fmaf(cf.x, hv, dG[q][4 * j4 + 0]); dh4.x = fmaf(cf.x, Gcol[q][4 * j4 + 0], dh4.x);
                                dG[q][4 * j4 + 1] = fmaf(cf.y, hv, dG[q][4 * j4 + 1]); dh4.y = fmaf(cf.y, Gcol[q][4 * j4 + 1], dh4.y);
                                dG[q][4 * j4 + 2] = fmaf(cf.z, hv, dG[q][4 * j4 + 2]); dh4.z = fmaf(cf.z, Gcol[q][4 * j4 + 2], dh4.z);
                                dG[q][4 * j4 + 3] = fmaf(cf.w, hv, dG[q][4 * j4 + 3]); dh4.w = fmaf(cf.w, Gcol[q][4 * j4 + 3], dh4.w);
                            }
                            const float dh = (dh4.x + dh4.y) + (dh4.z + dh4.w);
                            const size_t o = occ0 + 1 + t0 + cw;
                            const size_t x = (size_t)sx[cw];
                            if (sfl[cw]) h[x * H + col] = hv - alpha * (dh + lambda * hv);
                            else GH[o * H + col] = dh;
                        }
                    }
                }
                __syncthreads();
            }
            // ---- the positive's backward: eps_p = -sum_k eps_k ----
            if (lane == 0) sc[8 + warp] = E_w;
            __syncthreads();
            float E = 0.f;
            for (int ww = 0; ww < 8; ++ww) E += sc[8 + ww];               // warp order: fixed
            if (tid == 0) { ga_acc -= (double)(E * sc[1]); gb_acc -= (double)(E * sc[2]); }
            {
                const bool single = single_h[occ0] != 0;
                const size_t x = (size_t)Phu[i + 1];
#pragma unroll
                for (int q = 0; q < NCOL; ++q) {
                    const int col = tid + 256 * q;
                    if (col < H) {
                        const float hv = Hp[col];
                        float dh = 0.f;
#pragma unroll
                        for (int j = 0; j < GEO_MAXN; ++j) {
                            if (j > i) break;
                            const float cf = -E * wP[j];
                            dG[q][j] = fmaf(cf, hv, dG[q][j]); dh = fmaf(cf, Gcol[q][j], dh);
                        }
                        if (single) h[x * H + col] = hv - alpha * (dh + lambda * hv);
                        else GH[occ0 * H + col] = dh;
                    }
                }
            }
            __syncthreads();
        }
        // ---- g rows of the user's history ----
#pragma unroll
        for (int q = 0; q < NCOL; ++q) {
            const int col = tid + 256 * q;
            if (col < H) {
#pragma unroll
                for (int j = 0; j < GEO_MAXN; ++j) {
                    if (j < n) {
                        const size_t og = (size_t)u * n + j;
                        if (single_g[og])
                            g[(size_t)Pu[j] * H + col] = Gcol[q][j] - alpha * (dG[q][j] + lambda * Gcol[q][j]);
                        else GG[og * H + col] = dG[q][j];
                    }
                }
            }
        }
        __syncthreads();          // Gs is rewritten by the next user
    }
    if (lane == 0) { sred[warp][0] = loss_acc; sred[warp][1] = ga_acc; sred[warp][2] = gb_acc; }
    __syncthreads();
    if (tid < 3) { double t = 0.0; for (int ww = 0; ww < 8; ++ww) t += sred[ww][tid]; part[(size_t)blockIdx.x * 3 + tid] = t; }
}

// single[o] = 1 iff occurrence o is the only one of its row in the batch
__global__ void k_mark_single(SegList seg, uint8_t* __restrict__ single) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= seg.n) return;
    const uint32_t sl = seg.seg_of_occ[q];
    single[q] = seg.seg_start[sl + 1] - seg.seg_start[sl] == 1u ? 1 : 0;
}

static size_t geoie_k_smem(int H) {
    return ((size_t)GEO_MAXN * (H + 4) + (size_t)2 * GEO_TILE * H + H + GEO_TILE * 32 + 32 + 16) * sizeof(float);
}

// a, b <- a, b - alpha * (sum over CTAs of the partials), loss out; fixed order
__global__ void k_geoie_k_finalize(const double* __restrict__ part, int nblocks, double* ab, float alpha, double* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double loss = 0.0, ga = 0.0, gb = 0.0;
    for (int i = 0; i < nblocks; ++i) { loss += part[(size_t)i * 3]; ga += part[(size_t)i * 3 + 1]; gb += part[(size_t)i * 3 + 2]; }
    out[0] = loss;
    if (!ab) { out[1] = ga; out[2] = gb; return; }   // multi-GPU: the caller sums the ranks' partials first
    ab[0] -= (double)alpha * ga;                     // params = [a, b], no L2 on them (GeoIE.py:91,172-173)
    ab[1] -= (double)alpha * gb;
}
