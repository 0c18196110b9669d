// prme_k.cuh -- PRME with K negatives per positive (BASELINE.json C3: "PRME ... neg=20").
//
// The reference draws ONE negative per check-in (PRME.py:172-219, prog_prme.py:191-197); K > 1 is the
// driver-defined generalisation SURVEY.md 8(a6) allows, stated in oracle/models.py:
//   pqidx = [p, q_1 .. q_K, prev]
//   D(x)  = ||du[u] - dp[x]||^2                                        if gap > thd
//         = w (cw ||du[u] - dp[x]||^2 + (1 - cw) ||ds[x] - ds[prev]||^2), w = (1 + dist)^0.25   otherwise
//   upq   = sum_k log sigmoid(D(q_k) - D(p));   objective upq - lambda/2 (|du|^2 + sum |dp[pq]|^2 + sum |ds[pq]|^2), ASCENT
// With c_0 = -G, c_k = g_k = sigmoid(-(D(q_k) - D(p))), G = sum_k g_k and (cp, cs) = (1, 0) or (w cw, w (1 - cw)):
//   d/d dp[x_j] = c_j 2 cp (dp[x_j] - du)        d/d ds[x_j] = c_j 2 cs (ds[x_j] - ds[prev])       (j = 0 .. K)
//   d/d du      = 2 cp (G dp[p] - sum_k g_k dp[q_k])        d/d ds[prev] = 2 cs (G ds[p] - sum_k g_k ds[q_k])
//   d/d dp[prev] = 0 (L2 only)
// Two kernels:
//   k_prme_seq_k   -- parity mode: the ordered list of check-ins inside ONE CTA, every step from the values the previous
//                     step left (exact sequential SGD); rows written back in pqidx order, last occurrence wins
//                     (PRME.py:206-208).  At K = 1 it is the reference step (checker: oracle obo_prme_train_k).
//   k_prme_score + k_prme_apply -- throughput mode (EXTENSION semantics, like SpatialGru): N check-ins from pre-update
//                     values, gradients summed over duplicate occurrences (the reference's own mini-batch rule,
//                     BPR.py:385-387).  Phase A (one CTA per check-in) streams the check-in's 2(K+2)+1 rows once and
//                     leaves only SCALARS per occurrence (c_j 2 cp, c_j 2 cs) plus three rows per check-in (a copy of
//                     ds[prev], d/d du, d/d ds[prev]); it writes no table row.  Phase B (one warp per UNIQUE row of the
//                     batch, dp and ds together -- they share the key list) rebuilds each occurrence's gradient from
//                     those scalars, sums the occurrences in ascending order (fixed order, no atomics) and does the one
//                     read-modify-write of the row.  HBM traffic: every gathered row read once, every unique row
//                     written once; the per-check-in rows phase B re-reads (du[u], the ds[prev] copy) are L2 hits.
#pragma once
#include "common.cuh"
#include "rows.cuh"

// ------------------------------------------------------------------------------------------------------------------
// parity mode
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_prme_seq_k(float* du, float* dp, float* ds, int d, int K, const int32_t* __restrict__ us,
             const int32_t* __restrict__ ps, const int32_t* __restrict__ qs, const int32_t* __restrict__ prevs,
             const double* __restrict__ dist, const int32_t* __restrict__ gap, int64_t n, int thd, float cw,
             float alpha, float lambda, double* __restrict__ loss) {
    extern __shared__ __align__(16) float prme_sm[];
    float* sm = prme_sm;
    const int R = K + 2, d4 = d >> 2;
    float* U = sm;                                  // du[u]                         [d]
    float* RP = U + d;                              // dp rows p, q_1..q_K, prev     [R x d]
    float* RS = RP + (size_t)R * d;                 // ds rows                       [R x d]
    float* Dv = RS + (size_t)R * d;                 // D(x_j), j = 0..K              [K + 1]  (padded to R)
    float* gv = Dv + R;                             // g_k (gv[0] = G)               [K + 1]
    float* lv = gv + R;                             // log sigmoid terms             [K + 1]
    int32_t* idx = reinterpret_cast<int32_t*>(lv + R);       // row ids               [R]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int64_t i = 0; i < n; ++i) {
        if (tid < R) idx[tid] = tid == 0 ? ps[i] : (tid <= K ? qs[i * K + tid - 1] : prevs[i]);
        __syncthreads();
        const int32_t uu = us[i];
        for (int f = tid; f < (2 * R + 1) * d4; f += 256) {
            const int r = f / d4, c = f - r * d4;
            const float* src = r == 0 ? du + (size_t)uu * d
                                      : (r <= R ? dp + (size_t)idx[r - 1] * d : ds + (size_t)idx[r - 1 - R] * d);
            reinterpret_cast<float4*>(sm)[f] = ld4(src + 4 * c);       // smem layout = [U | RP | RS], r-major
        }
        __syncthreads();
        const bool far = gap[i] > thd;                                  // ifelse(T.gt(tidx, thd)), PRME.py:192
        const float w = (float)pow(1.0 + dist[i], 0.25);                // PRME.py:191
        const float cp = far ? 1.f : w * cw, cs = far ? 0.f : w * (1.f - cw);
        const float* SL = RS + (size_t)(K + 1) * d;
        for (int j = warp; j <= K; j += 8) {
            const float* rp = RP + (size_t)j * d; const float* rs = RS + (size_t)j * d;
            float a = 0.f, b = 0.f;
            for (int c = lane; c < d; c += 32) {
                const float x = U[c] - rp[c], y = rs[c] - SL[c];
                a = fmaf(x, x, a); b = fmaf(y, y, b);
            }
            a = warp_sum(a); b = warp_sum(b);
            if (lane == 0) Dv[j] = cp * a + cs * b;
        }
        __syncthreads();
        if (tid >= 1 && tid <= K) { const float x = Dv[tid] - Dv[0]; gv[tid] = sigmoidf_(-x); lv[tid] = logsigmoidf_(x); }
        __syncthreads();
        if (tid == 0) {
            float G = 0.f; double ls = 0.0;
            for (int k = 1; k <= K; ++k) { G += gv[k]; ls += (double)lv[k]; }
            gv[0] = G; loss[i] = ls;
        }
        __syncthreads();
        const float G = gv[0];
        for (int c = tid; c < d; c += 256) {
            const float u = U[c], pp = RP[c], sp = RS[c], sl = SL[c], pl = RP[(size_t)(K + 1) * d + c];
            float au = G * pp, as = G * sp;
            for (int k = 1; k <= K; ++k) { au = fmaf(-gv[k], RP[(size_t)k * d + c], au); as = fmaf(-gv[k], RS[(size_t)k * d + c], as); }
            du[(size_t)uu * d + c] = u + alpha * (2.f * cp * au - lambda * u);
            // rows go back in pqidx order [p, q_1..q_K, prev]; a later duplicate overwrites an earlier one (same thread,
            // same address, program order): last writer wins
            dp[(size_t)idx[0] * d + c] = pp + alpha * (2.f * G * cp * (u - pp) - lambda * pp);
            for (int k = 1; k <= K; ++k) {
                const float pq = RP[(size_t)k * d + c];
                dp[(size_t)idx[k] * d + c] = pq + alpha * (-2.f * gv[k] * cp * (u - pq) - lambda * pq);
            }
            dp[(size_t)idx[K + 1] * d + c] = pl + alpha * (-lambda * pl);
            ds[(size_t)idx[0] * d + c] = sp + alpha * (-2.f * G * cs * (sp - sl) - lambda * sp);
            for (int k = 1; k <= K; ++k) {
                const float sq = RS[(size_t)k * d + c];
                ds[(size_t)idx[k] * d + c] = sq + alpha * (2.f * gv[k] * cs * (sq - sl) - lambda * sq);
            }
            ds[(size_t)idx[K + 1] * d + c] = sl + alpha * (2.f * cs * as - lambda * sl);
        }
        __syncthreads();          // this step's rows are visible to the loads of the next one
    }
}

static size_t prme_seq_k_smem(int d, int K) {
    const int R = K + 2;
    return ((size_t)(2 * R + 1) * d + 4 * (size_t)R) * sizeof(float) + 16;
}

// ------------------------------------------------------------------------------------------------------------------
// throughput mode
// ------------------------------------------------------------------------------------------------------------------
struct PrmeBatchIdx {
    const int32_t* u; const int32_t* p; const int32_t* q; const int32_t* prev;    // device: [N], [N], [N x K], [N]
    const float* dist; const int32_t* gap;                                        // device: [N] km, [N] minutes
    int N, K;
};

// occurrence o = i * (K + 2) + j  ->  row id (j = 0: p, 1..K: q_k, K+1: prev): the key list shared by dp and ds
__global__ void k_prme_keys(PrmeBatchIdx b, uint32_t* __restrict__ keys) {
    const int R = b.K + 2;
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= (int64_t)b.N * R) return;
    const int i = (int)(o / R), j = (int)(o - (int64_t)i * R);
    keys[o] = (uint32_t)(j == 0 ? b.p[i] : (j <= b.K ? b.q[(size_t)i * b.K + j - 1] : b.prev[i]));
}

constexpr int PRME_MAXK = 127;

// Phase A.  grid-stride over check-ins, 8 warps per CTA: warp w takes candidates j = w, w + 8, ... (j = 0: p, 1..K: q_k);
// a lane owns NCH float4 columns of every row.  Outputs per check-in i (R = K + 2 occurrences o = i R + j):
//   KP[o], KS[o]   c_j 2 cp, c_j 2 cs  (j <= K; the prev occurrence j = K + 1 gets KP = KS = 0)
//   SL[i]          copy of ds[prev_i] (pre-update)           [N x d]
//   GU[i]          -(d upq / d du[u_i])                       [N x d]  (descent form for rows.cuh)
//   GL[i]          d upq / d ds[prev_i]                       [N x d]
template <int NCH>
__global__ void __launch_bounds__(256)
k_prme_score(const float* __restrict__ du, const float* __restrict__ dp, const float* __restrict__ ds, int d4, PrmeBatchIdx b,
             int thd, float cw, float* __restrict__ KP, float* __restrict__ KS, float* __restrict__ SL,
             float* __restrict__ GU, float* __restrict__ GL, double* __restrict__ part) {
    extern __shared__ __align__(16) float4 prme_red[];
    float4* red = prme_red;                         // [8][2][d4]: per-warp partial sums for du and ds[prev]
    __shared__ float sD[PRME_MAXK + 1], sg[PRME_MAXK + 1], sl_[PRME_MAXK + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = b.K, R = K + 2;
    double loss_acc = 0.0;
    for (int i = blockIdx.x; i < b.N; i += gridDim.x) {
        const int32_t uu = b.u[i], xp = b.p[i], xl = b.prev[i];
        const bool far = b.gap[i] > thd;
        const float w = sqrtf(sqrtf(1.0f + b.dist[i]));
        const float cp = far ? 1.f : w * cw, cs = far ? 0.f : w * (1.f - cw);
        float4 u[NCH], sl[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int c = lane + 32 * k;
            if (c < d4) { u[k] = ldg4(du + ((size_t)uu * d4 + c) * 4); sl[k] = ldg4(ds + ((size_t)xl * d4 + c) * 4); }
        }
        // ---- pass 1: D(x_j) ----
        for (int j = warp; j <= K; j += 8) {
            const int32_t x = j == 0 ? xp : b.q[(size_t)i * K + j - 1];
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int c = lane + 32 * k;
                if (c < d4) {
                    const float4 a = f4sub(u[k], ldg4(dp + ((size_t)x * d4 + c) * 4));
                    const float4 s = f4sub(ldg4(ds + ((size_t)x * d4 + c) * 4), sl[k]);
                    acc += cp * (a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w) + cs * (s.x * s.x + s.y * s.y + s.z * s.z + s.w * s.w);
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) sD[j] = acc;
        }
        __syncthreads();
        if (tid >= 1 && tid <= K) { const float x = sD[tid] - sD[0]; sg[tid] = sigmoidf_(-x); sl_[tid] = logsigmoidf_(x); }
        __syncthreads();
        float G = 0.f;
        for (int k = 1; k <= K; ++k) G += sg[k];                       // same fixed order in every thread
        if (tid == 0) { double ls = 0.0; for (int k = 1; k <= K; ++k) ls += (double)sl_[k]; loss_acc += ls; }
        if (tid <= K + 1) {                                             // the scalars phase B needs
            const float cj = tid == 0 ? -G : (tid <= K ? sg[tid] : 0.f);
            KP[(size_t)i * R + tid] = cj * 2.f * cp; KS[(size_t)i * R + tid] = cj * 2.f * cs;
        }
        // ---- pass 2 (rows come back from L1 / L2): d/d du = sum_j c_j 2 cp (du - dp[x_j]), d/d ds[prev] likewise ----
        float4 au[NCH], as[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k) { au[k] = f4zero(); as[k] = f4zero(); }
        for (int j = warp; j <= K; j += 8) {
            const int32_t x = j == 0 ? xp : b.q[(size_t)i * K + j - 1];
            const float cj = j == 0 ? -G : sg[j];
            const float kp = cj * 2.f * cp, ks = cj * 2.f * cs;
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int c = lane + 32 * k;
                if (c < d4) {
                    const float4 ep = f4sub(u[k], ldg4(dp + ((size_t)x * d4 + c) * 4));
                    const float4 es = f4sub(sl[k], ldg4(ds + ((size_t)x * d4 + c) * 4));
                    au[k] = f4fma(kp, ep, au[k]); as[k] = f4fma(ks, es, as[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int c = lane + 32 * k;
            if (c < d4) { red[(warp * 2 + 0) * d4 + c] = au[k]; red[(warp * 2 + 1) * d4 + c] = as[k]; }
        }
        __syncthreads();
        if (warp < 2) {
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int c = lane + 32 * k;
                if (c >= d4) continue;
                float4 t = red[(0 * 2 + warp) * d4 + c];
#pragma unroll
                for (int ww = 1; ww < 8; ++ww) t = f4add(t, red[(ww * 2 + warp) * d4 + c]);     // warp order: fixed
                if (warp == 0) st4(GU + ((size_t)i * d4 + c) * 4, make_float4(-t.x, -t.y, -t.z, -t.w));
                else { st4(GL + ((size_t)i * d4 + c) * 4, t); st4(SL + ((size_t)i * d4 + c) * 4, sl[k]); }
            }
        }
        __syncthreads();          // red / sD / sg are reused by the next check-in
    }
    if (tid == 0) part[blockIdx.x] = loss_acc;
}

// Phase A, one WARP per check-in, one pass, no block barrier (the CTA-per-check-in kernels above / below spend most of their
// time in the four __syncthreads of a check-in: ncu, barrier-stall bound at 29 % of the issue slots).  What makes one pass
// possible: c_j = sigmoid(D_0 - D_j) needs only the positive's distance, and
//     sum_j c_j 2cp (du - dp[x_j])  with  c_0 = -G, G = sum_{j>=1} c_j      =   2cp sum_{j>=1} c_j (dp[x_0] - dp[x_j])
// (likewise for ds[prev]), so with the positive's two rows kept in registers every negative is loaded, scored and folded
// into the two gradient rows while it is still in registers: each row is read exactly once, nothing is staged in shared
// memory.  A lane owns NCH float4 columns; the K candidate ids are read lane-parallel and broadcast by shuffles; UN
// candidates (4 NCH UN float4 per lane) are in flight together.  Same outputs as k_prme_score.
template <int NCH>
__global__ void __launch_bounds__(256, NCH <= 2 ? 2 : 1)
k_prme_score_warp(const float* __restrict__ du, const float* __restrict__ dp, const float* __restrict__ ds, int d4, PrmeBatchIdx b,
                  int thd, float cw, float* __restrict__ KP, float* __restrict__ KS, float* __restrict__ SL,
                  float* __restrict__ GU, float* __restrict__ GL, double* __restrict__ part) {
    __shared__ double sloss[8];
    constexpr int UN = NCH <= 2 ? 2 : 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = b.K, R = K + 2;
    double loss_acc = 0.0;                                  // lane 0 only
    for (int i = blockIdx.x * 8 + warp; i < b.N; i += gridDim.x * 8) {
        const int32_t uu = b.u[i], xp = b.p[i], xl = b.prev[i];
        const bool far = b.gap[i] > thd;
        const float w = sqrtf(sqrtf(1.0f + b.dist[i]));
        const float cp = far ? 1.f : w * cw, cs = far ? 0.f : w * (1.f - cw);
        int32_t qid[(PRME_MAXK + 31) / 32];
#pragma unroll
        for (int t = 0; t < (PRME_MAXK + 31) / 32; ++t) qid[t] = lane + 32 * t < K ? b.q[(size_t)i * K + lane + 32 * t] : 0;
        float4 u[NCH], sl[NCH], p0[NCH], s0[NCH], au[NCH], as[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int c = lane + 32 * k;
            au[k] = f4zero(); as[k] = f4zero();
            if (c < d4) {
                u[k] = ldg4(du + ((size_t)uu * d4 + c) * 4); sl[k] = ldg4(ds + ((size_t)xl * d4 + c) * 4);
                p0[k] = ldg4(dp + ((size_t)xp * d4 + c) * 4); s0[k] = ldg4(ds + ((size_t)xp * d4 + c) * 4);
            }
        }
        float D0 = 0.f;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            if (lane + 32 * k < d4) {
                const float4 a = f4sub(u[k], p0[k]), q = f4sub(s0[k], sl[k]);
                D0 += cp * (a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w) + cs * (q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
            }
        }
        D0 = warp_sum(D0);
        float G = 0.f;
#pragma unroll
        for (int t = 0; t < (PRME_MAXK + 31) / 32; ++t) {
            if (32 * t >= K) break;
            const int nl = min(32, K - 32 * t);
            for (int l0 = 0; l0 < nl; l0 += UN) {
                float4 P[UN][NCH], S[UN][NCH];
#pragma unroll
                for (int v = 0; v < UN; ++v) {
                    const int32_t x = __shfl_sync(0xffffffffu, qid[t], min(l0 + v, nl - 1));
#pragma unroll
                    for (int k = 0; k < NCH; ++k) {
                        const int c = lane + 32 * k;
                        if (c < d4) { P[v][k] = ldg4(dp + ((size_t)x * d4 + c) * 4); S[v][k] = ldg4(ds + ((size_t)x * d4 + c) * 4); }
                    }
                }
#pragma unroll
                for (int v = 0; v < UN; ++v) {
                    if (l0 + v >= nl) break;
                    float acc = 0.f;
#pragma unroll
                    for (int k = 0; k < NCH; ++k) {
                        if (lane + 32 * k < d4) {
                            const float4 a = f4sub(u[k], P[v][k]), q = f4sub(S[v][k], sl[k]);
                            acc += cp * (a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w) + cs * (q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
                        }
                    }
                    acc = warp_sum(acc);
                    const float xx = acc - D0, g = sigmoidf_(-xx);
                    const int j = 32 * t + l0 + v + 1;
                    if (lane == 0) {
                        loss_acc += (double)logsigmoidf_(xx);
                        KP[(size_t)i * R + j] = g * 2.f * cp; KS[(size_t)i * R + j] = g * 2.f * cs;
                    }
                    G += g;
#pragma unroll
                    for (int k = 0; k < NCH; ++k) {
                        if (lane + 32 * k < d4) { au[k] = f4fma(g, f4sub(p0[k], P[v][k]), au[k]); as[k] = f4fma(g, f4sub(s0[k], S[v][k]), as[k]); }
                    }
                }
            }
        }
        if (lane == 0) {
            KP[(size_t)i * R] = -G * 2.f * cp; KS[(size_t)i * R] = -G * 2.f * cs;
            KP[(size_t)i * R + K + 1] = 0.f; KS[(size_t)i * R + K + 1] = 0.f;
        }
        const float kp2 = 2.f * cp, ks2 = 2.f * cs;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int c = lane + 32 * k;
            if (c < d4) {
                st4(GU + ((size_t)i * d4 + c) * 4, make_float4(-kp2 * au[k].x, -kp2 * au[k].y, -kp2 * au[k].z, -kp2 * au[k].w));
                st4(GL + ((size_t)i * d4 + c) * 4, make_float4(ks2 * as[k].x, ks2 * as[k].y, ks2 * as[k].z, ks2 * as[k].w));
                st4(SL + ((size_t)i * d4 + c) * 4, sl[k]);
            }
        }
    }
    if (lane == 0) sloss[warp] = loss_acc;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int ww = 0; ww < 8; ++ww) t += sloss[ww]; part[blockIdx.x] = t; }
}

// Phase B.  One warp per unique row r of the batch; dp[r] and ds[r] are updated together (same key list).  Occurrence
// o = i R + j of r contributes  KP[o] (dp[r] - du[u_i])  to d upq / d dp[r]  and  KS[o] (ds[r] - SL[i])  (j <= K) or GL[i]
// (j = K + 1, the prev occurrence) to d upq / d ds[r]; ascent with the L2 term once per occurrence:
//   row <- row + alpha (sum_occ grad - lambda cnt row).
// du is read-only here (its own update runs afterwards), SL is the pre-update copy: every term is from pre-update values.
template <int NCH>
__global__ void __launch_bounds__(256)
k_prme_apply(SegList seg, const float* __restrict__ du, float* __restrict__ dp, float* __restrict__ ds, int d4, PrmeBatchIdx b,
             const float* __restrict__ KP, const float* __restrict__ KS, const float* __restrict__ SL,
             const float* __restrict__ GL, float alpha, float lambda) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t nu = *seg.n_unique;
    const int R = b.K + 2;
    for (int64_t sg = warp; sg < nu; sg += nwarps) {
        const uint32_t s0 = seg.seg_start[sg], s1 = seg.seg_start[sg + 1];
        const size_t r = seg.uniq[sg];
        float4 vp[NCH], vs[NCH], gp[NCH], gs[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int c = lane + 32 * k;
            gp[k] = f4zero(); gs[k] = f4zero();
            if (c < d4) { vp[k] = ld4(dp + (r * d4 + c) * 4); vs[k] = ld4(ds + (r * d4 + c) * 4); }
        }
        // Occurrences in ascending id (fixed summation order), 32 at a time: LANE l fetches the metadata of occurrence l
        // (record number -> scalars, user) so that those dependent loads run in parallel across the occurrences; the rows
        // of UN occurrences are then requested together before they are consumed in order.
        for (uint32_t base = s0; base < s1; base += 32) {
            const int cnt = (int)min(32u, s1 - base);
            uint32_t mi = 0; int mtype = 0; float mkp = 0.f, mks = 0.f; uint32_t mu = 0;
            if (lane < cnt) {
                const uint32_t o = seg.vals[base + lane];
                mi = o / (uint32_t)R;
                const uint32_t j = o - mi * (uint32_t)R;
                if ((int)j == R - 1) mtype = 1;                            // the prev occurrence: gradient row GL[i] for ds, none for dp
                else { mkp = KP[o]; mks = KS[o]; mu = (uint32_t)b.u[mi]; }
            }
            constexpr int UN = NCH <= 1 ? 4 : 2;                           // occurrences whose rows are in flight together
            for (int q0 = 0; q0 < cnt; q0 += UN) {
                float4 ra[UN][NCH], rb[UN][NCH];
                float kp[UN], ks[UN]; int ty[UN];
#pragma unroll
                for (int t = 0; t < UN; ++t) {
                    const int q = min(q0 + t, cnt - 1);
                    const uint32_t i = __shfl_sync(0xffffffffu, mi, q), uu = __shfl_sync(0xffffffffu, mu, q);
                    ty[t] = q0 + t < cnt ? __shfl_sync(0xffffffffu, mtype, q) : 2;       // 2 = past the end
                    kp[t] = __shfl_sync(0xffffffffu, mkp, q); ks[t] = __shfl_sync(0xffffffffu, mks, q);
#pragma unroll
                    for (int k = 0; k < NCH; ++k) {
                        const int c = lane + 32 * k;
                        if (c < d4 && ty[t] != 2) {
                            if (ty[t] == 1) rb[t][k] = ldg4(GL + ((size_t)i * d4 + c) * 4);
                            else {
                                ra[t][k] = ldg4(du + ((size_t)uu * d4 + c) * 4);
                                if (ks[t] != 0.f) rb[t][k] = ldg4(SL + ((size_t)i * d4 + c) * 4);
                            }
                        }
                    }
                }
#pragma unroll
                for (int t = 0; t < UN; ++t) {
                    if (ty[t] == 2) continue;
#pragma unroll
                    for (int k = 0; k < NCH; ++k) {
                        const int c = lane + 32 * k;
                        if (c < d4) {
                            if (ty[t] == 1) gs[k] = f4add(gs[k], rb[t][k]);
                            else {
                                gp[k] = f4fma(kp[t], f4sub(vp[k], ra[t][k]), gp[k]);
                                if (ks[t] != 0.f) gs[k] = f4fma(ks[t], f4sub(vs[k], rb[t][k]), gs[k]);
                            }
                        }
                    }
                }
            }
        }
        const float lc = lambda * (float)(s1 - s0);
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int c = lane + 32 * k;
            if (c < d4) {
                const float4 a = vp[k], z = vs[k];
                st4(dp + (r * d4 + c) * 4, make_float4(a.x + alpha * (gp[k].x - lc * a.x), a.y + alpha * (gp[k].y - lc * a.y),
                                                        a.z + alpha * (gp[k].z - lc * a.z), a.w + alpha * (gp[k].w - lc * a.w)));
                st4(ds + (r * d4 + c) * 4, make_float4(z.x + alpha * (gs[k].x - lc * z.x), z.y + alpha * (gs[k].y - lc * z.y),
                                                        z.z + alpha * (gs[k].z - lc * z.z), z.w + alpha * (gs[k].w - lc * z.w)));
            }
        }
    }
}

// fixed-order sum of n doubles by one warp: lane l adds part[l], part[l + 32], ... in order, then a fixed shuffle tree
__global__ void k_sum_partials_warp(const double* __restrict__ part, int n, double* __restrict__ out) {
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) t += part[i];
    t = warp_sum_d(t);
    if (threadIdx.x == 0) *out = t;
}
