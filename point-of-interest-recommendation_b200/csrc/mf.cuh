// mf.cuh -- pairwise-ranking embedding models trained one check-in at a time:
//   OboBpr  (BPR.py:201-241)   u . (lt[p] - lt[q]),        descent
//   OboPrme (PRME.py:172-219)  metric embedding, gated by time gap, ascent
// The reference performs one Theano call per check-in, each seeing the previous update.  Here the
// whole ordered list of check-ins runs inside ONE warp: lane l owns the same columns of every row,
// so a row written in step i and read in step i+1 is a same-thread RAW through memory -> exact
// sequential-SGD semantics with no synchronisation, and "last writer wins" for duplicate indices
// falls out of program order (BPR.py:228-230, PRME.py:206-208).
#pragma once
#include "common.cuh"

template <int NCH>
__global__ void __launch_bounds__(32)
k_bpr_seq(float* ux, float* lt, int d4, const int32_t* __restrict__ us, const int32_t* __restrict__ ps,
          const int32_t* __restrict__ qs, int64_t n, float alpha, float lambda, double* __restrict__ loss) {
    const int lane = threadIdx.x;
    for (int64_t i = 0; i < n; ++i) {
        float* ru = ux + (size_t)us[i] * d4 * 4;
        float* rp = lt + (size_t)ps[i] * d4 * 4;
        float* rq = lt + (size_t)qs[i] * d4 * 4;
        float4 u[NCH], p[NCH], q[NCH];
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            int c = lane + 32 * k;
            if (c < d4) {
                u[k] = ld4(ru + 4 * c); p[k] = ld4(rp + 4 * c); q[k] = ld4(rq + 4 * c);
                dot += u[k].x * (p[k].x - q[k].x) + u[k].y * (p[k].y - q[k].y) + u[k].z * (p[k].z - q[k].z) + u[k].w * (p[k].w - q[k].w);
            }
        }
        dot = warp_sum(dot);
        const float g = sigmoidf_(-dot);                 // -d(-log sigmoid(dot))/d dot
        if (lane == 0) loss[i] = -(double)logsigmoidf_(dot);
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            int c = lane + 32 * k;
            if (c < d4) {
                float4 nu, np, nq;
#define BPR_UPD(f)                                                         \
                nu.f = u[k].f - alpha * (-g * (p[k].f - q[k].f) + lambda * u[k].f); \
                np.f = p[k].f - alpha * (-g * u[k].f + lambda * p[k].f);   \
                nq.f = q[k].f - alpha * (g * u[k].f + lambda * q[k].f);
                BPR_UPD(x) BPR_UPD(y) BPR_UPD(z) BPR_UPD(w)
#undef BPR_UPD
                st4(ru + 4 * c, nu); st4(rp + 4 * c, np); st4(rq + 4 * c, nq);   // q last: last writer wins
            }
        }
    }
}

template <int NCH>
__global__ void __launch_bounds__(32)
k_prme_seq(float* du, float* dp, float* ds, int d4, const int32_t* __restrict__ us,
           const int32_t* __restrict__ ps, const int32_t* __restrict__ qs, const int32_t* __restrict__ prevs,
           const double* __restrict__ dist, const int32_t* __restrict__ gap, int64_t n,
           int thd, float cw, float alpha, float lambda, double* __restrict__ loss) {
    const int lane = threadIdx.x;
    for (int64_t i = 0; i < n; ++i) {
        float* ru = du + (size_t)us[i] * d4 * 4;
        const size_t op = (size_t)ps[i] * d4 * 4, oq = (size_t)qs[i] * d4 * 4, ol = (size_t)prevs[i] * d4 * 4;
        float4 u[NCH], pp[NCH], pq[NCH], pl[NCH], sp[NCH], sq[NCH], sl[NCH];
        float Dpp = 0.f, Dpq = 0.f, Dsp = 0.f, Dsq = 0.f;
#define SQ4(a, b) ((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z) + (a.w - b.w) * (a.w - b.w))
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            int c = lane + 32 * k;
            if (c < d4) {
                u[k] = ld4(ru + 4 * c);
                pp[k] = ld4(dp + op + 4 * c); pq[k] = ld4(dp + oq + 4 * c); pl[k] = ld4(dp + ol + 4 * c);
                sp[k] = ld4(ds + op + 4 * c); sq[k] = ld4(ds + oq + 4 * c); sl[k] = ld4(ds + ol + 4 * c);
                Dpp += SQ4(u[k], pp[k]); Dpq += SQ4(u[k], pq[k]);
                Dsp += SQ4(sp[k], sl[k]); Dsq += SQ4(sq[k], sl[k]);
            }
        }
#undef SQ4
        Dpp = warp_sum(Dpp); Dpq = warp_sum(Dpq); Dsp = warp_sum(Dsp); Dsq = warp_sum(Dsq);
        const bool far = gap[i] > thd;                                   // ifelse(T.gt(tidx, thd)), PRME.py:192
        const float w = (float)pow(1.0 + dist[i], 0.25);                 // PRME.py:191
        const float cp = far ? 1.f : w * cw, cs = far ? 0.f : w * (1.f - cw);
        const float x = (cp * Dpq + cs * Dsq) - (cp * Dpp + cs * Dsp);   // -Dp + Dq
        const float g = sigmoidf_(-x);                                   // d log sigmoid(x) / dx
        if (lane == 0) loss[i] = (double)logsigmoidf_(x);
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            int c = lane + 32 * k;
            if (c < d4) {
                float4 nu, npp, npq, npl, nsp, nsq, nsl;
#define PRME_UPD(f)                                                                                   \
                nu.f  = u[k].f  + alpha * (2.f * g * cp * ((u[k].f - pq[k].f) - (u[k].f - pp[k].f)) - lambda * u[k].f);   \
                npp.f = pp[k].f + alpha * (2.f * g * cp * (u[k].f - pp[k].f) - lambda * pp[k].f);     \
                npq.f = pq[k].f + alpha * (-2.f * g * cp * (u[k].f - pq[k].f) - lambda * pq[k].f);    \
                npl.f = pl[k].f + alpha * (-lambda * pl[k].f);                                        \
                nsp.f = sp[k].f + alpha * (-2.f * g * cs * (sp[k].f - sl[k].f) - lambda * sp[k].f);   \
                nsq.f = sq[k].f + alpha * (2.f * g * cs * (sq[k].f - sl[k].f) - lambda * sq[k].f);    \
                nsl.f = sl[k].f + alpha * (2.f * g * cs * ((sp[k].f - sl[k].f) - (sq[k].f - sl[k].f)) - lambda * sl[k].f);
                PRME_UPD(x) PRME_UPD(y) PRME_UPD(z) PRME_UPD(w)
#undef PRME_UPD
                st4(ru + 4 * c, nu);
                st4(dp + op + 4 * c, npp); st4(dp + oq + 4 * c, npq); st4(dp + ol + 4 * c, npl);   // [p, q, prev] order
                st4(ds + op + 4 * c, nsp); st4(ds + oq + 4 * c, nsq); st4(ds + ol + 4 * c, nsl);
            }
        }
    }
}
