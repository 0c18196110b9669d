// geoie.cuh -- GeoIE train step (reference public/GeoIE.py:129-194), one user per call.
//
//   sp_i = (sum_j msk_ij (g[p_j].h[p_{i+1}]) a d+_ij^b) / n_h_i + t[u].z[p_{i+1}]       (GeoIE.py:155-158)
//   sq_i = same with h[q_{i+1}], d-_ij and the SAME t.z[p_{i+1}] term (so t gets no gradient)
//   loss = sum_i log sigmoid(sp_i - sq_i);  cost = -loss + lambda/2 (|G|^2+|Hp|^2+|Hq|^2+|Zp|^2+|Zq|^2)
//
// Theano's type promotion makes this graph float64 from the first product on (int32 mask *
// float32 rows, float64 scalars a, b), so everything between the gathers and the row updates is
// computed in double here as well.  0 * inf = NaN for padded distances when b < 0 is reproduced,
// not guarded.
#pragma once
#include "common.cuh"

// rows: G = g[p[0:n]], Hp = h[p[1:n+1]], Hq = h[q[1:n+1]]  ->  dense [n x H] float buffers
__global__ void k_geoie_gather(const float* __restrict__ g, const float* __restrict__ h,
                               const int32_t* __restrict__ p, const int32_t* __restrict__ q, int n, int H,
                               float* __restrict__ G, float* __restrict__ Hp, float* __restrict__ Hq) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * H) return;
    int i = (int)(idx / H), c = (int)(idx % H);
    G[idx] = g[(size_t)p[i] * H + c];
    Hp[idx] = h[(size_t)p[i + 1] * H + c];
    Hq[idx] = h[(size_t)q[i + 1] * H + c];
}

// one CTA per target i: scores, loss term, d cost/d(sp-sq), coefficient rows for the backward
__global__ void __launch_bounds__(128)
k_geoie_fwd(const float* __restrict__ G, const float* __restrict__ Hp, const float* __restrict__ Hq, int n, int H,
            const float* __restrict__ dpos, const float* __restrict__ dneg, const int32_t* __restrict__ msk,
            const double* __restrict__ ab, double* __restrict__ CWp, double* __restrict__ CWq,
            double* __restrict__ per_i /* [n][4]: loss, ga, gb, unused */) {
    extern __shared__ double sm[];                 // Hp_i [H], Hq_i [H], then reduction scratch [4*4]
    double* hp = sm; double* hq = sm + H; double* red = sm + 2 * H;
    const int i = blockIdx.x, tid = threadIdx.x;
    for (int c = tid; c < H; c += blockDim.x) { hp[c] = Hp[(size_t)i * H + c]; hq[c] = Hq[(size_t)i * H + c]; }
    __syncthreads();
    const double a = ab[0], b = ab[1];
    double A = 0.0, Bq = 0.0, Sa = 0.0, Sb = 0.0, nh = 0.0;
    for (int j = tid; j < n; j += blockDim.x) {
        const double m = (double)msk[(size_t)i * n + j];
        double dp_ = 0.0, dq_ = 0.0;
        const float* gj = G + (size_t)j * H;
        for (int c = 0; c < H; ++c) { double gm = (double)gj[c] * m; dp_ += gm * hp[c]; dq_ += gm * hq[c]; }
        const double xp = (double)dpos[(size_t)i * n + j], xq = (double)dneg[(size_t)i * n + j];
        const double pp = pow(xp, b), pq = pow(xq, b);               // d ** b
        A += dp_ * (a * pp); Bq += dq_ * (a * pq);
        // d f/d a = d**b ; d f/d b = a d**b log d with Theano's switch(eq(d,0), 0, .) on the log part
        Sa += dp_ * pp - dq_ * pq;
        Sb += (xp == 0.0 ? 0.0 : dp_ * a * pp * log(xp)) - (xq == 0.0 ? 0.0 : dq_ * a * pq * log(xq));
        nh += m;
        CWp[(size_t)i * n + j] = m * (a * pp);                       // scaled by c_i below
        CWq[(size_t)i * n + j] = m * (a * pq);
    }
    double v[5] = {A, Bq, Sa, Sb, nh};
    for (int k = 0; k < 5; ++k) v[k] = warp_sum_d(v[k]);
    if ((tid & 31) == 0) for (int k = 0; k < 5; ++k) red[(tid >> 5) * 5 + k] = v[k];
    __syncthreads();
    __shared__ double s_c;
    if (tid == 0) {
        double t[5] = {0, 0, 0, 0, 0};
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) for (int k = 0; k < 5; ++k) t[k] += red[w * 5 + k];
        const double x = t[0] / t[4] - t[1] / t[4];                  // t.z cancels (GeoIE.py:155-159)
        const double ls = -(fmax(-x, 0.0) + log1p(exp(-fabs(x))));
        const double gx = -1.0 / (1.0 + exp(x));                     // d(-log sigmoid x)/dx = -sigmoid(-x)
        const double c_i = gx / t[4];
        per_i[(size_t)i * 4 + 0] = ls; per_i[(size_t)i * 4 + 1] = c_i * t[2]; per_i[(size_t)i * 4 + 2] = c_i * t[3];
        s_c = c_i;
    }
    __syncthreads();
    const double c_i = s_c;
    for (int j = tid; j < n; j += blockDim.x) { CWp[(size_t)i * n + j] *= c_i; CWq[(size_t)i * n + j] *= c_i; }
}

// per target i: d cost/d Hp_i = sum_j CWp_ij G_j + lambda Hp_i ; d cost/d Hq_i = -sum_j CWq_ij G_j + lambda Hq_i
// written as occurrence-gradient rows: position i+1 of the p half / q half of the [2*lmax] list
__global__ void __launch_bounds__(128)
k_geoie_bwd_h(const float* __restrict__ G, const float* __restrict__ Hp, const float* __restrict__ Hq, int n, int H,
              const double* __restrict__ CWp, const double* __restrict__ CWq, float lambda, int lmax,
              float* __restrict__ GH) {
    const int i = blockIdx.x;
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        double sp = 0.0, sq = 0.0;
        for (int j = 0; j < n; ++j) {
            double gj = (double)G[(size_t)j * H + c];
            sp += CWp[(size_t)i * n + j] * gj; sq += CWq[(size_t)i * n + j] * gj;
        }
        GH[(size_t)(i + 1) * H + c] = (float)(sp + (double)lambda * (double)Hp[(size_t)i * H + c]);
        GH[(size_t)(lmax + i + 1) * H + c] = (float)(-sq + (double)lambda * (double)Hq[(size_t)i * H + c]);
    }
}

// per history j: d cost/d G_j = sum_i (CWp_ij Hp_i - CWq_ij Hq_i) + lambda G_j  -> occurrence j of p
__global__ void __launch_bounds__(128)
k_geoie_bwd_g(const float* __restrict__ G, const float* __restrict__ Hp, const float* __restrict__ Hq, int n, int H,
              const double* __restrict__ CWp, const double* __restrict__ CWq, float lambda,
              float* __restrict__ GG) {
    const int j = blockIdx.x;
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        double s = 0.0;
        for (int i = 0; i < n; ++i)
            s += CWp[(size_t)i * n + j] * (double)Hp[(size_t)i * H + c] - CWq[(size_t)i * n + j] * (double)Hq[(size_t)i * H + c];
        GG[(size_t)j * H + c] = (float)(s + (double)lambda * (double)G[(size_t)j * H + c]);
    }
}

// z rows only decay: d cost/d z[p_{i+1}] = lambda z, same for q (GeoIE.py:164-171; t.z cancels)
__global__ void k_geoie_zgrad(const float* __restrict__ z, const int32_t* __restrict__ p, const int32_t* __restrict__ q,
                              int n, int H, int lmax, float lambda, float* __restrict__ GZ) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * H) return;
    int i = (int)(idx / H), c = (int)(idx % H);
    GZ[(size_t)(i + 1) * H + c] = lambda * z[(size_t)p[i + 1] * H + c];
    GZ[(size_t)(lmax + i + 1) * H + c] = lambda * z[(size_t)q[i + 1] * H + c];
}

__global__ void k_geoie_finalize(const double* __restrict__ per_i, int n, double* ab, float alpha, double* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double loss = 0.0, ga = 0.0, gb = 0.0;
    for (int i = 0; i < n; ++i) { loss += per_i[(size_t)i * 4]; ga += per_i[(size_t)i * 4 + 1]; gb += per_i[(size_t)i * 4 + 2]; }
    out[0] = loss;
    ab[0] -= (double)alpha * ga;                     // params = [a, b], no L2 on them (GeoIE.py:91,172-173)
    ab[1] -= (double)alpha * gb;
}
