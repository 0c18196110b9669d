// gru.cuh -- the GRU-family train / predict step (OboGru, Gru, OboSpatialGru = Distance2Pre).
//
// What one call computes is the reference's `seq_train` graph (GRU.py:313-389,407-498;
// GRU_Spatial.py:127-229): gather -> scan(GRU cell [, interval softmax head]) -> BPR (+survival)
// loss -> full BPTT -> dense SGD on the weights -> sparse SGD on the touched table rows, all from
// pre-update values.  How it is computed is new: time-major [t][b] activations, every
// time-independent contraction hoisted out of the recurrence into one large GEMM over all
// (t, b), two GEMMs per time step for the recurrence itself (the r-gate dependency forces two),
// epilogue-fused gate math, and a sort-based deterministic sparse update (rows.cuh).
//
// Pair/step convention (see oracle/explicit.py): step j computes h_j from x_j and h_{j-1};
// pair j scores x_{j+1} with h_j and is valid iff j+1 < L_b.  T = max_b L_b - 1.
#pragma once
#include "common.cuh"
#include "sort.cuh"
#include "rows.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "gru_fused.cuh"
#include "gru_small.cuh"
#include <string.h>

// ---------------------------------------------------------------------------------------------
// index preparation: user-major [n_user x lmax] rows -> time-major [lmax x B]
// ---------------------------------------------------------------------------------------------
__global__ void k_slice_indices(const int32_t* __restrict__ P, const int32_t* __restrict__ Q,
                                const int32_t* __restrict__ DP, const int32_t* __restrict__ DQ,
                                const int32_t* __restrict__ lens, int lmax,
                                const int32_t* __restrict__ uidx, int B,
                                int32_t* __restrict__ PQt, int32_t* __restrict__ DPt,
                                int32_t* __restrict__ DQt, int32_t* __restrict__ lensB) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)B * lmax) return;
    int b = (int)(idx / lmax), t = (int)(idx % lmax);
    int64_t u = uidx ? uidx[b] : b;
    int64_t src = u * lmax + t, dst = (int64_t)t * B + b, LB = (int64_t)lmax * B;
    PQt[dst] = P[src];
    PQt[LB + dst] = Q[src];
    if (DP) { DPt[dst] = DP[src]; DQt[dst] = DQ[src]; }
    if (t == 0) lensB[b] = lens[u];
}

// ---------------------------------------------------------------------------------------------
// fused gather: X[t,b] = [lt[p] ; di[dp]] for t < T ; XDiff[j,b] = lt[p[j+1]] - lt[q[j+1]]
//   xps = self.lt[xpidxs]; xqs = self.lt[xqidxs]; xds = self.di[dpidxs]; xs = concatenate(xps, xds)
//   (GRU_Spatial.py:144-147, GRU.py:327,424-425)
// one warp per (t, b); lanes cover the row with 128-bit accesses
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_gather_inputs(const float* __restrict__ lt, const float* __restrict__ di,
                const int32_t* __restrict__ PQt, const int32_t* __restrict__ DPt,
                int B, int T, int64_t LB, int d4, int din4, int want_diff,
                float* __restrict__ X, float* __restrict__ XDiff) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t rows = (int64_t)(T + (want_diff ? 1 : 0)) * B;
    const float4* lt4 = reinterpret_cast<const float4*>(lt);
    const float4* di4 = reinterpret_cast<const float4*>(di);
    float4* X4 = reinterpret_cast<float4*>(X);
    float4* D4 = reinterpret_cast<float4*>(XDiff);
    const int64_t TB = (int64_t)T * B;
    for (int64_t row = warp; row < rows; row += nwarps) {
        const int64_t p = PQt[row];
        const bool in_x = row < TB;
        const bool in_d = want_diff && row >= B;
        const int64_t q = in_d ? PQt[LB + row] : 0;
        const int64_t dpi = (di && in_x) ? DPt[row] : 0;
        for (int c = lane; c < d4; c += 32) {
            float4 xp = __ldg(lt4 + p * d4 + c);
            if (in_x) {
                X4[row * din4 + c] = xp;
                if (di) X4[row * din4 + d4 + c] = __ldg(di4 + dpi * d4 + c);
            }
            if (in_d) D4[(row - B) * d4 + c] = f4sub(xp, __ldg(lt4 + q * d4 + c));
        }
    }
}

// out[c, r] = in[r, c] for r < R, c < Cc; out has leading dimension ldo >= R, zero padded
__global__ void k_transpose_pad(const float* __restrict__ in, int R, int Cc, float* __restrict__ out, int ldo) {
    __shared__ float tile[32][33];
    int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < Cc) ? in[(size_t)r * Cc + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (c < Cc && r < ldo) out[(size_t)c * ldo + r] = tile[threadIdx.x][i];
    }
}

static int launch_transpose_pad(poi_engine* e, const float* in, int R, int Cc, float* out, int ldo) {
    dim3 grid((unsigned)poi_cdiv(Cc, 32), (unsigned)poi_cdiv(ldo, 32));
    POI_CAT(e, CAT_ELTWISE, 0, 0);
    POI_LAUNCH(e, k_transpose_pad, grid, dim3(32, 8), 0, in, R, Cc, out, ldo);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// GEMM epilogues (4 consecutive columns of one row each)
// ---------------------------------------------------------------------------------------------
struct EpiBiasStore {               // C = acc + bias          (input projection, head logits, DX)
    float* C; int ldc; const float* bias; int N;
    __device__ __forceinline__ void operator()(int m, int n, const float (&v)[4]) const {
        float* c = C + (size_t)m * ldc + n;
        if (n + 3 < N) {
            float4 b = bias ? ldg4(bias + n) : f4zero();
            st4(c, make_float4(v[0] + b.x, v[1] + b.y, v[2] + b.z, v[3] + b.w));
        } else {
            for (int i = 0; i < 4 && n + i < N; ++i) c[i] = v[i] + (bias ? bias[n + i] : 0.f);
        }
    }
};

struct EpiZR {   // z_r = sigmoid(ui[:2].x + wh[:2].h + bi[:2])   (GRU_Spatial.py:173-175); also r*h
    const float* AXj; const float* hp; float* Z; float* R; float* RH; int H;
    int fast;        // tensor-core modes: __expf / __fdividef forms (the same ones the fused H <= 128 kernels use); fp32 FMA mode: expf
    __device__ __forceinline__ float sg(float x) const { return fast ? __fdividef(1.f, 1.f + __expf(-x)) : sigmoidf_(x); }
    struct Pre { float4 ax, h; };
    __device__ __forceinline__ Pre pre(int m, int n) const {
        // both loads unconditional (the z half reads an h it does not use): a predicated load next to a register clear made
        // every request wait for the previous one (scoreboard aliasing, seen in the ncu source view)
        Pre p; p.ax = ldg4(AXj + (size_t)m * 3 * H + n);
        p.h = ldg4(hp + (size_t)m * H + (n < H ? n : n - H));
        return p;
    }
    __device__ __forceinline__ void operator()(int m, int n, const float (&v)[4], const Pre& p) const {
        const float4 ax = p.ax;
        float4 s = make_float4(sg(v[0] + ax.x), sg(v[1] + ax.y), sg(v[2] + ax.z), sg(v[3] + ax.w));
        if (n < H) {
            st4(Z + (size_t)m * H + n, s);
        } else {
            size_t o = (size_t)m * H + (n - H);
            st4(R + o, s);
            st4(RH + o, make_float4(s.x * p.h.x, s.y * p.h.y, s.z * p.h.z, s.w * p.h.w));
        }
    }
    __device__ __forceinline__ void operator()(int m, int n, const float (&v)[4]) const { (*this)(m, n, v, pre(m, n)); }
};

struct EpiC {    // c = tanh(ui[2].x + wh[2].(r*h) + bi[2]); h_t = (1-z)*h + z*c   (GRU_Spatial.py:176-178)
    const float* AXj; const float* hp; const float* Z; float* C; float* Hn; int H;
    int fast;
    __device__ __forceinline__ float th(float x) const { return fast ? 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)) : tanhf(x); }
    struct Pre { float4 ax, z, h; };
    __device__ __forceinline__ Pre pre(int m, int n) const {
        Pre p; size_t o = (size_t)m * H + n;
        p.ax = ldg4(AXj + (size_t)m * 3 * H + 2 * H + n); p.z = ldg4(Z + o); p.h = ldg4(hp + o);
        return p;
    }
    __device__ __forceinline__ void operator()(int m, int n, const float (&v)[4]) const { (*this)(m, n, v, pre(m, n)); }
    __device__ __forceinline__ void operator()(int m, int n, const float (&v)[4], const Pre& p) const {
        const float4 ax = p.ax, z = p.z, h = p.h;
        size_t o = (size_t)m * H + n;
        float4 c = make_float4(th(v[0] + ax.x), th(v[1] + ax.y), th(v[2] + ax.z), th(v[3] + ax.w));
        st4(C + o, c);
        st4(Hn + o, make_float4((1.f - z.x) * h.x + z.x * c.x, (1.f - z.y) * h.y + z.y * c.y,
                                (1.f - z.z) * h.z + z.z * c.z, (1.f - z.w) * h.w + z.w * c.w));
    }
};

struct EpiDHl {  // d cost/d h_j from the loss: Vs^T.do_j + e_j (xp_{j+1} - xq_{j+1})
    float* DHl; const float* ev; const float* XDiff; int H;
    struct Pre { float4 x; float e; };
    __device__ __forceinline__ Pre pre(int m, int n) const {
        Pre p; p.e = __ldg(ev + m); p.x = ldg4(XDiff + (size_t)m * H + n);
        return p;
    }
    __device__ __forceinline__ void operator()(int m, int n, const float (&v)[4]) const { (*this)(m, n, v, pre(m, n)); }
    __device__ __forceinline__ void operator()(int m, int n, const float (&v)[4], const Pre& p) const {
        size_t o = (size_t)m * H + n;
        const float e_ = p.e; const float4 x = p.x;
        st4(DHl + o, make_float4(fmaf(e_, x.x, v[0]), fmaf(e_, x.y, v[1]), fmaf(e_, x.z, v[2]), fmaf(e_, x.w, v[3])));
    }
};

struct EpiM {    // m = Wh[2]^T.da_c ; dr = m*h_prev ; da_r = dr*r(1-r) ; dh_keep += m*r
    const float* hp; const float* R; float* DAj; float* DHK; int H;
    struct Pre { float4 h, r, k; };
    __device__ __forceinline__ Pre pre(int m, int n) const {
        Pre p; size_t o = (size_t)m * H + n;
        p.h = ldg4(hp + o); p.r = ldg4(R + o); p.k = ld4(DHK + o);
        return p;
    }
    __device__ __forceinline__ void operator()(int m, int n, const float (&v)[4]) const { (*this)(m, n, v, pre(m, n)); }
    __device__ __forceinline__ void operator()(int m, int n, const float (&v)[4], const Pre& p) const {
        size_t o = (size_t)m * H + n;
        const float4 h = p.h, r = p.r, k = p.k;
        st4(DAj + (size_t)m * 3 * H + H + n,
            make_float4(v[0] * h.x * r.x * (1.f - r.x), v[1] * h.y * r.y * (1.f - r.y),
                        v[2] * h.z * r.z * (1.f - r.z), v[3] * h.w * r.w * (1.f - r.w)));
        st4(DHK + o, make_float4(fmaf(v[0], r.x, k.x), fmaf(v[1], r.y, k.y), fmaf(v[2], r.z, k.z), fmaf(v[3], r.w, k.w)));
    }
};

__device__ __forceinline__ void bwd_gate_math(float dht, float z, float c, float hp, float& da_z, float& da_c, float& keep) {
    da_c = dht * z * (1.f - c * c);
    da_z = dht * (c - hp) * z * (1.f - z);
    keep = dht * (1.f - z);
}

struct EpiDH {   // dh_{j-1} = dh_keep + [da_z,da_r].Wh[:2] ; then the gate math of step j-1
    float* DHK; const float* DHlp; const float* Zp; const float* Cp; const float* HPp; float* DAp; int H;
    struct Pre { float4 k, l, z, c, h; };
    __device__ __forceinline__ Pre pre(int m, int n) const {
        Pre p; size_t o = (size_t)m * H + n;
        p.k = ld4(DHK + o); p.l = ldg4(DHlp + o); p.z = ldg4(Zp + o); p.c = ldg4(Cp + o); p.h = ldg4(HPp + o);
        return p;
    }
    __device__ __forceinline__ void operator()(int m, int n, const float (&v)[4]) const { (*this)(m, n, v, pre(m, n)); }
    __device__ __forceinline__ void operator()(int m, int n, const float (&v)[4], const Pre& p) const {
        size_t o = (size_t)m * H + n;
        const float4 k = p.k, l = p.l, z = p.z, c = p.c, h = p.h;
        float4 daz, dac, kp;
        bwd_gate_math(v[0] + k.x + l.x, z.x, c.x, h.x, daz.x, dac.x, kp.x);
        bwd_gate_math(v[1] + k.y + l.y, z.y, c.y, h.y, daz.y, dac.y, kp.y);
        bwd_gate_math(v[2] + k.z + l.z, z.z, c.z, h.z, daz.z, dac.z, kp.z);
        bwd_gate_math(v[3] + k.w + l.w, z.w, c.w, h.w, daz.w, dac.w, kp.w);
        st4(DAp + (size_t)m * 3 * H + n, daz);
        st4(DAp + (size_t)m * 3 * H + 2 * H + n, dac);
        st4(DHK + o, kp);
    }
};

// gate math of the last step (incoming dh = 0)
__global__ void k_bwd_prep(const float* __restrict__ DHl, const float* __restrict__ Z, const float* __restrict__ C,
                           const float* __restrict__ HP, float* __restrict__ DA, float* __restrict__ DHK,
                           int B, int H) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)B * H / 4) return;
    int m = (int)(idx / (H / 4)), n = (int)(idx % (H / 4)) * 4;
    size_t o = (size_t)m * H + n;
    float4 l = ld4(DHl + o), z = ld4(Z + o), c = ld4(C + o), h = ld4(HP + o);
    float4 daz, dac, kp;
    bwd_gate_math(l.x, z.x, c.x, h.x, daz.x, dac.x, kp.x);
    bwd_gate_math(l.y, z.y, c.y, h.y, daz.y, dac.y, kp.y);
    bwd_gate_math(l.z, z.z, c.z, h.z, daz.z, dac.z, kp.z);
    bwd_gate_math(l.w, z.w, c.w, h.w, daz.w, dac.w, kp.w);
    st4(DA + (size_t)m * 3 * H + n, daz);
    st4(DA + (size_t)m * 3 * H + 2 * H + n, dac);
    st4(DHK + o, kp);
}

// plain GRU: d cost / d h_j = e_j (xp_{j+1} - xq_{j+1})
__global__ void k_dhl_nohead(const float* __restrict__ ev, const float* __restrict__ XDiff, float* __restrict__ DHl,
                             int64_t rows, int H4) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * H4) return;
    float e_ = ev[idx / H4];
    float4 x = reinterpret_cast<const float4*>(XDiff)[idx];
    reinterpret_cast<float4*>(DHl)[idx] = make_float4(e_ * x.x, e_ * x.y, e_ * x.z, e_ * x.w);
}

// ---------------------------------------------------------------------------------------------
// loss head: one warp per pair (j, b).
//   s = softmax(vs.h + bs); upq = h.(xp - xq) + wd (s[P] - s[Q]); bpr = log sigmoid(upq);
//   sur = sum(s[:P+1]) - log s[P]                                   (GRU_Spatial.py:180-189)
//   plain GRU: upq = h.(xp - xq)                                      (GRU.py:352-353)
// Writes e = d cost/d upq and, over the logits in place, do = d cost / d logits.
// Block partial sums of (sur, bpr, gwd) go to part[block][4] in fp64 (fixed order).
// ---------------------------------------------------------------------------------------------
constexpr int LOSS_RK = 8;          // logits per lane kept in registers by k_loss_head (rows up to 256 wide)

__global__ void __launch_bounds__(256, 5)
k_loss_head(float* __restrict__ S, int nD, int nDp, const float* __restrict__ Hc,
            const float* __restrict__ XDiff, int H4, const int32_t* __restrict__ DPt,
            const int32_t* __restrict__ DQt, const int32_t* __restrict__ lensB,
            const float* __restrict__ scal, int B, int T, float scale, int head,
            float* __restrict__ ev, double* __restrict__ part) {
    __shared__ double sh[8][3];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * 8 + w;
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    const int64_t rows = (int64_t)T * B;
    float w0 = 0.f, w1 = 1.f, wd = 0.f;
    if (head) {
        float a = scal[1], b = scal[2], mx = fmaxf(a, b);
        float ea = expf(a - mx), eb = expf(b - mx);
        w0 = ea / (ea + eb); w1 = eb / (ea + eb); wd = scal[0];
    }
    double a_sur = 0.0, a_bpr = 0.0, a_gwd = 0.0;
    for (int64_t m = warp; m < rows; m += nwarps) {
        const int j = (int)(m / B), b = (int)(m % B);
        // every load of the row is issued before anything depends on one of them (length, interval ids, the two dot operands,
        // the logits): one memory round trip per row instead of the three the branch structure used to serialise
        const int len_b = lensB[b];
        const bool in_regs = head && nDp <= 32 * LOSS_RK;
        int P = 0, Q = 0;
        if (head) { P = DPt[(int64_t)(j + 1) * B + b]; Q = DQt[(int64_t)(j + 1) * B + b]; }
        float* row = S + (size_t)m * nDp;
        float v[LOSS_RK];
        if (in_regs) {
#pragma unroll
            for (int i = 0; i < LOSS_RK; ++i) { const int k = lane + 32 * i; v[i] = k < nD ? row[k] : -INFINITY; }
        }
        const bool valid = (j + 1) < len_b;
        const float4* h4 = reinterpret_cast<const float4*>(Hc) + m * H4;
        const float4* x4 = reinterpret_cast<const float4*>(XDiff) + m * H4;
        float dot = 0.f;
        for (int c = lane; c < H4; c += 32) {
            float4 h = h4[c], x = x4[c];
            dot += h.x * x.x + h.y * x.y + h.z * x.z + h.w * x.w;
        }
        dot = warp_sum(dot);
        float e_ = 0.f, bpr = 0.f, sur = 0.f, gwd = 0.f;
        if (head) {
            if (!valid) {
                for (int k = lane; k < nDp; k += 32) row[k] = 0.f;
            } else {
                if (in_regs) {
                    // the whole row lives in registers (<= LOSS_RK values per lane): one read, one exp per logit, one write
                    float mx = -INFINITY;
#pragma unroll
                    for (int i = 0; i < LOSS_RK; ++i) mx = fmaxf(mx, v[i]);
                    mx = warp_max(mx);
                    float sum = 0.f, cumr = 0.f, eP = 0.f, eQ = 0.f;
#pragma unroll
                    for (int i = 0; i < LOSS_RK; ++i) {
                        const int k = lane + 32 * i;
                        const float ex = k < nD ? expf(v[i] - mx) : 0.f;
                        v[i] = ex;
                        sum += ex;
                        if (k <= P) cumr += ex;
                        if (k == P) eP = ex;
                        if (k == Q) eQ = ex;
                    }
                    sum = warp_sum(sum); cumr = warp_sum(cumr); eP = warp_sum(eP); eQ = warp_sum(eQ);
                    const float inv = 1.f / sum;
                    const float sP = eP * inv, sQ = eQ * inv, cum = cumr * inv;
                    const float u = dot + wd * (sP - sQ);
                    e_ = -w1 * sigmoidf_(-u) * scale;
                    bpr = logsigmoidf_(u);
                    sur = cum - logf(sP);
                    gwd = e_ * (sP - sQ);
                    const float Ac = w0 * scale, Bc = e_ * wd;
                    const float gs = Ac * cum - Ac + Bc * (sP - sQ);
                    const float aP = Ac / sP;
#pragma unroll
                    for (int i = 0; i < LOSS_RK; ++i) {
                        const int k = lane + 32 * i;
                        if (k < nDp) {
                            float o = 0.f;
                            if (k < nD) {
                                const float sk = v[i] * inv;
                                const float g = (k <= P ? Ac : 0.f) - (k == P ? aP : 0.f) + Bc * ((k == P ? 1.f : 0.f) - (k == Q ? 1.f : 0.f));
                                o = sk * (g - gs);
                            }
                            row[k] = o;
                        }
                    }
                } else {
                float mx = -INFINITY;
                for (int k = lane; k < nD; k += 32) mx = fmaxf(mx, row[k]);
                mx = warp_max(mx);
                float sum = 0.f, cumr = 0.f, eP = 0.f, eQ = 0.f;
                for (int k = lane; k < nD; k += 32) {
                    float ex = expf(row[k] - mx);
                    sum += ex;
                    if (k <= P) cumr += ex;
                    if (k == P) eP = ex;
                    if (k == Q) eQ = ex;
                }
                sum = warp_sum(sum); cumr = warp_sum(cumr); eP = warp_sum(eP); eQ = warp_sum(eQ);
                const float inv = 1.f / sum;
                const float sP = eP * inv, sQ = eQ * inv, cum = cumr * inv;
                const float u = dot + wd * (sP - sQ);
                e_ = -w1 * sigmoidf_(-u) * scale;
                bpr = logsigmoidf_(u);
                sur = cum - logf(sP);
                gwd = e_ * (sP - sQ);
                const float Ac = w0 * scale, Bc = e_ * wd;
                const float gs = Ac * cum - Ac + Bc * (sP - sQ);
                const float aP = Ac / sP;
                for (int k = lane; k < nDp; k += 32) {
                    float o = 0.f;
                    if (k < nD) {
                        float s = expf(row[k] - mx) * inv;
                        float g = (k <= P ? Ac : 0.f) - (k == P ? aP : 0.f) + Bc * ((k == P ? 1.f : 0.f) - (k == Q ? 1.f : 0.f));
                        o = s * (g - gs);
                    }
                    row[k] = o;
                }
                }
            }
        } else if (valid) {
            e_ = -sigmoidf_(-dot) * scale;
            bpr = logsigmoidf_(dot);
        }
        if (lane == 0) {
            ev[m] = e_;
            a_sur += (double)sur; a_bpr += (double)bpr; a_gwd += (double)gwd;
        }
    }
    if (lane == 0) { sh[w][0] = a_sur; sh[w][1] = a_bpr; sh[w][2] = a_gwd; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
        part[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
    }
}

// Final scalars + the scalar parameters' SGD step (wd, loss_weight; GRU_Spatial.py:210-211).
// out[0..4] = los, sur, upq, w0, w1 (plain GRU: out[0] = upq)
__device__ void finalize_apply(double sur, double bpr, double gwd, float* scal, int head, double extra_upq,
                               double scale, float alpha, float lambda, double* out) {
    double upq = -bpr + extra_upq;
    if (!head) { out[0] = upq; out[1] = 0.0; out[2] = upq; out[3] = 0.0; out[4] = 1.0; return; }
    float a = scal[1], b = scal[2], mx = fmaxf(a, b);
    float ea = expf(a - mx), eb = expf(b - mx);
    double w0 = ea / (ea + eb), w1 = eb / (ea + eb), wd = scal[0];
    out[0] = w0 * sur + w1 * upq; out[1] = sur; out[2] = upq; out[3] = w0; out[4] = w1;
    double g_wd = (double)lambda * wd + gwd;
    double dw0 = sur * scale + (double)lambda * w0, dw1 = upq * scale + (double)lambda * w1;
    double dot = w0 * dw0 + w1 * dw1;
    scal[0] = (float)(wd - (double)alpha * g_wd);
    scal[1] = (float)((double)a - (double)alpha * w0 * (dw0 - dot));
    scal[2] = (float)((double)b - (double)alpha * w1 * (dw1 - dot));
}

// sums_out != NULL (multi-GPU): only publish the three partial sums; the step is applied after the all-reduce
__global__ void __launch_bounds__(256)
k_finalize_gru(const double* __restrict__ part, int nblocks, float* scal, int head,
               double extra_upq, double scale, float alpha, float lambda,
               double* __restrict__ out, double* __restrict__ sums_out) {
    // fixed-shape tree over the block partials: thread t sums entries t, t+256, ... then a fixed
    // shuffle/shared tree -> deterministic for a given grid size
    __shared__ double sh[8][3];
    double s3[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < nblocks; i += 256)
        for (int k = 0; k < 3; ++k) s3[k] += part[(size_t)i * 4 + k];
    for (int k = 0; k < 3; ++k) s3[k] = warp_sum_d(s3[k]);
    if ((threadIdx.x & 31) == 0) for (int k = 0; k < 3; ++k) sh[threadIdx.x >> 5][k] = s3[k];
    __syncthreads();
    if (threadIdx.x != 0) return;
    double sur = 0.0, bpr = 0.0, gwd = 0.0;
    for (int w = 0; w < 8; ++w) { sur += sh[w][0]; bpr += sh[w][1]; gwd += sh[w][2]; }
    if (sums_out) { sums_out[0] = sur; sums_out[1] = bpr; sums_out[2] = gwd; return; }
    finalize_apply(sur, bpr, gwd, scal, head, extra_upq, scale, alpha, lambda, out);
}

__global__ void k_finalize_from_sums(const double* __restrict__ sums, float* scal, int head, double extra_upq,
                                     double scale, float alpha, float lambda, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0)
        finalize_apply(sums[0], sums[1], sums[2], scal, head, extra_upq, scale, alpha, lambda, out);
}

// ---------------------------------------------------------------------------------------------
// GEMM dispatch: SIMT fp32 (mode 0) or tcgen05 (modes 1, 2; gemm_tc.cuh)
// ---------------------------------------------------------------------------------------------
template <class Epi>
static int gemm_tn(poi_engine* e, const float* A, int lda, const float* W, int ldw,
                   int64_t M, int N, int K, const Epi& epi) {
    if (e->gemm_mode != 0 && tc_gemm_supported(M, N, K, lda, ldw))
        return launch_gemm_tn_tc(e, A, lda, W, ldw, M, N, K, epi, e->gemm_mode == 1);
    return launch_gemm_tn(e, A, lda, W, ldw, M, N, K, epi);
}

// ---------------------------------------------------------------------------------------------
// the step
// ---------------------------------------------------------------------------------------------
// multi-GPU step context: users are sharded over ranks, the item table row-sharded.
// The step gathers from `rows` (the unique rows this rank needs, fetched from their owners, sorted
// by row id) and, instead of applying updates, emits gradients to be exchanged / all-reduced.
struct MgCtx {
    const float* rows;       // [n_unique x d]
    int global_batch;        // sum of B over ranks (loss normalisation)
    float* dense_grads;      // flat buffer, layout MgLayout
    float* row_grads;        // [n_unique x d] duplicate-summed loss gradient per unique row
    float* row_cnt;          // [n_unique] occurrences (L2 multiplicity)
    double* loss_sums;       // device double[3]: sur, bpr, gwd
    bool no_sync = false;    // poi_gru_step_mg: the caller continues on the stream, nothing is read back here
};
struct MgLayout { int64_t ui, wh, bi, vs, bs, di, dicnt, total; };
struct GruIdx;
static MgLayout mg_layout(int H, int din, int nD, int d) {
    MgLayout L; int64_t o = 0;
    auto take = [&](int64_t n) { int64_t at = o; o += (n + 3) / 4 * 4; return at; };
    L.ui = take((int64_t)3 * H * din); L.wh = take((int64_t)3 * H * H); L.bi = take(3 * H);
    L.vs = take((int64_t)nD * H); L.bs = take(nD); L.di = take((int64_t)nD * d); L.dicnt = take(nD);
    L.total = o;
    return L;
}

struct GruIdx {              // time-major device index arrays for this batch
    int32_t* PQt; int32_t* DPt; int32_t* DQt; int32_t* lensB;
};

static int gru_check_params(poi_engine* e, const poi_gru_params* p) {
    if (!p || !p->lt || !p->ui || !p->wh || !p->bi) POI_FAIL(e, "gru params: null pointer");
    if (p->d <= 0 || p->d % 4 || p->H % 4) POI_FAIL(e, "n_in (%d) and n_hidden (%d) must be multiples of 4", p->d, p->H);
    if (p->d != p->H) POI_FAIL(e, "n_in (%d) must equal n_hidden (%d): the loss is h.(xp-xq) (GRU.py:352)", p->d, p->H);
    if (p->di && (!p->vs || !p->bs || !p->scal || p->n_rows_di <= 0)) POI_FAIL(e, "Distance2Pre params incomplete");
    return 0;
}

static int gru_forward(poi_engine* e, const poi_gru_params* p, const GruIdx& ix, int B, int lmax, int T,
                       bool training, float** X_, float** XDiff_, float** AX_, float** Hs_, float** Z_,
                       float** R_, float** C_, float** RH_) {
    const bool head = p->di != nullptr;
    const int d = p->d, H = p->H, din = head ? 2 * d : d;
    const int64_t LB = (int64_t)lmax * B, TB = (int64_t)T * B;
    float *X, *XDiff = nullptr, *AX, *Hs, *Z, *R, *C, *RH;
    POI_TRY(arena_get(e, (size_t)std::max<int64_t>(TB, 1) * din, &X));
    if (training) POI_TRY(arena_get(e, (size_t)std::max<int64_t>(TB, 1) * d, &XDiff));
    POI_TRY(arena_get(e, (size_t)std::max<int64_t>(TB, 1) * 3 * H, &AX));
    POI_TRY(arena_get(e, (size_t)(TB + B) * H, &Hs));
    POI_TRY(arena_get(e, (size_t)std::max<int64_t>(TB, 1) * H, &Z));
    POI_TRY(arena_get(e, (size_t)std::max<int64_t>(TB, 1) * H, &R));
    POI_TRY(arena_get(e, (size_t)std::max<int64_t>(TB, 1) * H, &C));
    POI_TRY(arena_get(e, (size_t)std::max<int64_t>(TB, 1) * H, &RH));
    *X_ = X; *XDiff_ = XDiff; *AX_ = AX; *Hs_ = Hs; *Z_ = Z; *R_ = R; *C_ = C; *RH_ = RH;
    POI_CK(e, cudaMemsetAsync(Hs, 0, (size_t)B * H * sizeof(float), e->stream));      // h_{-1} = h0 = 0 (GRU.py:63)
    if (T <= 0) return 0;
    {
        int64_t rows = (int64_t)(T + (training ? 1 : 0)) * B;
        unsigned grid = (unsigned)std::min<int64_t>(poi_cdiv(rows * 32, 256), (int64_t)e->num_sms * 16);
        // algorithmic bytes: rows read from the tables + the dense tiles written + the indices
        double gb = (double)rows * d * 4 + (training ? (double)TB * d * 4 * 2 : 0.0) + (head ? (double)TB * d * 4 : 0.0)
                  + (double)TB * din * 4 + 4.0 * (double)rows * (head ? 3 : 2);
        POI_CAT(e, CAT_GATHER, 0, gb);
        POI_LAUNCH(e, k_gather_inputs, grid, 256, 0, p->lt, p->di, ix.PQt, ix.DPt, B, T, LB, d / 4, din / 4,
                   training ? 1 : 0, X, XDiff);
    }
    phase_mark(e, 2);
    // hoisted input projection: AX = X . ui^T + bi over every (t, b)
    POI_TRY(gemm_tn(e, X, din, p->ui, din, TB, 3 * H, din, EpiBiasStore{AX, 3 * H, p->bi, 3 * H}));
    if (e->small_batch_path && small::supported(B, H)) {
        // tiny batches (the reference's one-by-one mode): one CTA per user, Wh resident in shared memory (gru_small.cuh)
        POI_TRY(small::launch_fwd(e, AX, p->wh, Hs, Z, R, C, RH, B, T, H));
        return 0;
    }
    if (e->gemm_mode != 0 && e->fuse_recurrence && fused::fwd_supported(H)) {
        // the whole recurrence in one persistent tcgen05 kernel (gru_fused.cuh)
        POI_TRY(fused::launch_gru_fwd_fused(e, AX, p->wh, Hs, Z, R, C, RH, B, T, H, e->gemm_mode == 1));
        return 0;
    }
    e->gemm_cat = CAT_RECUR_FWD;
    for (int j = 0; j < T; ++j) {
        const float* hp = Hs + (size_t)j * B * H;
        const float* AXj = AX + (size_t)j * B * 3 * H;
        size_t o = (size_t)j * B * H;
        POI_TRY(gemm_tn(e, hp, H, p->wh, H, B, 2 * H, j == 0 ? 0 : H, EpiZR{AXj, hp, Z + o, R + o, RH + o, H, e->gemm_mode != 0 ? 1 : 0}));
        POI_TRY(gemm_tn(e, RH + o, H, p->wh + (size_t)2 * H * H, H, B, H, j == 0 ? 0 : H,
                        EpiC{AXj, hp, Z + o, C + o, Hs + o + (size_t)B * H, H, e->gemm_mode != 0 ? 1 : 0}));
    }
    e->gemm_cat = -1;
    return 0;
}

// segments built by an earlier call in the same arena epoch (multi-GPU: poi_gru_mg_prepare)
struct PreSeg { const SegList* lt; const SegList* di; };

static int gru_train_core(poi_engine* e, const poi_gru_params* p, const GruIdx& ix, int B, int lmax,
                          int max_len, int64_t n_nonempty, float alpha, float lambda, double* out_host,
                          const MgCtx* mg = nullptr, const PreSeg* pre = nullptr) {
    const bool head = p->di != nullptr;
    const int d = p->d, H = p->H, din = head ? 2 * d : d;
    const int nD = head ? p->n_rows_di : 0, nDp = (nD + 3) / 4 * 4;
    const int T = std::max(std::min(max_len, lmax) - 1, 0);
    const int64_t LB = (int64_t)lmax * B, TB = (int64_t)T * B;
    const float scale = 1.0f / (float)(mg ? mg->global_batch : B);
    const MgLayout ML = mg_layout(H, din, nD, d);

    // ---- integer work: sorted-unique segments of the gathered row ids (pad rows included) ----
    SegList seg_lt, seg_di;
    if (pre) { seg_lt = *pre->lt; if (head) seg_di = *pre->di; }
    else {
        POI_TRY(build_segments(e, reinterpret_cast<const uint32_t*>(ix.PQt), 2 * LB, (uint32_t)p->n_rows_lt, mg != nullptr, &seg_lt));
        if (head) POI_TRY(build_segments(e, reinterpret_cast<const uint32_t*>(ix.DPt), LB, (uint32_t)nD, false, &seg_di));
    }
    phase_mark(e, 1);

    // ---- forward ----
    float *X, *XDiff, *AX, *Hs, *Z, *R, *C, *RH;
    poi_gru_params pf = *p; GruIdx ixf = ix;
    if (mg) { pf.lt = const_cast<float*>(mg->rows); ixf.PQt = reinterpret_cast<int32_t*>(seg_lt.seg_of_occ); }   // gather by slot
    POI_TRY(gru_forward(e, &pf, ixf, B, lmax, T, true, &X, &XDiff, &AX, &Hs, &Z, &R, &C, &RH));
    if (T <= 0) phase_mark(e, 2);
    phase_mark(e, 3);
    const float* Hc = Hs + (size_t)B * H;
    float *S = nullptr, *ev, *DHl, *DA, *DHK, *DX;
    const size_t TB1 = (size_t)std::max<int64_t>(TB, 1);
    POI_TRY(arena_get(e, TB1, &ev));
    POI_TRY(arena_get(e, TB1 * H, &DHl));
    if (head) {
        POI_TRY(arena_get(e, TB1 * nDp, &S));
        POI_TRY(gemm_tn(e, Hc, H, p->vs, H, TB, nD, H, EpiBiasStore{S, nDp, p->bs, nD}));
    }
    static int loss_occ = 0;            // one wave of resident CTAs (a grid-stride loop over more would run its tail half empty)
    if (!loss_occ) { int o = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_loss_head, 256, 0); loss_occ = std::max(o, 1); }
    int loss_blocks = (int)std::max<int64_t>(1, std::min<int64_t>(poi_cdiv(TB, 8), (int64_t)e->num_sms * loss_occ));
    double *part, *out_dev;
    POI_TRY(arena_get(e, (size_t)loss_blocks * 4, &part));
    POI_TRY(arena_get(e, 8, &out_dev));
    POI_CAT(e, CAT_LOSS, 0, (double)TB * ((head ? 2.0 * nDp : 0.0) + 2.0 * H) * 4);
    POI_LAUNCH(e, k_loss_head, loss_blocks, 256, 0, S, nD, nDp, Hc, XDiff, H / 4, ix.DPt, ix.DQt, ix.lensB,
               p->scal, B, T, scale, head ? 1 : 0, ev, part);
    phase_mark(e, 4);

    // ---- backward ----
    float *W2cT, *W2zrT, *U2T, *vsT = nullptr;
    POI_TRY(arena_get(e, (size_t)H * H, &W2cT));
    POI_TRY(arena_get(e, (size_t)H * 2 * H, &W2zrT));
    POI_TRY(arena_get(e, (size_t)din * 3 * H, &U2T));
    POI_TRY(launch_transpose_pad(e, p->wh + (size_t)2 * H * H, H, H, W2cT, H));
    POI_TRY(launch_transpose_pad(e, p->wh, 2 * H, H, W2zrT, 2 * H));
    POI_TRY(launch_transpose_pad(e, p->ui, 3 * H, din, U2T, 3 * H));
    POI_TRY(arena_get(e, TB1 * 3 * H, &DA));
    POI_TRY(arena_get(e, (size_t)B * H, &DHK));
    POI_TRY(arena_get(e, TB1 * din, &DX));
    if (T > 0) {
        if (head) {
            POI_TRY(arena_get(e, (size_t)H * nDp, &vsT));
            POI_TRY(launch_transpose_pad(e, p->vs, nD, H, vsT, nDp));
            POI_TRY(gemm_tn(e, S, nDp, vsT, nDp, TB, H, nDp, EpiDHl{DHl, ev, XDiff, H}));
        } else {
            POI_CAT(e, CAT_ELTWISE, 0, 0);
            POI_LAUNCH(e, k_dhl_nohead, (unsigned)poi_cdiv(TB * (H / 4), 256), 256, 0, ev, XDiff, DHl, TB, H / 4);
        }
        if (e->small_batch_path && small::supported(B, H)) {
            POI_TRY(small::launch_bwd(e, DHl, Z, R, C, Hs, p->wh, DA, B, T, H));
        } else if (e->gemm_mode != 0 && e->fuse_recurrence && fused::fwd_supported(H) && H <= 128) {
            // BPTT through the cell as one persistent tcgen05 kernel (gru_fused.cuh); dh never leaves the SM
            POI_TRY(fused::launch_gru_bwd_fused(e, DHl, Z, R, C, Hs, p->wh, DA, B, T, H, e->gemm_mode == 1));
        } else {
            {
                size_t o = (size_t)(T - 1) * B * H;
                POI_CAT(e, CAT_ELTWISE, 0, 0);
                POI_LAUNCH(e, k_bwd_prep, (unsigned)poi_cdiv((int64_t)B * H / 4, 256), 256, 0, DHl + o, Z + o, C + o,
                           Hs + o, DA + (size_t)(T - 1) * B * 3 * H, DHK, B, H);
            }
            e->gemm_cat = CAT_RECUR_BWD;
            for (int j = T - 1; j >= 0; --j) {
                size_t o = (size_t)j * B * H;
                float* DAj = DA + (size_t)j * B * 3 * H;
                POI_TRY(gemm_tn(e, DAj + 2 * H, 3 * H, W2cT, H, B, H, H, EpiM{Hs + o, R + o, DAj, DHK, H}));
                if (j > 0) {
                    size_t op = (size_t)(j - 1) * B * H;
                    POI_TRY(gemm_tn(e, DAj, 3 * H, W2zrT, 2 * H, B, H, 2 * H,
                                    EpiDH{DHK, DHl + op, Z + op, C + op, Hs + op, DA + (size_t)(j - 1) * B * 3 * H, H}));
                }
            }
            e->gemm_cat = -1;
        }
        POI_TRY(gemm_tn(e, DA, 3 * H, U2T, 3 * H, TB, din, 3 * H, EpiBiasStore{DX, din, nullptr, din}));
    }
    phase_mark(e, 5);

    // ---- weight gradients (split reductions) + dense SGD, all from pre-update values ----
    AtbPlan g_ui, g_whzr, g_whc, g_bi, g_vs, g_bs;
    const int64_t Mp = (TB + 3) / 4 * 4;
    bool have_colsums = false;
    if (e->gemm_mode != 0 && e->wgrad_mn && TB >= 2048) {
        // tensor-core path, MN-major operands: the activation matrices are read as they lie (time-major rows =
        // the contraction dimension), bias gradients (column sums of DA and dO) come out of the same pass
        const bool s3 = e->gemm_mode == 1;
        POI_TRY(launch_gemm_atb_tc_mn(e, DA, 3 * H, X, din, TB, 3 * H, din, s3, &g_ui, &g_bi));
        POI_TRY(launch_gemm_atb_tc_mn(e, DA, 3 * H, Hs, H, TB, 2 * H, H, s3, &g_whzr));
        POI_TRY(launch_gemm_atb_tc_mn(e, DA + 2 * H, 3 * H, RH, H, TB, H, H, s3, &g_whc));
        if (head) POI_TRY(launch_gemm_atb_tc_mn(e, S, nDp, Hc, H, TB, nDp, H, s3, &g_vs, &g_bs));
        have_colsums = true;
    } else if (e->gemm_mode != 0 && TB >= 2048 && Mp < 0x7fffffffLL) {
        // tensor-core path: one transpose per activation matrix, then split-K UMMA GEMMs
        const bool s3 = e->gemm_mode == 1;
        float *DAt, *Xt, *Hpt, *RHt;
        POI_TRY(arena_get(e, (size_t)3 * H * Mp, &DAt)); POI_TRY(arena_get(e, (size_t)din * Mp, &Xt));
        POI_TRY(arena_get(e, (size_t)H * Mp, &Hpt)); POI_TRY(arena_get(e, (size_t)H * Mp, &RHt));
        POI_TRY(launch_transpose_ld(e, DA, 3 * H, TB, 3 * H, DAt, Mp));
        POI_TRY(launch_transpose_ld(e, X, din, TB, din, Xt, Mp));
        POI_TRY(launch_transpose_ld(e, Hs, H, TB, H, Hpt, Mp));
        POI_TRY(launch_transpose_ld(e, RH, H, TB, H, RHt, Mp));
        POI_TRY(launch_gemm_atb_tc(e, DAt, Xt, Mp, 3 * H, din, s3, &g_ui));
        POI_TRY(launch_gemm_atb_tc(e, DAt, Hpt, Mp, 2 * H, H, s3, &g_whzr));
        POI_TRY(launch_gemm_atb_tc(e, DAt + (size_t)2 * H * Mp, RHt, Mp, H, H, s3, &g_whc));
        if (head) {
            float *St, *Hct;
            POI_TRY(arena_get(e, (size_t)nDp * Mp, &St)); POI_TRY(arena_get(e, (size_t)H * Mp, &Hct));
            POI_TRY(launch_transpose_ld(e, S, nDp, TB, nDp, St, Mp));
            POI_TRY(launch_transpose_ld(e, Hc, H, TB, H, Hct, Mp));
            POI_TRY(launch_gemm_atb_tc(e, St, Hct, Mp, nDp, H, s3, &g_vs));
        }
    } else {
        POI_TRY(launch_gemm_atb(e, DA, 3 * H, X, din, TB, 3 * H, din, &g_ui));
        POI_TRY(launch_gemm_atb(e, DA, 3 * H, Hs, H, TB, 2 * H, H, &g_whzr));
        POI_TRY(launch_gemm_atb(e, DA + 2 * H, 3 * H, RH, H, TB, H, H, &g_whc));
        if (head) POI_TRY(launch_gemm_atb(e, S, nDp, Hc, H, TB, nDp, H, &g_vs));
    }
    if (!have_colsums) {
        POI_TRY(launch_colsum(e, DA, 3 * H, TB, 3 * H, &g_bi));
        if (head) POI_TRY(launch_colsum(e, S, nDp, TB, nDp, &g_bs));
    }
    float* dg = mg ? mg->dense_grads : nullptr;
    ReduceSet rs; memset(&rs, 0, sizeof(rs));
    reduce_set_add(rs, g_ui, p->ui, din, 3 * H, din, dg ? dg + ML.ui : nullptr);
    reduce_set_add(rs, g_whzr, p->wh, H, 2 * H, H, dg ? dg + ML.wh : nullptr);
    reduce_set_add(rs, g_whc, p->wh + (size_t)2 * H * H, H, H, H, dg ? dg + ML.wh + (size_t)2 * H * H : nullptr);
    reduce_set_add(rs, g_bi, p->bi, 3 * H, 1, 3 * H, dg ? dg + ML.bi : nullptr);
    if (head) {
        reduce_set_add(rs, g_vs, p->vs, H, nD, H, dg ? dg + ML.vs : nullptr);
        reduce_set_add(rs, g_bs, p->bs, nD, 1, nD, dg ? dg + ML.bs : nullptr);
    }
    POI_TRY(launch_reduce_set(e, rs, alpha, lambda));
    POI_CAT(e, CAT_REDUCE, 0, 0);
    POI_LAUNCH(e, k_finalize_gru, 1, 256, 0, part, loss_blocks, p->scal, head ? 1 : 0,
               (double)n_nonempty * 0.6931471805599453, (double)scale, alpha, lambda, out_dev,
               mg ? mg->loss_sums : (double*)nullptr);
    phase_mark(e, 6);

    // ---- sparse row SGD: lt[unique(p u q)], di[unique(dp)] ----
    RowSrc src; memset(&src, 0, sizeof(src));
    src.mode = SRC_GRU_LT; src.grads = nullptr; src.DX = DX; src.ldx = din; src.Hc = Hc; src.ev = ev;
    src.B = B; src.T = T; src.LB = LB; src.dim = d;
    if (mg) { src.emit_rows = mg->row_grads; src.emit_cnt = mg->row_cnt; src.emit_by_key = 0; }
    // algorithmic bytes of the sparse step: 2 table rows per check-in, read + written,
    // plus the gradient rows that feed them (dx, and e*h for p and q)
    POI_TRY(launch_rows_update(e, seg_lt, p->lt, d, alpha, lambda, src, ROW_LONG_THRESH,
                               (double)TB * d * 4 * (4.0 + 3.0) + 8.0 * (double)LB));
    if (head) {
        src.mode = SRC_GRU_DI;
        if (mg) {   // di is replicated: emit a dense [nD x d] gradient + counts (all-reduced by the caller)
            POI_CK(e, cudaMemsetAsync(mg->dense_grads + ML.di, 0, (size_t)(ML.total - ML.di) * sizeof(float), e->stream));
            src.emit_rows = mg->dense_grads + ML.di; src.emit_cnt = mg->dense_grads + ML.dicnt; src.emit_by_key = 1;
        }
        POI_TRY(launch_rows_update(e, seg_di, p->di, d, alpha, lambda, src, ROW_LONG_THRESH,
                                   (double)TB * d * 4 + 4.0 * (double)LB, nD));
    }
    phase_mark(e, 7);

    POI_CK(e, cudaMemcpyAsync(e->h_out, out_dev, 8 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    phase_mark(e, 8);
    if (e->capturing) return 0;          // graph capture (poi_gru_train): the caller launches the graph and synchronises
    if (mg && mg->no_sync) return 0;
    POI_CK(e, cudaStreamSynchronize(e->stream));
    if (e->kprof) prof_harvest(e);
    if (out_host) for (int i = 0; i < 5; ++i) out_host[i] = e->h_out[i];
    if (e->timing) {
        cudaEventElapsedTime(&e->phase_ms[0], e->ev[0], e->ev[8]);
        for (int i = 1; i <= 7; ++i) cudaEventElapsedTime(&e->phase_ms[i], e->ev[i - 1], e->ev[i]);
    }
    return 0;
}

static int gru_alloc_idx(poi_engine* e, int B, int lmax, bool head, GruIdx* ix) {
    const size_t LB = (size_t)lmax * B;
    POI_TRY(arena_get(e, 2 * LB, &ix->PQt));
    ix->DPt = ix->DQt = nullptr;
    if (head) { POI_TRY(arena_get(e, LB, &ix->DPt)); POI_TRY(arena_get(e, LB, &ix->DQt)); }
    POI_TRY(arena_get(e, (size_t)B, &ix->lensB));
    return 0;
}

static int gru_upload_i32(poi_engine* e, const int32_t* host, size_t n, int32_t** dev, size_t* stage_off) {
    POI_TRY(arena_get(e, n, dev));
    // page-locked caller memory goes to the device directly (the call synchronises before it returns, so the
    // buffer outlives the copy); pageable memory is staged through the engine's pinned buffer
    cudaPointerAttributes at;
    // (while a graph is being captured everything goes through the engine's own pinned buffer: the replay refreshes it)
    if (!e->capturing && cudaPointerGetAttributes(&at, host) == cudaSuccess && at.type == cudaMemoryTypeHost) {
        POI_CK(e, cudaMemcpyAsync(*dev, host, n * 4, cudaMemcpyHostToDevice, e->stream));
        return 0;
    }
    cudaGetLastError();
    char* st = e->h_stage + *stage_off;
    memcpy(st, host, n * 4);
    POI_CK(e, cudaMemcpyAsync(*dev, st, n * 4, cudaMemcpyHostToDevice, e->stream));
    *stage_off += poi_align_up(n * 4, 256);
    return 0;
}
