// sampling.cuh -- SURVEY.md 8(f2): the per-epoch host loops of the reference on the device.
//
//   fun_random_neg_masks_tra / _tes (Load_Data_by_length.py:127-162): for every valid position of a user's padded
//       row draw j uniformly from [0, n_item) and redraw while j occurs in the user's training row (test variant:
//       training OR test row); positions from the first pad on get the pad id n_item.
//   fun_compute_dist_neg (Load_Data_by_length.py:165-180): interval(q[t], p[t-1]) for 1 <= t < L with cal_dis's
//       haversine (:24-42), dist_num at t = 0 and on the padding.
//
// The reference draws from Python's global Mersenne Twister, one sequential stream over all users: that stream cannot
// be reproduced in parallel, so the device sampler is a NEW, counter-based stream (Philox4x32-10, Salmon et al. 2011):
// word = philox(counter = (user, position, attempt / 4, epoch), key = seed)[attempt % 4], j = (word * n_item) >> 32.
// It is independent of launch geometry and restated bit-exactly in oracle/sampling.py; the reference's own sampler
// stays available on the host for parity runs.  The rejection test is a binary search in the user's sorted row(s).
#pragma once
#include "common.cuh"

struct Philox4 { uint32_t v[4]; };

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

__device__ __forceinline__ bool row_contains(const int32_t* __restrict__ sorted_row, int n, int32_t key) {
    int lo = 0, hi = n;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (sorted_row[m] < key) lo = m + 1; else hi = m; }
    return lo < n && sorted_row[lo] == key;
}

// one thread per (user, position); rows: the rows the negatives are drawn FOR (train rows, or test rows); sorted_a /
// sorted_b: the user's forbidden rows sorted ascending (b may be NULL)
__global__ void __launch_bounds__(256)
k_sample_negatives(const int32_t* __restrict__ rows, int lrow, const int32_t* __restrict__ sorted_a, int la,
                   const int32_t* __restrict__ sorted_b, int lb, int n_user, int n_item, uint32_t k0, uint32_t k1,
                   uint32_t epoch, int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_user * lrow) return;
    const int u = (int)(i / lrow), t = (int)(i % lrow);
    // valid = before the first pad of the row (the reference breaks at the first pad id, :134-136)
    if (rows[i] == n_item) { out[i] = n_item; return; }
    const int32_t* sa = sorted_a + (int64_t)u * la;
    const int32_t* sb = sorted_b ? sorted_b + (int64_t)u * lb : nullptr;
    for (uint32_t blk = 0;; ++blk) {
        const Philox4 r = philox4x32_10((uint32_t)u, (uint32_t)t, blk, epoch, k0, k1);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int32_t j = (int32_t)(((uint64_t)r.v[w] * (uint64_t)(uint32_t)n_item) >> 32);
            if (!row_contains(sa, la, j) && !(sb && row_contains(sb, lb, j))) { out[i] = j; return; }
        }
    }
}

// cal_dis (Load_Data_by_length.py:24-42) in fp64
__device__ __forceinline__ int haversine_interval(double lat1, double lon1, double lat2, double lon2, double dd, int dist_num) {
    const double d = 12742.0, p = 0.017453292519943295;
    const double a = (lat1 - lat2) * p, b = (lon1 - lon2) * p;
    double c = (1.0 - cos(a)) / 2 + cos(lat1 * p) * cos(lat2 * p) * (1.0 - cos(b)) / 2;
    c = c < 0.0 ? 0.0 : (c > 1.0 ? 1.0 : c);       // rounding outside [0, 1] would make asin NaN and the interval id negative
    const double dist = d * asin(sqrt(c));
    const double iv = dist * 1000 / dd;
    const int interval = iv >= 2147483647.0 ? 2147483647 : (iv > 0.0 ? (int)iv : 0);
    return interval < dist_num ? interval : dist_num;
}

__global__ void __launch_bounds__(256)
k_neg_intervals(const int32_t* __restrict__ p, const int32_t* __restrict__ q, const int32_t* __restrict__ lens, int n_user,
                int lmax, const double* __restrict__ coords, double dd, int dist_num, int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_user * lmax) return;
    const int u = (int)(i / lmax), t = (int)(i % lmax);
    if (t == 0 || t >= lens[u]) { out[i] = dist_num; return; }
    const int32_t pre = p[i - 1], cur = q[i];
    out[i] = haversine_interval(coords[2 * (int64_t)cur], coords[2 * (int64_t)cur + 1], coords[2 * (int64_t)pre],
                                coords[2 * (int64_t)pre + 1], dd, dist_num);
}
