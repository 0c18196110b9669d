// mg_step.cuh -- one multi-GPU mini-batch step as ONE engine call with no collective library on its path and no host
// synchronisation before the final read-back (poi_gru_step_mg in include/poi_engine.h).
//
// Users are sharded over ranks, the item table is row-sharded (owner = row % world, local row = row / world), dense
// weights and `di` are replicated.  Every rank owns a set of peer-visible buffers (poi_peer_alloc + cudaIpc, mapped by
// every rank once): its shard, its gradient OUTBOX (sorted unique ids, duplicate-summed gradient rows, counts, a
// permutation grouping the records by owner, the group sizes), its dense-gradient buffer and loss sums, and a flag
// array the peers write step counters into.  Step k on every rank, all on the engine stream:
//
//   slice + sort the batch's row ids                                   (local)
//   wait   flags[APPLIED][*] >= k-1                                     every owner has applied step k-1
//   gather the unique rows straight out of the owners' shards           (NVLink loads, k_gather_rows_sharded_dev)
//   forward + backward -> own outbox / dense buffer / loss sums         (gru_train_core, emit mode)
//   group the outbox records by owner                                   (one radix pass on id % world)
//   signal flags[READY][me] = k on every peer;  wait flags[READY][*] >= k
//   all-reduce of the dense gradients + loss sums: every rank reads every peer's buffer and adds in RANK ORDER
//        -> bit-identical dense weights on all ranks, no NCCL            (k_mg_allreduce)
//   dense SGD; sparse SGD of the own shard: records addressed to this owner are located through a direct-address table
//        (local row x source rank -> record number), the first source that holds a row sums the row's records over the
//        sources in rank order (reading the peers' outboxes over NVLink) and applies the step -- the same fixed order
//        as the all-to-all formulation, without the owner-side sort                       (k_mg_scatter_slots / k_mg_apply_rows)
//   signal flags[APPLIED][me] = k on every peer
//
// A waiting kernel gives up after POI_MG_TIMEOUT_MS (default 20 000) and raises an error flag that fails the call on
// the host: a rank that died or diverged turns into an error on its peers, not a hang.
#pragma once
#include "common.cuh"
#include "peer.cuh"

__device__ __forceinline__ unsigned long long mg_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// thread r < W polls flags[r] (written by peer r) until it reaches `want`
__global__ void k_mg_wait(const uint32_t* flags, int W, uint32_t want, int* err, unsigned long long timeout_ns, int code) {
    const int r = threadIdx.x;
    if (r >= W) return;
    const volatile uint32_t* f = flags + r;
    const unsigned long long t0 = mg_globaltimer();
    while (*f < want) {
        if (mg_globaltimer() - t0 > timeout_ns) { atomicExch(err, code + r); break; }
        __nanosleep(200);
    }
}

struct MgFlagPtrs { uint32_t* p[POI_MAX_PEERS]; };
// everything this rank wrote before (earlier kernels on the stream) is visible system-wide, then flags[slot] = value on
// every peer (and on itself)
__global__ void k_mg_signal(MgFlagPtrs fp, int W, int slot, uint32_t value) {
    __threadfence_system();
    const int r = threadIdx.x;
    if (r < W) *reinterpret_cast<volatile uint32_t*>(fp.p[r] + slot) = value;
    __threadfence_system();
}

template <int LPR, int UNR>
__global__ void __launch_bounds__(256)
k_gather_rows_sharded_dev(PeerTable pt, int dim4, const uint32_t* __restrict__ ids, const uint32_t* __restrict__ n_dev,
                          float* __restrict__ out) {
    const int64_t n_idx = *n_dev;
    const int lane = threadIdx.x % LPR;
    const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    const int64_t n_groups = (int64_t)gridDim.x * blockDim.x / LPR;
    float4* out4 = reinterpret_cast<float4*>(out);
    for (int64_t r0 = group * UNR; r0 < n_idx; r0 += n_groups * UNR) {
        const float4* src[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            src[u] = nullptr;
            if (r0 + u < n_idx) {
                const uint32_t id = ids[r0 + u];
                src[u] = reinterpret_cast<const float4*>(pt.shard[id % pt.world]) + (int64_t)(id / pt.world) * dim4;
            }
        }
        for (int c = lane; c < dim4; c += LPR) {
            float4 v[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u)
                if (src[u]) v[u] = src[u][c];
#pragma unroll
            for (int u = 0; u < UNR; ++u)
                if (src[u]) out4[(r0 + u) * dim4 + c] = v[u];
        }
    }
}

// outbox ids + owner keys over the whole capacity: entries past n get the sentinel key `world` (sorted to the end)
__global__ void k_mg_owner_keys(const uint32_t* __restrict__ uniq, const uint32_t* __restrict__ n_dev, int64_t cap, int world,
                                int32_t* __restrict__ ob_ids, uint32_t* __restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    if (i < *n_dev) { const uint32_t id = uniq[i]; ob_ids[i] = (int32_t)id; keys[i] = id % (uint32_t)world; }
    else keys[i] = (uint32_t)world;
}
// perm = record numbers grouped by owner; meta[o] = records for owner o, meta[W] = n_unique, meta[W+1] = B
__global__ void k_mg_perm_meta(const uint32_t* __restrict__ keys_sorted, const uint32_t* __restrict__ vals_sorted, int64_t cap, int world,
                               const uint32_t* __restrict__ n_dev, int B, int32_t* __restrict__ perm, int32_t* __restrict__ meta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) perm[i] = (int32_t)vals_sorted[i];
    if (blockIdx.x == 0 && threadIdx.x <= world + 1) {
        const int t = threadIdx.x;
        if (t == world) meta[t] = (int32_t)*n_dev;
        else if (t == world + 1) meta[t] = B;
        else {
            auto lower = [&](uint32_t key) { int64_t lo = 0, hi = cap; while (lo < hi) { int64_t m = (lo + hi) >> 1; if (keys_sorted[m] < key) lo = m + 1; else hi = m; } return lo; };
            meta[t] = (int32_t)(lower((uint32_t)t + 1) - lower((uint32_t)t));
        }
    }
}

struct MgDensePtrs { const float* dense[POI_MAX_PEERS]; const double* sums[POI_MAX_PEERS]; const int32_t* meta[POI_MAX_PEERS]; };
// out[i] = sum over ranks (rank order) of dense_r[i]; sums likewise (4 doubles); every rank must have used the same B
__global__ void __launch_bounds__(256)
k_mg_allreduce(MgDensePtrs dp, int W, int64_t n4, float4* __restrict__ out, double* __restrict__ sums_out, int B, int* err) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n4; i += nth) {
        float4 a = reinterpret_cast<const float4*>(dp.dense[0])[i];
        for (int r = 1; r < W; ++r) a = f4add(a, reinterpret_cast<const float4*>(dp.dense[r])[i]);
        out[i] = a;
    }
    if (tid < 4) {
        double s = 0.0;
        for (int r = 0; r < W; ++r) s += dp.sums[r][tid];
        sums_out[tid] = s;
    }
    if (tid >= 32 && tid < 32 + W) { if (dp.meta[tid - 32][W + 1] != B) atomicExch(err, 900); }
}

struct MgPull {
    const int32_t* perm[POI_MAX_PEERS]; const int32_t* ids[POI_MAX_PEERS]; const float* grads[POI_MAX_PEERS];
    const float* cnts[POI_MAX_PEERS]; const int32_t* meta[POI_MAX_PEERS];
    int world, me;
};
__device__ __forceinline__ void mg_group(const MgPull& pl, int p, int* off, int* cnt) {
    int o = 0;
    for (int q = 0; q < pl.me; ++q) o += pl.meta[p][q];
    *off = o; *cnt = pl.meta[p][pl.me];
}
// slot_tab[local row][source rank] = record number in that source's outbox (value < 0: -1 everywhere between steps)
__global__ void k_mg_scatter_slots(MgPull pl, int32_t* __restrict__ slot_tab, int clear) {
    const int p = blockIdx.y;
    int off, cnt; mg_group(pl, p, &off, &cnt);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += gridDim.x * blockDim.x) {
        const int s = pl.perm[p][off + t];
        const int x = pl.ids[p][s];
        slot_tab[(size_t)(x / pl.world) * pl.world + p] = clear ? -1 : s;
    }
}
template <int NCH>
__global__ void __launch_bounds__(256)
k_mg_apply_rows(MgPull pl, const int32_t* __restrict__ slot_tab, float* __restrict__ shard, int dim4, float alpha, float lambda) {
    const int lane = threadIdx.x & 31;
    const int gw = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nw = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
    const int W = pl.world;
    for (int p = 0; p < W; ++p) {
        int off, cnt; mg_group(pl, p, &off, &cnt);
        for (int t = gw; t < cnt; t += nw) {
            const int s = pl.perm[p][off + t];
            const size_t row = (size_t)(pl.ids[p][s] / W);
            const int32_t* tab = slot_tab + row * W;
            bool leader = true;
            for (int q = 0; q < p; ++q) leader = leader && tab[q] < 0;
            if (!leader) continue;                   // an earlier source holds this row too: its warp does the sum
            float4 acc[NCH];
#pragma unroll
            for (int k = 0; k < NCH; ++k) acc[k] = f4zero();
            float cntf = 0.f;
            for (int q = p; q < W; ++q) {            // sources in rank order = the arrival order of the all-to-all formulation
                const int s2 = q == p ? s : tab[q];
                if (s2 < 0) continue;
                if (pl.cnts[q]) cntf += pl.cnts[q][s2];
                const float* g = pl.grads[q] + (size_t)s2 * dim4 * 4;
#pragma unroll
                for (int k = 0; k < NCH; ++k) { const int c = lane + 32 * k; if (c < dim4) acc[k] = f4add(acc[k], ld4(g + 4 * c)); }
            }
            const float lc = lambda * cntf;
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int c = lane + 32 * k;
                if (c < dim4) {
                    float* rp = shard + (row * dim4 + c) * 4;
                    float4 r = ld4(rp);
                    r.x -= alpha * (acc[k].x + lc * r.x); r.y -= alpha * (acc[k].y + lc * r.y);
                    r.z -= alpha * (acc[k].z + lc * r.z); r.w -= alpha * (acc[k].w + lc * r.w);
                    st4(rp, r);
                }
            }
        }
    }
}

// loss scalars + wd / loss_weight SGD from the all-reduced sums; sums[3] = n_nonempty_global * ln 2 (plain GRU's t = 0 term)
__global__ void k_finalize_from_sums4(const double* __restrict__ sums, float* scal, int head, double scale, float alpha, float lambda,
                                      double* __restrict__ out, const int* __restrict__ err) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        finalize_apply(sums[0], sums[1], sums[2], scal, head, sums[3], scale, alpha, lambda, out);
        out[7] = (double)*err;
    }
}
__global__ void k_mg_set_extra(double* sums, double v) { if (threadIdx.x == 0 && blockIdx.x == 0) sums[3] = v; }

// index rows either resident on the device (uidx_host = the B row numbers) or supplied from host memory (rows_host != NULL:
// [B x lmax] each, copied inside the call -- the end-to-end path)
struct MgHostRows { const int32_t* p; const int32_t* q; const int32_t* dp; const int32_t* dq; const int32_t* lens; };

static int gru_step_mg_body(poi_engine* e, const poi_gru_params* p, const poi_seq_index* index, const int32_t* uidx_host,
                            const MgHostRows* rows_host, int32_t B, int32_t lmax, int32_t max_len, const poi_mg_peers* pr,
                            int64_t step, float alpha, float lambda, double* out_host) {
    e->prep_valid = false;
    POI_TRY(begin_call(e));
    if (!p || !p->ui || !p->wh || !p->bi) POI_FAIL(e, "gru params: null pointer");
    if (p->d <= 0 || p->d % 4 || p->H != p->d) POI_FAIL(e, "n_in must equal n_hidden and be a multiple of 4");
    const bool head = p->di != nullptr;
    if (!rows_host && (!index || !index->p || !index->q || !index->lens || (head && (!index->dp || !index->dq)))) POI_FAIL(e, "index matrices missing");
    if (rows_host && (!rows_host->p || !rows_host->q || !rows_host->lens || (head && (!rows_host->dp || !rows_host->dq)))) POI_FAIL(e, "host index rows missing");
    if (!pr || pr->world < 1 || pr->world > POI_MAX_PEERS || pr->rank < 0 || pr->rank >= pr->world) POI_FAIL(e, "bad peer table");
    if (B <= 0 || lmax <= 0 || step < 1) POI_FAIL(e, "bad batch size / step number");
    const int W = pr->world, me = pr->rank, d = p->d, H = p->H, din = head ? 2 * d : d, nD = head ? p->n_rows_di : 0;
    const int64_t LB = (int64_t)lmax * B;
    if (2 * LB > pr->cap) POI_FAIL(e, "batch needs up to %lld outbox records, capacity is %lld", (long long)(2 * LB), (long long)pr->cap);
    const MgLayout ML = mg_layout(H, din, nD, d);
    unsigned long long timeout_ns = 20000ull * 1000000ull;
    if (const char* tm = getenv("POI_MG_TIMEOUT_MS")) timeout_ns = strtoull(tm, nullptr, 10) * 1000000ull;

    int* err = nullptr;
    POI_TRY(arena_get(e, 4, &err));
    POI_CK(e, cudaMemsetAsync(err, 0, 4, e->stream));
    phase_mark(e, 0);
    GruIdx ix;
    int64_t n_nonempty = 0;
    if (rows_host) {
        POI_TRY(stage_reserve(e, (head ? 4 : 2) * ((size_t)LB * 4 + 256) + (size_t)B * 4 + 256));
        size_t so = 0;
        int32_t *P = nullptr, *Q = nullptr, *DP = nullptr, *DQ = nullptr, *lens = nullptr;
        POI_TRY(gru_upload_i32(e, rows_host->p, (size_t)LB, &P, &so));
        POI_TRY(gru_upload_i32(e, rows_host->q, (size_t)LB, &Q, &so));
        if (head) { POI_TRY(gru_upload_i32(e, rows_host->dp, (size_t)LB, &DP, &so)); POI_TRY(gru_upload_i32(e, rows_host->dq, (size_t)LB, &DQ, &so)); }
        POI_TRY(gru_upload_i32(e, rows_host->lens, (size_t)B, &lens, &so));
        max_len = 0;
        for (int b = 0; b < B; ++b) { max_len = std::max(max_len, rows_host->lens[b]); n_nonempty += rows_host->lens[b] >= 1; }
        POI_TRY(gru_alloc_idx(e, B, lmax, head, &ix));
        POI_CAT(e, CAT_INDEX, 0, 0);
        POI_LAUNCH(e, k_slice_indices, (unsigned)poi_cdiv(LB, 256), 256, 0, P, Q, DP, DQ, lens, lmax, (const int32_t*)nullptr, B,
                   ix.PQt, ix.DPt, ix.DQt, ix.lensB);
    } else {
        POI_TRY(stage_reserve(e, (size_t)B * 4 + 256));
        size_t so = 0;
        int32_t* uidx_dev = nullptr;
        POI_TRY(gru_upload_i32(e, uidx_host, (size_t)B, &uidx_dev, &so));
        POI_TRY(gru_alloc_idx(e, B, lmax, head, &ix));
        POI_CAT(e, CAT_INDEX, 0, 0);
        POI_LAUNCH(e, k_slice_indices, (unsigned)poi_cdiv(LB, 256), 256, 0, index->p, index->q, head ? index->dp : nullptr,
                   head ? index->dq : nullptr, index->lens, lmax, uidx_dev, B, ix.PQt, ix.DPt, ix.DQt, ix.lensB);
        n_nonempty = B;                                  // reference data: every user has L >= 1 (GRU.py:352)
    }
    if (head) n_nonempty = 0;
    SegList seg_lt, seg_di;
    POI_TRY(build_segments(e, reinterpret_cast<const uint32_t*>(ix.PQt), 2 * LB, (uint32_t)p->n_rows_lt, true, &seg_lt));
    if (head) POI_TRY(build_segments(e, reinterpret_cast<const uint32_t*>(ix.DPt), LB, (uint32_t)nD, false, &seg_di));

    // ---- every owner has applied the previous step -> the shards may be read ----
    POI_CAT(e, CAT_OTHER, 0, 0);
    POI_LAUNCH(e, k_mg_wait, 1, 32, 0, pr->flags[me] + W, W, (uint32_t)(step - 1), err, timeout_ns, 100);
    float* rows = nullptr;
    POI_TRY(arena_get(e, (size_t)2 * LB * d, &rows));
    {
        PeerTable pt; memset(&pt, 0, sizeof(pt));
        pt.world = W;
        for (int r = 0; r < W; ++r) pt.shard[r] = pr->shard[r];
        const int dim4 = d / 4;
        POI_CAT(e, CAT_GATHER, 0, 0);
        if (peer_gather_mode() == 1 && d * 4 * PG_WARPS * PG_ST <= (200 << 10)) {
            POI_TRY(launch_gather_bulk(e, pt, d, seg_lt.uniq, seg_lt.n_unique, 0, 2 * LB, rows));
        } else {
        const int lpr = dim4 <= 8 ? 8 : (dim4 <= 16 ? 16 : 32);
        const int64_t threads_needed = poi_cdiv(2 * LB, 4) * lpr;
        unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(poi_cdiv(threads_needed, 256), (int64_t)e->num_sms * 16));
        if (lpr == 8)       POI_LAUNCH(e, (k_gather_rows_sharded_dev<8, 4>), grid, 256, 0, pt, dim4, seg_lt.uniq, seg_lt.n_unique, rows);
        else if (lpr == 16) POI_LAUNCH(e, (k_gather_rows_sharded_dev<16, 4>), grid, 256, 0, pt, dim4, seg_lt.uniq, seg_lt.n_unique, rows);
        else                POI_LAUNCH(e, (k_gather_rows_sharded_dev<32, 4>), grid, 256, 0, pt, dim4, seg_lt.uniq, seg_lt.n_unique, rows);
        }
    }
    // ---- forward + backward into the own outbox / dense buffer / loss sums (nothing is updated) ----
    MgCtx mg; mg.rows = rows; mg.global_batch = B * W; mg.dense_grads = pr->dense[me];
    mg.row_grads = pr->ob_grads[me]; mg.row_cnt = pr->ob_cnts[me]; mg.loss_sums = pr->sums[me]; mg.no_sync = true;
    PreSeg pre{&seg_lt, &seg_di};
    POI_TRY(gru_train_core(e, p, ix, B, lmax, max_len, 0, 0.f, 0.f, nullptr, &mg, &pre));
    POI_CAT(e, CAT_REDUCE, 0, 0);
    POI_LAUNCH(e, k_mg_set_extra, 1, 32, 0, pr->sums[me], (double)n_nonempty * 0.6931471805599453);
    // ---- outbox: ids, permutation grouped by owner, group sizes ----
    POI_CAT(e, CAT_INDEX, 0, 0);
    uint32_t *okeys = nullptr, *ks = nullptr, *vs = nullptr;
    POI_TRY(arena_get(e, (size_t)pr->cap, &okeys));
    POI_LAUNCH(e, k_mg_owner_keys, (unsigned)poi_cdiv(pr->cap, 256), 256, 0, seg_lt.uniq, seg_lt.n_unique, pr->cap, W, pr->ob_ids[me], okeys);
    POI_TRY(sort_pairs(e, okeys, pr->cap, (uint32_t)W + 1, &ks, &vs));
    POI_LAUNCH(e, k_mg_perm_meta, (unsigned)poi_cdiv(pr->cap, 256), 256, 0, ks, vs, pr->cap, W, seg_lt.n_unique, (int)B, pr->ob_perm[me], pr->ob_meta[me]);
    // ---- exchange point: my outbox is written; wait for everybody's ----
    MgFlagPtrs fp; memset(&fp, 0, sizeof(fp));
    for (int r = 0; r < W; ++r) fp.p[r] = pr->flags[r];
    POI_CAT(e, CAT_OTHER, 0, 0);
    POI_LAUNCH(e, k_mg_signal, 1, 32, 0, fp, W, me, (uint32_t)step);
    POI_LAUNCH(e, k_mg_wait, 1, 32, 0, pr->flags[me], W, (uint32_t)step, err, timeout_ns, 200);
    // ---- dense all-reduce in rank order + dense SGD ----
    float* dred = nullptr; double* sred = nullptr; double* out_dev = nullptr;
    POI_TRY(arena_get(e, (size_t)ML.total, &dred));
    POI_TRY(arena_get(e, 4, &sred));
    POI_TRY(arena_get(e, 8, &out_dev));
    {
        MgDensePtrs dpz; memset(&dpz, 0, sizeof(dpz));
        for (int r = 0; r < W; ++r) { dpz.dense[r] = pr->dense[r]; dpz.sums[r] = pr->sums[r]; dpz.meta[r] = pr->ob_meta[r]; }
        const int64_t n4 = ML.total / 4;
        POI_CAT(e, CAT_WGRAD, 0, 0);
        unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(poi_cdiv(n4, 256), (int64_t)e->num_sms * 4));
        POI_LAUNCH(e, k_mg_allreduce, grid, 256, 0, dpz, W, n4, reinterpret_cast<float4*>(dred), sred, (int)B, err);
    }
    auto apply = [&](float* theta, int64_t off, int64_t n) -> int {
        if (n <= 0) return 0;
        POI_LAUNCH(e, k_dense_apply, (unsigned)poi_cdiv(n, 256), 256, 0, theta, dred + off, n, alpha, lambda);
        return 0;
    };
    POI_TRY(apply(p->ui, ML.ui, (int64_t)3 * H * din));
    POI_TRY(apply(p->wh, ML.wh, (int64_t)3 * H * H));
    POI_TRY(apply(p->bi, ML.bi, 3 * H));
    if (head) {
        POI_TRY(apply(p->vs, ML.vs, (int64_t)nD * H));
        POI_TRY(apply(p->bs, ML.bs, nD));
        POI_LAUNCH(e, k_di_apply, (unsigned)poi_cdiv((int64_t)nD * d, 256), 256, 0, p->di, dred + ML.di, dred + ML.dicnt, nD, d, alpha, lambda);
    }
    // ---- sparse SGD of the own shard from every outbox ----
    {
        MgPull pl; memset(&pl, 0, sizeof(pl));
        pl.world = W; pl.me = me;
        for (int r = 0; r < W; ++r) { pl.perm[r] = pr->ob_perm[r]; pl.ids[r] = pr->ob_ids[r]; pl.grads[r] = pr->ob_grads[r]; pl.cnts[r] = pr->ob_cnts[r]; pl.meta[r] = pr->ob_meta[r]; }
        POI_CAT(e, CAT_ROWS, 0, 0);
        dim3 sg((unsigned)std::min<int64_t>(poi_cdiv(pr->cap, 256), 256), (unsigned)W);
        POI_LAUNCH(e, k_mg_scatter_slots, sg, 256, 0, pl, pr->slot_tab, 0);
        const int dim4 = d / 4;
        unsigned grid = (unsigned)((int64_t)e->num_sms * 8);
        if (dim4 <= 32)       POI_LAUNCH(e, (k_mg_apply_rows<1>), grid, 256, 0, pl, pr->slot_tab, pr->shard[me], dim4, alpha, lambda);
        else if (dim4 <= 64)  POI_LAUNCH(e, (k_mg_apply_rows<2>), grid, 256, 0, pl, pr->slot_tab, pr->shard[me], dim4, alpha, lambda);
        else if (dim4 <= 128) POI_LAUNCH(e, (k_mg_apply_rows<4>), grid, 256, 0, pl, pr->slot_tab, pr->shard[me], dim4, alpha, lambda);
        else if (dim4 <= 256) POI_LAUNCH(e, (k_mg_apply_rows<8>), grid, 256, 0, pl, pr->slot_tab, pr->shard[me], dim4, alpha, lambda);
        else POI_FAIL(e, "row dim %d too large (max 1024)", d);
        POI_LAUNCH(e, k_mg_scatter_slots, sg, 256, 0, pl, pr->slot_tab, 1);
    }
    POI_CAT(e, CAT_REDUCE, 0, 0);
    POI_LAUNCH(e, k_finalize_from_sums4, 1, 32, 0, sred, p->scal, head ? 1 : 0, 1.0 / (double)(B * W), alpha, lambda, out_dev, err);
    POI_CAT(e, CAT_OTHER, 0, 0);
    POI_LAUNCH(e, k_mg_signal, 1, 32, 0, fp, W, W + me, (uint32_t)step);
    POI_CK(e, cudaMemcpyAsync(e->h_out, out_dev, 8 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    POI_CK(e, cudaStreamSynchronize(e->stream));
    if (e->kprof) prof_harvest(e);
    const int ec = (int)e->h_out[7];
    if (ec >= 100 && ec < 300) POI_FAIL(e, "multi-GPU step %lld: timed out waiting for rank %d (%s)", (long long)step, ec % 100,
                                        ec < 200 ? "has not applied the previous step" : "has not published its outbox");
    if (ec == 900) POI_FAIL(e, "multi-GPU step %lld: the ranks passed different batch sizes; this step's update is wrong", (long long)step);
    for (int i = 0; i < 5; ++i) out_host[i] = e->h_out[i];
    return 0;
}

extern "C" int poi_gru_step_mg(poi_engine* e, const poi_gru_params* p, const poi_seq_index* index, const int32_t* uidx_host,
                               int32_t B, int32_t max_len, const poi_mg_peers* pr, int64_t step, float alpha, float lambda,
                               double* out_host) {
    if (!index) POI_FAIL(e, "index matrices missing");
    return gru_step_mg_body(e, p, index, uidx_host, nullptr, B, index->lmax, max_len, pr, step, alpha, lambda, out_host);
}

extern "C" int poi_gru_step_mg_host_rows(poi_engine* e, const poi_gru_params* p, const int32_t* p_host, const int32_t* q_host,
                                         const int32_t* dp_host, const int32_t* dq_host, const int32_t* lens_host, int32_t B,
                                         int32_t lmax, const poi_mg_peers* pr, int64_t step, float alpha, float lambda,
                                         double* out_host) {
    MgHostRows hr{p_host, q_host, dp_host, dq_host, lens_host};
    return gru_step_mg_body(e, p, nullptr, nullptr, &hr, B, lmax, 0, pr, step, alpha, lambda, out_host);
}
