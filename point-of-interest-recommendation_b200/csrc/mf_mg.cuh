// mf_mg.cuh -- row-sharded multi-GPU step for the pairwise embedding models (last bullet; BASELINE.json C4:
// "GeoIE ... 2 x B200 row-sharded").  Users are split over ranks, the item tables g / h / z are row-sharded (owner = row %
// world, local row = row / world), a and b are replicated.  Same peer-memory protocol as the GRU step (mg_step.cuh): no
// collective library on the path, flags for "my outbox is written" / "I have applied the step", waiting kernels that time
// out into an error.  Per step and rank:
//
//   keys + segments of the batch (bit-exact integer work, as on one GPU)
//   wait (every owner has applied the previous step)
//   COMPACT copies of the rows the batch touches, gathered straight out of the owners' shards over NVLink (one compact
//        table per sharded table; slot = position in the sorted unique id list) + compact coordinate lists
//   the single-GPU mini-batch step (geoie_batch_core) on the compact tables with slot indices: every term from
//        pre-update values, duplicates inside the rank's batch summed in fixed order
//   outbox: per unique row  delta = (old - new) / alpha  (old re-read from the owner's shard, still untouched) -- this IS
//        the rank's summed gradient including its L2 share, because new = old - alpha (grad + lambda cnt old)
//   group the outbox by owner, signal / wait
//   a, b: every rank adds the ranks' partial gradients in rank order and applies the same step
//   owner: row -= alpha * sum over ranks (rank order) of delta -- located through the direct-address table of mg_step.cuh
//   signal "applied"
// G ranks x batch B equals one rank x batch G B up to float32 rounding of the deltas (tests/test_gpu_multigpu.py).
#pragma once
#include "mg_step.cuh"

__global__ void k_mf_gather_coords(const float4* __restrict__ coords, const uint32_t* __restrict__ uniq, const uint32_t* __restrict__ n_dev,
                                   int64_t cap, float4* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap && i < *n_dev) out[i] = coords[uniq[i]];
}
__global__ void k_mf_iota(uint32_t* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)i;
}
// slot index arrays for the compact tables: Pg[b, j] (j < n), Ph[b, i+1], Qs[b, i+1, k]
__global__ void k_geoie_slots(const uint32_t* __restrict__ slot_h, const uint32_t* __restrict__ slot_g, int Bu, int L, int K,
                              int32_t* __restrict__ Pg, int32_t* __restrict__ Ph, int32_t* __restrict__ Qs) {
    const int n = L - 1, C = K + 1;
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o < (int64_t)Bu * n * C) {
        const int c = (int)(o % C); const int64_t bi = o / C; const int i = (int)(bi % n); const int b = (int)(bi / n);
        if (c == 0) Ph[(size_t)b * L + i + 1] = (int32_t)slot_h[o];
        else Qs[((size_t)b * L + i + 1) * K + c - 1] = (int32_t)slot_h[o];
    }
    if (o < (int64_t)Bu * n) { const int j = (int)(o % n), b = (int)(o / n); Pg[(size_t)b * L + j] = (int32_t)slot_g[o]; }
    if (o < Bu) { Pg[(size_t)o * L + L - 1] = 0; Ph[(size_t)o * L] = 0; }
}
// outbox gradient rows: (row as it still is in the owner's shard - row after this rank's step) / alpha
__global__ void __launch_bounds__(256)
k_mf_delta(PeerTable pt, int dim4, const uint32_t* __restrict__ uniq, const uint32_t* __restrict__ n_dev, const float* __restrict__ cnew,
           float inv_alpha, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n = *n_dev;
    for (int64_t r = gw; r < n; r += nw) {
        const uint32_t id = uniq[r];
        const float4* src = reinterpret_cast<const float4*>(pt.shard[id % pt.world]) + (int64_t)(id / pt.world) * dim4;
        for (int c = lane; c < dim4; c += 32) {
            const float4 o = src[c], w = reinterpret_cast<const float4*>(cnew)[r * dim4 + c];
            reinterpret_cast<float4*>(out)[r * dim4 + c] = make_float4((o.x - w.x) * inv_alpha, (o.y - w.y) * inv_alpha, (o.z - w.z) * inv_alpha, (o.w - w.w) * inv_alpha);
        }
    }
}
struct MfSumPtrs { const double* sums[POI_MAX_PEERS]; };
// loss and the a, b gradients summed over the ranks in rank order; a, b <- a, b - alpha * gradient (GeoIE.py:91,172-173)
__global__ void k_mf_sums_apply(MfSumPtrs sp, int W, double* ab, float alpha, double* out, const int* err) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double loss = 0.0, ga = 0.0, gb = 0.0;
    for (int r = 0; r < W; ++r) { loss += sp.sums[r][0]; ga += sp.sums[r][1]; gb += sp.sums[r][2]; }
    out[0] = loss;
    if (ab) { ab[0] -= (double)alpha * ga; ab[1] -= (double)alpha * gb; }
    out[7] = (double)*err;
}

static int mf_gather_sharded(poi_engine* e, float* const* shards, int W, int dim, const uint32_t* uniq, const uint32_t* n_dev,
                             int64_t cap, float* out) {
    PeerTable pt; memset(&pt, 0, sizeof(pt));
    pt.world = W;
    for (int r = 0; r < W; ++r) pt.shard[r] = shards[r];
    const int dim4 = dim / 4;
    POI_CAT(e, CAT_GATHER, 0, 0);
    const int lpr = dim4 <= 8 ? 8 : (dim4 <= 16 ? 16 : 32);
    const int64_t threads_needed = poi_cdiv(cap, 4) * lpr;
    unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(poi_cdiv(threads_needed, 256), (int64_t)e->num_sms * 16));
    if (lpr == 8)       POI_LAUNCH(e, (k_gather_rows_sharded_dev<8, 4>), grid, 256, 0, pt, dim4, uniq, n_dev, out);
    else if (lpr == 16) POI_LAUNCH(e, (k_gather_rows_sharded_dev<16, 4>), grid, 256, 0, pt, dim4, uniq, n_dev, out);
    else                POI_LAUNCH(e, (k_gather_rows_sharded_dev<32, 4>), grid, 256, 0, pt, dim4, uniq, n_dev, out);
    return 0;
}

static int mf_apply_set(poi_engine* e, const poi_mf_peers* pr, int set, const int* tables, int n_tables, int dim, float alpha) {
    const int W = pr->world, me = pr->rank;
    MgPull pl; memset(&pl, 0, sizeof(pl));
    pl.world = W; pl.me = me;
    for (int r = 0; r < W; ++r) { pl.perm[r] = pr->ob_perm[set][r]; pl.ids[r] = pr->ob_ids[set][r]; pl.meta[r] = pr->ob_meta[set][r]; pl.cnts[r] = nullptr; }
    POI_CAT(e, CAT_ROWS, 0, 0);
    dim3 sg((unsigned)std::min<int64_t>(poi_cdiv(pr->cap[set], 256), 256), (unsigned)W);
    POI_LAUNCH(e, k_mg_scatter_slots, sg, 256, 0, pl, pr->slot_tab[set], 0);
    const int dim4 = dim / 4;
    unsigned grid = (unsigned)((int64_t)e->num_sms * 8);
    for (int ti = 0; ti < n_tables; ++ti) {
        const int t = tables[ti];
        for (int r = 0; r < W; ++r) pl.grads[r] = pr->ob_grads[t][r];
        // row -= alpha * (sum of the ranks' deltas); the L2 share is inside the deltas (lambda = 0 here)
        if (dim4 <= 32)       POI_LAUNCH(e, (k_mg_apply_rows<1>), grid, 256, 0, pl, pr->slot_tab[set], pr->shard[t][me], dim4, alpha, 0.f);
        else if (dim4 <= 64)  POI_LAUNCH(e, (k_mg_apply_rows<2>), grid, 256, 0, pl, pr->slot_tab[set], pr->shard[t][me], dim4, alpha, 0.f);
        else if (dim4 <= 128) POI_LAUNCH(e, (k_mg_apply_rows<4>), grid, 256, 0, pl, pr->slot_tab[set], pr->shard[t][me], dim4, alpha, 0.f);
        else                  POI_LAUNCH(e, (k_mg_apply_rows<8>), grid, 256, 0, pl, pr->slot_tab[set], pr->shard[t][me], dim4, alpha, 0.f);
    }
    POI_LAUNCH(e, k_mg_scatter_slots, sg, 256, 0, pl, pr->slot_tab[set], 1);
    return 0;
}

static int mf_publish_set(poi_engine* e, const poi_mf_peers* pr, int set, const SegList& seg, int Bu) {
    const int W = pr->world, me = pr->rank;
    uint32_t *okeys = nullptr, *ks = nullptr, *vs = nullptr;
    POI_TRY(arena_get(e, (size_t)pr->cap[set], &okeys));
    POI_CAT(e, CAT_INDEX, 0, 0);
    POI_LAUNCH(e, k_mg_owner_keys, (unsigned)poi_cdiv(pr->cap[set], 256), 256, 0, seg.uniq, seg.n_unique, pr->cap[set], W, pr->ob_ids[set][me], okeys);
    POI_TRY(sort_pairs(e, okeys, pr->cap[set], (uint32_t)W + 1, &ks, &vs));
    POI_LAUNCH(e, k_mg_perm_meta, (unsigned)poi_cdiv(pr->cap[set], 256), 256, 0, ks, vs, pr->cap[set], W, seg.n_unique, Bu, pr->ob_perm[set][me], pr->ob_meta[set][me]);
    return 0;
}

extern "C" int poi_geoie_step_mg(poi_engine* e, const poi_geoie_params* prm, const int32_t* P, const int32_t* Q,
                                 const float* coords_dev, int32_t Bu, int32_t L, int32_t K, int32_t on_host,
                                 const poi_mf_peers* pr, int64_t step, float alpha, float lambda, double* loss_host) {
    POI_TRY(begin_call(e));
    if (!prm || !prm->ab) POI_FAIL(e, "geoie params: null pointer");
    if (prm->H <= 0 || prm->H % 4 || prm->H > 512) POI_FAIL(e, "n_hidden must be a multiple of 4, <= 512");
    if (L < 2 || L - 1 > GEO_MAXN) POI_FAIL(e, "sequence length must be in [2, %d] for the mini-batch kernel", GEO_MAXN + 1);
    if (K < 1 || K > 128 || !coords_dev) POI_FAIL(e, "1 <= K <= 128 and a coordinate table are required");
    if (!pr || pr->world < 1 || pr->world > POI_MAX_PEERS || pr->rank < 0 || pr->rank >= pr->world) POI_FAIL(e, "bad peer table");
    if (Bu <= 0 || step < 1) POI_FAIL(e, "bad batch size / step number");
    const int H = prm->H, W = pr->world, me = pr->rank, n = L - 1, C = K + 1;
    const int64_t n_occ = (int64_t)Bu * n * C, n_g = (int64_t)Bu * n;
    if (n_occ >= (int64_t)1 << 31) POI_FAIL(e, "batch too large");
    if (n_g > pr->cap[0] || n_occ > pr->cap[1]) POI_FAIL(e, "batch exceeds the outbox capacity (%lld / %lld records)", (long long)pr->cap[0], (long long)pr->cap[1]);
    unsigned long long timeout_ns = 20000ull * 1000000ull;
    if (const char* tm = getenv("POI_MG_TIMEOUT_MS")) timeout_ns = strtoull(tm, nullptr, 10) * 1000000ull;
    int* err = nullptr;
    POI_TRY(arena_get(e, 4, &err));
    POI_CK(e, cudaMemsetAsync(err, 0, 4, e->stream));

    GeoBatch gg; gg.Bu = Bu; gg.L = L; gg.K = K;                         // global ids
    gg.coords_g = gg.coords_h = reinterpret_cast<const float4*>(coords_dev);
    if (on_host) {
        const void* hs[2] = {P, Q}; size_t bs[2] = {(size_t)Bu * L * 4, (size_t)Bu * L * K * 4}; void* dv[2];
        POI_TRY(upload_many(e, hs, bs, 2, dv));
        gg.P = (const int32_t*)dv[0]; gg.Q = (const int32_t*)dv[1];
    } else { gg.P = P; gg.Q = Q; }
    gg.Ph = gg.P;
    uint32_t *keys_h = nullptr, *keys_g = nullptr;
    POI_TRY(arena_get(e, (size_t)n_occ, &keys_h));
    POI_TRY(arena_get(e, (size_t)n_g, &keys_g));
    POI_CAT(e, CAT_INDEX, 0, 0);
    POI_LAUNCH(e, k_geoie_keys, (unsigned)poi_cdiv(n_occ, 256), 256, 0, gg, keys_h, keys_g);
    SegList seg_h, seg_g;
    POI_TRY(build_segments(e, keys_h, n_occ, (uint32_t)prm->n_rows, true, &seg_h));      // prm->n_rows = GLOBAL row count
    POI_TRY(build_segments(e, keys_g, n_g, (uint32_t)prm->n_rows, true, &seg_g));

    // ---- every owner has applied the previous step -> gather the compact tables from the owners' shards ----
    POI_CAT(e, CAT_OTHER, 0, 0);
    POI_LAUNCH(e, k_mg_wait, 1, 32, 0, pr->flags[me] + W, W, (uint32_t)(step - 1), err, timeout_ns, 100);
    float *cg = nullptr, *ch = nullptr, *cz = nullptr; float4 *xg = nullptr, *xh = nullptr;
    POI_TRY(arena_get(e, (size_t)n_g * H, &cg)); POI_TRY(arena_get(e, (size_t)n_occ * H, &ch)); POI_TRY(arena_get(e, (size_t)n_occ * H, &cz));
    POI_TRY(arena_get(e, (size_t)n_g, &xg)); POI_TRY(arena_get(e, (size_t)n_occ, &xh));
    POI_TRY(mf_gather_sharded(e, pr->shard[0], W, H, seg_g.uniq, seg_g.n_unique, n_g, cg));
    POI_TRY(mf_gather_sharded(e, pr->shard[1], W, H, seg_h.uniq, seg_h.n_unique, n_occ, ch));
    POI_TRY(mf_gather_sharded(e, pr->shard[2], W, H, seg_h.uniq, seg_h.n_unique, n_occ, cz));
    POI_LAUNCH(e, k_mf_gather_coords, (unsigned)poi_cdiv(n_g, 256), 256, 0, gg.coords_g, seg_g.uniq, seg_g.n_unique, n_g, xg);
    POI_LAUNCH(e, k_mf_gather_coords, (unsigned)poi_cdiv(n_occ, 256), 256, 0, gg.coords_g, seg_h.uniq, seg_h.n_unique, n_occ, xh);
    // ---- slot indices + segments whose keys are slots ----
    int32_t *Pg = nullptr, *Ph = nullptr, *Qs = nullptr; uint32_t* iota = nullptr;
    POI_TRY(arena_get(e, (size_t)Bu * L, &Pg)); POI_TRY(arena_get(e, (size_t)Bu * L, &Ph)); POI_TRY(arena_get(e, (size_t)Bu * L * K, &Qs));
    POI_TRY(arena_get(e, (size_t)n_occ, &iota));
    POI_CAT(e, CAT_INDEX, 0, 0);
    POI_CK(e, cudaMemsetAsync(Qs, 0, (size_t)Bu * L * K * 4, e->stream));
    POI_LAUNCH(e, k_geoie_slots, (unsigned)poi_cdiv(n_occ, 256), 256, 0, seg_h.seg_of_occ, seg_g.seg_of_occ, Bu, L, K, Pg, Ph, Qs);
    POI_LAUNCH(e, k_mf_iota, (unsigned)poi_cdiv(n_occ, 256), 256, 0, iota, n_occ);
    GeoBatch gs = gg; gs.P = Pg; gs.Ph = Ph; gs.Q = Qs; gs.coords_g = xg; gs.coords_h = xh;
    SegList sh = seg_h, sg = seg_g; sh.uniq = iota; sg.uniq = iota;
    double* out_dev = nullptr;
    POI_TRY(arena_get(e, 8, &out_dev));
    POI_TRY(geoie_batch_core(e, cg, ch, cz, prm->ab, nullptr, H, gs, sh, sg, alpha, lambda, out_dev));
    POI_CK(e, cudaMemcpyAsync(pr->sums[me], out_dev, 3 * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
    // ---- outbox: deltas per unique row, grouped by owner ----
    {
        const float inv_alpha = 1.0f / alpha;
        PeerTable pt; memset(&pt, 0, sizeof(pt)); pt.world = W;
        POI_CAT(e, CAT_ROWS, 0, 0);
        unsigned grid = (unsigned)((int64_t)e->num_sms * 8);
        for (int r = 0; r < W; ++r) pt.shard[r] = pr->shard[0][r];
        POI_LAUNCH(e, k_mf_delta, grid, 256, 0, pt, H / 4, seg_g.uniq, seg_g.n_unique, cg, inv_alpha, pr->ob_grads[0][me]);
        for (int r = 0; r < W; ++r) pt.shard[r] = pr->shard[1][r];
        POI_LAUNCH(e, k_mf_delta, grid, 256, 0, pt, H / 4, seg_h.uniq, seg_h.n_unique, ch, inv_alpha, pr->ob_grads[1][me]);
        for (int r = 0; r < W; ++r) pt.shard[r] = pr->shard[2][r];
        POI_LAUNCH(e, k_mf_delta, grid, 256, 0, pt, H / 4, seg_h.uniq, seg_h.n_unique, cz, inv_alpha, pr->ob_grads[2][me]);
    }
    POI_TRY(mf_publish_set(e, pr, 0, seg_g, Bu));
    POI_TRY(mf_publish_set(e, pr, 1, seg_h, Bu));
    MgFlagPtrs fp; memset(&fp, 0, sizeof(fp));
    for (int r = 0; r < W; ++r) fp.p[r] = pr->flags[r];
    POI_CAT(e, CAT_OTHER, 0, 0);
    POI_LAUNCH(e, k_mg_signal, 1, 32, 0, fp, W, me, (uint32_t)step);
    POI_LAUNCH(e, k_mg_wait, 1, 32, 0, pr->flags[me], W, (uint32_t)step, err, timeout_ns, 200);
    // ---- replicated scalars, then the owner-side sparse step ----
    double* fin = nullptr;
    POI_TRY(arena_get(e, 8, &fin));
    {
        MfSumPtrs sp; memset(&sp, 0, sizeof(sp));
        for (int r = 0; r < W; ++r) sp.sums[r] = pr->sums[r];
        POI_CAT(e, CAT_REDUCE, 0, 0);
        POI_LAUNCH(e, k_mf_sums_apply, 1, 32, 0, sp, W, prm->ab, alpha, fin, err);
    }
    const int t_g[1] = {0}, t_hz[2] = {1, 2};
    POI_TRY(mf_apply_set(e, pr, 0, t_g, 1, H, alpha));
    POI_TRY(mf_apply_set(e, pr, 1, t_hz, 2, H, alpha));
    POI_CAT(e, CAT_OTHER, 0, 0);
    POI_LAUNCH(e, k_mg_signal, 1, 32, 0, fp, W, W + me, (uint32_t)step);
    POI_CK(e, cudaMemcpyAsync(e->h_out, fin, 8 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    POI_CK(e, cudaStreamSynchronize(e->stream));
    if (e->kprof) prof_harvest(e);
    const int ec = (int)e->h_out[7];
    if (ec >= 100 && ec < 300) POI_FAIL(e, "multi-GPU GeoIE step %lld: timed out waiting for rank %d (%s)", (long long)step, ec % 100,
                                        ec < 200 ? "has not applied the previous step" : "has not published its outbox");
    if (loss_host) *loss_host = e->h_out[0];
    return 0;
}
