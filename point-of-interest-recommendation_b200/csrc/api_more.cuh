// api_more.cuh -- remaining extern "C" entry points (Bpr mini-batch, GeoIE, score+top-K)
#pragma once

// per-occurrence loss gradients of the Bpr mini-batch (BPR.py:372-383), L2 handled by the row update
__global__ void __launch_bounds__(256)
k_bpr_batch_grads(const float* __restrict__ ux, const float* __restrict__ lt, int d4,
                  const int32_t* __restrict__ us, const int32_t* __restrict__ ps, const int32_t* __restrict__ qs,
                  const int32_t* __restrict__ mask, int64_t n, float* __restrict__ GU, float* __restrict__ GPQ,
                  double* __restrict__ part) {
    __shared__ double sh[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * 8 + w, nwarps = (int64_t)gridDim.x * 8;
    double acc = 0.0;
    for (int64_t i = warp; i < n; i += nwarps) {
        const float4* u4 = reinterpret_cast<const float4*>(ux) + (size_t)us[i] * d4;
        const float4* p4 = reinterpret_cast<const float4*>(lt) + (size_t)ps[i] * d4;
        const float4* q4 = reinterpret_cast<const float4*>(lt) + (size_t)qs[i] * d4;
        float dot = 0.f;
        for (int c = lane; c < d4; c += 32) {
            float4 u = u4[c], p = p4[c], q = q4[c];
            dot += u.x * (p.x - q.x) + u.y * (p.y - q.y) + u.z * (p.z - q.z) + u.w * (p.w - q.w);
        }
        dot = warp_sum(dot);
        const float m = (float)mask[i];
        const float g = sigmoidf_(-dot) * m;
        if (lane == 0) acc += (double)(logsigmoidf_(dot) * m);
        for (int c = lane; c < d4; c += 32) {
            float4 u = u4[c], p = p4[c], q = q4[c];
            reinterpret_cast<float4*>(GU)[(size_t)i * d4 + c] = make_float4(-g * (p.x - q.x), -g * (p.y - q.y), -g * (p.z - q.z), -g * (p.w - q.w));
            reinterpret_cast<float4*>(GPQ)[(size_t)i * d4 + c] = make_float4(-g * u.x, -g * u.y, -g * u.z, -g * u.w);
            reinterpret_cast<float4*>(GPQ)[(size_t)(n + i) * d4 + c] = make_float4(g * u.x, g * u.y, g * u.z, g * u.w);
        }
    }
    if (lane == 0) sh[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int i = 0; i < 8; ++i) t += sh[i]; part[blockIdx.x] = t; }
}

extern "C" {

int poi_bpr_train_batch(poi_engine* e, float* ux, int64_t n_user, float* lt, int64_t n_rows_lt, int32_t d,
                        const int32_t* p, const int32_t* q, const int32_t* mask, const int32_t* u, int64_t n,
                        float alpha, float lambda, double* loss_host) {
    POI_TRY(begin_call(e));
    if (d <= 0 || d % 4) POI_FAIL(e, "d must be a positive multiple of 4");
    if (n <= 0) { *loss_host = 0.0; return 0; }
    // device copies: [p ; q] contiguous (the key vector of the lt segments), u, mask
    int32_t *pq_dev = nullptr, *u_dev = nullptr, *m_dev = nullptr;
    POI_TRY(arena_get(e, (size_t)2 * n, &pq_dev));
    POI_TRY(arena_get(e, (size_t)n, &u_dev));
    POI_TRY(arena_get(e, (size_t)n, &m_dev));
    POI_TRY(stage_reserve(e, (size_t)n * 16 + 1024));
    int32_t* st = reinterpret_cast<int32_t*>(e->h_stage);
    memcpy(st, p, n * 4); memcpy(st + n, q, n * 4); memcpy(st + 2 * n, u, n * 4); memcpy(st + 3 * n, mask, n * 4);
    POI_CK(e, cudaMemcpyAsync(pq_dev, st, (size_t)2 * n * 4, cudaMemcpyHostToDevice, e->stream));
    POI_CK(e, cudaMemcpyAsync(u_dev, st + 2 * n, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
    POI_CK(e, cudaMemcpyAsync(m_dev, st + 3 * n, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
    float *GU = nullptr, *GPQ = nullptr; double *part = nullptr, *out_dev = nullptr;
    POI_TRY(arena_get(e, (size_t)n * d, &GU));
    POI_TRY(arena_get(e, (size_t)2 * n * d, &GPQ));
    int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(poi_cdiv(n, 8), (int64_t)e->num_sms * 8));
    POI_TRY(arena_get(e, (size_t)blocks, &part));
    POI_TRY(arena_get(e, 1, &out_dev));
    SegList seg_u, seg_pq;
    POI_TRY(build_segments(e, reinterpret_cast<const uint32_t*>(u_dev), n, (uint32_t)n_user, false, &seg_u));
    POI_TRY(build_segments(e, reinterpret_cast<const uint32_t*>(pq_dev), 2 * n, (uint32_t)n_rows_lt, false, &seg_pq));
    POI_CAT(e, CAT_MF, 0, (double)n * 3 * d * 4 * 2);
    POI_LAUNCH(e, k_bpr_batch_grads, blocks, 256, 0, ux, lt, d / 4, u_dev, pq_dev, pq_dev + n, m_dev, n, GU, GPQ, part);
    POI_LAUNCH(e, k_sum_partials_d, 1, 32, 0, part, blocks, out_dev, 1, 1);
    RowSrc src; memset(&src, 0, sizeof(src));
    src.mode = SRC_DENSE_GRADS; src.dim = d;
    src.grads = GU;
    POI_TRY(launch_rows_update(e, seg_u, ux, d, alpha, lambda, src, ROW_LONG_THRESH));
    src.grads = GPQ;
    POI_TRY(launch_rows_update(e, seg_pq, lt, d, alpha, lambda, src, ROW_LONG_THRESH));
    POI_CK(e, cudaMemcpyAsync(e->h_out, out_dev, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    POI_CK(e, cudaStreamSynchronize(e->stream));
    if (e->kprof) prof_harvest(e);
    *loss_host = -e->h_out[0];
    return 0;
}

int poi_geoie_train(poi_engine* e, const poi_geoie_params* prm, int32_t uidx, const int32_t* p_row,
                    const int32_t* q_row, int32_t lmax, const float* dist_pos, const float* dist_neg,
                    const int32_t* msk, int32_t n, float alpha, float lambda, double* out_host) {
    POI_TRY(begin_call(e));
    if (!prm || !prm->g || !prm->h || !prm->z || !prm->t || !prm->ab) POI_FAIL(e, "geoie params: null pointer");
    const int H = prm->H;
    if (H <= 0 || H % 4) POI_FAIL(e, "n_hidden must be a positive multiple of 4");
    if (n <= 0 || n + 1 > lmax) POI_FAIL(e, "bad n (%d) for lmax %d", n, lmax);
    (void)uidx;     // t[uidx] receives an exactly-zero gradient (t.z cancels): its "update" is a no-op
    const size_t nn = (size_t)n * n;
    const void* hs[5] = {p_row, q_row, dist_pos, dist_neg, msk};
    size_t bs[5] = {(size_t)lmax * 4, (size_t)lmax * 4, nn * 4, nn * 4, nn * 4};
    void* ds[5];
    // [p_row ; q_row] must be contiguous on the device: it is the key vector of the h/z segments
    int32_t* pq_dev = nullptr;
    POI_TRY(arena_get(e, (size_t)2 * lmax, &pq_dev));
    POI_TRY(upload_many(e, hs, bs, 5, ds));
    POI_CK(e, cudaMemcpyAsync(pq_dev, ds[0], (size_t)lmax * 4, cudaMemcpyDeviceToDevice, e->stream));
    POI_CK(e, cudaMemcpyAsync(pq_dev + lmax, ds[1], (size_t)lmax * 4, cudaMemcpyDeviceToDevice, e->stream));
    const int32_t* p_dev = pq_dev; const int32_t* q_dev = pq_dev + lmax;
    const float* dpos = (const float*)ds[2]; const float* dneg = (const float*)ds[3]; const int32_t* mk = (const int32_t*)ds[4];
    float *G, *Hp, *Hq, *GG, *GH, *GZ; double *CWp, *CWq, *per_i, *out_dev;
    POI_TRY(arena_get(e, (size_t)n * H, &G)); POI_TRY(arena_get(e, (size_t)n * H, &Hp)); POI_TRY(arena_get(e, (size_t)n * H, &Hq));
    POI_TRY(arena_get(e, (size_t)lmax * H, &GG));
    POI_TRY(arena_get(e, (size_t)2 * lmax * H, &GH)); POI_TRY(arena_get(e, (size_t)2 * lmax * H, &GZ));
    POI_TRY(arena_get(e, nn, &CWp)); POI_TRY(arena_get(e, nn, &CWq));
    POI_TRY(arena_get(e, (size_t)n * 4, &per_i)); POI_TRY(arena_get(e, 1, &out_dev));
    POI_CK(e, cudaMemsetAsync(GG, 0, (size_t)lmax * H * 4, e->stream));
    POI_CK(e, cudaMemsetAsync(GH, 0, (size_t)2 * lmax * H * 4, e->stream));
    POI_CK(e, cudaMemsetAsync(GZ, 0, (size_t)2 * lmax * H * 4, e->stream));
    SegList seg_p, seg_pq;
    POI_TRY(build_segments(e, reinterpret_cast<const uint32_t*>(p_dev), lmax, (uint32_t)prm->n_rows, false, &seg_p));
    POI_TRY(build_segments(e, reinterpret_cast<const uint32_t*>(pq_dev), 2 * lmax, (uint32_t)prm->n_rows, false, &seg_pq));
    POI_CAT(e, CAT_GEOIE, 0, 0);
    unsigned gnh = (unsigned)poi_cdiv((int64_t)n * H, 256);
    POI_LAUNCH(e, k_geoie_gather, gnh, 256, 0, prm->g, prm->h, p_dev, q_dev, n, H, G, Hp, Hq);
    POI_LAUNCH(e, k_geoie_fwd, n, 128, (size_t)(2 * H + 20) * sizeof(double), G, Hp, Hq, n, H, dpos, dneg, mk, prm->ab, CWp, CWq, per_i);
    POI_LAUNCH(e, k_geoie_bwd_h, n, 128, 0, G, Hp, Hq, n, H, CWp, CWq, lambda, lmax, GH);
    POI_LAUNCH(e, k_geoie_bwd_g, n, 128, 0, G, Hp, Hq, n, H, CWp, CWq, lambda, GG);
    POI_LAUNCH(e, k_geoie_zgrad, gnh, 256, 0, prm->z, p_dev, q_dev, n, H, lmax, lambda, GZ);
    POI_LAUNCH(e, k_geoie_finalize, 1, 32, 0, per_i, n, prm->ab, alpha, out_dev);
    // sparse updates: g[unique(p_full)], h/z[unique(p_full u q_full)] (GeoIE.py:147-153,174-181); the L2
    // terms are already inside the occurrence gradients (they cover only the n gathered rows), so lambda = 0
    RowSrc src; memset(&src, 0, sizeof(src));
    src.mode = SRC_DENSE_GRADS; src.dim = H;
    src.grads = GG; POI_TRY(launch_rows_update(e, seg_p, prm->g, H, alpha, 0.f, src, ROW_LONG_THRESH));
    src.grads = GH; POI_TRY(launch_rows_update(e, seg_pq, prm->h, H, alpha, 0.f, src, ROW_LONG_THRESH));
    src.grads = GZ; POI_TRY(launch_rows_update(e, seg_pq, prm->z, H, alpha, 0.f, src, ROW_LONG_THRESH));
    POI_CK(e, cudaMemcpyAsync(e->h_out, out_dev, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    POI_CK(e, cudaStreamSynchronize(e->stream));
    if (e->kprof) prof_harvest(e);
    *out_host = e->h_out[0];
    return 0;
}

// theta <- theta - alpha (g + lambda theta) over a flat region
__global__ void k_dense_apply(float* __restrict__ theta, const float* __restrict__ g, int64_t n, float alpha, float lambda) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float t = theta[i]; theta[i] = t - alpha * (g[i] + lambda * t); }
}
// di[k,:] <- di[k,:] - alpha (G[k,:] + lambda cnt[k] di[k,:])
__global__ void k_di_apply(float* __restrict__ di, const float* __restrict__ g, const float* __restrict__ cnt,
                           int nD, int d, float alpha, float lambda) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)nD * d) { float t = di[i]; di[i] = t - alpha * (g[i] + lambda * cnt[i / d] * t); }
}

int poi_gru_mg_dense_size(const poi_gru_params* p, int64_t* n_floats) {
    const bool head = p->di != nullptr;
    *n_floats = mg_layout(p->H, head ? 2 * p->d : p->d, head ? p->n_rows_di : 0, p->d).total;
    return 0;
}

struct MgPrep { GruIdx ix; SegList seg_lt, seg_di; };

__global__ void k_copy_u32_to_i32(const uint32_t* __restrict__ src, const uint32_t* __restrict__ n, int32_t* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *n) dst[i] = (int32_t)src[i];
}

int poi_gru_mg_prepare(poi_engine* e, const poi_gru_params* p, const poi_seq_index* index, const int32_t* uidx_host,
                       int32_t B, int32_t* uniq_out_dev, int64_t* n_unique_host) {
    e->prep_valid = false;
    POI_TRY(begin_call(e));
    const bool head = p->di != nullptr;
    if (!index || !index->p || !index->q || !index->lens || (head && (!index->dp || !index->dq))) POI_FAIL(e, "index matrices missing");
    if (B <= 0) POI_FAIL(e, "empty batch");
    if (!e->prep_state) e->prep_state = new MgPrep();
    MgPrep* st = static_cast<MgPrep*>(e->prep_state);
    POI_TRY(stage_reserve(e, (size_t)B * 4 + 256));
    size_t so = 0;
    int32_t* uidx_dev = nullptr;
    POI_TRY(gru_upload_i32(e, uidx_host, (size_t)B, &uidx_dev, &so));
    POI_TRY(gru_alloc_idx(e, B, index->lmax, head, &st->ix));
    POI_CAT(e, CAT_INDEX, 0, 0);
    POI_LAUNCH(e, k_slice_indices, (unsigned)poi_cdiv((int64_t)B * index->lmax, 256), 256, 0, index->p, index->q,
               head ? index->dp : nullptr, head ? index->dq : nullptr, index->lens, index->lmax, uidx_dev, B,
               st->ix.PQt, st->ix.DPt, st->ix.DQt, st->ix.lensB);
    const int64_t LB = (int64_t)index->lmax * B;
    POI_TRY(build_segments(e, reinterpret_cast<const uint32_t*>(st->ix.PQt), 2 * LB, (uint32_t)p->n_rows_lt, true, &st->seg_lt));
    if (head) POI_TRY(build_segments(e, reinterpret_cast<const uint32_t*>(st->ix.DPt), LB, (uint32_t)p->n_rows_di, false, &st->seg_di));
    POI_LAUNCH(e, k_copy_u32_to_i32, (unsigned)poi_cdiv(2 * LB, 256), 256, 0, st->seg_lt.uniq, st->seg_lt.n_unique, uniq_out_dev);
    uint32_t nu = 0;
    POI_CK(e, cudaMemcpyAsync(&nu, st->seg_lt.n_unique, 4, cudaMemcpyDeviceToHost, e->stream));
    POI_CK(e, cudaStreamSynchronize(e->stream));
    *n_unique_host = nu;
    e->prep_valid = true; e->prep_B = B; e->prep_lmax = index->lmax;
    return 0;
}

int poi_gru_train_mg(poi_engine* e, const poi_gru_params* p, const poi_seq_index* index, const int32_t* uidx_host,
                     int32_t B, int32_t max_len, int32_t global_batch, const float* rows_dev, int64_t n_unique,
                     float* dense_grads, float* row_grads, float* row_cnt, double* loss_sums) {
    if (e->prep_valid && e->prep_B == B && index && e->prep_lmax == index->lmax) {
        // continue in the arena epoch of poi_gru_mg_prepare: indices are sliced and sorted already
        e->prep_valid = false;
        POI_CK(e, cudaSetDevice(e->device));
        if (!rows_dev || !dense_grads || !row_grads || !row_cnt || !loss_sums) POI_FAIL(e, "null exchange buffer");
        MgPrep* st = static_cast<MgPrep*>(e->prep_state);
        MgCtx mg; mg.rows = rows_dev; mg.global_batch = global_batch; mg.dense_grads = dense_grads;
        mg.row_grads = row_grads; mg.row_cnt = row_cnt; mg.loss_sums = loss_sums;
        PreSeg pre{&st->seg_lt, &st->seg_di};
        phase_mark(e, 0);
        return gru_train_core(e, p, st->ix, B, index->lmax, max_len, 0, 0.f, 0.f, nullptr, &mg, &pre);
    }
    POI_TRY(begin_call(e));
    if (!p || !p->ui || !p->wh || !p->bi) POI_FAIL(e, "gru params: null pointer");
    if (p->d <= 0 || p->d % 4 || p->H != p->d) POI_FAIL(e, "n_in must equal n_hidden and be a multiple of 4");
    const bool head = p->di != nullptr;
    if (!index || !index->p || !index->q || !index->lens || (head && (!index->dp || !index->dq))) POI_FAIL(e, "index matrices missing");
    if (B <= 0 || global_batch < B) POI_FAIL(e, "bad batch sizes");
    if (!rows_dev || !dense_grads || !row_grads || !row_cnt || !loss_sums) POI_FAIL(e, "null exchange buffer");
    (void)n_unique;
    phase_mark(e, 0);
    POI_TRY(stage_reserve(e, (size_t)B * 4 + 256));
    size_t so = 0;
    int32_t* uidx_dev = nullptr;
    POI_TRY(gru_upload_i32(e, uidx_host, (size_t)B, &uidx_dev, &so));
    GruIdx ix;
    POI_TRY(gru_alloc_idx(e, B, index->lmax, head, &ix));
    POI_CAT(e, CAT_INDEX, 0, 0);
    POI_LAUNCH(e, k_slice_indices, (unsigned)poi_cdiv((int64_t)B * index->lmax, 256), 256, 0, index->p, index->q,
               head ? index->dp : nullptr, head ? index->dq : nullptr, index->lens, index->lmax, uidx_dev, B,
               ix.PQt, ix.DPt, ix.DQt, ix.lensB);
    MgCtx mg; mg.rows = rows_dev; mg.global_batch = global_batch; mg.dense_grads = dense_grads;
    mg.row_grads = row_grads; mg.row_cnt = row_cnt; mg.loss_sums = loss_sums;
    return gru_train_core(e, p, ix, B, index->lmax, max_len, 0, 0.f, 0.f, nullptr, &mg);
}

int poi_gru_apply_mg(poi_engine* e, const poi_gru_params* p, const float* dg, const double* loss_sums,
                     int32_t global_batch, int64_t n_nonempty_global, float* lt_local, int64_t n_local_rows,
                     const int32_t* recv_ids, const float* recv_grads, const float* recv_cnts, int64_t n_recv,
                     float alpha, float lambda, double* out_host) {
    POI_TRY(begin_call(e));
    const bool head = p->di != nullptr;
    const int d = p->d, H = p->H, din = head ? 2 * d : d, nD = head ? p->n_rows_di : 0;
    const MgLayout ML = mg_layout(H, din, nD, d);
    POI_CAT(e, CAT_WGRAD, 0, 0);
    auto apply = [&](float* theta, int64_t off, int64_t n) -> int {
        if (n <= 0) return 0;
        POI_LAUNCH(e, k_dense_apply, (unsigned)poi_cdiv(n, 256), 256, 0, theta, dg + off, n, alpha, lambda);
        return 0;
    };
    POI_TRY(apply(p->ui, ML.ui, (int64_t)3 * H * din));
    POI_TRY(apply(p->wh, ML.wh, (int64_t)3 * H * H));
    POI_TRY(apply(p->bi, ML.bi, 3 * H));
    if (head) {
        POI_TRY(apply(p->vs, ML.vs, (int64_t)nD * H));
        POI_TRY(apply(p->bs, ML.bs, nD));
        POI_LAUNCH(e, k_di_apply, (unsigned)poi_cdiv((int64_t)nD * d, 256), 256, 0, p->di, dg + ML.di, dg + ML.dicnt, nD, d, alpha, lambda);
    }
    double* out_dev = nullptr;
    POI_TRY(arena_get(e, 8, &out_dev));
    POI_CAT(e, CAT_REDUCE, 0, 0);
    POI_LAUNCH(e, k_finalize_from_sums, 1, 32, 0, loss_sums, p->scal, head ? 1 : 0,
               (double)n_nonempty_global * 0.6931471805599453, 1.0 / (double)global_batch, alpha, lambda, out_dev);
    if (n_recv > 0) {
        SegList seg;
        POI_TRY(build_segments(e, reinterpret_cast<const uint32_t*>(recv_ids), n_recv, (uint32_t)n_local_rows, false, &seg));
        RowSrc src; memset(&src, 0, sizeof(src));
        src.mode = SRC_DENSE_GRADS; src.grads = recv_grads; src.dim = d; src.weights = recv_cnts;
        POI_TRY(launch_rows_update(e, seg, lt_local, d, alpha, lambda, src, ROW_LONG_THRESH, (double)n_recv * d * 4 * 3));
    }
    POI_CK(e, cudaMemcpyAsync(e->h_out, out_dev, 8 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    POI_CK(e, cudaStreamSynchronize(e->stream));
    if (e->kprof) prof_harvest(e);
    for (int i = 0; i < 5; ++i) out_host[i] = e->h_out[i];
    return 0;
}

int poi_gemm_tn(poi_engine* e, const float* A, int lda, const float* W, int ldw, int64_t M, int N, int K,
                const float* bias, float* C, int ldc, int mode) {
    POI_TRY(begin_call(e));
    if (mode < 0 || mode > 2) POI_FAIL(e, "bad gemm mode");
    if ((K % 4) || (lda % 4) || (ldw % 4) || (ldc % 4)) POI_FAIL(e, "K, lda, ldw, ldc must be multiples of 4");
    int saved = e->gemm_mode;
    e->gemm_mode = mode;
    int rc = gemm_tn(e, A, lda, W, ldw, M, N, K, EpiBiasStore{C, ldc, bias, N});
    e->gemm_mode = saved;
    if (rc) return rc;
    POI_CK(e, cudaStreamSynchronize(e->stream));
    if (e->kprof) prof_harvest(e);
    return 0;
}

__global__ void k_reduce_plain(const float* __restrict__ part, int splits, int64_t n, float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = 0.f;
    for (int s = 0; s < splits; ++s) g += part[(size_t)s * n + i];
    out[i] = g;
}

int poi_gemm_atb(poi_engine* e, const float* A, int lda, const float* B, int ldb, int64_t M, int N1, int N2,
                 float* C, int mode) {
    POI_TRY(begin_call(e));
    if ((N1 % 4) || (N2 % 4) || (lda % 4) || (ldb % 4)) POI_FAIL(e, "N1, N2, lda, ldb must be multiples of 4");
    AtbPlan plan;
    if (mode == 0) POI_TRY(launch_gemm_atb(e, A, lda, B, ldb, M, N1, N2, &plan));
    else POI_TRY(launch_gemm_atb_tc_mn(e, A, lda, B, ldb, M, N1, N2, mode == 1, &plan));
    const int64_t n = (int64_t)N1 * N2;
    POI_LAUNCH(e, k_reduce_plain, (unsigned)poi_cdiv(n, 256), 256, 0, plan.part, plan.splits, n, C);
    POI_CK(e, cudaStreamSynchronize(e->stream));
    if (e->kprof) prof_harvest(e);
    return 0;
}

static int score_topk_body(poi_engine* e, const float* users, int32_t B, const float* items, int64_t n_item, int32_t H,
                           const float* prob, const float* sts, int32_t nD, const double* ucoord, const double* icoord,
                           double dd, int32_t dist_num, float wd, int32_t top_k, int32_t* topk_dev) {
    POI_TRY(begin_call(e));
    if (B <= 0 || n_item <= 0) return 0;
    if (H % 4) POI_FAIL(e, "H must be a multiple of 4");
    if (top_k <= 0 || top_k > 256 || top_k > n_item) POI_FAIL(e, "top_k out of range");
    const int chunk = (int)std::min<int64_t>(n_item, 8192);
    const int lds = (chunk + 3) / 4 * 4;
    float *S = nullptr, *best_val = nullptr;
    POI_TRY(arena_get(e, (size_t)B * lds, &S));
    POI_TRY(arena_get(e, (size_t)B * top_k, &best_val));
    size_t smem = (size_t)(chunk + top_k) * 4 + (size_t)top_k * 4;
    if (smem > 48 * 1024)
        POI_CK(e, cudaFuncSetAttribute(k_topk_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int64_t i0 = 0; i0 < n_item; i0 += chunk) {
        int cur = (int)std::min<int64_t>(chunk, n_item - i0);
        // users . items^T on the tensor cores (3xTF32, fp32-faithful) when the shape allows, scores finished in the epilogue
        if (sts) POI_TRY(gemm_tn(e, users, H, items + (size_t)i0 * H, H, B, cur, H,
                                 EpiScoreGeo{S, lds, sts, nD, ucoord, icoord, dd, dist_num, i0, wd, cur}));
        else POI_TRY(gemm_tn(e, users, H, items + (size_t)i0 * H, H, B, cur, H, EpiScore{S, lds, prob, n_item, i0, wd, cur}));
        POI_CAT(e, CAT_EVAL, 0, 0);
        POI_LAUNCH(e, k_topk_merge, B, 256, smem, S, lds, cur, i0, top_k, best_val, topk_dev, i0 == 0 ? 1 : 0);
    }
    POI_CK(e, cudaStreamSynchronize(e->stream));
    if (e->kprof) prof_harvest(e);
    return 0;
}

int poi_score_topk(poi_engine* e, const float* users, int32_t B, const float* items, int64_t n_item, int32_t H,
                   const float* prob, float wd, int32_t top_k, int32_t* topk_dev) {
    return score_topk_body(e, users, B, items, n_item, H, prob, nullptr, 0, nullptr, nullptr, 1.0, 0, wd, top_k, topk_dev);
}

int poi_score_topk_geo(poi_engine* e, const float* users, int32_t B, const float* items, int64_t n_item, int32_t H,
                       const float* sts, int32_t n_dist_rows, const double* user_coords, const double* item_coords, double dd,
                       int32_t dist_num, float wd, int32_t top_k, int32_t* topk_dev) {
    if (!sts || !user_coords || !item_coords || !(dd > 0) || n_dist_rows <= dist_num - 1) POI_FAIL(e, "poi_score_topk_geo: bad arguments");
    return score_topk_body(e, users, B, items, n_item, H, nullptr, sts, n_dist_rows, user_coords, item_coords, dd, dist_num, wd, top_k, topk_dev);
}

}  // extern "C"

// ---- NVLink peer memory (peer.cuh) ---------------------------------------------------------------
extern "C" int poi_peer_alloc(poi_engine* e, int64_t bytes, void** ptr_out, unsigned char* handle_out64) {
    POI_CK(e, cudaSetDevice(e->device));
    if (bytes <= 0 || !ptr_out || !handle_out64) POI_FAIL(e, "poi_peer_alloc: bad arguments");
    void* p = nullptr;
    POI_CK(e, cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t st = cudaIpcGetMemHandle(&h, p);
    if (st != cudaSuccess) { cudaFree(p); POI_FAIL(e, "cudaIpcGetMemHandle: %s", cudaGetErrorString(st)); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle_out64, &h, 64);
    *ptr_out = p;
    return 0;
}
extern "C" int poi_peer_free(poi_engine* e, void* ptr) {
    POI_CK(e, cudaSetDevice(e->device));
    POI_CK(e, cudaFree(ptr));
    return 0;
}
extern "C" int poi_peer_open(poi_engine* e, const unsigned char* handle64, void** ptr_out) {
    POI_CK(e, cudaSetDevice(e->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    POI_CK(e, cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int poi_peer_close(poi_engine* e, void* ptr) {
    POI_CK(e, cudaSetDevice(e->device));
    POI_CK(e, cudaIpcCloseMemHandle(ptr));
    return 0;
}

extern "C" int poi_gather_rows_sharded(poi_engine* e, const float* const* shards_host, int world, int dim,
                                       const int32_t* ids_dev, int64_t n_idx, float* out_dev) {
    POI_TRY(begin_call(e));
    if (world < 1 || world > POI_MAX_PEERS) POI_FAIL(e, "world must be in [1, %d]", POI_MAX_PEERS);
    if (dim <= 0 || dim % 4) POI_FAIL(e, "dim must be a positive multiple of 4");
    if (n_idx <= 0) return 0;
    PeerTable pt; memset(&pt, 0, sizeof(pt));
    pt.world = world;
    for (int r = 0; r < world; ++r) pt.shard[r] = shards_host[r];
    const int dim4 = dim / 4;
    // bytes: every row read once (world-1 of world over NVLink) and written once, plus the ids
    POI_CAT(e, CAT_GATHER, 0, 2.0 * (double)n_idx * dim * 4 + 4.0 * (double)n_idx);
    if (peer_gather_mode() == 1 && dim * 4 * PG_WARPS * PG_ST <= (200 << 10))
        return launch_gather_bulk(e, pt, dim, reinterpret_cast<const uint32_t*>(ids_dev), nullptr, n_idx, n_idx, out_dev);
    const int lpr = dim4 <= 8 ? 8 : (dim4 <= 16 ? 16 : 32);
    const int64_t threads_needed = poi_cdiv(n_idx, 4) * lpr;
    unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(poi_cdiv(threads_needed, 256), (int64_t)e->num_sms * 16));
    if (lpr == 8)       POI_LAUNCH(e, (k_gather_rows_sharded<8, 4>), grid, 256, 0, pt, dim4, ids_dev, n_idx, out_dev);
    else if (lpr == 16) POI_LAUNCH(e, (k_gather_rows_sharded<16, 4>), grid, 256, 0, pt, dim4, ids_dev, n_idx, out_dev);
    else                POI_LAUNCH(e, (k_gather_rows_sharded<32, 4>), grid, 256, 0, pt, dim4, ids_dev, n_idx, out_dev);
    return 0;
}

extern "C" int poi_pull_segments(poi_engine* e, int world, int dim, const int32_t* const* perm_host,
                                 const int32_t* const* ids_host, const float* const* grads_host,
                                 const float* const* cnts_host, const int64_t* src_off_host, const int64_t* n_host,
                                 int32_t* recv_local_ids_dev, float* recv_grads_dev, float* recv_cnts_dev) {
    POI_TRY(begin_call(e));
    if (world < 1 || world > POI_MAX_PEERS) POI_FAIL(e, "world must be in [1, %d]", POI_MAX_PEERS);
    if (dim <= 0 || dim % 4) POI_FAIL(e, "dim must be a positive multiple of 4");
    PullTable pt; memset(&pt, 0, sizeof(pt));
    pt.world = world;
    int64_t tot = 0;
    for (int r = 0; r < world; ++r) {
        pt.perm[r] = perm_host[r]; pt.ids[r] = ids_host[r]; pt.grads[r] = grads_host[r]; pt.cnts[r] = cnts_host[r];
        pt.src_off[r] = src_off_host[r]; pt.dst_off[r] = tot;
        if (n_host[r] < 0) POI_FAIL(e, "negative segment size");
        tot += n_host[r];
    }
    pt.dst_off[world] = tot;
    if (tot == 0) return 0;
    POI_CAT(e, CAT_ROWS, 0, 2.0 * (double)tot * dim * 4 + 20.0 * (double)tot);
    unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(poi_cdiv(tot * 32, 256), (int64_t)e->num_sms * 16));
    POI_LAUNCH(e, k_pull_segments, grid, 256, 0, pt, dim / 4, recv_local_ids_dev, recv_grads_dev, recv_cnts_dev);
    return 0;
}


extern "C" int poi_group_by_owner(poi_engine* e, const int32_t* ids_dev, int64_t n, int world, int32_t* perm_out_dev,
                                  double* counts_out_dev) {
    POI_TRY(begin_call(e));
    if (world < 1 || world > POI_MAX_PEERS) POI_FAIL(e, "world must be in [1, %d]", POI_MAX_PEERS);
    if (n < 0) POI_FAIL(e, "negative count");
    POI_CAT(e, CAT_INDEX, 0, 0);
    uint32_t* keys = nullptr;
    POI_TRY(arena_get(e, (size_t)std::max<int64_t>(n, 1), &keys));
    if (n > 0) POI_LAUNCH(e, k_owner_keys, (unsigned)poi_cdiv(n, 256), 256, 0, ids_dev, n, world, keys);
    uint32_t *ks = nullptr, *vs = nullptr;
    POI_TRY(sort_pairs(e, keys, n, (uint32_t)world, &ks, &vs));
    POI_LAUNCH(e, k_owner_perm_counts, (unsigned)std::max<int64_t>(1, poi_cdiv(n, 256)), 256, 0, ks, vs, n, world, perm_out_dev, counts_out_dev);
    return 0;
}

// ---- per-epoch negative sampling and negative-distance binning on the device (sampling.cuh) ----
extern "C" int poi_sample_negatives(poi_engine* e, const int32_t* rows_dev, int32_t lrow, const int32_t* sorted_a_dev, int32_t la,
                                    const int32_t* sorted_b_dev, int32_t lb, int32_t n_user, int32_t n_item, uint64_t seed,
                                    uint32_t epoch, int32_t* out_dev) {
    POI_TRY(begin_call(e));
    if (!rows_dev || !sorted_a_dev || !out_dev) POI_FAIL(e, "poi_sample_negatives: null pointer");
    if (n_user <= 0 || lrow <= 0 || la <= 0 || n_item <= 0) POI_FAIL(e, "poi_sample_negatives: bad sizes");
    if ((int64_t)la + (sorted_b_dev ? lb : 0) >= n_item) POI_FAIL(e, "a user's rows could cover the whole catalogue: the rejection loop may not end");
    POI_CAT(e, CAT_INDEX, 0, 0);
    const int64_t n = (int64_t)n_user * lrow;
    POI_LAUNCH(e, k_sample_negatives, (unsigned)poi_cdiv(n, 256), 256, 0, rows_dev, lrow, sorted_a_dev, la, sorted_b_dev, lb,
               n_user, n_item, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), epoch, out_dev);
    return 0;
}
extern "C" int poi_neg_intervals(poi_engine* e, const int32_t* p_dev, const int32_t* q_dev, const int32_t* lens_dev, int32_t n_user,
                                 int32_t lmax, const double* coords_dev, double dd, int32_t dist_num, int32_t* out_dev) {
    POI_TRY(begin_call(e));
    if (!p_dev || !q_dev || !lens_dev || !coords_dev || !out_dev) POI_FAIL(e, "poi_neg_intervals: null pointer");
    if (n_user <= 0 || lmax <= 0 || !(dd > 0)) POI_FAIL(e, "poi_neg_intervals: bad sizes");
    POI_CAT(e, CAT_INDEX, 0, 0);
    const int64_t n = (int64_t)n_user * lmax;
    POI_LAUNCH(e, k_neg_intervals, (unsigned)poi_cdiv(n, 256), 256, 0, p_dev, q_dev, lens_dev, n_user, lmax, coords_dev, dd, dist_num, out_dev);
    return 0;
}

#ifdef POI_FUSED_TRACE
// debugging build only (tools/fused_trace.py): clock stamps of CTA 0 of the fused recurrence kernels
extern "C" int poi_debug_fused_trace(int dir, long long* out, int n, int clear) {
    if (clear) { static long long zero[512 * 16]; return (int)cudaMemcpyToSymbol(fused::g_trace, zero, sizeof(zero), (size_t)dir * sizeof(zero)); }
    return (int)cudaMemcpyFromSymbol(out, fused::g_trace, (size_t)n * sizeof(long long), (size_t)dir * 512 * 16 * sizeof(long long));
}
#endif

// ---- K-negative PRME (prme_k.cuh) ------------------------------------------------------------------
extern "C" int poi_prme_train_seq_k(poi_engine* e, float* du, float* dp, float* ds_, int32_t d, const int32_t* u,
                                    const int32_t* p, const int32_t* q, const int32_t* prev, const double* dist,
                                    const int32_t* gap, int64_t n, int32_t K, int32_t threshold, double cw, float alpha,
                                    float lambda, double* loss_host) {
    POI_TRY(begin_call(e));
    if (d <= 0 || d % 4) POI_FAIL(e, "d must be a positive multiple of 4");
    if (K < 1 || K > PRME_MAXK) POI_FAIL(e, "K must be in [1, %d]", PRME_MAXK);
    if (n <= 0) return 0;
    const size_t smem = prme_seq_k_smem(d, K);
    if (smem > 227 * 1024) POI_FAIL(e, "K * d too large for the sequential kernel (%zu bytes of shared memory)", smem);
    const void* hs[6] = {u, p, q, prev, dist, gap};
    size_t bs[6] = {(size_t)n * 4, (size_t)n * 4, (size_t)n * K * 4, (size_t)n * 4, (size_t)n * 8, (size_t)n * 4};
    void* dv[6];
    POI_TRY(upload_many(e, hs, bs, 6, dv));
    double* loss_dev = nullptr;
    POI_TRY(arena_get(e, (size_t)n, &loss_dev));
    POI_CK(e, cudaFuncSetAttribute(k_prme_seq_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    POI_CAT(e, CAT_MF, 0, (double)n * (2 * (K + 2) + 1) * d * 4 * 2);
    POI_LAUNCH(e, k_prme_seq_k, 1, 256, smem, du, dp, ds_, (int)d, (int)K, (const int32_t*)dv[0], (const int32_t*)dv[1],
               (const int32_t*)dv[2], (const int32_t*)dv[3], (const double*)dv[4], (const int32_t*)dv[5], n, (int)threshold,
               (float)cw, alpha, lambda, loss_dev);
    POI_CK(e, cudaMemcpyAsync(loss_host, loss_dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    POI_CK(e, cudaStreamSynchronize(e->stream));
    if (e->kprof) prof_harvest(e);
    return 0;
}

extern "C" int poi_prme_train_batch_k(poi_engine* e, float* du, int64_t n_user, float* dp, float* ds_, int64_t n_rows, int32_t d,
                                      const int32_t* u, const int32_t* p, const int32_t* q, const int32_t* prev,
                                      const float* dist, const int32_t* gap, int64_t n, int32_t K, int32_t on_host,
                                      int32_t threshold, double cw, float alpha, float lambda, double* loss_sum_host) {
    POI_TRY(begin_call(e));
    if (d <= 0 || d % 4 || d > 1024) POI_FAIL(e, "d must be a multiple of 4, <= 1024");
    if (K < 1 || K > PRME_MAXK) POI_FAIL(e, "K must be in [1, %d]", PRME_MAXK);
    if (n <= 0) { if (loss_sum_host) *loss_sum_host = 0.0; return 0; }
    if (n * (K + 2) >= (int64_t)1 << 31) POI_FAIL(e, "batch too large");
    PrmeBatchIdx b;
    b.N = (int)n; b.K = K;
    if (on_host) {       // end-to-end path: the step's index arrays come from host memory inside the call
        const void* hs[6] = {u, p, q, prev, dist, gap};
        size_t bs[6] = {(size_t)n * 4, (size_t)n * 4, (size_t)n * K * 4, (size_t)n * 4, (size_t)n * 4, (size_t)n * 4};
        void* dv[6];
        POI_TRY(upload_many(e, hs, bs, 6, dv));
        b.u = (const int32_t*)dv[0]; b.p = (const int32_t*)dv[1]; b.q = (const int32_t*)dv[2]; b.prev = (const int32_t*)dv[3];
        b.dist = (const float*)dv[4]; b.gap = (const int32_t*)dv[5];
    } else { b.u = u; b.p = p; b.q = q; b.prev = prev; b.dist = dist; b.gap = gap; }
    const int R = K + 2, d4 = d / 4;
    const int64_t n_occ = n * R;
    uint32_t* keys = nullptr;
    POI_TRY(arena_get(e, (size_t)n_occ, &keys));
    POI_CAT(e, CAT_INDEX, 0, 0);
    POI_LAUNCH(e, k_prme_keys, (unsigned)poi_cdiv(n_occ, 256), 256, 0, b, keys);
    SegList seg, seg_u;
    POI_TRY(build_segments(e, keys, n_occ, (uint32_t)n_rows, false, &seg));
    POI_TRY(build_segments(e, reinterpret_cast<const uint32_t*>(b.u), n, (uint32_t)n_user, false, &seg_u));
    float *KP = nullptr, *KS = nullptr, *SL = nullptr, *GU = nullptr, *GL = nullptr; double *part = nullptr, *out_dev = nullptr;
    POI_TRY(arena_get(e, (size_t)n_occ, &KP));
    POI_TRY(arena_get(e, (size_t)n_occ, &KS));
    POI_TRY(arena_get(e, (size_t)n * d, &SL));
    POI_TRY(arena_get(e, (size_t)n * d, &GU));
    POI_TRY(arena_get(e, (size_t)n * d, &GL));
    // one warp per check-in, single pass (k_prme_score_warp); d > 512: the CTA-per-check-in kernel (rows do not fit the registers).
    // POI_PRME_SCORE=cta forces the latter (A/B).  Measured on c3: 0.172 ms against 0.335 ms for the CTA kernels (register
    // version and a TMA-staged version, removed), which spend their time in the four block barriers of a check-in.
    const char* score_env = getenv("POI_PRME_SCORE");
    const bool use_warp = (!score_env || score_env[0] == 'w') && d4 <= 128;
    const int blocks = use_warp ? (int)std::min<int64_t>(poi_cdiv(n, 8), (int64_t)e->num_sms * 2)
                                : (int)std::min<int64_t>(n, (int64_t)e->num_sms * 8);
    POI_TRY(arena_get(e, (size_t)blocks, &part));
    POI_TRY(arena_get(e, 1, &out_dev));
    const size_t smem = (size_t)8 * 2 * d4 * sizeof(float4);
    // algorithmic bytes: every gathered row read once and written once + the index words.  Phase A is
    // booked with the reads (category "mf"), phase B with the write-back (category "rows").
    const double algo = (double)n * ((2.0 * (2 * R + 1)) * d * 4 + 4.0 * (K + 5));
    int64_t awarps = std::min<int64_t>(n_occ, (int64_t)e->num_sms * 64);
    unsigned agrid = (unsigned)std::max<int64_t>(poi_cdiv(awarps * 32, 256), 1);
    POI_CAT(e, CAT_MF, 0, 0.5 * algo);
#define PRME_BK(NCH)                                                                                                        \
    do {                                                                                                                    \
        if (use_warp) {                                                                                                     \
            POI_LAUNCH(e, (k_prme_score_warp<(NCH <= 4 ? NCH : 4)>), blocks, 256, 0, du, dp, ds_, d4, b, (int)threshold, (float)cw, KP, KS,  \
                       SL, GU, GL, part);                                                                                   \
        } else {                                                                                                            \
            POI_CK(e, cudaFuncSetAttribute(k_prme_score<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
            POI_LAUNCH(e, (k_prme_score<NCH>), blocks, 256, smem, du, dp, ds_, d4, b, (int)threshold, (float)cw, KP, KS,    \
                       SL, GU, GL, part);                                                                                   \
        }                                                                                                                   \
        POI_CAT(e, CAT_ROWS, 0, 0.5 * algo);                                                                                \
        POI_LAUNCH(e, (k_prme_apply<NCH>), agrid, 256, 0, seg, du, dp, ds_, d4, b, KP, KS, SL, GL, alpha, lambda);          \
    } while (0)
    if (d4 <= 32) PRME_BK(1); else if (d4 <= 64) PRME_BK(2); else if (d4 <= 128) PRME_BK(4); else PRME_BK(8);
#undef PRME_BK
    POI_CAT(e, CAT_REDUCE, 0, 0);
    POI_LAUNCH(e, k_sum_partials_warp, 1, 32, 0, part, blocks, out_dev);
    // du[u]: one step per unique user, the check-ins of a user summed in fixed order (rows.cuh)
    RowSrc src; memset(&src, 0, sizeof(src));
    src.mode = SRC_DENSE_GRADS; src.dim = d; src.grads = GU;
    POI_TRY(launch_rows_update(e, seg_u, du, d, alpha, lambda, src, ROW_LONG_THRESH));
    POI_CK(e, cudaMemcpyAsync(e->h_out, out_dev, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    POI_CK(e, cudaStreamSynchronize(e->stream));
    if (e->kprof) prof_harvest(e);
    if (loss_sum_host) *loss_sum_host = e->h_out[0];
    return 0;
}

// ---- K-negative GeoIE mini-batch (geoie_k.cuh) ------------------------------------------------------
// One GeoIE mini-batch step on the tables given (the real ones on one GPU; compact copies of the touched rows with slot
// indices in `gb` under mf_mg.cuh).  seg_h / seg_g: segments of the h / z and g occurrences (inverse map included) whose
// `uniq` entries are the row numbers IN THE TABLES PASSED.  ab_apply == NULL: a, b are left alone and out_dev[1..2] receive
// this rank's d cost / d a, d cost / d b.
static int geoie_batch_core(poi_engine* e, float* g, float* h, float* z, const double* ab_read, double* ab_apply, int H,
                            const GeoBatch& gb, const SegList& seg_h, const SegList& seg_g, float alpha, float lambda,
                            double* out_dev) {
    const int Bu = gb.Bu, n = gb.L - 1, C = gb.K + 1;
    const int64_t n_occ = (int64_t)Bu * n * C, n_g = (int64_t)Bu * n;
    uint8_t *single_h = nullptr, *single_g = nullptr;
    POI_TRY(arena_get(e, (size_t)n_occ, &single_h));
    POI_TRY(arena_get(e, (size_t)n_g, &single_g));
    POI_CAT(e, CAT_INDEX, 0, 0);
    POI_LAUNCH(e, k_mark_single, (unsigned)poi_cdiv(n_occ, 256), 256, 0, seg_h, single_h);
    POI_LAUNCH(e, k_mark_single, (unsigned)poi_cdiv(n_g, 256), 256, 0, seg_g, single_g);
    float *GH = nullptr, *GG = nullptr; double* part = nullptr;
    POI_TRY(arena_get(e, (size_t)n_occ * H, &GH));
    POI_TRY(arena_get(e, (size_t)n_g * H, &GG));
    const int blocks = (int)std::min<int64_t>(Bu, (int64_t)e->num_sms * 2);
    POI_TRY(arena_get(e, (size_t)blocks * 3, &part));
    const size_t smem = geoie_k_smem(H);
    // algorithmic bytes: per target 1 g row + (h, z) x (1 + K) rows, each read once and written once
    POI_CAT(e, CAT_GEOIE, 0, (double)Bu * n * (2.0 * (1 + 2 * C) * H * 4 + 4.0 * C));
    if (H <= 256) {
        POI_CK(e, cudaFuncSetAttribute(k_geoie_batch_k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        POI_LAUNCH(e, (k_geoie_batch_k<1>), blocks, 256, smem, g, h, ab_read, H, gb, single_h, single_g, alpha, lambda, GH, GG, part);
    } else {
        POI_CK(e, cudaFuncSetAttribute(k_geoie_batch_k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        POI_LAUNCH(e, (k_geoie_batch_k<2>), blocks, 256, smem, g, h, ab_read, H, gb, single_h, single_g, alpha, lambda, GH, GG, part);
    }
    POI_CAT(e, CAT_REDUCE, 0, 0);
    POI_LAUNCH(e, k_geoie_k_finalize, 1, 32, 0, part, blocks, ab_apply, alpha, out_dev);
    // rows that occur several times in the batch: duplicate-summed, L2 once per occurrence
    RowSrc src; memset(&src, 0, sizeof(src));
    src.mode = SRC_DENSE_GRADS; src.dim = H; src.skip_single = 1;
    src.grads = GH; POI_TRY(launch_rows_update(e, seg_h, h, H, alpha, lambda, src, ROW_LONG_THRESH));
    // z: no loss gradient, L2 decay of every unique row (lambda x its occurrence count), single or not
    src.grads = nullptr; src.skip_single = 0;
    POI_TRY(launch_rows_update(e, seg_h, z, H, alpha, lambda, src, ROW_LONG_THRESH));
    src.skip_single = 1;
    src.grads = GG; POI_TRY(launch_rows_update(e, seg_g, g, H, alpha, lambda, src, ROW_LONG_THRESH));
    return 0;
}

static int geoie_check(poi_engine* e, const poi_geoie_params* prm, int L, int K, const float* coords_dev) {
    if (!prm || !prm->g || !prm->h || !prm->z || !prm->ab) POI_FAIL(e, "geoie params: null pointer");
    if (prm->H <= 0 || prm->H % 4 || prm->H > 512) POI_FAIL(e, "n_hidden must be a multiple of 4, <= 512");
    if (L < 2 || L - 1 > GEO_MAXN) POI_FAIL(e, "sequence length must be in [2, %d] for the mini-batch kernel", GEO_MAXN + 1);
    if (K < 1 || K > 128 || !coords_dev) POI_FAIL(e, "1 <= K <= 128 and a coordinate table are required");
    return 0;
}

extern "C" int poi_geoie_train_batch_k(poi_engine* e, const poi_geoie_params* prm, const int32_t* P, const int32_t* Q,
                                       const float* coords_dev, int32_t Bu, int32_t L, int32_t K, int32_t on_host,
                                       float alpha, float lambda, double* loss_host) {
    POI_TRY(begin_call(e));
    POI_TRY(geoie_check(e, prm, L, K, coords_dev));
    const int H = prm->H;
    if (Bu <= 0) { if (loss_host) *loss_host = 0.0; return 0; }
    const int n = L - 1, C = K + 1;
    const int64_t n_occ = (int64_t)Bu * n * C, n_g = (int64_t)Bu * n;
    if (n_occ >= (int64_t)1 << 31) POI_FAIL(e, "batch too large");
    GeoBatch gb; gb.Bu = Bu; gb.L = L; gb.K = K;
    gb.coords_g = gb.coords_h = reinterpret_cast<const float4*>(coords_dev);
    if (on_host) {
        const void* hs[2] = {P, Q}; size_t bs[2] = {(size_t)Bu * L * 4, (size_t)Bu * L * K * 4}; void* dv[2];
        POI_TRY(upload_many(e, hs, bs, 2, dv));
        gb.P = (const int32_t*)dv[0]; gb.Q = (const int32_t*)dv[1];
    } else { gb.P = P; gb.Q = Q; }
    gb.Ph = gb.P;
    uint32_t *keys_h = nullptr, *keys_g = nullptr;
    POI_TRY(arena_get(e, (size_t)n_occ, &keys_h));
    POI_TRY(arena_get(e, (size_t)n_g, &keys_g));
    POI_CAT(e, CAT_INDEX, 0, 0);
    POI_LAUNCH(e, k_geoie_keys, (unsigned)poi_cdiv(n_occ, 256), 256, 0, gb, keys_h, keys_g);
    SegList seg_h, seg_g;
    POI_TRY(build_segments(e, keys_h, n_occ, (uint32_t)prm->n_rows, true, &seg_h));
    POI_TRY(build_segments(e, keys_g, n_g, (uint32_t)prm->n_rows, true, &seg_g));
    double* out_dev = nullptr;
    POI_TRY(arena_get(e, 4, &out_dev));
    POI_TRY(geoie_batch_core(e, prm->g, prm->h, prm->z, prm->ab, prm->ab, H, gb, seg_h, seg_g, alpha, lambda, out_dev));
    POI_CK(e, cudaMemcpyAsync(e->h_out, out_dev, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    POI_CK(e, cudaStreamSynchronize(e->stream));
    if (e->kprof) prof_harvest(e);
    if (loss_host) *loss_host = e->h_out[0];
    return 0;
}
