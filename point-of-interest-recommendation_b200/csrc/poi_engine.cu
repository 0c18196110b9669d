// poi_engine.cu -- extern "C" surface of the engine (see include/poi_engine.h).
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include "common.cuh"
#include "sort.cuh"
#include "rows.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "gru.cuh"
#include "mf.cuh"
#include "geoie.cuh"
#include "prme_k.cuh"
#include "geoie_k.cuh"
#include "peer.cuh"
#include "sampling.cuh"
#include "eval.cuh"

extern "C" {

int poi_engine_create(int device, poi_engine** out) {
    if (!out) return -1;
    *out = nullptr;
    int ndev = 0;
    cudaError_t s = cudaGetDeviceCount(&ndev);
    if (s != cudaSuccess || ndev <= 0) {
        g_create_err = std::string("poi_engine_create: no CUDA device (") + cudaGetErrorString(s) + ")";
        return -2;
    }
    if (device < 0 || device >= ndev) { g_create_err = "poi_engine_create: bad device index"; return -3; }
    s = cudaSetDevice(device);
    if (s != cudaSuccess) { g_create_err = std::string("cudaSetDevice: ") + cudaGetErrorString(s); return -4; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10) {
        g_create_err = "poi_engine_create: this library is built for sm_100a (B200) only; found sm_" +
                       std::to_string(prop.major) + std::to_string(prop.minor);
        return -5;
    }
    poi_engine* e = new poi_engine();
    e->device = device;
    e->num_sms = prop.multiProcessorCount;
    { const char* v = getenv("POI_GEMM_CSPLIT"); if (v && v[0] == '0') e->gemm_csplit = false; }
    if (cudaMallocHost((void**)&e->h_out, 64 * sizeof(double)) != cudaSuccess) {
        g_create_err = "cudaMallocHost failed"; delete e; return -6;
    }
    for (auto& ev : e->ev) cudaEventCreate(&ev);
    if (cudaMalloc((void**)&e->grid_bar, 256) != cudaSuccess || cudaMemset(e->grid_bar, 0, 256) != cudaSuccess) {
        g_create_err = "cudaMalloc (grid barrier) failed"; cudaFreeHost(e->h_out); delete e; return -6;
    }
    *out = e;
    return 0;
}

void poi_engine_destroy(poi_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    for (auto& kv : e->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
    for (auto& c : e->chunks) cudaFree(c.ptr);
    if (e->h_out) cudaFreeHost(e->h_out);
    if (e->grid_bar) cudaFree(e->grid_bar);
    if (e->h_stage) cudaFreeHost(e->h_stage);
    for (auto& ev : e->ev) if (ev) cudaEventDestroy(ev);
    for (auto& r : e->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    delete e;
}

const char* poi_last_error(poi_engine* e) { return e ? e->err.c_str() : g_create_err.c_str(); }

int poi_set_stream(poi_engine* e, void* s) { e->stream = (cudaStream_t)s; return 0; }
int poi_sync(poi_engine* e) { POI_CK(e, cudaStreamSynchronize(e->stream)); return 0; }
int poi_launch_count(poi_engine* e, int64_t* out) { *out = e->launches; return 0; }
int poi_last_phase_ms(poi_engine* e, float* out8) { for (int i = 0; i < 8; ++i) out8[i] = e->phase_ms[i]; return 0; }
int poi_enable_phase_timing(poi_engine* e, int on) { e->timing = on != 0; return 0; }
int poi_set_gemm_mode(poi_engine* e, int mode) {
    if (mode < 0 || mode > 2) POI_FAIL(e, "gemm mode must be 0, 1 or 2");
    e->gemm_mode = mode; return 0;
}
int poi_get_gemm_mode(poi_engine* e, int* mode) { *mode = e->gemm_mode; return 0; }
int poi_set_fused_recurrence(poi_engine* e, int on) { e->fuse_recurrence = on != 0; return 0; }
int poi_set_small_batch_path(poi_engine* e, int on) { e->small_batch_path = on != 0; return 0; }
int poi_set_graph_mode(poi_engine* e, int on) { e->graph_mode = on != 0; return 0; }
int poi_graph_replays(poi_engine* e, int64_t* out) { *out = e->graph_replays; return 0; }
int poi_set_fused_sort(poi_engine* e, int on) { e->fused_sort = on != 0; return 0; }
int poi_set_fused_cluster(poi_engine* e, int cl) {
    if (cl != 0 && cl != 1 && cl != 2 && cl != 4) POI_FAIL(e, "poi_set_fused_cluster: %d (0 auto, 1, 2, 4)", cl);
    e->fused_cluster = cl; return 0;
}
int poi_set_wgrad_mn(poi_engine* e, int on) { e->wgrad_mn = on != 0; return 0; }
int poi_kprof_enable(poi_engine* e, int on) { e->kprof = on != 0; return 0; }
int poi_kprof_reset(poi_engine* e) {
    for (int c = 0; c < POI_NCAT; ++c) { e->cat_ms[c] = e->cat_flops[c] = e->cat_bytes[c] = 0.0; e->cat_launches[c] = 0; }
    return 0;
}
int poi_kprof_get(poi_engine* e, double* out) {
    POI_CK(e, cudaStreamSynchronize(e->stream));
    prof_harvest(e);
    for (int c = 0; c < POI_NCAT; ++c) {
        out[c * 4 + 0] = e->cat_ms[c]; out[c * 4 + 1] = (double)e->cat_launches[c];
        out[c * 4 + 2] = e->cat_flops[c]; out[c * 4 + 3] = e->cat_bytes[c];
    }
    return 0;
}

static int begin_call(poi_engine* e) {
    POI_CK(e, cudaSetDevice(e->device));
    if (e->nrec > 4096) { POI_CK(e, cudaStreamSynchronize(e->stream)); prof_harvest(e); }
    e->gemm_cat = -1;
    POI_CAT(e, CAT_OTHER, 0, 0);
    // between poi_gru_mg_prepare and poi_gru_train_mg the arena holds the prepared segments: append, don't reset
    if (!e->prep_valid) POI_TRY(arena_reset(e));
    return 0;
}

// ---- first-slice kernels ----------------------------------------------------------------------
int poi_gather_rows(poi_engine* e, const float* table, int64_t n_rows, int dim, const int32_t* idx,
                    int64_t n_idx, float* out) {
    POI_TRY(begin_call(e));
    if (dim <= 0 || dim % 4) POI_FAIL(e, "dim must be a positive multiple of 4");
    (void)n_rows;
    return launch_gather_rows(e, table, dim, idx, n_idx, out);
}

__global__ void k_seg_counts(SegList seg, int32_t* uniq_out, int32_t* count_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *seg.n_unique) {
        uniq_out[i] = (int32_t)seg.uniq[i];
        count_out[i] = (int32_t)(seg.seg_start[i + 1] - seg.seg_start[i]);
    }
}

int poi_unique(poi_engine* e, const int32_t* idx, int64_t n, int32_t key_bound, int32_t* uniq,
               int32_t* count, int64_t* n_unique_host) {
    POI_TRY(begin_call(e));
    SegList seg;
    POI_TRY(build_segments(e, reinterpret_cast<const uint32_t*>(idx), n, (uint32_t)key_bound, false, &seg));
    if (n > 0) POI_LAUNCH(e, k_seg_counts, (unsigned)poi_cdiv(n, 256), 256, 0, seg, uniq, count);
    uint32_t nu = 0;
    POI_CK(e, cudaMemcpyAsync(&nu, seg.n_unique, 4, cudaMemcpyDeviceToHost, e->stream));
    POI_CK(e, cudaStreamSynchronize(e->stream));
    *n_unique_host = nu;
    return 0;
}

int poi_scatter_sgd(poi_engine* e, float* table, int64_t n_rows, int dim, const int32_t* idx, int64_t n,
                    const float* grad, float alpha, float lambda) {
    POI_TRY(begin_call(e));
    if (dim <= 0 || dim % 4) POI_FAIL(e, "dim must be a positive multiple of 4");
    SegList seg;
    POI_TRY(build_segments(e, reinterpret_cast<const uint32_t*>(idx), n, (uint32_t)n_rows, false, &seg));
    RowSrc src; memset(&src, 0, sizeof(src));
    src.mode = SRC_DENSE_GRADS; src.grads = grad; src.dim = dim;
    return launch_rows_update(e, seg, table, dim, alpha, lambda, src, ROW_LONG_THRESH);
}

int poi_sumsq(poi_engine* e, const float* x, int64_t n, double* out_host) {
    POI_TRY(begin_call(e));
    double* od = nullptr;
    POI_TRY(arena_get(e, 1, &od));
    POI_TRY(launch_sumsq(e, x, n, od));
    POI_CK(e, cudaMemcpyAsync(e->h_out, od, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    POI_CK(e, cudaStreamSynchronize(e->stream));
    *out_host = e->h_out[0];
    return 0;
}

// ---- GRU family ---------------------------------------------------------------------------------
static inline uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

static int gru_train_body(poi_engine* e, const poi_gru_params* p, const poi_seq_index* index, const int32_t* uidx_host,
                          int32_t B, int32_t max_len, float alpha, float lambda, double* out_host) {
    const bool head = p->di != nullptr;
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
    // the plain GRU's t = 0 term: log sigmoid(0) for every non-empty user (GRU.py:352 with h_{-1} = 0).
    // Users in the reference data always have L >= 1, so this is B; callers with empty rows pass lens.
    int64_t n_nonempty = head ? 0 : B;
    return gru_train_core(e, p, ix, B, index->lmax, max_len, n_nonempty, alpha, lambda, out_host);
}

// The reference trains one user per call (GRU.py:388-389, GRU_Spatial.py:290-292): ~120 tiny kernels whose launch
// overhead is the whole cost.  A call with B <= 8 is therefore captured into a CUDA graph the second time its shape
// is seen (same parameter / index pointers, B, max length, hyper-parameters, kernel modes) and replayed afterwards:
// the user index travels through the pinned staging buffer, every other argument is identical by construction
// (the bump arena hands out the same addresses for the same sizes; a re-allocated arena invalidates the graphs).
int poi_gru_train(poi_engine* e, const poi_gru_params* p, const poi_seq_index* index,
                  const int32_t* uidx_host, int32_t B, int32_t max_len, float alpha, float lambda,
                  double* out_host) {
    POI_TRY(begin_call(e));
    POI_TRY(gru_check_params(e, p));
    if (!index || !index->p || !index->q || !index->lens) POI_FAIL(e, "index matrices missing");
    const bool head = p->di != nullptr;
    if (head && (!index->dp || !index->dq)) POI_FAIL(e, "Distance2Pre needs dp/dq index matrices");
    if (B <= 0) POI_FAIL(e, "empty batch");
    const bool graphable = e->graph_mode && B <= 8 && !e->kprof && !e->timing;
    if (!graphable) return gru_train_body(e, p, index, uidx_host, B, max_len, alpha, lambda, out_host);

    // cache key = every field the captured launches depend on, listed explicitly (no struct padding), stored in the entry
    // and compared on lookup: a 64-bit hash collision can not replay a graph built for other pointers or another T
    std::vector<uint64_t> kv = {(uint64_t)(uintptr_t)p->lt, (uint64_t)p->n_rows_lt, (uint64_t)p->d, (uint64_t)p->H, (uint64_t)(uintptr_t)p->ui,
                                (uint64_t)(uintptr_t)p->wh, (uint64_t)(uintptr_t)p->bi, (uint64_t)(uintptr_t)p->di, (uint64_t)p->n_rows_di,
                                (uint64_t)(uintptr_t)p->vs, (uint64_t)(uintptr_t)p->bs, (uint64_t)(uintptr_t)p->scal,
                                (uint64_t)(uintptr_t)index->p, (uint64_t)(uintptr_t)index->q, (uint64_t)(uintptr_t)index->dp,
                                (uint64_t)(uintptr_t)index->dq, (uint64_t)(uintptr_t)index->lens, (uint64_t)index->lmax, (uint64_t)index->n_user,
                                (uint64_t)B, (uint64_t)max_len, (uint64_t)e->gemm_mode,
                                (uint64_t)((int)e->fuse_recurrence | ((int)e->persistent_gemm << 1) | ((int)e->wgrad_mn << 2) | ((int)e->small_batch_path << 3)),
                                (uint64_t)e->fused_cluster, 0, 0, (uint64_t)(uintptr_t)e->stream};
    memcpy(&kv[24], &alpha, 4); memcpy(&kv[25], &lambda, 4);
    uint64_t key = fnv1a(1469598103934665603ull, kv.data(), kv.size() * sizeof(uint64_t));
    {
        auto it = e->graphs.find(key);
        while (it != e->graphs.end() && !it->second.key.empty() && it->second.key != kv) it = e->graphs.find(++key);   // collision: probe
    }
    if (e->graphs.size() >= 128 && e->graphs.find(key) == e->graphs.end()) {      // bounded: drop the least recently used entry
        auto lru = e->graphs.begin();
        for (auto it = e->graphs.begin(); it != e->graphs.end(); ++it) if (it->second.last_use < lru->second.last_use) lru = it;
        if (lru->second.exec) cudaGraphExecDestroy(lru->second.exec);
        e->graphs.erase(lru);
    }
    poi_engine::GraphEntry& ge = e->graphs[key];
    if (ge.key.empty()) ge.key = kv;
    ge.last_use = ++e->graph_clock;
    auto finish = [&]() -> int {
        POI_CK(e, cudaStreamSynchronize(e->stream));
        if (out_host) for (int i = 0; i < 5; ++i) out_host[i] = e->h_out[i];
        return 0;
    };
    if (ge.exec && ge.arena_gen == e->arena_gen) {                       // ---- replay ----
        memcpy(e->h_stage, uidx_host, (size_t)B * 4);
        POI_CK(e, cudaGraphLaunch(ge.exec, e->stream));
        e->launches += ge.n_launch; e->graph_replays++;
        return finish();
    }
    if (ge.exec) { cudaGraphExecDestroy(ge.exec); ge.exec = nullptr; ge.warm = 0; }
    if (ge.warm < 1 || e->chunks.size() != 1) {                          // ---- first sight: plain call (sizes the arena) ----
        ge.warm++;
        return gru_train_body(e, p, index, uidx_host, B, max_len, alpha, lambda, out_host);
    }
    // ---- capture ----
    if (!e->cap_stream) POI_CK(e, cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
    POI_TRY(stage_reserve(e, (size_t)B * 4 + 256));
    POI_CK(e, cudaStreamSynchronize(e->stream));
    cudaStream_t user = e->stream;
    const int64_t launches0 = e->launches;
    const uint64_t gen0 = e->arena_gen;
    cudaError_t st = cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeRelaxed);
    int rc = -1;
    cudaGraph_t graph = nullptr;
    if (st == cudaSuccess) {
        e->stream = e->cap_stream; e->capturing = true;
        rc = gru_train_body(e, p, index, uidx_host, B, max_len, alpha, lambda, nullptr);
        e->stream = user; e->capturing = false;
        st = cudaStreamEndCapture(e->cap_stream, &graph);
    }
    if (rc == 0 && st == cudaSuccess && graph && e->arena_gen == gen0) st = cudaGraphInstantiate(&ge.exec, graph, 0);
    else if (st == cudaSuccess) st = cudaErrorUnknown;
    if (graph) cudaGraphDestroy(graph);
    if (st != cudaSuccess || !ge.exec) {
        // capture is an optimisation: fall back to kernel-by-kernel launches for good and redo this call
        cudaGetLastError();
        ge.exec = nullptr; e->graph_mode = 0; e->launches = launches0;
        POI_TRY(begin_call(e));
        return gru_train_body(e, p, index, uidx_host, B, max_len, alpha, lambda, out_host);
    }
    ge.arena_gen = e->arena_gen; ge.n_launch = e->launches - launches0;
    POI_CK(e, cudaGraphLaunch(ge.exec, e->stream));
    e->graph_replays++;
    return finish();
}

int poi_gru_train_host_rows(poi_engine* e, const poi_gru_params* p, const int32_t* p_host,
                            const int32_t* q_host, const int32_t* dp_host, const int32_t* dq_host,
                            const int32_t* lens_host, int32_t B, int32_t lmax, float alpha, float lambda,
                            double* out_host) {
    POI_TRY(begin_call(e));
    POI_TRY(gru_check_params(e, p));
    const bool head = p->di != nullptr;
    if (!p_host || !q_host || !lens_host || (head && (!dp_host || !dq_host))) POI_FAIL(e, "host index rows missing");
    if (B <= 0 || lmax <= 0) POI_FAIL(e, "empty batch");
    phase_mark(e, 0);
    const size_t LB = (size_t)B * lmax;
    POI_TRY(stage_reserve(e, (head ? 4 : 2) * (LB * 4 + 256) + (size_t)B * 4 + 256));
    size_t so = 0;
    int32_t *P = nullptr, *Q = nullptr, *DP = nullptr, *DQ = nullptr, *lens = nullptr;
    POI_TRY(gru_upload_i32(e, p_host, LB, &P, &so));
    POI_TRY(gru_upload_i32(e, q_host, LB, &Q, &so));
    if (head) { POI_TRY(gru_upload_i32(e, dp_host, LB, &DP, &so)); POI_TRY(gru_upload_i32(e, dq_host, LB, &DQ, &so)); }
    POI_TRY(gru_upload_i32(e, lens_host, (size_t)B, &lens, &so));
    int max_len = 0; int64_t n_nonempty = 0;
    for (int b = 0; b < B; ++b) { max_len = std::max(max_len, lens_host[b]); n_nonempty += lens_host[b] >= 1; }
    GruIdx ix;
    POI_TRY(gru_alloc_idx(e, B, lmax, head, &ix));
    POI_CAT(e, CAT_INDEX, 0, 0);
    POI_LAUNCH(e, k_slice_indices, (unsigned)poi_cdiv((int64_t)LB, 256), 256, 0, P, Q, DP, DQ, lens, lmax,
               (const int32_t*)nullptr, B, ix.PQt, ix.DPt, ix.DQt, ix.lensB);
    return gru_train_core(e, p, ix, B, lmax, max_len, head ? 0 : n_nonempty, alpha, lambda, out_host);
}

__global__ void k_pick_last(const float* __restrict__ Hs, const int32_t* __restrict__ lensB, int B, int H4,
                            float* __restrict__ hts) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)B * H4) return;
    int b = (int)(idx / H4), c = (int)(idx % H4);
    int L = lensB[b];
    // hs[arange(B), sum(mask)-1] (GRU.py:190-193): h_{L-1} = Hs[L]
    reinterpret_cast<float4*>(hts)[idx] = reinterpret_cast<const float4*>(Hs)[((int64_t)L * B + b) * H4 + c];
}

__global__ void k_softmax_rows(const float* __restrict__ logits, int ld, int n, int64_t rows, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= rows) return;
    const float* r = logits + (size_t)warp * ld;
    float mx = -INFINITY;
    for (int k = lane; k < n; k += 32) mx = fmaxf(mx, r[k]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int k = lane; k < n; k += 32) sum += expf(r[k] - mx);
    sum = warp_sum(sum);
    for (int k = lane; k < n; k += 32) out[(size_t)warp * n + k] = expf(r[k] - mx) / sum;
}

int poi_gru_predict(poi_engine* e, const poi_gru_params* p, const poi_seq_index* index,
                    const int32_t* uidx_host, int32_t B, int32_t max_len, float* hts_dev, float* sts_dev) {
    POI_TRY(begin_call(e));
    POI_TRY(gru_check_params(e, p));
    const bool head = p->di != nullptr;
    if (B <= 0) POI_FAIL(e, "empty batch");
    if (head && !sts_dev) POI_FAIL(e, "sts output required for Distance2Pre");
    POI_TRY(stage_reserve(e, (size_t)B * 4 + 256));
    size_t so = 0;
    int32_t* uidx_dev = nullptr;
    POI_TRY(gru_upload_i32(e, uidx_host, (size_t)B, &uidx_dev, &so));
    GruIdx ix;
    POI_TRY(gru_alloc_idx(e, B, index->lmax, head, &ix));
    // predict has no negatives: reuse p for the q slot of the slicer
    POI_CAT(e, CAT_INDEX, 0, 0);
    POI_LAUNCH(e, k_slice_indices, (unsigned)poi_cdiv((int64_t)B * index->lmax, 256), 256, 0, index->p, index->p,
               head ? index->dp : nullptr, head ? index->dp : nullptr, index->lens, index->lmax, uidx_dev, B,
               ix.PQt, ix.DPt, ix.DQt, ix.lensB);
    const int T = std::min(max_len, index->lmax);      // seq_length = max L (GRU.py:162), all L steps run
    float *X, *XDiff, *AX, *Hs, *Z, *R, *C, *RH;
    POI_TRY(gru_forward(e, p, ix, B, index->lmax, T, false, &X, &XDiff, &AX, &Hs, &Z, &R, &C, &RH));
    const int H = p->H;
    POI_CAT(e, CAT_ELTWISE, 0, 0);
    POI_LAUNCH(e, k_pick_last, (unsigned)poi_cdiv((int64_t)B * H / 4, 256), 256, 0, Hs, ix.lensB, B, H / 4, hts_dev);
    if (head) {
        const int nD = p->n_rows_di, nDp = (nD + 3) / 4 * 4;
        float* lg = nullptr;
        POI_TRY(arena_get(e, (size_t)B * nDp, &lg));
        POI_TRY(gemm_tn(e, hts_dev, H, p->vs, H, B, nD, H, EpiBiasStore{lg, nDp, p->bs, nD}));
        POI_CAT(e, CAT_LOSS, 0, 0);
        POI_LAUNCH(e, k_softmax_rows, (unsigned)poi_cdiv((int64_t)B * 32, 256), 256, 0, lg, nDp, nD, (int64_t)B, sts_dev);
    }
    POI_CK(e, cudaStreamSynchronize(e->stream));
    return 0;
}

// ---- BPR / PRME -----------------------------------------------------------------------------------
static int upload_many(poi_engine* e, const void* const* hosts, const size_t* bytes, int k, void** devs) {
    size_t tot = 0;
    for (int i = 0; i < k; ++i) tot += poi_align_up(bytes[i], 256);
    POI_TRY(stage_reserve(e, tot + 256));
    size_t so = 0;
    for (int i = 0; i < k; ++i) {
        POI_TRY(arena_alloc(e, bytes[i], &devs[i]));
        // page-locked caller memory goes to the device directly (every caller synchronises before it returns)
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, hosts[i]) == cudaSuccess && at.type == cudaMemoryTypeHost) {
            POI_CK(e, cudaMemcpyAsync(devs[i], hosts[i], bytes[i], cudaMemcpyHostToDevice, e->stream));
            continue;
        }
        cudaGetLastError();
        memcpy(e->h_stage + so, hosts[i], bytes[i]);
        POI_CK(e, cudaMemcpyAsync(devs[i], e->h_stage + so, bytes[i], cudaMemcpyHostToDevice, e->stream));
        so += poi_align_up(bytes[i], 256);
    }
    return 0;
}

int poi_bpr_train_seq(poi_engine* e, float* ux, float* lt, int32_t d, const int32_t* u, const int32_t* p,
                      const int32_t* q, int64_t n, float alpha, float lambda, double* loss_host) {
    POI_TRY(begin_call(e));
    if (d <= 0 || d % 4 || d > 1024) POI_FAIL(e, "d must be a multiple of 4, <= 1024");
    if (n <= 0) return 0;
    const void* hs[3] = {u, p, q}; size_t bs[3] = {(size_t)n * 4, (size_t)n * 4, (size_t)n * 4}; void* ds[3];
    POI_TRY(upload_many(e, hs, bs, 3, ds));
    double* loss_dev = nullptr;
    POI_TRY(arena_get(e, (size_t)n, &loss_dev));
    const int d4 = d / 4;
    POI_CAT(e, CAT_MF, 0, (double)n * 3 * d * 4 * 2);
#define BPR_GO(N) POI_LAUNCH(e, (k_bpr_seq<N>), 1, 32, 0, ux, lt, d4, (const int32_t*)ds[0], (const int32_t*)ds[1], (const int32_t*)ds[2], n, alpha, lambda, loss_dev)
    if (d4 <= 32) BPR_GO(1); else if (d4 <= 64) BPR_GO(2); else if (d4 <= 128) BPR_GO(4); else BPR_GO(8);
#undef BPR_GO
    POI_CK(e, cudaMemcpyAsync(loss_host, loss_dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    POI_CK(e, cudaStreamSynchronize(e->stream));
    return 0;
}

int poi_prme_train_seq(poi_engine* e, float* du, float* dp, float* ds_, int32_t d, const int32_t* u,
                       const int32_t* p, const int32_t* q, const int32_t* prev, const double* dist,
                       const int32_t* gap, int64_t n, int32_t threshold, double cw, float alpha, float lambda,
                       double* loss_host) {
    POI_TRY(begin_call(e));
    if (d <= 0 || d % 4 || d > 512) POI_FAIL(e, "d must be a multiple of 4, <= 512");
    if (n <= 0) return 0;
    const void* hs[6] = {u, p, q, prev, dist, gap};
    size_t bs[6] = {(size_t)n * 4, (size_t)n * 4, (size_t)n * 4, (size_t)n * 4, (size_t)n * 8, (size_t)n * 4};
    void* ds[6];
    POI_TRY(upload_many(e, hs, bs, 6, ds));
    double* loss_dev = nullptr;
    POI_TRY(arena_get(e, (size_t)n, &loss_dev));
    const int d4 = d / 4;
    POI_CAT(e, CAT_MF, 0, (double)n * 7 * d * 4 * 2);
#define PRME_GO(N) POI_LAUNCH(e, (k_prme_seq<N>), 1, 32, 0, du, dp, ds_, d4, (const int32_t*)ds[0], (const int32_t*)ds[1], (const int32_t*)ds[2], (const int32_t*)ds[3], (const double*)ds[4], (const int32_t*)ds[5], n, (int)threshold, (float)cw, alpha, lambda, loss_dev)
    if (d4 <= 32) PRME_GO(1); else if (d4 <= 64) PRME_GO(2); else PRME_GO(4);
#undef PRME_GO
    POI_CK(e, cudaMemcpyAsync(loss_host, loss_dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    POI_CK(e, cudaStreamSynchronize(e->stream));
    return 0;
}

}  // extern "C"

#include "api_more.cuh"
#include "mg_step.cuh"
#include "mf_mg.cuh"
