/*
 * poi_engine.h -- C-ABI of the B200-native next-POI training engine.
 *
 * The reference (tangrizzly/Point-of-Interest-Recommendation) has no FFI: its
 * seam is the Python model-class surface that prog_*.py calls, i.e. the objects
 * that wrap `theano.function`.  Each entry point below
 * names the compiled Theano function / shared-variable operation it replaces.
 * The Python classes in point-of-interest-recommendation_b200/public/ bind these
 * through ctypes (see INTEGRATION.md for the stub a reference maintainer adds).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (message: poi_last_error);
 *     nothing throws across the ABI;
 *   - "dev" pointers are device memory BORROWED from the caller (torch tensors on
 *     the engine's device); the engine never frees them.  "host" pointers are
 *     ordinary host memory.  No torch types appear in any signature;
 *   - one engine per (process, GPU); not thread-safe; work is enqueued on the
 *     engine's stream (poi_set_stream) and every train/predict call synchronises
 *     before returning host scalars, like a theano.function call does;
 *   - tables are fp32 row-major [(rows) x dim]; index matrices are int32
 *     row-major [n_user x lmax] exactly as the reference's theano.shared int32
 *     masks (GRU.py:50-55, GRU_Spatial.py:46-49); pad POI index = n_item, pad
 *     interval index = n_dist;
 *   - dim, n_hidden must be multiples of 4 (128-bit vector access).
 */
#ifndef POI_ENGINE_H_
#define POI_ENGINE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct poi_engine poi_engine;

/* ---- lifetime / plumbing --------------------------------------------------------------- */
int         poi_engine_create(int device, poi_engine** out);
void        poi_engine_destroy(poi_engine* e);
const char* poi_last_error(poi_engine* e);          /* also valid with e == NULL (create errors) */
int         poi_set_stream(poi_engine* e, void* cuda_stream);   /* cudaStream_t; NULL = default  */
int         poi_sync(poi_engine* e);
/* kernels launched by this engine since creation (bench.py's gpu_launches claim) */
int         poi_launch_count(poi_engine* e, int64_t* out);
/* time of the most recent train call's phases in ms (CUDA events on the engine stream):
 * out[0]=whole call, [1]=index prep+sort, [2]=gather, [3]=forward GEMMs+recurrence,
 * [4]=loss head, [5]=backward, [6]=weight grads + dense update, [7]=sparse row update */
int         poi_last_phase_ms(poi_engine* e, float* out8);
int         poi_enable_phase_timing(poi_engine* e, int on);
/* Built-in per-launch profiler: when enabled every kernel the engine launches is bracketed by CUDA
 * events on the engine stream and accumulated per category.  out = double[POI_KPROF_NCAT][4] =
 * {ms, launches, algorithmic flops, algorithmic bytes} for categories
 * 0 other, 1 index prep/sort/unique, 2 gather, 3 GEMM (TN: input projection, head, dgrad), 4 weight-gradient GEMM +
 * dense update, 5 loss head, 6 elementwise, 7 sparse row update, 8 BPR/PRME, 9 GeoIE, 10 eval, 11 reductions,
 * 12 forward recurrence kernel(s), 13 backward recurrence kernel(s). */
#define POI_KPROF_NCAT 14
int         poi_kprof_enable(poi_engine* e, int on);
int         poi_kprof_reset(poi_engine* e);
int         poi_kprof_get(poi_engine* e, double* out /* [POI_KPROF_NCAT * 4] */);
/* 0 = SIMT fp32 FMA GEMMs, 1 = tcgen05 3xTF32 (fp32-faithful), 2 = tcgen05 1xTF32 */
int         poi_set_gemm_mode(poi_engine* e, int mode);
int         poi_get_gemm_mode(poi_engine* e, int* mode);
/* tensor-core modes only: 1 (default) = the recurrence runs as one persistent fused kernel per
 * direction, 0 = two GEMM launches per time step (kept for A/B measurements) */
int         poi_set_fused_recurrence(poi_engine* e, int on);
/* CTAs per 128 users in the fused recurrence kernels (thread-block cluster splitting the gate columns, next
 * operand exchanged through distributed shared memory): 0 = auto (default), 1, 2 or 4. */
int         poi_set_fused_cluster(poi_engine* e, int cl);
/* index lists longer than 4096 keys: 1 (default) = all radix passes and the segment arrays in ONE persistent launch
 * (grid barrier between the phases, csrc/sort.cuh), 0 = one launch per phase (kept for A/B measurements) */
int         poi_set_fused_sort(poi_engine* e, int on);
/* poi_gru_train calls with B <= 8 (the reference's one-by-one mode) are captured into a CUDA graph the second time a
 * shape is seen and replayed afterwards (1 = default); 0 = always launch kernel by kernel.  poi_graph_replays: how
 * many calls were served by a graph launch. */
int         poi_set_graph_mode(poi_engine* e, int on);
/* Batches of <= 8 users (the reference's one-by-one mode) run the recurrence on SIMT kernels that keep Wh in shared
 * memory (csrc/gru_small.cuh; exact fp32 FMA) instead of a 128-row tensor-core tile: 1 = default, 0 = off. */
int         poi_set_small_batch_path(poi_engine* e, int on);
int         poi_graph_replays(poi_engine* e, int64_t* out);
/* tensor-core modes only: 1 (default) = the weight-gradient GEMMs read the activation matrices as they lie in
 * memory (MN-major tcgen05 operands; bias gradients fused into the same pass), 0 = transposed copies + K-major
 * operands + a separate column-sum pass (kept for A/B measurements) */
int         poi_set_wgrad_mn(poi_engine* e, int on);

/* ---- first-slice kernels, individually testable ------------ */

/* out[i,:] = table[idx[i],:]   -- Theano AdvancedSubtensor1, e.g. `self.lt[pidxs]`
 * (GRU.py:327, GRU_Spatial.py:144-146, BPR.py:207-208, PRME.py:178-180, GeoIE.py:143-145). */
int poi_gather_rows(poi_engine* e, const float* table_dev, int64_t n_rows, int dim,
                    const int32_t* idx_dev, int64_t n_idx, float* out_dev);

/* Sorted unique of an int32 vector -- Theano `Unique(False,False,False)` (GRU.py:329-331,
 * GRU_Spatial.py:149-153, GeoIE.py:147-153).  uniq_dev/count_dev need room for n entries;
 * *n_unique_host receives the number of distinct values. */
int poi_unique(poi_engine* e, const int32_t* idx_dev, int64_t n, int32_t key_bound,
               int32_t* uniq_dev, int32_t* count_dev, int64_t* n_unique_host);

/* table[U] -= alpha * (sum over duplicate occurrences of grad rows + lambda*count*table[U]),
 * U = unique(idx) -- `T.set_subtensor(uiq_x, uiq_x - lr * T.grad(cost, self.lt)[uiq_pqs])`
 * (GRU.py:372-373, GRU_Spatial.py:212-215, GeoIE.py:174-181); grad_dev is one row per
 * occurrence [n x dim] (may be NULL = zero).  Summation order is fixed -> bit-reproducible. */
int poi_scatter_sgd(poi_engine* e, float* table_dev, int64_t n_rows, int dim,
                    const int32_t* idx_dev, int64_t n, const float* grad_dev,
                    float alpha, float lambda);

/* C[m, n] = sum_k A[m*lda + k] * W[n*ldw + k] (+ bias[n]) -- the dense contraction every
 * `T.dot(ui, x)`, `T.dot(wh, h)`, `T.dot(vs, h)` of the reference lowers to (GRU.py:346-350,
 * GRU_Spatial.py:173-180), batched over rows.  mode as in poi_set_gemm_mode (0 fp32 FMA, 1 tcgen05
 * 3xTF32, 2 tcgen05 1xTF32).  Exposed so the tensor-core kernels can be tested in isolation. */
int poi_gemm_tn(poi_engine* e, const float* A_dev, int lda, const float* W_dev, int ldw,
                int64_t M, int N, int K, const float* bias_dev, float* C_dev, int ldc, int mode);

/* C[i, j] = sum_m A[m*lda + i] * B[m*ldb + j]  (i < N1, j < N2; N1, N2 multiples of 4) -- the weight-gradient
 * contraction over all (t, b) rows (T.grad w.r.t. ui / wh / vs, GRU.py:370, GRU_Spatial.py:210).  mode 0 = fp32
 * FMA split-K kernel, 1/2 = tcgen05 with MN-major operands (no transposed copies).  Test hook. */
int poi_gemm_atb(poi_engine* e, const float* A_dev, int lda, const float* B_dev, int ldb,
                 int64_t M, int N1, int N2, float* C_dev, int mode);

/* sum of squares of n floats, fp64 accumulation -- building block of `model.l2.eval()`
 * (GRU.py:305-309, GRU_Spatial.py:83-88, BPR.py:195-198, PRME.py:166-169, GeoIE.py:92-98). */
int poi_sumsq(poi_engine* e, const float* x_dev, int64_t n, double* out_host);

/* ---- GRU family: OboGru / Gru / OboSpatialGru (Distance2Pre) --------------------------- */
typedef struct {
    float*   lt;        /* [(n_item+1) x d]   item table, last row = pad   (GRU.py:60,66)         */
    int64_t  n_rows_lt; /* n_item + 1                                                             */
    int32_t  d;         /* n_in                                                                   */
    int32_t  H;         /* n_hidden                                                               */
    float*   ui;        /* [3 x H x din], din = d (GRU.py:61) or 2d (GRU_Spatial.py:51)           */
    float*   wh;        /* [3 x H x H]                                      (GRU.py:62)           */
    float*   bi;        /* [3 x H]                                          (GRU.py:64)           */
    /* Distance2Pre only (all NULL / 0 for the plain GRU): */
    float*   di;        /* [(n_dist+1) x d]   interval table               (GRU_Spatial.py:57)   */
    int32_t  n_rows_di; /* n_dist + 1                                                             */
    float*   vs;        /* [(n_dist+1) x H]                                 (GRU_Spatial.py:60)   */
    float*   bs;        /* [n_dist+1]                                       (GRU_Spatial.py:61)   */
    float*   scal;      /* dev float[3] = {wd, loss_weight[0], loss_weight[1]} (GRU_Spatial.py:66-71) */
} poi_gru_params;

typedef struct {
    const int32_t* p;    /* tra_buys_masks      [n_user x lmax]  (GRU.py:50)                      */
    const int32_t* q;    /* tra_buys_neg_masks  [n_user x lmax]  (GRU.py:54)                      */
    const int32_t* dp;   /* tra_dist_masks      (GRU_Spatial.py:47)   NULL for the plain GRU      */
    const int32_t* dq;   /* tra_dist_neg_masks  (GRU_Spatial.py:49)   NULL for the plain GRU      */
    const int32_t* lens; /* [n_user] = tra_masks.sum(1); masks are prefix masks (Load_Data_by_length.py:123) */
    int32_t lmax;
    int32_t n_user;
} poi_seq_index;

/* One `seq_train` call: OboGru.seq_train(uidx) (GRU.py:378-389) / Gru.seq_train(start_end)
 * (GRU.py:481-498) when params->di == NULL; OboSpatialGru.seq_train(uidx)
 * (GRU_Spatial.py:220-229,290-292) otherwise.  B users are one mini-batch step
 * (B=1 = the reference's one-by-one semantics; B>1 for Distance2Pre = the mini-batch
 * extension of SURVEY.md 3.6).  Index matrices are DEVICE resident (like theano.shared);
 * uidx_host are the B row numbers (`givens`, GRU.py:382-385).  max_len_host = max lens
 * over the batch.  Updates params in place.  out_host (double[5]):
 *   plain GRU   : out[0] = -sum(loss)            (GRU.py:380,483)
 *   Distance2Pre: out[0..4] = los, sur, upq, softmax(loss_weight)[0..1] (GRU_Spatial.py:222) */
int poi_gru_train(poi_engine* e, const poi_gru_params* params, const poi_seq_index* index,
                  const int32_t* uidx_host, int32_t B, int32_t max_len_host,
                  float alpha, float lambda, double* out_host);

/* Same step with the batch's index ROWS supplied from HOST memory ([B x lmax] each, pinned
 * or pageable; dp/dq NULL for the plain GRU): the host->device copies are part of the call.
 * This is the end-to-end entry bench.py times (`e2e`). */
int poi_gru_train_host_rows(poi_engine* e, const poi_gru_params* params,
                            const int32_t* p_host, const int32_t* q_host,
                            const int32_t* dp_host, const int32_t* dq_host,
                            const int32_t* lens_host, int32_t B, int32_t lmax,
                            float alpha, float lambda, double* out_host);

/* `seq_predict(start_end)` (GRU.py:197-205, GRU_Spatial.py:282-288): forward over the padded
 * training sequences using the trained_* copies passed in `params` (lt=trained_items,
 * di=trained_dists).  hts_dev [B x H]; sts_dev [B x (n_dist+1)] or NULL for the plain GRU. */
int poi_gru_predict(poi_engine* e, const poi_gru_params* params, const poi_seq_index* index,
                    const int32_t* uidx_host, int32_t B, int32_t max_len_host,
                    float* hts_dev, float* sts_dev);

/* ---- multi-GPU mini-batch step (no reference counterpart -- the reference is single
 * process).  Users are sharded over ranks, the item table `lt` is row-sharded (owner = row % world),
 * dense weights and `di` are replicated.  The caller (Python, torch.distributed over NCCL) moves
 * rows and gradients between ranks; the engine does the arithmetic in two calls:
 *
 *  poi_gru_train_mg : forward + backward for this rank's B users, gathering item rows from
 *      rows_dev = the rows of sorted-unique(p u q of the batch) fetched from their owners
 *      ([n_unique x d], same order as poi_unique returns).  Nothing is updated; it emits
 *        dense_grads_dev  float[poi_gru_mg_dense_size]: loss gradients of ui, wh, bi, vs, bs, then a
 *                         dense [n_dist+1 x d] gradient of di and its [n_dist+1] occurrence counts
 *        row_grads_dev    [n_unique x d] duplicate-summed loss gradient per unique row
 *        row_cnt_dev      float[n_unique] occurrence counts (L2 multiplicity)
 *        loss_sums_dev    double[3] = sum sur, sum log-sigmoid, sum d/d wd
 *      (params->lt may be NULL; params->n_rows_lt must be the GLOBAL row count.)
 *  poi_gru_apply_mg : after all-reduce(dense_grads, loss_sums) and the all-to-all of (row id, row_grads,
 *      row_cnt) to the owners: dense SGD, scalar SGD, and the owner's sparse SGD on its shard
 *      lt_local_dev [n_local_rows x d] with recv_local_ids (= global id / world), duplicate ids from
 *      different ranks summed in arrival order.  out_host as poi_gru_train (global sums). */
int poi_gru_mg_dense_size(const poi_gru_params* params, int64_t* n_floats);
/* Optional first phase: slice + sort the batch's row ids once and hand the sorted unique ids to the
 * caller (uniq_out_dev: room for 2*B*lmax int32).  A poi_gru_train_mg call that follows directly (same
 * B, same index) reuses the sorted segments instead of sorting again. */
int poi_gru_mg_prepare(poi_engine* e, const poi_gru_params* params, const poi_seq_index* index,
                       const int32_t* uidx_host, int32_t B, int32_t* uniq_out_dev, int64_t* n_unique_host);
int poi_gru_train_mg(poi_engine* e, const poi_gru_params* params, const poi_seq_index* index,
                     const int32_t* uidx_host, int32_t B, int32_t max_len_host, int32_t global_batch,
                     const float* rows_dev, int64_t n_unique, float* dense_grads_dev,
                     float* row_grads_dev, float* row_cnt_dev, double* loss_sums_dev);
int poi_gru_apply_mg(poi_engine* e, const poi_gru_params* params, const float* dense_grads_dev,
                     const double* loss_sums_dev, int32_t global_batch, int64_t n_nonempty_global,
                     float* lt_local_dev, int64_t n_local_rows, const int32_t* recv_local_ids_dev,
                     const float* recv_grads_dev, const float* recv_cnts_dev, int64_t n_recv,
                     float alpha, float lambda, double* out_host);

/* ---- the same exchanges over NVLink peer memory, no collective on the data path (csrc/peer.cuh).
 * poi_peer_alloc: device memory exportable to the other ranks of the box (cudaMalloc + cudaIpcGetMemHandle; the
 * 64-byte handle is sent through torch.distributed once); poi_peer_open maps a peer's buffer (peer access is
 * enabled on first use).  poi_gather_rows_sharded: out[i] = row ids[i] of the row-sharded table, read from the
 * OWNER's shard (owner = id % world, local row = id / world); shards_host = host array of `world` device
 * pointers (own shard + opened peers).  poi_pull_segments: rank r's outbox is (ids, gradient rows, counts) exactly
 * as poi_gru_mg_prepare / poi_gru_train_mg wrote them, plus perm = the record numbers grouped by owner; the
 * owner copies the n_host[r] records perm[src_off_host[r] ...] of every rank into its receive buffers, grouped by
 * source rank in rank order, ids converted to local rows -- the input poi_gru_apply_mg expects. */
int poi_peer_alloc(poi_engine* e, int64_t bytes, void** ptr_out, unsigned char* handle_out64);
int poi_peer_free(poi_engine* e, void* ptr);
int poi_peer_open(poi_engine* e, const unsigned char* handle64, void** ptr_out);
int poi_peer_close(poi_engine* e, void* ptr);
int poi_gather_rows_sharded(poi_engine* e, const float* const* shards_host, int world, int dim,
                            const int32_t* ids_dev, int64_t n_idx, float* out_dev);
/* perm_out_dev[n] = record numbers 0..n-1 grouped by owner (ids[i] % world), original order inside a group (one
 * stable radix pass); counts_out_dev[world] (double) = records per owner. */
int poi_group_by_owner(poi_engine* e, const int32_t* ids_dev, int64_t n, int world, int32_t* perm_out_dev,
                       double* counts_out_dev);
int poi_pull_segments(poi_engine* e, int world, int dim, const int32_t* const* perm_host,
                      const int32_t* const* ids_host, const float* const* grads_host,
                      const float* const* cnts_host, const int64_t* src_off_host, const int64_t* n_host,
                      int32_t* recv_local_ids_dev, float* recv_grads_dev, float* recv_cnts_dev);

/* ---- the whole multi-GPU step as ONE call: no collective library on its path, no host synchronisation before the final
 * read-back (csrc/mg_step.cuh).  Every rank owns peer-visible buffers (poi_peer_alloc, mapped by all ranks with
 * poi_peer_open) and passes the table of everybody's pointers; entry [rank] is the rank's own (writable) memory:
 *   shard     [rows_r x d]  item-table shard (owner = row % world, local row = row / world)
 *   ob_ids    [cap] int32, ob_grads [cap x d], ob_cnts [cap], ob_perm [cap] int32: the gradient outbox of the rank's batch
 *             (sorted unique global row ids, duplicate-summed gradient rows, occurrence counts, record numbers grouped by owner)
 *   ob_meta   int32 [world + 2]: records per owner, then n_unique and B
 *   dense     float [poi_gru_mg_dense_size], sums double[4]: dense gradients and loss sums of the rank's batch
 *   flags     uint32 [2 * world], ZERO before the first step: [r] = last step whose outbox rank r published, [world + r] = last
 *             step rank r applied -- written by rank r into every peer's array
 *   slot_tab  own memory, int32 [n_local_rows x world], all -1 (the step leaves it so)
 * cap >= 2 * B * lmax.  Steps are numbered 1, 2, ... and every rank must make the same sequence of calls with the same B.
 * Dense gradients are summed in rank order by every rank itself (identical weights everywhere); a rank that waits longer
 * than POI_MG_TIMEOUT_MS (default 20 000) for a peer fails the call instead of hanging.  out_host as poi_gru_train
 * (GLOBAL sums). */
#define POI_MG_MAX_RANKS 16
typedef struct {
    int32_t   world, rank;
    int64_t   cap;
    int64_t   n_local_rows;
    float*    shard[POI_MG_MAX_RANKS];
    int32_t*  ob_ids[POI_MG_MAX_RANKS];
    float*    ob_grads[POI_MG_MAX_RANKS];
    float*    ob_cnts[POI_MG_MAX_RANKS];
    int32_t*  ob_perm[POI_MG_MAX_RANKS];
    int32_t*  ob_meta[POI_MG_MAX_RANKS];
    float*    dense[POI_MG_MAX_RANKS];
    double*   sums[POI_MG_MAX_RANKS];
    uint32_t* flags[POI_MG_MAX_RANKS];
    int32_t*  slot_tab;
} poi_mg_peers;
int poi_gru_step_mg(poi_engine* e, const poi_gru_params* params, const poi_seq_index* index, const int32_t* uidx_host,
                    int32_t B, int32_t max_len_host, const poi_mg_peers* peers, int64_t step, float alpha, float lambda,
                    double* out_host);
/* the same step with the batch's index ROWS supplied from host memory ([B x lmax] each; the copies are part of the call):
 * the end-to-end entry, as poi_gru_train_host_rows is for one GPU */
int poi_gru_step_mg_host_rows(poi_engine* e, const poi_gru_params* params, const int32_t* p_host, const int32_t* q_host,
                              const int32_t* dp_host, const int32_t* dq_host, const int32_t* lens_host, int32_t B, int32_t lmax,
                              const poi_mg_peers* peers, int64_t step, float alpha, float lambda, double* out_host);

/* ---- the reference's per-epoch host loops on the device (csrc/sampling.cuh).
 * poi_sample_negatives = fun_random_neg_masks_tra / _tes (Load_Data_by_length.py:127-162): out[u][t] = a uniform
 * draw from [0, n_item) redrawn while it occurs in the user's forbidden rows (sorted_a [n_user x la], optionally
 * sorted_b [n_user x lb], each row sorted ascending), for every t with rows[u][t] != n_item; pad positions get n_item.
 * Counter-based stream: Philox4x32-10(counter = (u, t, attempt / 4, epoch), key = seed), j = (word * n_item) >> 32 --
 * a NEW stream (the reference's sequential Mersenne Twister cannot be reproduced in parallel), restated bit-exactly
 * in oracle/sampling.py.  poi_neg_intervals = fun_compute_dist_neg (:165-180): out[u][t] = cal_dis interval between
 * q[u][t] and p[u][t-1] (fp64 haversine, coords [n_item x 2] = lat, lon) for 1 <= t < lens[u], dist_num elsewhere. */
int poi_sample_negatives(poi_engine* e, const int32_t* rows_dev, int32_t lrow, const int32_t* sorted_a_dev, int32_t la,
                         const int32_t* sorted_b_dev, int32_t lb, int32_t n_user, int32_t n_item, uint64_t seed,
                         uint32_t epoch, int32_t* out_dev);
int poi_neg_intervals(poi_engine* e, const int32_t* p_dev, const int32_t* q_dev, const int32_t* lens_dev, int32_t n_user,
                      int32_t lmax, const double* coords_dev, double dd, int32_t dist_num, int32_t* out_dev);

/* ---- BPR-MF: OboBpr.bpr_train (BPR.py:234-241) / Bpr.bpr_train (BPR.py:389-397) --------- */
/* n sequential (u, p, q) SGD steps in the order given -- exactly n back-to-back
 * `model.train(uidx, [p, q])` calls (prog_bpr_gru_spatial.py:240-244); per-occurrence
 * gradient, last writer wins (BPR.py:228-230).  loss_host[n] = -log sigmoid(u_i). */
int poi_bpr_train_seq(poi_engine* e, float* ux_dev, float* lt_dev, int32_t d,
                      const int32_t* u_host, const int32_t* p_host, const int32_t* q_host,
                      int64_t n, float alpha, float lambda, double* loss_host);
/* one mini-batch step with duplicate-summed unique-row updates (Bpr, BPR.py:351-397) */
int poi_bpr_train_batch(poi_engine* e, float* ux_dev, int64_t n_user, float* lt_dev,
                        int64_t n_rows_lt, int32_t d,
                        const int32_t* p_host, const int32_t* q_host, const int32_t* mask_host,
                        const int32_t* u_host, int64_t n, float alpha, float lambda,
                        double* loss_host);

/* ---- PRME: OboPrme.prme_train (PRME.py:212-219) ----------------------------------------- */
/* n sequential (u, [p, q, prev], dist_km, gap) ASCENT steps, = n back-to-back
 * `model.train(uidx, [p, q, prev], dist, gap)` calls (prog_prme.py:191-197).
 * loss_host[n] = log sigmoid(Dq - Dp). */
int poi_prme_train_seq(poi_engine* e, float* du_dev, float* dp_dev, float* ds_dev, int32_t d,
                       const int32_t* u_host, const int32_t* p_host, const int32_t* q_host,
                       const int32_t* prev_host, const double* dist_host, const int32_t* gap_host,
                       int64_t n, int32_t threshold, double component_weight,
                       float alpha, float lambda, double* loss_host);

/* K negatives per positive (BASELINE.json C3 "PRME ... neg=20").  The reference draws one negative (PRME.py:172-219);
 * this is the driver-defined generalisation of SURVEY.md 8(a6): pqidx = [p, q_1..q_K, prev], upq = sum_k log
 * sigmoid(D(q_k) - D(p)), L2 over every gathered row; q_host is [n x K] row-major.
 *  poi_prme_train_seq_k   : n sequential ascent steps (parity mode; K = 1 is exactly poi_prme_train_seq), rows written
 *                           back in pqidx order, last occurrence wins.  loss_host[n] = upq of each step.
 *  poi_prme_train_batch_k : ONE mini-batch step over n check-ins (throughput mode, EXTENSION semantics): every term from
 *                           pre-update values, gradients summed over duplicate occurrences, one step per unique row --
 *                           the reference's own mini-batch rule (Bpr, BPR.py:351-397) applied to PRME.  The six index
 *                           arrays are device resident (on_host = 0) or host memory copied inside the call (on_host = 1,
 *                           the end-to-end path); dist is float km, gap int32 minutes.  *loss_sum_host = sum of upq. */
int poi_prme_train_seq_k(poi_engine* e, float* du_dev, float* dp_dev, float* ds_dev, int32_t d,
                         const int32_t* u_host, const int32_t* p_host, const int32_t* q_host,
                         const int32_t* prev_host, const double* dist_host, const int32_t* gap_host,
                         int64_t n, int32_t K, int32_t threshold, double component_weight,
                         float alpha, float lambda, double* loss_host);
int poi_prme_train_batch_k(poi_engine* e, float* du_dev, int64_t n_user, float* dp_dev, float* ds_dev, int64_t n_rows,
                           int32_t d, const int32_t* u, const int32_t* p, const int32_t* q, const int32_t* prev,
                           const float* dist, const int32_t* gap, int64_t n, int32_t K, int32_t on_host,
                           int32_t threshold, double component_weight, float alpha, float lambda,
                           double* loss_sum_host);

/* ---- GeoIE: GeoIE.seq_train (GeoIE.py:185-194) ------------------------------------------ */
typedef struct {
    float*  g;  float* h;  float* z;   /* [(n_item+1) x H]  (GeoIE.py:65-72) */
    float*  t;                         /* [n_user x H]                        */
    double* ab;                        /* dev double[2] = {a, b}  (GeoIE.py:74-78) */
    int64_t n_rows;                    /* n_item + 1 */
    int32_t H;
} poi_geoie_params;
/* p_row_host/q_row_host: the user's rows of tra_buys_masks / tra_buys_neg_masks [lmax];
 * dist_pos/dist_neg (float) and msk (int32) are the (n x n) host matrices the driver
 * passes (prog_geoie.py:179-183).  out_host[0] = sum log sigmoid(sp - sq). */
int poi_geoie_train(poi_engine* e, const poi_geoie_params* params, int32_t uidx,
                    const int32_t* p_row_host, const int32_t* q_row_host, int32_t lmax,
                    const float* dist_pos_host, const float* dist_neg_host,
                    const int32_t* msk_host, int32_t n, float alpha, float lambda,
                    double* out_host);

/* K negatives per target, ONE mini-batch step over Bu users (BASELINE.json C4 "GeoIE ... neg=100"; throughput mode,
 * EXTENSION semantics -- the reference trains one user per call with one negative; mini-batch rule as Bpr, BPR.py:351-397):
 * every term from pre-update values, g / h / z rows duplicate-summed over the batch, a and b dense SGD, t untouched (its
 * gradient is exactly zero).  P [Bu x L] POI sequences WITHOUT padding (2 <= L <= 33), Q [Bu x L x K] negatives (position 0
 * unused), device resident (on_host = 0) or host memory copied inside the call (on_host = 1); coords_dev float
 * [n_rows x 4] = lat, lon (degrees), cos(lat), 0: the pairwise distances the reference's driver precomputes
 * (Load_Data_GeoIE.py:143-156) are recomputed on the fly.  *loss_host = sum log sigmoid(sp - sq). */
int poi_geoie_train_batch_k(poi_engine* e, const poi_geoie_params* params, const int32_t* P, const int32_t* Q,
                            const float* coords_dev, int32_t Bu, int32_t L, int32_t K, int32_t on_host,
                            float alpha, float lambda, double* loss_host);

/* ---- row-sharded multi-GPU step of the pairwise models (csrc/mf_mg.cuh; BASELINE.json C4 "GeoIE ... 2 x B200 row-sharded").
 * Same peer-memory protocol as poi_gru_step_mg.  Key set 0 = the g occurrences (history), key set 1 = the h / z occurrences
 * (candidates); table 0 = g (set 0), 1 = h, 2 = z (set 1).  Per rank, peer-visible (poi_peer_alloc): the three shards, per key
 * set ob_ids / ob_perm [cap[set]] and ob_meta [world + 2], per table ob_grads [cap[set of the table] x H], sums double[4],
 * flags uint32 [2 * world] (zero before step 1); slot_tab[set]: own int32 [n_local_rows x world], all -1.
 * cap[0] >= Bu (L - 1), cap[1] >= Bu (L - 1)(K + 1).  params->g/h/z are ignored (the shards come from `peers`), params->n_rows
 * is the GLOBAL row count, params->ab the replicated scalars.  P, Q hold GLOBAL row ids of this rank's users. */
typedef struct {
    int32_t   world, rank;
    int64_t   cap[2];
    int64_t   n_local_rows;
    float*    shard[3][POI_MG_MAX_RANKS];
    int32_t*  ob_ids[2][POI_MG_MAX_RANKS];
    int32_t*  ob_perm[2][POI_MG_MAX_RANKS];
    int32_t*  ob_meta[2][POI_MG_MAX_RANKS];
    float*    ob_grads[3][POI_MG_MAX_RANKS];
    double*   sums[POI_MG_MAX_RANKS];
    uint32_t* flags[POI_MG_MAX_RANKS];
    int32_t*  slot_tab[2];
} poi_mf_peers;
int poi_geoie_step_mg(poi_engine* e, const poi_geoie_params* params, const int32_t* P, const int32_t* Q, const float* coords_dev,
                      int32_t Bu, int32_t L, int32_t K, int32_t on_host, const poi_mf_peers* peers, int64_t step,
                      float alpha, float lambda, double* loss_host);

/* ---- evaluation helpers (scoring + top-K) ---------------------------- */
/* scores[b, i] = users[b,:] . items[i,:] (+ wd * prob[b, i] if prob_dev != NULL), then the
 * indices of the top_k largest scores per row in descending order -- the product of
 * compute_sub_all_scores (GRU.py:93-96, GRU_Spatial.py:117-125) fused with
 * Valuate.py:91-100,133-146.  items excludes the pad row.  topk_dev [B x top_k] int32. */
int poi_score_topk(poi_engine* e, const float* users_dev, int32_t B, const float* items_dev,
                   int64_t n_item, int32_t H, const float* prob_dev, float wd,
                   int32_t top_k, int32_t* topk_dev);

/* Distance2Pre scoring + top-K WITHOUT the U x I `prob` / `ulptai` matrices (GRU_Spatial.py:77-78,117-125,
 * Load_Data_by_length.py:183-235, Valuate.py:133-146): score[b, i] = users[b].items[i] + wd * sts[b, iv] * [iv < dist_num]
 * where iv = cal_dis interval (Load_Data_by_length.py:24-42, fp64) between user b's last training POI (user_coords
 * [B x 2] = lat, lon) and item i (item_coords [n_item x 2]) is computed on the fly; sts [B x n_dist_rows] are the
 * interval distributions `predict` returns. */
int poi_score_topk_geo(poi_engine* e, const float* users_dev, int32_t B, const float* items_dev, int64_t n_item, int32_t H,
                       const float* sts_dev, int32_t n_dist_rows, const double* user_coords_dev, const double* item_coords_dev,
                       double dd, int32_t dist_num, float wd, int32_t top_k, int32_t* topk_dev);

#ifdef __cplusplus
}
#endif
#endif /* POI_ENGINE_H_ */
