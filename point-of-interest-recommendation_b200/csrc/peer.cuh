// peer.cuh -- the multi-GPU exchange over NVLink peer memory, no collective on the data path.
//
// The item table is row-sharded (owner = row % world, local row = row / world).  Every rank allocates its shard
// and its gradient "outbox" with poi_peer_alloc (plain cudaMalloc -> exportable with cudaIpcGetMemHandle), the
// handles travel once through torch.distributed, and every rank maps every peer's buffers (poi_peer_open).
// After that the two exchanges of a training step are device code reading peer memory through NVLink/NVSwitch:
//
//   k_gather_rows_sharded  rows of the batch's unique ids straight out of the OWNERS' shards (replaces
//                          all-to-all(ids) -> owner gather -> all-to-all(rows));
//   k_pull_segments        the owner copies, from every peer's outbox, the (id, gradient row, count) records
//                          addressed to it (replaces all-to-all(gradient rows) and all-to-all(counts)); the
//                          outbox is the peer's own result buffers as the backward pass wrote them plus a
//                          permutation list that groups the record numbers by owner -- nothing is re-packed;
//                          records land grouped by source rank, ascending id inside a rank -> the duplicate sum
//                          of poi_gru_apply_mg runs in the same fixed order as before (deterministic).
//
// Ordering between ranks comes from the two small NCCL all-reduces the step needs anyway (dense gradients; the
// step-start barrier), see dist.py.
#pragma once
#include <stdlib.h>
#ifndef POI_PEER_GATHER_DEFAULT
#define POI_PEER_GATHER_DEFAULT 0
#endif
#include "common.cuh"
#include "sort.cuh"
#include "gemm_tc.cuh"

constexpr int POI_MAX_PEERS = 16;

struct PeerTable {                       // by value in kernel arguments
    const float* shard[POI_MAX_PEERS];   // peer r's shard of the item table [rows_r x d]
    int world;
};

template <int LPR, int UNR>
__global__ void __launch_bounds__(256)
k_gather_rows_sharded(PeerTable pt, int dim4, const int32_t* __restrict__ ids, int64_t n_idx, float* __restrict__ out) {
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
                const int32_t id = ids[r0 + u];
                src[u] = reinterpret_cast<const float4*>(pt.shard[id % pt.world]) + (int64_t)(id / pt.world) * dim4;
            }
        }
        for (int c = lane; c < dim4; c += LPR) {
            float4 v[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u)          // all peer loads of the group in flight before the first store
                if (src[u]) v[u] = src[u][c];
#pragma unroll
            for (int u = 0; u < UNR; ++u)
                if (src[u]) out4[(r0 + u) * dim4 + c] = v[u];
        }
    }
}

// The same gather with the copy engines instead of the load/store units: one thread per warp drives a ring of PG_ST row
// buffers in shared memory -- cp.async.bulk peer/global -> shared (completion on an mbarrier), then cp.async.bulk shared ->
// global -- so that a warp has PG_ST - 1 whole rows in flight over NVLink whatever its register budget.
// ids: uint32 row ids; n_dev (device scalar) overrides n_host when given.
constexpr int PG_WARPS = 8, PG_ST = 4;
__global__ void __launch_bounds__(PG_WARPS * 32)
k_gather_rows_sharded_bulk(PeerTable pt, uint32_t row_bytes, const uint32_t* __restrict__ ids, const uint32_t* __restrict__ n_dev,
                           int64_t n_host, float* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t pg_smem[];          // [PG_WARPS][PG_ST][row_bytes]
    __shared__ uint64_t bar[PG_WARPS][PG_ST];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane != 0) return;
    const int64_t n = n_dev ? (int64_t)*n_dev : n_host;
    for (int s2 = 0; s2 < PG_ST; ++s2) tc::mbar_init(&bar[w][s2], 1);
    tc::fence_barrier_init();
    const int64_t gw = (int64_t)blockIdx.x * PG_WARPS + w, nw = (int64_t)gridDim.x * PG_WARPS;
    const uint32_t sm0 = tc::smem_u32(pg_smem) + (uint32_t)(w * PG_ST) * row_bytes;
    const uint32_t W = (uint32_t)pt.world;
    auto issue_load = [&](int st, uint32_t id) {
        const char* src = reinterpret_cast<const char*>(pt.shard[id % W]) + (size_t)(id / W) * row_bytes;
        const uint32_t barp = tc::smem_u32(&bar[w][st]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barp), "r"(row_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(sm0 + (uint32_t)st * row_bytes), "l"(src), "r"(row_bytes), "r"(barp) : "memory");
    };
    // prologue: PG_ST - 1 rows in flight, the id of the next one in a register
    for (int s2 = 0; s2 < PG_ST - 1; ++s2) { const int64_t r = gw + s2 * nw; if (r < n) issue_load(s2, ids[r]); }
    int64_t rn = gw + (PG_ST - 1) * nw;                         // next row to request
    uint32_t idn = rn < n ? ids[rn] : 0u;
    int it = 0;
    for (int64_t r = gw; r < n; r += nw, ++it) {
        const int st = it % PG_ST;
        tc::mbar_wait(&bar[w][st], (uint32_t)(it / PG_ST) & 1u);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(reinterpret_cast<char*>(out) + (size_t)r * row_bytes), "r"(sm0 + (uint32_t)st * row_bytes), "r"(row_bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // the buffer stored one iteration ago is free once every group but the newest has been read out of shared memory
        if (rn < n) {
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            issue_load((it + PG_ST - 1) % PG_ST, idn);
            rn += nw;
            idn = rn < n ? ids[rn] : 0u;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static inline int peer_gather_mode() {
    static int mode = -1;
    if (mode < 0) { const char* v = getenv("POI_PEER_GATHER"); mode = (v && v[0] == '1') ? 1 : ((v && v[0] == '0') ? 0 : POI_PEER_GATHER_DEFAULT); }
    return mode;
}
static int launch_gather_bulk(poi_engine* e, const PeerTable& pt, int dim, const uint32_t* ids, const uint32_t* n_dev, int64_t n_host,
                              int64_t n_max, float* out) {
    const uint32_t row_bytes = (uint32_t)dim * 4u;
    const size_t smem = (size_t)PG_WARPS * PG_ST * row_bytes;
    static size_t attr = 0;
    if (smem > attr) {
        POI_CK(e, cudaFuncSetAttribute(k_gather_rows_sharded_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, ((size_t)200 << 10) / std::max<size_t>(smem, 1)));
    unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(poi_cdiv(n_max, PG_WARPS), (int64_t)e->num_sms * per_sm));
    POI_LAUNCH(e, k_gather_rows_sharded_bulk, grid, PG_WARPS * 32, smem, pt, row_bytes, ids, n_dev, n_host, out);
    return 0;
}

struct PullTable {                        // one entry per source rank, in rank order
    const int32_t* perm[POI_MAX_PEERS];   // record numbers grouped by owner (ascending id inside a group)
    const int32_t* ids[POI_MAX_PEERS];    // outbox: sorted unique global row ids of the peer's batch
    const float* grads[POI_MAX_PEERS];    // [n x d] duplicate-summed gradient rows, same order as ids
    const float* cnts[POI_MAX_PEERS];     // [n] occurrence counts
    int64_t src_off[POI_MAX_PEERS];       // first entry of this rank's group in peer r's perm list
    int64_t dst_off[POI_MAX_PEERS + 1];   // where peer r's records go in the receive buffers (prefix sums)
    int world;
};

__global__ void __launch_bounds__(256)
k_pull_segments(PullTable pt, int dim4, int32_t* __restrict__ recv_local_ids, float* __restrict__ recv_grads,
                float* __restrict__ recv_cnts) {
    constexpr int LPR = 32;
    const int lane = threadIdx.x % LPR;
    const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    const int64_t n_groups = (int64_t)gridDim.x * blockDim.x / LPR;
    const int64_t n = pt.dst_off[pt.world];
    float4* out4 = reinterpret_cast<float4*>(recv_grads);
    for (int64_t i = group; i < n; i += n_groups) {
        int r = 0;
        while (r + 1 < pt.world && i >= pt.dst_off[r + 1]) ++r;
        const int64_t s = pt.perm[r][pt.src_off[r] + (i - pt.dst_off[r])];
        const float4* g = reinterpret_cast<const float4*>(pt.grads[r]) + s * dim4;
        for (int c = lane; c < dim4; c += LPR) out4[i * dim4 + c] = g[c];
        if (lane == 0) {
            recv_local_ids[i] = pt.ids[r][s] / pt.world;
            recv_cnts[i] = pt.cnts[r][s];
        }
    }
}

// ---- outbox permutation: record numbers grouped by owner (= one stable radix pass on id % world) ----
__global__ void k_owner_keys(const int32_t* __restrict__ ids, int64_t n, int world, uint32_t* __restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = (uint32_t)(ids[i] % world);
}
// perm = the sorted occurrence ids; counts[o] = number of keys equal to o (binary search in the sorted keys)
__global__ void k_owner_perm_counts(const uint32_t* __restrict__ keys_sorted, const uint32_t* __restrict__ vals_sorted, int64_t n,
                                    int world, int32_t* __restrict__ perm, double* __restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] = (int32_t)vals_sorted[i];
    if (blockIdx.x == 0 && threadIdx.x < world) {
        auto lower = [&](uint32_t key) { int64_t lo = 0, hi = n; while (lo < hi) { int64_t m = (lo + hi) >> 1; if (keys_sorted[m] < key) lo = m + 1; else hi = m; } return lo; };
        counts[threadIdx.x] = (double)(lower(threadIdx.x + 1) - lower(threadIdx.x));
    }
}
