// common.cuh -- engine object, workspace arena, launch accounting, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <unordered_map>
#include <algorithm>

#include "../../include/poi_engine.h"

#define POI_WARP 32

struct PoiChunk { char* ptr; size_t cap; };

// kernel categories for the built-in per-launch profiler (bench.py's roofline numbers)
enum PoiCat { CAT_OTHER = 0, CAT_INDEX = 1, CAT_GATHER = 2, CAT_GEMM = 3, CAT_WGRAD = 4, CAT_LOSS = 5,
              CAT_ELTWISE = 6, CAT_ROWS = 7, CAT_MF = 8, CAT_GEOIE = 9, CAT_EVAL = 10, CAT_REDUCE = 11,
              CAT_RECUR_FWD = 12, CAT_RECUR_BWD = 13, POI_NCAT = POI_KPROF_NCAT };
struct ProfRec { cudaEvent_t a, b; int cat; double flops, bytes; };

struct poi_engine {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int gemm_mode = 1;               // default: tcgen05 3xTF32 (fp32-faithful); 0 = fp32 FMA, 2 = 1xTF32
    bool persistent_gemm = true;     // large tensor-core GEMMs: persistent CTAs, epilogue overlapped with the next tile
    bool wgrad_mn = true;            // tensor-core weight gradients read the activations as they lie (MN-major UMMA operands, no transposes)
    bool small_batch_path = true;    // B <= 8: SIMT recurrence kernels with Wh resident in shared memory (gru_small.cuh)
    int fused_cluster = 0;           // CTAs per 128 users in the fused recurrence: 0 = auto, else 1 / 2 / 4
    bool fuse_recurrence = true;
    bool gemm_csplit = true;         // per-tile tensor-core GEMMs with few tiles and long K: the two K halves as a 2-CTA cluster (POI_GEMM_CSPLIT=0 disables)
    bool fused_sort = true;          // n > 4096 keys: radix passes + segment arrays in one persistent launch (sort.cuh)
    uint32_t* grid_bar = nullptr;    // {arrivals, generation} of that kernel's grid barrier; zero between launches     // tensor-core modes: forward recurrence as one persistent fused kernel (gru_fused.cuh)
    // bump arena (device scratch owned by the engine); reset at the start of every call
    std::vector<PoiChunk> chunks;
    size_t cur_chunk = 0, cur_off = 0, high_water = 0, call_bytes = 0;
    // pinned host staging
    double*  h_out = nullptr;        // 64 doubles
    char*    h_stage = nullptr;      // pinned staging for small host->device index uploads
    size_t   h_stage_cap = 0;
    // per-launch profiler: CUDA events around every kernel on the engine stream
    bool kprof = false;
    int cur_cat = CAT_OTHER;
    int gemm_cat = -1;               // >= 0: GEMM launches are booked under this category (per-step recurrence GEMMs)
    double cur_flops = 0.0, cur_bytes = 0.0;
    std::vector<ProfRec> recs;
    size_t nrec = 0;
    double cat_ms[POI_NCAT] = {}, cat_flops[POI_NCAT] = {}, cat_bytes[POI_NCAT] = {};
    int64_t cat_launches[POI_NCAT] = {};
    // multi-GPU: index prep + segments of poi_gru_mg_prepare kept alive (same arena epoch) for poi_gru_train_mg
    bool prep_valid = false;
    int prep_B = 0, prep_lmax = 0;
    void* prep_state = nullptr;      // MgPrep*, owned
    // CUDA-graph replay of small-batch train calls (the reference's one-by-one mode): one instantiated graph per
    // (parameter pointers, batch size, max length, modes); valid while the arena has not been re-allocated
    struct GraphEntry { cudaGraphExec_t exec = nullptr; uint64_t arena_gen = 0; int warm = 0; int64_t n_launch = 0;
                        std::vector<uint64_t> key; uint64_t last_use = 0; };
    uint64_t graph_clock = 0;
    std::unordered_map<uint64_t, GraphEntry> graphs;
    cudaStream_t cap_stream = nullptr;
    bool capturing = false;
    int graph_mode = 1;              // 1 = replay train calls with B <= 8 as CUDA graphs, 0 = always launch kernel by kernel
    uint64_t arena_gen = 0;          // bumped whenever the arena is (re)allocated
    int64_t graph_replays = 0;
    // phase timing
    bool timing = false;
    cudaEvent_t ev[9] = {};
    float phase_ms[8] = {};
};

static thread_local std::string g_create_err;

#define POI_FAIL(e, ...)                                                         \
    do {                                                                         \
        char _b[512];                                                            \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                                   \
        (e)->err = std::string(_b) + " [" + __FILE__ + ":" + std::to_string(__LINE__) + "]"; \
        return -1;                                                               \
    } while (0)

#define POI_CK(e, call)                                                          \
    do {                                                                         \
        cudaError_t _s = (call);                                                 \
        if (_s != cudaSuccess) POI_FAIL(e, "CUDA error %s: %s", #call, cudaGetErrorString(_s)); \
    } while (0)

#define POI_TRY(call)                                                            \
    do {                                                                         \
        int _r = (call);                                                         \
        if (_r != 0) return _r;                                                  \
    } while (0)

static inline ProfRec* prof_begin(poi_engine* e);

// every kernel launch goes through here so that poi_launch_count is exact; `cat`/work set by
// POI_CAT apply to the launches that follow
#define POI_CAT(e, c, fl, by) do { (e)->cur_cat = ((c) == CAT_GEMM && (e)->gemm_cat >= 0) ? (e)->gemm_cat : (c); (e)->cur_flops = (double)(fl); (e)->cur_bytes = (double)(by); } while (0)
#define POI_LAUNCH(e, kern, grid, block, smem, ...)                              \
    do {                                                                         \
        ProfRec* _pr = (e)->kprof ? prof_begin(e) : nullptr;                     \
        kern<<<(grid), (block), (smem), (e)->stream>>>(__VA_ARGS__);             \
        if (_pr) cudaEventRecord(_pr->b, (e)->stream);                           \
        (e)->cur_flops = 0.0; (e)->cur_bytes = 0.0;                              \
        (e)->launches++;                                                         \
        cudaError_t _s = cudaPeekAtLastError();                                  \
        if (_s != cudaSuccess) POI_FAIL(e, "launch %s failed: %s", #kern, cudaGetErrorString(_s)); \
    } while (0)

static inline ProfRec* prof_begin(poi_engine* e) {
    if (e->nrec == e->recs.size()) {
        ProfRec r; cudaEventCreate(&r.a); cudaEventCreate(&r.b); r.cat = 0; r.flops = r.bytes = 0;
        e->recs.push_back(r);
    }
    ProfRec* r = &e->recs[e->nrec++];
    r->cat = e->cur_cat; r->flops = e->cur_flops; r->bytes = e->cur_bytes;
    cudaEventRecord(r->a, e->stream);
    return r;
}
// call after the stream has been synchronised
static inline void prof_harvest(poi_engine* e) {
    for (size_t i = 0; i < e->nrec; ++i) {
        ProfRec& r = e->recs[i];
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            e->cat_ms[r.cat] += ms; e->cat_flops[r.cat] += r.flops; e->cat_bytes[r.cat] += r.bytes;
            e->cat_launches[r.cat]++;
        }
    }
    e->nrec = 0;
}

static inline size_t poi_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t poi_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// arena: bump allocation out of a list of cudaMalloc'ed chunks.  When a call needed more than
// one chunk, the next reset consolidates into a single chunk of the high-water size, so the
// steady state is allocation-free.
// ---------------------------------------------------------------------------------------------
static int arena_reset(poi_engine* e) {
    if (e->chunks.size() > 1) {
        POI_CK(e, cudaStreamSynchronize(e->stream));
        size_t total = 0;
        for (auto& c : e->chunks) { total += c.cap; cudaFree(c.ptr); }
        e->chunks.clear();
        total = poi_align_up(std::max(total, e->high_water) + (total >> 3), (size_t)1 << 20);
        char* p = nullptr;
        POI_CK(e, cudaMalloc(&p, total));
        e->chunks.push_back({p, total});
        e->arena_gen++;
    }
    e->cur_chunk = 0; e->cur_off = 0; e->call_bytes = 0;
    return 0;
}

static int arena_alloc(poi_engine* e, size_t bytes, void** out) {
    bytes = poi_align_up(std::max<size_t>(bytes, 16), 256);
    while (true) {
        if (e->cur_chunk < e->chunks.size()) {
            PoiChunk& c = e->chunks[e->cur_chunk];
            if (e->cur_off + bytes <= c.cap) {
                *out = c.ptr + e->cur_off;
                e->cur_off += bytes;
                e->call_bytes += bytes;
                e->high_water = std::max(e->high_water, e->call_bytes);
                return 0;
            }
            e->cur_chunk++; e->cur_off = 0;
            continue;
        }
        size_t cap = poi_align_up(std::max(bytes, (size_t)64 << 20), (size_t)1 << 20);
        char* p = nullptr;
        if (e->capturing) POI_FAIL(e, "arena would grow during graph capture");
        POI_CK(e, cudaMalloc(&p, cap));
        e->chunks.push_back({p, cap});
        e->arena_gen++;
    }
}

template <typename T>
static int arena_get(poi_engine* e, size_t n, T** out) {
    void* p = nullptr;
    POI_TRY(arena_alloc(e, n * sizeof(T), &p));
    *out = reinterpret_cast<T*>(p);
    return 0;
}

static int stage_reserve(poi_engine* e, size_t bytes) {
    if (bytes <= e->h_stage_cap) return 0;
    POI_CK(e, cudaStreamSynchronize(e->stream));
    if (e->h_stage) cudaFreeHost(e->h_stage);
    e->h_stage = nullptr; e->h_stage_cap = 0;
    size_t cap = poi_align_up(bytes + (bytes >> 2), 4096);
    POI_CK(e, cudaMallocHost((void**)&e->h_stage, cap));
    e->h_stage_cap = cap;
    e->arena_gen++;          // captured graphs copy from the old staging buffer: invalidate them
    return 0;
}

static inline void phase_mark(poi_engine* e, int i) {
    if (e->timing) cudaEventRecord(e->ev[i], e->stream);
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// log(sigmoid(x)) = -softplus(-x), the form Theano rewrites to
__device__ __forceinline__ float logsigmoidf_(float x) {
    return -(fmaxf(-x, 0.0f) + log1pf(expf(-fabsf(x))));
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// read-only data path: lets the compiler batch the loads of several epilogue calls ahead of their stores
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 f4fma(float s, float4 a, float4 c) { return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w)); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
