// C-ABI implementation of the B200-native exact simplex engine (include/relp_gpu.h).
// Host orchestration only: every arithmetic step runs in the kernels of kernels.cuh.
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <vector>
#include <unordered_map>

#include <dlfcn.h>
#include <nccl.h>   // types only: the library is opened lazily so that single-GPU use has no NCCL dependency

#include "../../include/relp_gpu.h"
#include "../../include/relp_gpu_test.h"
#include "kernels.cuh"

// ------------------------------------------------------------------------------------------------
// NCCL, resolved at run time from whichever libnccl.so.2 the process already holds (torch's) or the
// system one.  Collectives used per pivot: all-gather of the ratio-test
// candidates, sum all-reduce of the staged pivot row (one non-zero contributor), all-gather of the
// work-vector partial sums.
// ------------------------------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return nullptr;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
    api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.AllReduce) return nullptr;
    api.handle = h;
    return &api;
}
#define NK(call)                                                                              \
    do {                                                                                      \
        ncclResult_t r__ = (call);                                                            \
        if (r__ != ncclSuccess) {                                                             \
            ctx->err = std::string(#call) + ": " + (nccl_api()->GetErrorString ? nccl_api()->GetErrorString(r__) : "nccl error"); \
            return RG_ERR_NCCL;                                                               \
        }                                                                                     \
    } while (0)

using namespace rg;

#define CK(call)                                                                        \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);             \
            return RG_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

#define RG_TRY(call)                     \
    do {                                 \
        int r__ = (call);                \
        if (r__ != RG_OK) return r__;    \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
// The limb-width ladder (RG_NWIDTHS entries, relp_gpu.h): powers of two up to 8 limbs, then steps of two limbs --
// above 512 bits a doubling would spend up to 4x the multiply-adds the numbers need (config 5 peaks at 783 bits).
static const int kWidths[RG_NWIDTHS] = {1, 2, 4, 8, 10, 12, 14, 16};
static inline int width_index(int L) { for (int k = 0; k < RG_NWIDTHS; ++k) if (kWidths[k] == L) return k; return -1; }
// pow2_only (RG_WIDTH_LADDER=pow2): the round-1 ladder 1, 2, 4, 8, 16 (tests use it to reach the 16-limb kernels early)
static inline bool width_allowed(int L, bool pow2_only) { return !pow2_only || (L & (L - 1)) == 0; }
static inline int next_width(int L, bool pow2_only = false) {
    for (int k = width_index(L) + 1; k >= 1 && k < RG_NWIDTHS; ++k) if (width_allowed(kWidths[k], pow2_only)) return kWidths[k];
    return 0;
}
static inline int prev_width(int L, bool pow2_only = false) {
    for (int k = width_index(L) - 1; k >= 0; --k) if (width_allowed(kWidths[k], pow2_only)) return kWidths[k];
    return 0;
}

#define DISPATCH_L(Lv, FN, ...)                                   \
    switch (Lv) {                                                 \
        case 1: FN<1>(__VA_ARGS__); break;                        \
        case 2: FN<2>(__VA_ARGS__); break;                        \
        case 4: FN<4>(__VA_ARGS__); break;                        \
        case 8: FN<8>(__VA_ARGS__); break;                        \
        case 10: FN<10>(__VA_ARGS__); break;                      \
        case 12: FN<12>(__VA_ARGS__); break;                      \
        case 14: FN<14>(__VA_ARGS__); break;                      \
        case 16: FN<16>(__VA_ARGS__); break;                      \
        default: break;                                           \
    }
// same, for launchers that return a status
#define DISPATCH_L_RET(Lv, FN, ...)                               \
    switch (Lv) {                                                 \
        case 1: return FN<1>(__VA_ARGS__);                        \
        case 2: return FN<2>(__VA_ARGS__);                        \
        case 4: return FN<4>(__VA_ARGS__);                        \
        case 8: return FN<8>(__VA_ARGS__);                        \
        case 10: return FN<10>(__VA_ARGS__);                      \
        case 12: return FN<12>(__VA_ARGS__);                      \
        case 14: return FN<14>(__VA_ARGS__);                      \
        default: return FN<16>(__VA_ARGS__);                      \
    }

// ------------------------------------------------------------------------------------------------
// allocation
// ------------------------------------------------------------------------------------------------
// Device memory comes from the stream-ordered pool of the device with an unlimited release threshold:
// after the first solve of a process, allocation and promotion (K9) never reach the driver allocator.
static void pool_setup(int device) {
    static bool done[64] = {false};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done[device] = true;
}
// Large buffers (>= 1 MiB) are additionally recycled by EXACT size in a per-device free list of the process: the
// repeated solves of a bench / a branch-and-bound driver allocate the same sizes in the same order, and the
// stream-ordered pool was measured to stall for ~0.8 s now and then when a 400 MB request met a fragmented pool
// (it maps fresh physical memory although enough is free).  A recycled buffer carries the event recorded on the
// stream that released it; a taker on another stream waits for it.
struct BigFree { void* p; size_t bytes; cudaEvent_t ev; cudaStream_t st; };
static std::mutex g_big_mutex;
static std::vector<BigFree> g_big_free[16];
static std::unordered_map<void*, size_t> g_big_live;       // buffers handed out through the recycler
static size_t g_big_cached[16] = {0};
static const size_t kBigMin = (size_t)1 << 20, kBigCacheMax = (size_t)48 << 30;
template <class T>
static cudaError_t dev_alloc(T** p, size_t bytes, cudaStream_t st) {
    if (bytes < kBigMin) return cudaMallocAsync((void**)p, bytes ? bytes : 8, st);
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(g_big_mutex);
        auto& fl = g_big_free[dev & 15];
        for (size_t k = fl.size(); k-- > 0;) {
            if (fl[k].bytes != bytes) continue;
            BigFree b = fl[k];
            fl.erase(fl.begin() + k);
            g_big_cached[dev & 15] -= bytes;
            if (b.st != st) cudaStreamWaitEvent(st, b.ev, 0);
            cudaEventDestroy(b.ev);
            g_big_live[b.p] = bytes;
            *p = (T*)b.p;
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMallocAsync((void**)p, bytes, st);
    if (e != cudaSuccess) {     // out of memory with buffers parked in the free list: release them and retry
        (void)cudaGetLastError();
        std::lock_guard<std::mutex> lock(g_big_mutex);
        auto& fl = g_big_free[dev & 15];
        for (auto& b : fl) { cudaStreamWaitEvent(st, b.ev, 0); cudaEventDestroy(b.ev); cudaFreeAsync(b.p, st); }
        fl.clear(); g_big_cached[dev & 15] = 0;
        cudaStreamSynchronize(st);
        e = cudaMallocAsync((void**)p, bytes, st);
    }
    if (e == cudaSuccess) { std::lock_guard<std::mutex> lock(g_big_mutex); g_big_live[(void*)*p] = bytes; }
    return e;
}
static void free_dev_on(void* p, cudaStream_t st) {
    if (!p) return;
    size_t bytes = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(g_big_mutex);
        auto it = g_big_live.find(p);
        if (it != g_big_live.end()) { bytes = it->second; g_big_live.erase(it); }
        if (bytes && g_big_cached[dev & 15] + bytes <= kBigCacheMax) {
            BigFree b{p, bytes, nullptr, st};
            if (cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventRecord(b.ev, st) == cudaSuccess) {
                g_big_free[dev & 15].push_back(b);
                g_big_cached[dev & 15] += bytes;
                return;
            }
            (void)cudaGetLastError();
        }
    }
    cudaFreeAsync(p, st);
}
static int alloc_carry(rg_context* ctx, bool list_mode);

// geometry of the tensor-core dense dots for a vector of LV limbs over this rank's dense columns
struct DenseGeom { int ncols, ntc, zs, ks, rps, rpitch; size_t rstride_k; };
static DenseGeom dense_geom(rg_context* ctx, int LV) {
    DenseGeom g;
    g.ncols = ctx->d1 - ctx->d0;
    const int nt_max = LV + 1;                         // n-tiles of 8 slice rows: ceil((8 LV + 1) / 8)
    g.ntc = nt_max <= 20 ? nt_max : (nt_max + 1) / 2;  // per CTA (accumulator registers)
    g.zs = cdiv(nt_max, g.ntc);
    const int mp = (int)ctx->dmp;
    int ks = 296 / std::max(1, cdiv(std::max(g.ncols, 1), 128) * g.zs);
    ks = std::max(1, std::min(8, ks));
    ks = std::max(ks, cdiv(mp, 32768));
    g.rps = (cdiv(mp, ks) + 127) / 128 * 128;        // whole 128-row K blocks (tcgen05 path) = whole 64-row chunks
    g.ks = cdiv(mp, g.rps);
    g.rpitch = std::max(g.ncols, 1);
    g.rstride_k = (size_t)nt_max * 8 * g.rpitch;
    return g;
}
// TMA descriptor of a row-major u8 matrix [outer][inner] (row pitch `pitch` bytes), box 128 x box_outer, 128-byte
// swizzle.  cuTensorMapEncodeTiled is resolved through the runtime (no link-time dependency on libcuda).
static bool make_map_u8(CUtensorMap* map, const void* base, size_t inner, size_t outer, size_t pitch, int box_outer) {
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (Fn)p;
        (void)cudaGetLastError();
    }
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {128u, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1u, 1u};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int alloc_dense_scratch(rg_context* ctx, int L) {
    if (ctx->nd <= 0) return RG_OK;
    ctx->dmp = ((size_t)ctx->m + 63) / 64 * 64;
    size_t words = 0;
    for (int LV : {L, LU_of(L), LU_of(L) + 1, LW_of(L)}) {
        DenseGeom g = dense_geom(ctx, LV);
        words = std::max(words, g.rstride_k * g.ks);
    }
    CK(dev_alloc(&ctx->dR, sizeof(int) * words, ctx->stream));
    ctx->dR_words = words;
    const size_t sl_rows = std::max<size_t>((size_t)(LW_of(L) + 1) * 8, UM_N), sl2_rows = std::max<size_t>((size_t)(L + 1) * 8, UM_N);
    CK(dev_alloc(&ctx->dSl, sl_rows * ctx->dmp, ctx->stream));
    CK(cudaMemsetAsync(ctx->dSl, 0, sl_rows * ctx->dmp, ctx->stream));
    ctx->dSl_primary = ctx->dSl;
    CK(dev_alloc(&ctx->dchunk, sizeof(int) * (ctx->dmp / 64), ctx->stream));
    {   // a second, smaller set for the pricing dot (L-limb cost row): it runs on the main stream while the
        // steepest-edge dots and the weight update still occupy the first set on their side stream
        DenseGeom g = dense_geom(ctx, L);
        CK(dev_alloc(&ctx->dR2, sizeof(int) * g.rstride_k * g.ks, ctx->stream));
        CK(dev_alloc(&ctx->dSl2, sl2_rows * ctx->dmp, ctx->stream));
        CK(cudaMemsetAsync(ctx->dSl2, 0, sl2_rows * ctx->dmp, ctx->stream));
        CK(dev_alloc(&ctx->dchunk2, sizeof(int) * (ctx->dmp / 64), ctx->stream));
    }
    // tcgen05 path (dense_umma.cuh): worthwhile from a few 128 x 128 tiles on; RG_NO_UMMA=1 keeps the mma.sync kernel
    const bool no_umma = getenv("RG_NO_UMMA") != nullptr;
    ctx->umma_ok = !no_umma && ctx->m >= 512 && ctx->nd >= 256 && ctx->ldc >= 128 &&
                   make_map_u8(&ctx->mapA, ctx->Acm, ctx->ldc, (size_t)ctx->nd, ctx->ldc, UM_M) &&
                   make_map_u8(&ctx->mapB, ctx->dSl, ctx->dmp, sl_rows, ctx->dmp, UM_N) &&
                   make_map_u8(&ctx->mapB2, ctx->dSl2, ctx->dmp, sl2_rows, ctx->dmp, UM_N);
    if (ctx->umma_ok) {
        static bool attr_done = false;
        if (!attr_done) {
            attr_done = cudaFuncSetAttribute(k_dense_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UM_SMEM) == cudaSuccess;
            if (!attr_done) { ctx->umma_ok = false; (void)cudaGetLastError(); }
        }
    }
    return RG_OK;
}
static int alloc_width_buffers(rg_context* ctx, int L) {
    const size_t ld = ctx->ld;
    const size_t n = ctx->n;
    CK(dev_alloc(&ctx->u, sizeof(u64) * LU_of(L) * ld, ctx->stream));
    CK(dev_alloc(&ctx->rowp, sizeof(u64) * L * ld, ctx->stream));
    CK(dev_alloc(&ctx->omega, sizeof(u64) * LW_of(L) * ld, ctx->stream));
    {
        // two partial slabs per row chunk (rows with a positive / a negative factor, see k_colsum1)
        size_t dense_slots = 2 * (size_t)ctx->work_chunks * ld;
        size_t list_slots = 2 * (size_t)ctx->list_chunks * ctx->list_pcols;
        CK(dev_alloc(&ctx->omega_part, sizeof(u64) * LW_of(L) * std::max(dense_slots, list_slots), ctx->stream));
    }
    CK(dev_alloc(&ctx->tmprow, sizeof(u64) * LU_of(L) * ld, ctx->stream));
    CK(dev_alloc(&ctx->bn, sizeof(u32) * (size_t)(ctx->nloc + 2) * (2 * (L + 8) + 1), ctx->stream));
    CK(dev_alloc(&ctx->us2, sizeof(u64) * (LU_of(L) + 1) * ld, ctx->stream));
    if (ctx->world > 1) {
        CK(dev_alloc(&ctx->ufull, sizeof(u64) * (LU_of(L) + 1) * ld, ctx->stream));
        CK(cudaMemsetAsync(ctx->ufull, 0, sizeof(u64) * (LU_of(L) + 1) * ld, ctx->stream));
    }
    if (ctx->nd > 0) RG_TRY(alloc_dense_scratch(ctx, L));
    CK(dev_alloc(&ctx->kappa, sizeof(u64) * LU_of(L) * n, ctx->stream));
    CK(dev_alloc(&ctx->nu, sizeof(u64) * LU_of(L) * n, ctx->stream));
    CK(dev_alloc(&ctx->sigma, sizeof(u64) * LS_of(L) * n, ctx->stream));
    if (ctx->nd > 0 || true) CK(dev_alloc(&ctx->tau, sizeof(u64) * (LU_of(L) + 3) * n, ctx->stream));
    if (ctx->world > 1) {   // exchange buffers sized once per width for every collective of the engine
        size_t words = std::max<size_t>((size_t)LW_of(L) * ld, (size_t)LG_of(L) * n);
        words = std::max<size_t>(words, (size_t)(LU_of(L) + 1) * (((size_t)ctx->m + ctx->world - 1) / ctx->world));
        words = std::max<size_t>(words, 128) * ctx->world;
        CK(dev_alloc(&ctx->xsend, words * sizeof(u64), ctx->stream));
        CK(dev_alloc(&ctx->xrecv, words * sizeof(u64), ctx->stream));
        ctx->xbytes = words * sizeof(u64);
    }
    CK(cudaMemsetAsync(ctx->u, 0, sizeof(u64) * LU_of(L) * ld, ctx->stream));
    CK(cudaMemsetAsync(ctx->rowp, 0, sizeof(u64) * L * ld, ctx->stream));
    CK(cudaMemsetAsync(ctx->kappa, 0, sizeof(u64) * LU_of(L) * n, ctx->stream));
    ctx->kappa_valid = false;
    return RG_OK;
}
static void free_width_buffers(rg_context* ctx) {
    free_dev_on(ctx->u, ctx->stream); free_dev_on(ctx->rowp, ctx->stream); free_dev_on(ctx->omega, ctx->stream); free_dev_on(ctx->omega_part, ctx->stream);
    free_dev_on(ctx->bn, ctx->stream); ctx->bn = nullptr;
    free_dev_on(ctx->tmprow, ctx->stream); free_dev_on(ctx->us2, ctx->stream); free_dev_on(ctx->ufull, ctx->stream); ctx->ufull = nullptr; free_dev_on(ctx->dR, ctx->stream); free_dev_on(ctx->dSl, ctx->stream); free_dev_on(ctx->dchunk, ctx->stream);
    free_dev_on(ctx->dR2, ctx->stream); free_dev_on(ctx->dSl2, ctx->stream); free_dev_on(ctx->dchunk2, ctx->stream);
    ctx->dR2 = nullptr; ctx->dSl2 = nullptr; ctx->dchunk2 = nullptr; free_dev_on(ctx->kappa, ctx->stream); free_dev_on(ctx->nu, ctx->stream); free_dev_on(ctx->sigma, ctx->stream); free_dev_on(ctx->tau, ctx->stream); ctx->tau = nullptr;
    free_dev_on(ctx->xsend, ctx->stream); free_dev_on(ctx->xrecv, ctx->stream);
    ctx->xsend = ctx->xrecv = nullptr; ctx->xbytes = 0;
    ctx->u = ctx->rowp = ctx->omega = ctx->omega_part = ctx->tmprow = ctx->us2 = nullptr; ctx->dR = nullptr; ctx->dSl = nullptr; ctx->dchunk = nullptr;
    ctx->kappa = ctx->nu = ctx->sigma = nullptr;
}

static double g_graph_prof[5] = {0, 0, 0, 0, 0};   // capture s, launch s, sync s, captures, launches
static std::mutex g_hm_mutex;
static std::vector<HostMirror*> g_hm_free[16];   // per device: recycled pinned mirrors
struct ProfEvents { cudaEvent_t e[10]; };
static std::vector<ProfEvents> g_prof_events[16];   // per device: recycled profiling event sets
static void drop_graphs(rg_context* ctx);
static void drop_graphs(rg_context* ctx) {
    for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
    ctx->graphs.clear();
}

extern "C" int rg_create(const rg_options* opts, rg_context** out) {
    if (!out) return RG_ERR_ARG;
    rg_context* ctx = new rg_context();
    ctx->device = opts ? opts->device : 0;
    ctx->rank = opts ? opts->rank : 0;
    ctx->world = (opts && opts->world > 0) ? opts->world : 1;
    if (ctx->world > 16) { delete ctx; return RG_ERR_ARG; }      // the candidate merges compare rank pairs in one block
    ctx->dense_carry_opt = opts ? opts->dense_carry : 0;
    ctx->use_graphs = getenv("RG_NO_GRAPH") == nullptr;
    { const char* e = getenv("RG_GRAPH_NCCL"); ctx->graph_nccl = e && atoi(e) != 0; }
    ctx->k1_items_prefetch = getenv("RG_K1_NOPF") == nullptr;
    ctx->kappa_recur = getenv("RG_NO_KAPPA_RECUR") == nullptr;
    ctx->ftran_overlap = getenv("RG_NO_FTRAN_OVERLAP") == nullptr;
    { const char* e = getenv("RG_K1_ITEMS_MINL"); if (e) ctx->k1_items_min_limbs = atoi(e); }
    { const char* e = getenv("RG_K1_ITEMS_ROWS"); if (e) ctx->k1_items_rows = std::min(32, std::max(1, atoi(e))); }
    { const char* e = getenv("RG_WIDTH_LADDER"); ctx->pow2_only = e && strcmp(e, "pow2") == 0; }
    { const char* e = getenv("RG_DEMOTE_MARGIN"); if (e) ctx->demote_margin = std::max(2, atoi(e)); }
    { const char* e = getenv("RG_DEMOTE_FLOOR"); if (e) ctx->demote_floor = std::max(1, atoi(e)); }   // 99 = never demote
    int L = (opts && opts->initial_limbs) ? opts->initial_limbs : 2;
    if (width_index(L) < 0) { delete ctx; return RG_ERR_ARG; }
    ctx->L = L;
    *out = ctx;
    CK(cudaSetDevice(ctx->device));
    pool_setup(ctx->device);
    // the main stream carries the critical path (work vector, K1): it gets the greatest priority so that its blocks
    // are placed before those of the side streams' dots and recurrences when both have blocks pending
    int prio_lo = 0, prio_hi = 0;
    const bool use_prio = getenv("RG_NO_PRIO") == nullptr &&
                          cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) == cudaSuccess && prio_lo != prio_hi;
    if (!use_prio) { prio_lo = prio_hi = 0; (void)cudaGetLastError(); }
    CK(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi));   // every copy is issued on this stream
    CK(cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, prio_lo));
    CK(cudaEventCreateWithFlags(&ctx->ev_side0, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_side1, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_side2, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_work, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_side3, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_nu, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_ft0, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_ft1, cudaEventDisableTiming));
    CK(cudaStreamCreateWithPriority(&ctx->side2, cudaStreamNonBlocking, prio_lo));
    CK(cudaStreamCreateWithPriority(&ctx->side3, cudaStreamNonBlocking, prio_lo));
    CK(dev_alloc(&ctx->sc, sizeof(Scalars), ctx->stream));
    CK(cudaMemsetAsync(ctx->sc, 0, sizeof(Scalars), ctx->stream));
    {   // pinned mirror: cudaHostAlloc / cudaFreeHost synchronise the whole device, so mirrors are recycled
        std::lock_guard<std::mutex> lock(g_hm_mutex);
        if (!g_hm_free[ctx->device & 15].empty()) {
            ctx->hm = g_hm_free[ctx->device & 15].back();
            g_hm_free[ctx->device & 15].pop_back();
        }
    }
    if (!ctx->hm) CK(cudaHostAlloc(&ctx->hm, sizeof(HostMirror), cudaHostAllocMapped));
    memset(ctx->hm, 0, sizeof(HostMirror));
    CK(cudaHostGetDevicePointer((void**)&ctx->hm_dev, ctx->hm, 0));
    if (ctx->world > 1) {
        if (ctx->rank < 0 || ctx->rank >= ctx->world || !opts->nccl_unique_id) {
            ctx->err = "rg_create: world > 1 needs a valid rank and nccl_unique_id";
            return RG_ERR_ARG;
        }
        NcclApi* api = nccl_api();
        if (!api) { ctx->err = "rg_create: libnccl.so.2 not found"; return RG_ERR_NCCL; }
        // One communicator per (process, world, rank), created by the first context and reused by
        // later ones: NCCL connects its channels lazily (~0.5 s on the first collective).  Every
        // rank takes the same branch because every rank issues the same call sequence.
        // The cache is keyed on (world, rank, device) only: a NEW unique id with the same key means the same peer
        // group re-creating its contexts (every rank takes the same branch, so no rank waits in CommInitRank
        // alone).  A process that talks to DIFFERENT peer groups must set RG_NCCL_NO_CACHE=1 (one communicator
        // per context, destroyed with it).  Guarded by a mutex: contexts may be created from several threads.
        static std::mutex comm_mutex;
        std::lock_guard<std::mutex> comm_lock(comm_mutex);
        static ncclComm_t cached = nullptr;
        static int cached_world = 0, cached_rank = -1, cached_dev = -1;
        const bool no_cache = getenv("RG_NCCL_NO_CACHE") != nullptr;
        if (no_cache || !cached || cached_world != ctx->world || cached_rank != ctx->rank || cached_dev != ctx->device) {
            ncclUniqueId id;
            memcpy(&id, opts->nccl_unique_id, sizeof(id));
            ncclComm_t comm;
            NK(api->CommInitRank(&comm, ctx->world, id, ctx->rank));
            if (no_cache) ctx->own_comm = true;
            else { cached = comm; cached_world = ctx->world; cached_rank = ctx->rank; cached_dev = ctx->device; }
            // warm the all-gather and all-reduce paths (connection setup) outside any solve
            u64* w = nullptr;
            CK(dev_alloc(&w, sizeof(u64) * 64 * (ctx->world + 1), ctx->stream));
            CK(cudaMemsetAsync(w, 0, sizeof(u64) * 64 * (ctx->world + 1), ctx->stream));
            NK(api->AllGather(w, w + 64, 64, ncclUint64, comm, ctx->stream));
            NK(api->AllReduce(w, w, 64, ncclUint64, ncclSum, comm, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            free_dev_on(w, ctx->stream);
            ctx->nccl_comm = comm;
        } else {
            ctx->nccl_comm = cached;
        }
    }
    return RG_OK;
}

extern "C" int rg_nccl_unique_id(void* out, int32_t bytes) {
    NcclApi* api = nccl_api();
    if (!api || !out || bytes < (int32_t)sizeof(ncclUniqueId)) return RG_ERR_NCCL;
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return RG_ERR_NCCL;
    memcpy(out, &id, sizeof(id));
    return (int)sizeof(id);
}

extern "C" int rg_destroy(rg_context* ctx) {
    if (!ctx) return RG_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    free_width_buffers(ctx);
    free_dev_on(ctx->carry, ctx->stream); free_dev_on(ctx->pk, ctx->stream); free_dev_on(ctx->A.colptr, ctx->stream); free_dev_on(ctx->A.rowidx, ctx->stream); free_dev_on(ctx->A.vals, ctx->stream);
    free_dev_on(ctx->cost, ctx->stream); free_dev_on(ctx->rhs, ctx->stream); free_dev_on(ctx->basis, ctx->stream); free_dev_on(ctx->inbasis, ctx->stream);
    free_dev_on(ctx->G, ctx->stream); free_dev_on(ctx->cand, ctx->stream); free_dev_on(ctx->score, ctx->stream); free_dev_on(ctx->sc, ctx->stream); free_dev_on(ctx->svec, ctx->stream);
    free_dev_on(ctx->triv, ctx->stream); free_dev_on(ctx->klist, ctx->stream); free_dev_on(ctx->nzrows, ctx->stream); free_dev_on(ctx->kpos, ctx->stream); free_dev_on(ctx->aq, ctx->stream); free_dev_on(ctx->row0_part, ctx->stream);
    free_dev_on(ctx->Acm, ctx->stream);
    free_dev_on(ctx->wf, ctx->stream); free_dev_on(ctx->wcol, ctx->stream); free_dev_on(ctx->artf, ctx->stream);
    free_dev_on(ctx->artcost, ctx->stream); free_dev_on(ctx->rowf, ctx->stream);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    drop_graphs(ctx);
    if (getenv("RG_HOSTPROF"))
        fprintf(stderr, "[hostprof] graphs: %.0f captures %.3f s, %.0f launches %.3f s enqueue + %.3f s sync (cumulative)\n",
                g_graph_prof[3], g_graph_prof[0], g_graph_prof[4], g_graph_prof[1], g_graph_prof[2]);
    if (ctx->hm) {
        std::lock_guard<std::mutex> lock(g_hm_mutex);
        g_hm_free[ctx->device & 15].push_back(ctx->hm);
    }
    if (ctx->ev0) {   // back to the per-device pool (see rg_set_profile)
        ProfEvents pe; pe.e[0] = ctx->ev0; pe.e[1] = ctx->ev1;
        for (int k = 0; k < 8; ++k) pe.e[2 + k] = ctx->evp[k];
        std::lock_guard<std::mutex> lock(g_hm_mutex);
        g_prof_events[ctx->device & 15].push_back(pe);
    }
    if (ctx->evt0) { cudaEventDestroy(ctx->evt0); cudaEventDestroy(ctx->evt1); }
    if (ctx->ev_side0) { cudaEventDestroy(ctx->ev_side0); cudaEventDestroy(ctx->ev_side1); cudaEventDestroy(ctx->ev_side2); cudaEventDestroy(ctx->ev_work); cudaEventDestroy(ctx->ev_side3); cudaEventDestroy(ctx->ev_nu); cudaEventDestroy(ctx->ev_ft0); cudaEventDestroy(ctx->ev_ft1); }
    if (ctx->side2) cudaStreamDestroy(ctx->side2);
    if (ctx->side3) cudaStreamDestroy(ctx->side3);
    // the communicator is process-cached (see rg_create) unless this context owns it
    if (ctx->own_comm && ctx->nccl_comm && nccl_api()) nccl_api()->CommDestroy((ncclComm_t)ctx->nccl_comm);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return RG_OK;
}

extern "C" int64_t rg_release_cached_memory(int32_t device) {
    if (device < 0 || device >= 16) return RG_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { (void)cudaGetLastError(); return RG_ERR_CUDA; }
    std::vector<BigFree> parked;
    {
        std::lock_guard<std::mutex> lock(g_big_mutex);
        parked.swap(g_big_free[device]);
        g_big_cached[device] = 0;
    }
    int64_t bytes = 0;
    for (auto& b : parked) {
        cudaEventSynchronize(b.ev);        // the stream that released it may be gone; the event is not
        cudaEventDestroy(b.ev);
        cudaFree(b.p);
        bytes += (int64_t)b.bytes;
    }
    (void)cudaGetLastError();
    return bytes;
}

extern "C" const char* rg_last_error(const rg_context* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// The partition of the row-sharded engine, pure host arithmetic shared by rg_load_csc /
// rg_load_dense_i8 and by the Python side (relp_b200/sharding.py calls these, so the gloo tests cover the very
// functions the device path uses): rank r owns the block [r*q, min(count, (r+1)*q)), q = ceil(count / world).
extern "C" int rg_shard_block(int32_t count, int32_t world, int32_t rank, int32_t* first, int32_t* number) {
    if (count < 0 || world < 1 || rank < 0 || rank >= world || !first || !number) return RG_ERR_ARG;
    const int q = (count + world - 1) / world;
    *first = std::min(count, rank * q);
    *number = std::max(0, std::min(count, (rank + 1) * q) - *first);
    return RG_OK;
}

extern "C" int rg_load_csc(rg_context* ctx, int32_t m, int32_t n, const int64_t* colptr,
                           const int32_t* rowidx, const int64_t* vals) {
    if (!ctx || m <= 0 || n <= 0 || !colptr) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->carry) { ctx->err = "rg_load_csc: context already holds a problem"; return RG_ERR_STATE; }
    ctx->m = m; ctx->n = n;
    ctx->ld = ((m + 1 + 15) / 16) * 16;
    {   // block row partition of the carry; block column partition of the pricing
        int32_t first = 0, number = 0;
        rg_shard_block(m, ctx->world, ctx->rank, &first, &number);
        ctx->row_lo = first; ctx->nloc = number;
        ctx->d0 = ctx->d1 = 0;          // no dense block yet: all columns are CSC columns
        rg_shard_block(n, ctx->world, ctx->rank, &first, &number);
        ctx->s0 = first; ctx->s1 = first + number;
    }
    long long nnz = colptr[n];
    ctx->A.nnz = nnz;
    if (colptr[0] != 0) { ctx->err = "rg_load_csc: colptr[0] must be 0"; return RG_ERR_ARG; }
    for (long long j = 0; j < n; ++j) {
        if (colptr[j + 1] < colptr[j]) { ctx->err = "rg_load_csc: colptr must be non-decreasing"; return RG_ERR_ARG; }
        for (long long k = colptr[j]; k < colptr[j + 1]; ++k) {
            if (rowidx[k] < 0 || rowidx[k] >= m) { ctx->err = "rg_load_csc: row index out of range"; return RG_ERR_ARG; }
            if (k > colptr[j] && rowidx[k] <= rowidx[k - 1]) { ctx->err = "rg_load_csc: row indices must ascend within a column"; return RG_ERR_ARG; }
        }
    }
    ctx->h_colptr.assign(colptr, colptr + n + 1);     // host copies of the structure: unit-column validation
    ctx->h_rowidx.assign(rowidx, rowidx + nnz);
    ctx->h_vals.assign(vals, vals + nnz);
    CK(dev_alloc(&ctx->A.colptr, sizeof(long long) * (n + 1), ctx->stream));
    CK(dev_alloc(&ctx->A.rowidx, sizeof(int) * std::max<long long>(nnz, 1), ctx->stream));
    CK(dev_alloc(&ctx->A.vals, sizeof(long long) * std::max<long long>(nnz, 1), ctx->stream));
    CK(cudaMemcpyAsync(ctx->A.colptr, colptr, sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpyAsync(ctx->A.rowidx, rowidx, sizeof(int) * nnz, cudaMemcpyHostToDevice, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpyAsync(ctx->A.vals, vals, sizeof(long long) * nnz, cudaMemcpyHostToDevice, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    CK(dev_alloc(&ctx->cost, sizeof(long long) * n, ctx->stream));
    CK(cudaMemsetAsync(ctx->cost, 0, sizeof(long long) * n, ctx->stream));
    CK(dev_alloc(&ctx->rhs, sizeof(long long) * m, ctx->stream));
    CK(cudaMemsetAsync(ctx->rhs, 0, sizeof(long long) * m, ctx->stream));
    CK(dev_alloc(&ctx->basis, sizeof(int) * m, ctx->stream));
    CK(dev_alloc(&ctx->inbasis, n, ctx->stream));
    CK(dev_alloc(&ctx->cand, sizeof(int) * 1024, ctx->stream));
    CK(dev_alloc(&ctx->score, sizeof(double) * std::max(m, n), ctx->stream));
    CK(dev_alloc(&ctx->svec, sizeof(u64) * ctx->ld, ctx->stream));
    CK(dev_alloc(&ctx->triv, (size_t)ctx->ld, ctx->stream));
    CK(dev_alloc(&ctx->klist, sizeof(int) * ctx->ld, ctx->stream));
    CK(dev_alloc(&ctx->nzrows, sizeof(int) * ctx->ld, ctx->stream));
    CK(dev_alloc(&ctx->kpos, sizeof(int) * ctx->ld, ctx->stream));
    CK(dev_alloc(&ctx->aq, sizeof(long long) * m, ctx->stream));
    CK(dev_alloc(&ctx->row0_part, sizeof(u64) * 32 * (RG_MAXL + 2), ctx->stream));
    CK(dev_alloc(&ctx->wf, sizeof(long long) * n, ctx->stream));
    CK(dev_alloc(&ctx->wcol, sizeof(long long) * n, ctx->stream));
    CK(dev_alloc(&ctx->artf, sizeof(long long) * m, ctx->stream));
    CK(dev_alloc(&ctx->artcost, sizeof(long long) * m, ctx->stream));
    CK(dev_alloc(&ctx->rowf, sizeof(long long) * m, ctx->stream));
    ctx->work_chunks = std::max(1, std::min(16, cdiv(std::max(ctx->nloc, 1), 256)));
    {   // rows per work-vector chunk in list mode (a chunk = one block row of k_colsum1 = two partial slabs)
        static const int chunk_rows = [] { const char* e = getenv("RG_LIST_CHUNK_ROWS"); int v = e ? atoi(e) : 64; return v >= 16 ? v : 64; }();
        ctx->list_chunks = std::max(1, cdiv(std::max(ctx->nloc, 1), chunk_rows));
    }
    ctx->list_pcols = ((m + 1) / 3 + 2 + 127) / 128 * 128;                    // the list never exceeds (m+1)/3 + 1
    RG_TRY(alloc_carry(ctx, true));   // cost row + packed block; rg_init_identity_basis picks the real mode
    CK(dev_alloc(&ctx->G, sizeof(u64) * LG_of(ctx->L) * n, ctx->stream));
    CK(cudaMemsetAsync(ctx->G, 0, sizeof(u64) * LG_of(ctx->L) * n, ctx->stream));
    RG_TRY(alloc_width_buffers(ctx, ctx->L));
    CK(cudaStreamSynchronize(ctx->stream));
    return RG_OK;
}

extern "C" int rg_set_rhs(rg_context* ctx, const int64_t* b) {
    if (!ctx || !ctx->rhs || !b) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->rhs, b, sizeof(long long) * ctx->m, cudaMemcpyHostToDevice, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    return RG_OK;
}

extern "C" int rg_load_dense_i8(rg_context* ctx, int32_t nd, const int8_t* colmajor) {
    if (!ctx || !ctx->carry || nd <= 0 || nd > ctx->n || !colmajor) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->nd) { ctx->err = "rg_load_dense_i8: dense block already loaded"; return RG_ERR_STATE; }
    const int m = ctx->m;
    ctx->ldc = ((size_t)m + 15) / 16 * 16;
    CK(dev_alloc(&ctx->Acm, ctx->ldc * nd + 64, ctx->stream));   // +64: the last column's 64-row chunk may overhang
    CK(cudaMemsetAsync(ctx->Acm, 0, ctx->ldc * nd + 64, ctx->stream));
    CK(cudaMemcpy2DAsync(ctx->Acm, ctx->ldc, colmajor, (size_t)m, (size_t)m, (size_t)nd, cudaMemcpyHostToDevice,
                         ctx->stream));
    ctx->nd = nd;
    {   // column ownership: the dense block and the CSC columns are split separately (balanced cost)
        int32_t first = 0, number = 0;
        rg_shard_block(nd, ctx->world, ctx->rank, &first, &number);
        ctx->d0 = first; ctx->d1 = first + number;
        rg_shard_block(ctx->n - nd, ctx->world, ctx->rank, &first, &number);
        ctx->s0 = nd + first; ctx->s1 = nd + first + number;
    }
    RG_TRY(alloc_dense_scratch(ctx, ctx->L));
    CK(cudaStreamSynchronize(ctx->stream));
    return RG_OK;
}

extern "C" int rg_set_weights(rg_context* ctx, const int64_t* colfac, const int64_t* artfac,
                              const int64_t* colw, const int64_t* artcost) {
    if (!ctx || !ctx->carry || !colfac || !artfac || !colw || !artcost) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    for (int j = 0; j < ctx->n; ++j)
        if (colfac[j] <= 0 || colfac[j] >= (1ll << 31) || colw[j] <= 0 || colw[j] >= (1ll << 31)) {
            ctx->err = "rg_set_weights: weights must be in [1, 2^31)"; return RG_ERR_ARG;
        }
    for (int i = 0; i < ctx->m; ++i)
        if (artfac[i] <= 0 || artfac[i] >= (1ll << 31) || artcost[i] < 0) {
            ctx->err = "rg_set_weights: artificial factors must be in [1, 2^31), costs >= 0"; return RG_ERR_ARG;
        }
    CK(cudaMemcpyAsync(ctx->wf, colfac, sizeof(long long) * ctx->n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->wcol, colw, sizeof(long long) * ctx->n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->artf, artfac, sizeof(long long) * ctx->m, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->artcost, artcost, sizeof(long long) * ctx->m, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->weighted = true;
    return RG_OK;
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
// every launch is checked; the first failure is kept (with the kernel's name) and reported by the next
// synchronising call (launch_failed), so a bad launch configuration can never pass silently
#define LAUNCH(kernel, grid, block, ...)                                   \
    do {                                                                   \
        kernel<<<grid, block, 0, ctx->stream>>>(__VA_ARGS__);              \
        ctx->launches++;                                                   \
        cudaError_t e__ = cudaPeekAtLastError();                           \
        if (e__ != cudaSuccess && ctx->launch_err.empty())                 \
            ctx->launch_err = std::string("launch of " #kernel " failed: ") + cudaGetErrorString(e__); \
    } while (0)
static int launch_failed(rg_context* ctx) {
    if (ctx->launch_err.empty()) return RG_OK;
    ctx->err = ctx->launch_err;
    ctx->launch_err.clear();
    (void)cudaGetLastError();
    return RG_ERR_CUDA;
}

// exchange buffers of the row-sharded engine (words of 8 bytes)
static int ensure_xbuf(rg_context* ctx, size_t send_words, size_t recv_words) {
    // the buffers are sized with the width buffers for every collective of the engine; growing them here is
    // a safety net only (never inside a stream capture: it synchronises and reallocates)
    size_t need = std::max(send_words, recv_words) * sizeof(u64);
    if (need <= ctx->xbytes) return RG_OK;
    if (ctx->capturing) { ctx->err = "exchange buffer too small inside a graph capture"; return RG_ERR_STATE; }
    CK(cudaStreamSynchronize(ctx->stream));
    free_dev_on(ctx->xsend, ctx->stream); free_dev_on(ctx->xrecv, ctx->stream);
    ctx->xsend = ctx->xrecv = nullptr;
    CK(dev_alloc(&ctx->xsend, need, ctx->stream));
    CK(dev_alloc(&ctx->xrecv, need, ctx->stream));
    ctx->xbytes = need;
    drop_graphs(ctx);   // captured pointers are stale
    return RG_OK;
}
static double g_nccl_host_s = 0;
static long long g_nccl_calls = 0;
static int all_gather(rg_context* ctx, const void* send, void* recv, size_t words_per_rank) {
    static const bool hostprof = getenv("RG_HOSTPROF") != nullptr;
    timespec a, b;
    if (hostprof) clock_gettime(CLOCK_MONOTONIC, &a);
    NK(nccl_api()->AllGather(send, recv, words_per_rank, ncclUint64, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    if (hostprof) {
        clock_gettime(CLOCK_MONOTONIC, &b);
        double dt = (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
        g_nccl_host_s += dt;
        if (dt > 0.001) fprintf(stderr, "[hostprof rank %d] all_gather call %lld words %zu took %.3f ms\n", ctx->rank, g_nccl_calls, words_per_rank, dt * 1e3);
        if (++g_nccl_calls % 400 == 0)
            fprintf(stderr, "[hostprof rank %d] %lld all_gather calls, %.3f s on the host\n", ctx->rank, g_nccl_calls, g_nccl_host_s);
    }
    (void)cudaGetLastError();   // NCCL probes pointers internally; benign failures must not look like ours
    return RG_OK;
}

// profiling events: inside a stream capture they must be external event-record nodes
static inline void rec_event(rg_context* ctx, cudaEvent_t ev) {
    cudaEventRecordWithFlags(ev, ctx->stream, ctx->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
}
static inline const int* klist_of(rg_context* ctx) { return ctx->list_mode ? ctx->klist : nullptr; }
static inline const unsigned char* triv_of(rg_context* ctx) { return ctx->list_mode ? ctx->triv : nullptr; }
static inline const int* kpos_of(rg_context* ctx) { return ctx->list_mode ? ctx->kpos : nullptr; }
// rows 1..nloc of the carry in the current mode: the packed active block (list mode) or the dense carry
struct BlockView { u64* base; size_t ps; int stride; };
static inline BlockView block_of(rg_context* ctx) {
    if (ctx->list_mode) return BlockView{ctx->pk, ctx->pplane, ctx->cap};
    return BlockView{ctx->carry, ctx->plane, ctx->ld};
}
// (re)allocates the carry for a mode at the current limb width, zero-filled.  List mode: the cost row
// (L planes x ld) plus the packed active block; dense mode: the full (nloc+1) x ld carry.
static int alloc_carry(rg_context* ctx, bool list_mode) {
    free_dev_on(ctx->carry, ctx->stream); ctx->carry = nullptr;
    free_dev_on(ctx->pk, ctx->stream); ctx->pk = nullptr;
    ctx->list_mode = list_mode;
    if (list_mode) {
        ctx->plane = (size_t)ctx->ld;
        ctx->cap = ctx->m >= 1024 ? 128 : 32;
        ctx->pplane = (size_t)(ctx->nloc + 1) * ctx->cap;
        CK(dev_alloc(&ctx->pk, sizeof(u64) * ctx->L * ctx->pplane, ctx->stream));
        CK(cudaMemsetAsync(ctx->pk, 0, sizeof(u64) * ctx->L * ctx->pplane, ctx->stream));
    } else {
        ctx->plane = (size_t)(ctx->nloc + 1) * ctx->ld;
        ctx->cap = 0; ctx->pplane = 0;
    }
    CK(dev_alloc(&ctx->carry, sizeof(u64) * ctx->L * ctx->plane, ctx->stream));
    CK(cudaMemsetAsync(ctx->carry, 0, sizeof(u64) * ctx->L * ctx->plane, ctx->stream));
    return RG_OK;
}
static void drop_graphs(rg_context* ctx);
// the list outgrew the packed block: double its capacity (2D copy, new slots zero)
static int grow_packed(rg_context* ctx, int need) {
    if (!ctx->list_mode || need <= ctx->cap) return RG_OK;
    int cap2 = ctx->cap;
    while (cap2 < need) cap2 *= 2;
    size_t pplane2 = (size_t)(ctx->nloc + 1) * cap2;
    u64* np = nullptr;
    CK(dev_alloc(&np, sizeof(u64) * ctx->L * pplane2, ctx->stream));
    CK(cudaMemsetAsync(np, 0, sizeof(u64) * ctx->L * pplane2, ctx->stream));
    CK(cudaMemcpy2DAsync(np, sizeof(u64) * cap2, ctx->pk, sizeof(u64) * ctx->cap, sizeof(u64) * ctx->cap,
                         (size_t)ctx->L * (ctx->nloc + 1), cudaMemcpyDeviceToDevice, ctx->stream));
    free_dev_on(ctx->pk, ctx->stream);
    ctx->pk = np; ctx->cap = cap2; ctx->pplane = pplane2;
    drop_graphs(ctx);   // captured pointers are stale
    return RG_OK;
}

static int sync_mirror(rg_context* ctx) {
    LAUNCH(k_mirror, 1, 1, ctx->sc, ctx->hm_dev, ctx->L);
    CK(cudaStreamSynchronize(ctx->stream));
    RG_TRY(launch_failed(ctx));
    CK(cudaGetLastError());
    ctx->t_cur = ctx->hm->t_next;
    ctx->nk_host = ctx->hm->nk;
    return RG_OK;
}

static void set_status(rg_context* ctx, int st) { LAUNCH(k_set_status, 1, 1, ctx->sc, st); }

// out_j = cmul cost_j D + vec[1..m] . a_j for every provider column: dense block + CSC remainder.
// `bits` points at the device-side bit-length maximum that bounds every entry of `vec`.
// mask (dense block only): 0 none, 1 keep the entries of trivial carry columns, 2 keep the listed ones
template <int LV, int LO>
static void launch_coldots(rg_context* ctx, const u64* vec, size_t vs, int cmul, u64* out, const int* bits,
                           int mask = 0, bool csc_part = true) {
    const int jd0 = ctx->d0, jd1 = ctx->d1;      // dense block slice
    const int j0 = ctx->s0, j1 = ctx->s1;        // CSC slice
    if (jd1 > jd0) {
        // tensor-core path: byte slices of the vector x int8 block (exact s32 accumulation), then recombination
        constexpr int NT_MAX = LV + 1;
        constexpr int NTC = NT_MAX <= 20 ? NT_MAX : (NT_MAX + 1) / 2;
        const DenseGeom g = dense_geom(ctx, LV);
        LAUNCH((k_dense_slices<LV>), (unsigned)(ctx->dmp / 64), 64, vec, vs, ctx->m, bits, ctx->dSl, ctx->dmp,
               ctx->dchunk, ctx->sc, mask ? (const unsigned char*)ctx->triv : (const unsigned char*)nullptr,
               mask == 1 ? 1 : 0);
        if (ctx->umma_ok) {
            // tcgen05: 128 columns x 160 slice rows per CTA, TMA-fed, accumulators in TMEM
            dim3 ugrid(cdiv(g.ncols, UM_M), g.ks, cdiv(NT_MAX * 8, UM_N));
            const CUtensorMap& mb = ctx->dSl == ctx->dSl_primary ? ctx->mapB : ctx->mapB2;
            k_dense_umma<<<ugrid, 128, UM_SMEM, ctx->stream>>>(ctx->mapA, mb, jd0, jd1, ctx->dmp, ctx->dchunk, bits, LV,
                                                              g.rps, ctx->dR, g.rstride_k, g.rpitch, ctx->sc);
            ctx->launches++;
            cudaError_t e__ = cudaPeekAtLastError();
            if (e__ != cudaSuccess && ctx->launch_err.empty())
                ctx->launch_err = std::string("launch of k_dense_umma failed: ") + cudaGetErrorString(e__);
        } else {
            dim3 grid(cdiv(g.ncols, 128), g.ks, g.zs);
            LAUNCH((k_dense_mma<NTC>), grid, 256, ctx->Acm, ctx->ldc, jd0, jd1, ctx->dSl, ctx->dmp, ctx->dchunk, bits, LV,
                   g.rps, ctx->dR, g.rstride_k, g.rpitch, ctx->sc);
        }
        LAUNCH((k_dense_combine<LV, LO>), cdiv(g.ncols, 64), 64, ctx->dR, g.rstride_k, g.ks, g.rpitch, ctx->n, jd0,
               jd1, bits, ctx->inbasis, ctx->cost, cmul, ctx->L, out, ctx->sc);
    }
    if (j1 > j0 && csc_part)
        LAUNCH((k_coldot<LV, LO>), cdiv(j1 - j0, 64), 64, vec, vs, ctx->n, j0, j1, ctx->A.colptr,
               ctx->A.rowidx, ctx->A.vals, ctx->inbasis, ctx->cost, cmul, ctx->L, out, ctx->sc);
}
template <int L>
static void launch_price_t(rg_context* ctx) {
    // pricing uses its own tensor-core scratch set (see alloc_dense_scratch)
    if (ctx->dR2) { std::swap(ctx->dR, ctx->dR2); std::swap(ctx->dSl, ctx->dSl2); std::swap(ctx->dchunk, ctx->dchunk2); }
    launch_coldots<L, L + 2>(ctx, ctx->carry, ctx->plane, 1, ctx->kappa, &ctx->sc->maxbits_carry);
    if (ctx->dR2) { std::swap(ctx->dR, ctx->dR2); std::swap(ctx->dSl, ctx->dSl2); std::swap(ctx->dchunk, ctx->dchunk2); }
}
static void launch_price(rg_context* ctx) { DISPATCH_L(ctx->L, launch_price_t, ctx); ctx->kappa_valid = true; }

template <class Cmp>
static void launch_argbest(rg_context* ctx, int off, int count, const Cmp& cmp, int mode) {
    int nb = std::max(1, std::min(1024, cdiv(count, 256)));
    LAUNCH((k_argbest1<Cmp>), nb, 256, off, count, cmp, ctx->cand, ctx->sc);
    LAUNCH((k_argbest2<Cmp>), 1, 256, nb, cmp, ctx->cand, mode, ctx->sc);
}
// column-sharded pricing: exchange the local candidates and reduce them identically on every rank
static int merge_columns(rg_context* ctx, int use_found) {
    RG_TRY(ensure_xbuf(ctx, RG_COLCAND_WORDS, (size_t)RG_COLCAND_WORDS * ctx->world));
    LAUNCH(k_column_pack, 1, 1, ctx->kappa, LU_of(ctx->L), ctx->G, LG_of(ctx->L), ctx->n,
           ctx->weighted ? ctx->wcol : nullptr, use_found, ctx->xsend, ctx->sc);
    RG_TRY(all_gather(ctx, ctx->xsend, ctx->xrecv, RG_COLCAND_WORDS));
    LAUNCH(k_column_merge, 1, 128, ctx->xrecv, ctx->world, ctx->rule, ctx->L, ctx->n, use_found, ctx->sc);
    return RG_OK;
}

static inline ColOwn own_of(rg_context* ctx) { return ColOwn{ctx->nd, ctx->d0, ctx->d1, ctx->s0, ctx->s1}; }

static int launch_select(rg_context* ctx) {
    PriceView v{ctx->kappa, LU_of(ctx->L), ctx->n, ctx->inbasis, own_of(ctx)};
    const int off = 0, cnt = ctx->n;
    const int mode = ctx->world == 1 ? 0 : 4;
    switch (ctx->rule) {
        case RG_RULE_FIRST_PROFITABLE: launch_argbest(ctx, off, cnt, CmpFirst{v}, mode); break;
        case RG_RULE_FIRST_PROFITABLE_WITH_MEMORY: launch_argbest(ctx, off, cnt, CmpFirstMem{v, ctx->sc}, mode); break;
        case RG_RULE_DANTZIG:
            LAUNCH(k_score_columns, cdiv(std::max(cnt, 1), 64), 64, ctx->n, own_of(ctx), 2, ctx->kappa,
                   LU_of(ctx->L), ctx->G, LG_of(ctx->L), ctx->inbasis, ctx->weighted ? ctx->wcol : nullptr,
                   ctx->score, ctx->sc);
            LAUNCH((k_select_scored<CmpDantzig>), 1, 1024, off, cnt,
                   (CmpDantzig{v, ctx->weighted ? ctx->wcol : nullptr}), ctx->score, mode, ctx->sc);
            break;
        default:
            LAUNCH(k_score_columns, cdiv(std::max(cnt, 1), 64), 64, ctx->n, own_of(ctx), 3, ctx->kappa,
                   LU_of(ctx->L), ctx->G, LG_of(ctx->L), ctx->inbasis, nullptr, ctx->score, ctx->sc);
            LAUNCH((k_select_scored<CmpSteepest>), 1, 1024, off, cnt, (CmpSteepest{v, ctx->G, LG_of(ctx->L)}),
                   ctx->score, mode, ctx->sc);
            break;
    }
    if (ctx->world > 1) RG_TRY(merge_columns(ctx, 0));
    return RG_OK;
}

template <int L>
static void launch_ftran_t(rg_context* ctx, int q) {
    if (ctx->list_mode) {
        LAUNCH(k_scatter_col, cdiv(ctx->m, 256), 256, ctx->aq, ctx->m, ctx->nd, ctx->A.colptr, ctx->A.rowidx,
               ctx->A.vals, ctx->Acm, ctx->ldc, q, ctx->sc);
        LAUNCH(k_scatter_col2, cdiv(ctx->m, 256), 256, ctx->aq, ctx->nd, ctx->A.colptr, ctx->A.rowidx, ctx->A.vals,
               q, ctx->sc);
        // the cost-row dot (a few blocks and a last-block fold: latency) runs on a side stream beside the list FTRAN
        // (which streams the packed block from HBM); both only read the scattered column
        const bool overlap = ctx->ftran_overlap && ctx->side3 && ctx->ev_ft0;
        cudaStream_t main_stream = ctx->stream;
        if (overlap) {
            cudaEventRecord(ctx->ev_ft0, main_stream);
            cudaStreamWaitEvent(ctx->side3, ctx->ev_ft0, 0);
            ctx->stream = ctx->side3;
        }
        LAUNCH((k_ftran_row0<L>), std::max(1, std::min(32, cdiv(ctx->m, 1024))), 256, ctx->carry, ctx->plane, ctx->m,
               ctx->aq, ctx->cost, q, ctx->u, (size_t)ctx->ld, ctx->row0_part, ctx->sc);
        if (overlap) {
            cudaEventRecord(ctx->ev_ft1, ctx->side3);
            ctx->stream = main_stream;
        }
        LAUNCH((k_ftran_list<L>), cdiv((long long)(ctx->nloc + 1) * 32, 256), 256, ctx->pk, ctx->pplane, ctx->cap,
               ctx->nloc + 1, ctx->m, ctx->aq, ctx->klist, ctx->triv, ctx->cost, q, ctx->u, (size_t)ctx->ld,
               ctx->sc);
        if (overlap) cudaStreamWaitEvent(main_stream, ctx->ev_ft1, 0);
        return;
    }
    LAUNCH((k_ftran<L>), cdiv((long long)(ctx->nloc + 1) * 32, 256), 256, ctx->carry, ctx->plane, ctx->ld,
           ctx->nloc + 1, ctx->A.colptr, ctx->A.rowidx, ctx->A.vals, ctx->cost, q, ctx->nd, ctx->u,
           (size_t)ctx->ld, ctx->sc);
    if (ctx->nd > 0)
        LAUNCH((k_ftran_dense<L>), cdiv((long long)(ctx->nloc + 1) * 32, 256), 256, ctx->carry, ctx->plane,
               ctx->ld, ctx->nloc + 1, ctx->m, ctx->Acm, ctx->ldc, ctx->cost, q, ctx->nd, ctx->u,
               (size_t)ctx->ld, ctx->sc);
}
static void launch_ftran(rg_context* ctx, int q) { DISPATCH_L(ctx->L, launch_ftran_t, ctx, q); }

static int launch_ratio(rg_context* ctx) {
    const BlockView bv = block_of(ctx);        // b = column 0 = list position 0
    CmpRatio c{bv.base, bv.ps, bv.stride, ctx->L, ctx->u, (size_t)ctx->ld, LU_of(ctx->L), ctx->basis,
               ctx->row_lo};
    const int cnt = std::max(ctx->nloc, 1);
    LAUNCH(k_score_rows, cdiv(cnt, 256), 256, ctx->nloc, bv.base, bv.ps, bv.stride, ctx->L, ctx->u,
           (size_t)ctx->ld, LU_of(ctx->L), ctx->score, ctx->sc);
    if (ctx->world == 1) {
        LAUNCH((k_select_scored<CmpRatio>), 1, 1024, 0, ctx->nloc, c, ctx->score, 1, ctx->sc, (const u64*)ctx->u,
               (size_t)ctx->ld, LU_of(ctx->L));    // also takes the pivot element a = u[p]
        return RG_OK;
    }
    // row-sharded: local candidate -> all-gather -> identical deterministic reduction on every rank
    LAUNCH((k_select_scored<CmpRatio>), 1, 1024, 0, ctx->nloc, c, ctx->score, 3, ctx->sc);
    RG_TRY(ensure_xbuf(ctx, RG_CAND_WORDS, (size_t)RG_CAND_WORDS * ctx->world));
    LAUNCH(k_ratio_pack, 1, 1, bv.base, bv.ps, bv.stride, ctx->L, ctx->u, (size_t)ctx->ld, ctx->basis,
           ctx->xsend, ctx->sc);
    RG_TRY(all_gather(ctx, ctx->xsend, ctx->xrecv, RG_CAND_WORDS));
    LAUNCH(k_ratio_merge, 1, 128, ctx->xrecv, ctx->world, ctx->L, ctx->sc);
    return RG_OK;
}
// pivot row given (artificial removal, trait-shaped bring_into_basis): set p / pg and the pivot element
static int launch_fixed_row(rg_context* ctx, int row) {
    int local = (row >= ctx->row_lo && row < ctx->row_lo + ctx->nloc) ? row - ctx->row_lo + 1 : -1;
    LAUNCH(k_set_rows, 1, 1, ctx->sc, local, row + 1);
    if (ctx->world == 1) {
        LAUNCH(k_take_a, 1, 1, ctx->u, (size_t)ctx->ld, ctx->L, ctx->sc);
        return RG_OK;
    }
    RG_TRY(ensure_xbuf(ctx, RG_CAND_WORDS, (size_t)RG_CAND_WORDS * ctx->world));
    const BlockView bv = block_of(ctx);
    LAUNCH(k_ratio_pack, 1, 1, bv.base, bv.ps, bv.stride, ctx->L, ctx->u, (size_t)ctx->ld, ctx->basis,
           ctx->xsend, ctx->sc);
    RG_TRY(all_gather(ctx, ctx->xsend, ctx->xrecv, RG_CAND_WORDS));
    LAUNCH(k_ratio_merge, 1, 128, ctx->xrecv, ctx->world, ctx->L, ctx->sc);
    return RG_OK;
}

template <int L>
static void launch_copyrow_t(rg_context* ctx) {
    const BlockView bv = block_of(ctx);
    LAUNCH((k_copyrow<L>), cdiv(ctx->ld, 256), 256, bv.base, bv.ps, ctx->ld, bv.stride, ctx->m, triv_of(ctx),
           kpos_of(ctx), ctx->rowp, (size_t)ctx->ld, ctx->sc);
}
template <int L>
static void launch_rowbits_t(rg_context* ctx) {
    LAUNCH((k_rowbits<L>), cdiv(ctx->ld, 256), 256, ctx->rowp, (size_t)ctx->ld, ctx->ld, ctx->sc);
}
static int launch_copyrow(rg_context* ctx) {
    DISPATCH_L(ctx->L, launch_copyrow_t, ctx);
    if (ctx->world > 1)   // exact: exactly one rank contributes non-zero words
    {
        NK(nccl_api()->AllReduce(ctx->rowp, ctx->rowp, (size_t)ctx->L * ctx->ld, ncclUint64, ncclSum,
                                 (ncclComm_t)ctx->nccl_comm, ctx->stream));
        (void)cudaGetLastError();
    }
    DISPATCH_L(ctx->L, launch_rowbits_t, ctx);
    return RG_OK;
}

// column-sum geometry of the current carry mode
struct ColsumGeom { int chunks, rpc, pcols, ncols; const int* klist; const int* kpos; const unsigned char* triv; };
static ColsumGeom colsum_geom(rg_context* ctx) {
    ColsumGeom g;
    if (ctx->list_mode) {
        g.chunks = ctx->list_chunks; g.pcols = ctx->list_pcols; g.ncols = ctx->nk_grid;
        g.klist = ctx->klist; g.kpos = ctx->kpos; g.triv = ctx->triv;
    } else {
        g.chunks = ctx->work_chunks; g.pcols = ctx->ld; g.ncols = ctx->ld;
        g.klist = nullptr; g.kpos = nullptr; g.triv = nullptr;
    }
    g.rpc = cdiv(std::max(ctx->nloc, 1), g.chunks);
    return g;
}
__global__ void k_gather(u64* out, const u64* base, size_t stride, size_t idx0, size_t step, int count, int nl);
// row-sharded runs: the factor vector (pivot column u, or its weighted form) is computed per row block; the split
// sigma dot needs it on every rank, so the blocks are all-gathered into `ufull` (planar, LSRCV limbs x ld)
__global__ void k_assemble_factor(u64* __restrict__ ufull, int ld, const u64* __restrict__ recv, int q, int nl, int m) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;     // global constraint row
    if (g >= m) return;
    const u64* src = recv + ((size_t)(g / q) * q + (g % q)) * nl;
    for (int l = 0; l < nl; ++l) ufull[(size_t)l * ld + 1 + g] = src[l];
}
static int gather_factor(rg_context* ctx, const u64* src, int nl) {
    const int q = cdiv(ctx->m, ctx->world);
    const size_t words = (size_t)q * nl;
    RG_TRY(ensure_xbuf(ctx, words, words * ctx->world));
    CK(cudaMemsetAsync(ctx->xsend, 0, words * sizeof(u64), ctx->stream));
    if (ctx->nloc > 0)
        LAUNCH(k_gather, cdiv(ctx->nloc, 256), 256, ctx->xsend, src, (size_t)ctx->ld, (size_t)1, (size_t)1, ctx->nloc, nl);
    RG_TRY(all_gather(ctx, ctx->xsend, ctx->xrecv, words));
    LAUNCH(k_assemble_factor, cdiv(ctx->m, 256), 256, ctx->ufull, ctx->ld, ctx->xrecv, q, nl, ctx->m);
    return RG_OK;
}

template <int L>
static int launch_work_t(rg_context* ctx) {
    constexpr int LU = L + 2, LW = 2 * L + 5;
    const ColsumGeom g = colsum_geom(ctx);
    dim3 grid(cdiv(g.ncols, 32), g.chunks);        // k_colsum1: 32 column slots x 4 row groups per block
    const u64* src = ctx->u;
    size_t words = (size_t)LW * ctx->ld;
    if (ctx->world > 1) RG_TRY(ensure_xbuf(ctx, words, words * ctx->world));
    u64* first_out = ctx->world == 1 ? ctx->omega : ctx->xsend;
    if (ctx->weighted) {
        LAUNCH((k_scale_u<L>), cdiv(ctx->nloc + 1, 256), 256, ctx->u, (size_t)ctx->ld, ctx->nloc, ctx->rowf,
               ctx->us2, (size_t)ctx->ld, ctx->sc);
        src = ctx->us2;
    }
    // the split sigma dot of a row-sharded run reads the whole factor vector
    ctx->ufull_valid = false;
    if (ctx->world > 1 && ctx->list_mode) {
        RG_TRY(gather_factor(ctx, src, ctx->weighted ? LU + 1 : LU));
        ctx->ufull_valid = true;
    }
    if (ctx->world > 1 && ctx->ufull_valid) {
        // row-sharded split mode: the implicit columns' sums come from the all-gathered factor vector on every rank,
        // and only the LISTED columns' partial sums are exchanged (nk_grid x (2L+5) words per rank instead of the
        // whole (2L+5) x ld vector)
        const BlockView bv2 = block_of(ctx);
        const size_t pwords = (size_t)LW * ctx->nk_grid;
        RG_TRY(ensure_xbuf(ctx, pwords, pwords * ctx->world));
        if (ctx->weighted) {
            LAUNCH((k_colsum1<L, LU + 1, LW>), grid, 128, bv2.base, bv2.ps, bv2.stride, ctx->nloc, g.rpc,
                   g.klist, ctx->us2, (size_t)ctx->ld, ctx->omega_part, g.pcols, ctx->sc);
            LAUNCH((k_colsum2<LW, LU + 1, L>), cdiv(ctx->ld, 64), 64, ctx->omega_part, ctx->ld, 2 * g.chunks, 0,
                   ctx->omega, ctx->sc, g.triv, (const u64*)ctx->ufull, (size_t)ctx->ld, L, g.kpos, g.pcols, 1, 1);
        } else {
            LAUNCH((k_colsum1<L, LU, LW>), grid, 128, bv2.base, bv2.ps, bv2.stride, ctx->nloc, g.rpc,
                   g.klist, ctx->u, (size_t)ctx->ld, ctx->omega_part, g.pcols, ctx->sc);
            LAUNCH((k_colsum2<LW, LU, L>), cdiv(ctx->ld, 64), 64, ctx->omega_part, ctx->ld, 2 * g.chunks, 0,
                   ctx->omega, ctx->sc, g.triv, (const u64*)ctx->ufull, (size_t)ctx->ld, L, g.kpos, g.pcols, 1, 1);
        }
        LAUNCH((k_colsum2_list<LW>), g.ncols, 128, ctx->omega_part, 2 * g.chunks, g.pcols, g.klist, ctx->ld, 0,
               ctx->xsend, ctx->sc, ctx->nk_grid);
        RG_TRY(all_gather(ctx, ctx->xsend, ctx->xrecv, pwords));
        LAUNCH((k_list_sum_scatter<LW>), cdiv(ctx->nk_grid, 128), 128, ctx->xrecv, ctx->world, ctx->nk_grid, g.klist,
               ctx->ld, ctx->omega, ctx->sc);
        return RG_OK;
    }
    // stage 1: thread = column (dense carry) or list position (packed active block: every load coalesced),
    // rows in chunks, rows with a zero factor skipped; stage 2 sums the chunks and adds the implicit
    // trivial columns (s_k * D)
    const BlockView bv = block_of(ctx);
    if (ctx->weighted) {
        LAUNCH((k_colsum1<L, LU + 1, LW>), grid, 128, bv.base, bv.ps, bv.stride, ctx->nloc, g.rpc,
               g.klist, ctx->us2, (size_t)ctx->ld, ctx->omega_part, g.pcols, ctx->sc);
        if (ctx->list_mode)
        {
            LAUNCH((k_colsum2<LW, LU + 1, L>), cdiv(ctx->ld, 64), 64, ctx->omega_part, ctx->ld, 2 * g.chunks, 0,
                   first_out, ctx->sc, g.triv, src, (size_t)ctx->ld, L, g.kpos, g.pcols, 1);
            LAUNCH((k_colsum2_list<LW>), g.ncols, 128, ctx->omega_part, 2 * g.chunks, g.pcols, g.klist, ctx->ld, 0,
                   first_out, ctx->sc);
        }
        else
            LAUNCH((k_colsum2<LW, LU + 1>), cdiv(ctx->ld, 64), 64, ctx->omega_part, ctx->ld, 2 * g.chunks, 0,
                   first_out, ctx->sc, g.triv, src, (size_t)ctx->ld, L, g.kpos, g.pcols, 1);
    } else {
        LAUNCH((k_colsum1<L, LU, LW>), grid, 128, bv.base, bv.ps, bv.stride, ctx->nloc, g.rpc,
               g.klist, ctx->u, (size_t)ctx->ld, ctx->omega_part, g.pcols, ctx->sc);
        if (ctx->list_mode)
        {
            LAUNCH((k_colsum2<LW, LU, L>), cdiv(ctx->ld, 64), 64, ctx->omega_part, ctx->ld, 2 * g.chunks, 0,
                   first_out, ctx->sc, g.triv, src, (size_t)ctx->ld, L, g.kpos, g.pcols, 1);
            LAUNCH((k_colsum2_list<LW>), g.ncols, 128, ctx->omega_part, 2 * g.chunks, g.pcols, g.klist, ctx->ld, 0,
                   first_out, ctx->sc);
        }
        else
            LAUNCH((k_colsum2<LW, LU>), cdiv(ctx->ld, 64), 64, ctx->omega_part, ctx->ld, 2 * g.chunks, 0,
                   first_out, ctx->sc, g.triv, src, (size_t)ctx->ld, L, g.kpos, g.pcols, 1);
    }
    if (ctx->world == 1) return RG_OK;
    RG_TRY(all_gather(ctx, ctx->xsend, ctx->xrecv, words));
    LAUNCH((k_colsum2<LW>), cdiv(ctx->ld, 64), 64, ctx->xrecv, ctx->ld, ctx->world, 0, ctx->omega, ctx->sc);
    return RG_OK;
}
static int launch_work(rg_context* ctx) {
    DISPATCH_L_RET(ctx->L, launch_work_t, ctx);
}

// K1 variants: E = extra limbs (>= ceil(ctz(D)/64)).  ctz(D) grows by about one bit per structural
// column in the basis, so wide carries need wide E; the generic run-time-width kernel is the last resort.
static int pick_update_variant(int L, int E_needed) {
    static const int opts[] = {0, 1, 2, 3, 4, 6, 8};
    int emax = L == 1 ? 1 : (L == 2 ? 2 : ((L == 4 || (L > 8 && L < 16)) ? 4 : 8));
    for (int e : opts) if (e >= E_needed && e <= emax) return e;
    return -1;   // generic kernel
}
// the fixed-width variants live in k1_variants.cu, one translation unit per limb width
namespace rg {
bool k1_launch_1(rg_context*, int);
bool k1_launch_2(rg_context*, int);
bool k1_launch_4(rg_context*, int);
bool k1_launch_8(rg_context*, int);
bool k1_launch_10(rg_context*, int);
bool k1_launch_12(rg_context*, int);
bool k1_launch_14(rg_context*, int);
bool k1_launch_16(rg_context*, int);
bool k1_kappa_launch_1(rg_context*, int);
bool k1_kappa_launch_2(rg_context*, int);
bool k1_kappa_launch_4(rg_context*, int);
bool k1_kappa_launch_8(rg_context*, int);
bool k1_kappa_launch_10(rg_context*, int);
bool k1_kappa_launch_12(rg_context*, int);
bool k1_kappa_launch_14(rg_context*, int);
bool k1_kappa_launch_16(rg_context*, int);
}
// reduced costs by recurrence (k_kappa_update) on ctx->stream; false when no fixed-width variant covers E
static bool launch_kappa_update(rg_context* ctx, int E) {
    switch (ctx->L) {
        case 1: return k1_kappa_launch_1(ctx, E);
        case 2: return k1_kappa_launch_2(ctx, E);
        case 4: return k1_kappa_launch_4(ctx, E);
        case 8: return k1_kappa_launch_8(ctx, E);
        case 10: return k1_kappa_launch_10(ctx, E);
        case 12: return k1_kappa_launch_12(ctx, E);
        case 14: return k1_kappa_launch_14(ctx, E);
        default: return k1_kappa_launch_16(ctx, E);
    }
}
static void launch_update(rg_context* ctx, int E) {
    bool ok = false;
    switch (ctx->L) {
        case 1: ok = k1_launch_1(ctx, E); break;
        case 2: ok = k1_launch_2(ctx, E); break;
        case 4: ok = k1_launch_4(ctx, E); break;
        case 8: ok = k1_launch_8(ctx, E); break;
        case 10: ok = k1_launch_10(ctx, E); break;
        case 12: ok = k1_launch_12(ctx, E); break;
        case 14: ok = k1_launch_14(ctx, E); break;
        default: ok = k1_launch_16(ctx, E); break;
    }
    if (ok) return;
    dim3 g2(cdiv(ctx->ld, 128), ctx->nloc + 1);
    LAUNCH(k_update_generic, g2, 128, ctx->carry, ctx->plane, ctx->ld, ctx->nloc + 1, ctx->L, ctx->u,
           (size_t)ctx->ld, ctx->rowp, (size_t)ctx->ld, ctx->sc);
}

// The steepest-edge dots, in three pieces so that they can run beside the work vector on their own streams:
//   nu_j = rowp . a_j          (needs the staged pivot row only; tensor-core scratch set 2)
//   tau_j = s . a_j            (split mode: needs the factor vector only; scratch set 1)
//   sigma_j = omega . a_j      (needs the work vector; split mode: omega on the listed columns + D tau)
static inline bool split_sigma(rg_context* ctx) {
    return ctx->list_mode && ctx->d1 > ctx->d0 && (ctx->world == 1 || ctx->ufull_valid);
}
template <int L>
static void launch_nu_dot_t(rg_context* ctx) {
    constexpr int LU = L + 2;
    if (ctx->dR2) { std::swap(ctx->dR, ctx->dR2); std::swap(ctx->dSl, ctx->dSl2); std::swap(ctx->dchunk, ctx->dchunk2); }
    launch_coldots<L, LU>(ctx, ctx->rowp, (size_t)ctx->ld, 0, ctx->nu, &ctx->sc->maxbits_rowp);
    if (ctx->dR2) { std::swap(ctx->dR, ctx->dR2); std::swap(ctx->dSl, ctx->dSl2); std::swap(ctx->dchunk, ctx->dchunk2); }
}
template <int L>
static void launch_tau_dot_t(rg_context* ctx) {
    constexpr int LU = L + 2;
    const u64* fvec = ctx->world == 1 ? (ctx->weighted ? ctx->us2 : ctx->u) : ctx->ufull;
    if (ctx->weighted) launch_coldots<LU + 1, LU + 3>(ctx, fvec, (size_t)ctx->ld, 0, ctx->tau, &ctx->sc->maxbits_s, 1, false);
    else launch_coldots<LU, LU + 2>(ctx, fvec, (size_t)ctx->ld, 0, ctx->tau, &ctx->sc->maxbits_u, 1, false);
}
template <int L>
static void launch_sigma_dot_t(rg_context* ctx, bool tau_done) {
    constexpr int LU = L + 2, LW = LW_of(L), LS = LS_of(L);
    if (split_sigma(ctx)) {
        // split sigma dot over the dense block: the work vector on the trivial carry columns
        // is D * s_k, so the wide (2L+5 limb) vector only has to cover the LISTED columns (a few 64-row chunks: the
        // others are skipped as zero) and the long dot runs on the (L+2)-limb factor vector s -- half the slices
        if (!tau_done) launch_tau_dot_t<L>(ctx);
        launch_coldots<LW, LS>(ctx, ctx->omega, (size_t)ctx->ld, 0, ctx->sigma, &ctx->sc->maxbits_tmp, 2, true);
        if (ctx->weighted)
            LAUNCH((k_sigma_add_dtau<LU + 3, L, LS>), cdiv(ctx->d1 - ctx->d0, 128), 128, ctx->tau, ctx->n, ctx->d0, ctx->d1,
                   ctx->inbasis, ctx->sigma, ctx->sc);
        else
            LAUNCH((k_sigma_add_dtau<LU + 2, L, LS>), cdiv(ctx->d1 - ctx->d0, 128), 128, ctx->tau, ctx->n, ctx->d0, ctx->d1,
                   ctx->inbasis, ctx->sigma, ctx->sc);
        return;
    }
    launch_coldots<LW, LS>(ctx, ctx->omega, (size_t)ctx->ld, 0, ctx->sigma, &ctx->sc->maxbits_tmp);
}
template <int L>
static void launch_gamma_update_t(rg_context* ctx) {
    LAUNCH((k_gamma_update_t<L>), cdiv(std::max(own_of(ctx).count(), 1), 64), 64, ctx->n, own_of(ctx), ctx->inbasis,
           ctx->nu, ctx->sigma, ctx->G, ctx->sc);
}
// LAUNCH goes to ctx->stream: swapped to the given side stream for the duration
static void launch_nu_dot(rg_context* ctx, cudaStream_t st) {
    cudaStream_t main_stream = ctx->stream;
    ctx->stream = st;
    DISPATCH_L(ctx->L, launch_nu_dot_t, ctx);
    ctx->stream = main_stream;
}
static void launch_tau_dot(rg_context* ctx, cudaStream_t st) {
    cudaStream_t main_stream = ctx->stream;
    ctx->stream = st;
    DISPATCH_L(ctx->L, launch_tau_dot_t, ctx);
    ctx->stream = main_stream;
}
static void launch_sigma_dot(rg_context* ctx, cudaStream_t st, bool tau_done) {
    cudaStream_t main_stream = ctx->stream;
    ctx->stream = st;
    DISPATCH_L(ctx->L, launch_sigma_dot_t, ctx, tau_done);
    ctx->stream = main_stream;
}
static void launch_se_update(rg_context* ctx) {
    if (ctx->profile >= 2) rec_event(ctx, ctx->evp[5]);    // after finalize + wait for the side streams
    int E2 = (2 * ctx->t_cur + 63) / 64;
    if (E2 <= 4 && ctx->L <= 8) {
        switch (ctx->L) {
            case 1: launch_gamma_update_t<1>(ctx); break;
            case 2: launch_gamma_update_t<2>(ctx); break;
            case 4: launch_gamma_update_t<4>(ctx); break;
            default: launch_gamma_update_t<8>(ctx); break;
        }
    } else {
        static const bool one_thread = getenv("RG_GAMMA1") != nullptr;
        static const bool three_warps = getenv("RG_GAMMA3") != nullptr;
        if (!one_thread && !three_warps) {
            // a warp per column (k_gamma_update_w): each lane owns CH consecutive 32-bit output limbs of every product.
            // A captured launch is replayed while ctz(D) -- and with it the division width -- may grow: widest chunk.
            const int W32 = 2 * (LG_of(ctx->L) + std::max(E2, 4));
            const int CH = ctx->capturing ? 5 : std::max(2, cdiv(W32, 32));
            const int W32max = 32 * CH, WP = W32max + 2;
            const size_t smem = (size_t)(3 + 4 * 6) * WP * sizeof(u32);
            const int blocks = cdiv(std::max(own_of(ctx).count(), 1), 4);
            switch (CH) {
                case 2: k_gamma_update_w<2><<<blocks, 128, smem, ctx->stream>>>(ctx->n, own_of(ctx), ctx->L, ctx->inbasis, ctx->nu, ctx->sigma, ctx->G, ctx->sc); break;
                case 3: k_gamma_update_w<3><<<blocks, 128, smem, ctx->stream>>>(ctx->n, own_of(ctx), ctx->L, ctx->inbasis, ctx->nu, ctx->sigma, ctx->G, ctx->sc); break;
                case 4: k_gamma_update_w<4><<<blocks, 128, smem, ctx->stream>>>(ctx->n, own_of(ctx), ctx->L, ctx->inbasis, ctx->nu, ctx->sigma, ctx->G, ctx->sc); break;
                default: k_gamma_update_w<5><<<blocks, 128, smem, ctx->stream>>>(ctx->n, own_of(ctx), ctx->L, ctx->inbasis, ctx->nu, ctx->sigma, ctx->G, ctx->sc); break;
            }
            ctx->launches++;
        } else if (one_thread) {
            LAUNCH(k_gamma_update, cdiv(std::max(own_of(ctx).count(), 1), 128), 128, ctx->n, own_of(ctx), ctx->L,
                   ctx->inbasis, ctx->nu, ctx->sigma, ctx->G, ctx->sc);
        } else {
            // three warps per 32 columns (k_gamma_update3); shared memory: two WX-limb results per column
            // sized for any ctz(D) < 64 L (a captured launch is replayed while the division width grows)
            const int WX = LG_of(ctx->L) + std::max(2 * ctx->L, 4);
            const size_t smem = (size_t)2 * WX * 32 * sizeof(u64);
            k_gamma_update3<<<cdiv(std::max(own_of(ctx).count(), 1), 32), 96, smem, ctx->stream>>>(
                ctx->n, own_of(ctx), ctx->L, ctx->inbasis, ctx->nu, ctx->sigma, ctx->G, ctx->sc);
            ctx->launches++;
        }
    }
    if (ctx->profile >= 2) rec_event(ctx, ctx->evp[6]);    // after the recurrence
}

template <int L>
static void launch_rowdot_t(rg_context* ctx) {   // nu_j = rowp . a_j
    launch_coldots<L, L + 2>(ctx, ctx->rowp, (size_t)ctx->ld, 0, ctx->nu, &ctx->sc->maxbits_rowp);
}

// ------------------------------------------------------------------------------------------------
// K9 promotion
// ------------------------------------------------------------------------------------------------
// Re-widen the persistent state (carry, packed block, steepest-edge weights) to Lnew limbs: wider = sign / zero
// extension (K9 promotion), narrower = dropping high limbs that hold only sign bits (demotion: the caller has
// checked the tracked bit lengths).  Everything else is per-pivot scratch and is reallocated at the new width.
static int change_width(rg_context* ctx, int Lnew) {
    const int Lold = ctx->L;
    if (Lnew < 1 || Lnew > RG_MAXL || width_index(Lnew) < 0) { ctx->err = "numerators exceed 16 limbs"; return RG_ERR_OVERFLOW; }
    const int Lcopy = std::min(Lold, Lnew);
    u64* nc = nullptr;
    CK(dev_alloc(&nc, sizeof(u64) * Lnew * ctx->plane, ctx->stream));
    CK(cudaMemcpyAsync(nc, ctx->carry, sizeof(u64) * Lcopy * ctx->plane, cudaMemcpyDeviceToDevice, ctx->stream));
    if (Lnew > Lold) LAUNCH(k_sign_extend, 148 * 8, 256, nc, ctx->plane, ctx->plane, Lold, Lnew);
    u64* npk = nullptr;
    if (ctx->list_mode) {   // the packed active block re-widens the same way
        CK(dev_alloc(&npk, sizeof(u64) * Lnew * ctx->pplane, ctx->stream));
        CK(cudaMemcpyAsync(npk, ctx->pk, sizeof(u64) * Lcopy * ctx->pplane, cudaMemcpyDeviceToDevice, ctx->stream));
        if (Lnew > Lold) LAUNCH(k_sign_extend, 148 * 8, 256, npk, ctx->pplane, ctx->pplane, Lold, Lnew);
    }
    u64* ng = nullptr;
    CK(dev_alloc(&ng, sizeof(u64) * LG_of(Lnew) * ctx->n, ctx->stream));
    if (Lnew > Lold) CK(cudaMemsetAsync(ng, 0, sizeof(u64) * LG_of(Lnew) * ctx->n, ctx->stream));
    CK(cudaMemcpyAsync(ng, ctx->G, sizeof(u64) * LG_of(Lcopy) * ctx->n, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    free_dev_on(ctx->carry, ctx->stream); ctx->carry = nc;
    if (ctx->list_mode) { free_dev_on(ctx->pk, ctx->stream); ctx->pk = npk; }
    free_dev_on(ctx->G, ctx->stream); ctx->G = ng;
    free_width_buffers(ctx);
    ctx->L = Lnew;
    RG_TRY(alloc_width_buffers(ctx, Lnew));
    ctx->have_column = false;
    drop_graphs(ctx);   // buffers moved: captured pointers are stale
    return RG_OK;
}
static int promote(rg_context* ctx) {
    const int Lnew = next_width(ctx->L, ctx->pow2_only);
    if (Lnew == 0) { ctx->err = "numerators exceed 16 limbs"; return RG_ERR_OVERFLOW; }
    RG_TRY(change_width(ctx, Lnew));
    ctx->promotions++;
    return RG_OK;
}

// ------------------------------------------------------------------------------------------------
// one basis change on the device.  q < 0: use the selected column sc->q.  fixed_row < 0: ratio test.
// ------------------------------------------------------------------------------------------------
static inline double now_s() {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
// leave the active-column mode: write every implicit column out and use the dense kernels from now on
static int switch_to_dense(rg_context* ctx) {
    if (!ctx->list_mode) return RG_OK;
    // the full (nloc+1) x ld carry: cost row copied, packed columns scattered, implicit columns written out
    const size_t fplane = (size_t)(ctx->nloc + 1) * ctx->ld;
    u64* full = nullptr;
    CK(dev_alloc(&full, sizeof(u64) * ctx->L * fplane, ctx->stream));
    CK(cudaMemsetAsync(full, 0, sizeof(u64) * ctx->L * fplane, ctx->stream));
    CK(cudaMemcpy2DAsync(full, sizeof(u64) * fplane, ctx->carry, sizeof(u64) * ctx->plane, sizeof(u64) * ctx->ld,
                         (size_t)ctx->L, cudaMemcpyDeviceToDevice, ctx->stream));
    dim3 ugrid(cdiv(ctx->nk_host + 1, 128), 64);
    LAUNCH(k_unpack, ugrid, 128, full, fplane, ctx->ld, ctx->pk, ctx->pplane, ctx->cap, ctx->L, ctx->klist, ctx->sc);
    dim3 grid(cdiv(ctx->m + 1, 128), 64);
    LAUNCH(k_materialise_all, grid, 128, full, fplane, ctx->ld, ctx->L, ctx->m, ctx->triv, ctx->sc);
    LAUNCH(k_clear_trivial, cdiv(ctx->ld, 256), 256, ctx->triv, ctx->ld);
    free_dev_on(ctx->carry, ctx->stream); free_dev_on(ctx->pk, ctx->stream);
    ctx->carry = full; ctx->plane = fplane; ctx->pk = nullptr; ctx->cap = 0; ctx->pplane = 0;
    ctx->list_mode = false;
    drop_graphs(ctx);   // buffers moved
    return RG_OK;
}

// the launch sequence of one iteration (no synchronisation): pivot column, ratio test, row staging, work
// vector, scalars, K1, bookkeeping, rule update, pricing of the next iteration
static int enqueue_iteration(rg_context* ctx, int q, int fixed_row, bool want_se, bool reselect, int E) {
    const bool prof = ctx->profile >= 2, prof1 = ctx->profile >= 1;
    // steepest edge: the next reduced costs follow from the current ones and the pivot-row dots nu (k_kappa_update)
    // when the current ones are valid and a fixed-width division variant is in use; otherwise they are priced
    // from the new cost row (launch_price)
    const bool kappa_recur = want_se && reselect && ctx->kappa_valid && ctx->kappa_recur &&
                             pick_update_variant(ctx->L, E) == E;
    if (prof) rec_event(ctx, ctx->evp[0]);
    LAUNCH(k_reset_iter, 1, 1, ctx->sc);
    launch_ftran(ctx, q);
    if (fixed_row < 0) RG_TRY(launch_ratio(ctx));
    else RG_TRY(launch_fixed_row(ctx, fixed_row));
    if (ctx->list_mode) {   // the pivot row's own column stops being trivial: materialise and list it
        LAUNCH(k_materialise_pivot_column, cdiv(std::max(ctx->nloc, 1), 256), 256, ctx->pk, ctx->pplane,
               ctx->cap, ctx->L, ctx->triv, ctx->sc);
        LAUNCH(k_activate_pivot_column, 1, 1, ctx->triv, ctx->klist, ctx->kpos, ctx->sc);
    }
    RG_TRY(launch_copyrow(ctx));
    if (prof) rec_event(ctx, ctx->evp[1]);
    if (want_se) {
        // the pivot scalars need only a, D and the tracked bit lengths: k_scalars and k_scalars_se run on
        // two side streams while the main stream builds the work vector; K1 joins after k_scalars, the
        // bookkeeping after k_scalars_se
        cudaEventRecord(ctx->ev_side0, ctx->stream);
        cudaStreamWaitEvent(ctx->side, ctx->ev_side0, 0);
        cudaStreamWaitEvent(ctx->side3, ctx->ev_side0, 0);
        k_scalars<<<1, 1, 0, ctx->side>>>(ctx->L, E, ctx->sc);
        cudaEventRecord(ctx->ev_side2, ctx->side);
        k_scalars_se<<<1, 32, 0, ctx->side3>>>(ctx->L, ctx->world == 1 ? ctx->G : nullptr, ctx->n, ctx->basis, ctx->sc);
        ctx->launches += 2;
        cudaEventRecord(ctx->ev_side1, ctx->side3);
        // nu = rowp . A needs only the staged pivot row: it streams the constraint block on the first side stream
        // (tensor-core scratch set 2) while the main stream builds the work vector; in the split mode of a
        // single-GPU unweighted run tau = s . A (the factor vector is the pivot column) does the same on side2
        launch_nu_dot(ctx, ctx->side);
        cudaEventRecord(ctx->ev_nu, ctx->side);
        const bool early_tau = ctx->list_mode && ctx->d1 > ctx->d0 && ctx->world == 1 && !ctx->weighted;
        if (early_tau) {
            cudaStreamWaitEvent(ctx->side2, ctx->ev_side0, 0);
            launch_tau_dot(ctx, ctx->side2);
        }
        RG_TRY(launch_work(ctx));
        if (prof) rec_event(ctx, ctx->evp[2]);
        cudaEventRecord(ctx->ev_work, ctx->stream);
        cudaStreamWaitEvent(ctx->side2, ctx->ev_work, 0);
        launch_sigma_dot(ctx, ctx->side2, early_tau);
        // the reduced-cost and weight recurrences follow on the same side stream (they need nu, the leaving column
        // and the steepest-edge scalars of k_scalars_se, and A, Dinv of k_scalars); the main stream runs K1 and the
        // bookkeeping meanwhile and joins before the column selection
        cudaStreamWaitEvent(ctx->side2, ctx->ev_side1, 0);
        cudaStreamWaitEvent(ctx->side2, ctx->ev_nu, 0);
        {
            cudaStream_t main_stream = ctx->stream;
            ctx->stream = ctx->side2;
            if (kappa_recur) {
                cudaStreamWaitEvent(ctx->side2, ctx->ev_side2, 0);
                launch_kappa_update(ctx, E);
            }
            launch_se_update(ctx);
            ctx->stream = main_stream;
        }
        cudaEventRecord(ctx->ev_side3, ctx->side2);
        cudaStreamWaitEvent(ctx->stream, ctx->ev_side2, 0);
    } else {
        if (prof) rec_event(ctx, ctx->evp[2]);
        LAUNCH(k_scalars, 1, 1, ctx->L, E, ctx->sc);
    }
    if (prof1) rec_event(ctx, ctx->ev0);
    launch_update(ctx, E);
    if (prof1) rec_event(ctx, ctx->ev1);
    if (want_se) cudaStreamWaitEvent(ctx->stream, ctx->ev_side1, 0);     // k_finalize stores Ghat_q of k_scalars_se
    LAUNCH(k_finalize, 1, 1, ctx->basis, ctx->inbasis, ctx->L, ctx->G, ctx->n, LG_of(ctx->L),
           want_se ? 1 : 0, ctx->weighted ? ctx->wf : nullptr, ctx->weighted ? ctx->rowf : nullptr, ctx->sc,
           ctx->hm_dev);
    if (prof) rec_event(ctx, ctx->evp[3]);
    if (reselect && !kappa_recur) {
        if (want_se) cudaStreamWaitEvent(ctx->stream, ctx->ev_nu, 0);      // scratch set 2 is the nu dot's
        launch_price(ctx);
    }
    ctx->kappa_valid = reselect;       // priced for the carry this iteration leaves behind (either route), or stale
    if (want_se) cudaStreamWaitEvent(ctx->stream, ctx->ev_side3, 0);     // dots + weight recurrence done
    if (reselect) RG_TRY(launch_select(ctx));
    if (prof) rec_event(ctx, ctx->evp[4]);
    return RG_OK;
}

// Demotion: the numerators of an exact simplex run shrink again once the basis stops growing (config 5: 783 bits at
// pivot 220, 590 at pivot 345), and every kernel of the iteration costs O(L) bytes and O(L^2) multiply-adds.  After
// a pivot the replicated upper bound of the carry's bit length (the overflow prediction of that pivot, and D) is
// known on the host; when it fits the next narrower width with `demote_margin` (16) bits to spare, the persistent state
// is narrowed before the next pivot.  A later overflow prediction promotes again (K9) -- the margin keeps the two
// from alternating.  Widths below `demote_floor` (default 8 limbs: pivots there are launch-latency bound) stay.
static int maybe_demote(rg_context* ctx) {
    const int need = ctx->demote_need;
    ctx->demote_need = 0;
    if (need <= 0) return RG_OK;
    int Lnew = ctx->L;
    for (int lower = prev_width(Lnew, ctx->pow2_only); lower >= ctx->demote_floor && lower >= 1 && need + ctx->demote_margin <= 64 * lower - 1;
         lower = prev_width(Lnew, ctx->pow2_only))
        Lnew = lower;
    if (Lnew == ctx->L) return RG_OK;
    RG_TRY(change_width(ctx, Lnew));
    ctx->demotions++;
    return RG_OK;
}

static int do_pivot(rg_context* ctx, int q, int fixed_row, bool want_se, bool reselect) {
    RG_TRY(maybe_demote(ctx));
    for (;;) {
        if (ctx->list_mode) {
            int E_need = (ctx->t_cur + 63) / 64;
            if (ctx->nk_host + 1 > (ctx->m + 1) / 3 || pick_update_variant(ctx->L, E_need) < 0)
                RG_TRY(switch_to_dense(ctx));
            else
                RG_TRY(grow_packed(ctx, ctx->nk_host + 2));   // this pivot may list one more column
        }
        ctx->hm->pivoted = 0;
        ctx->nk_grid = (ctx->nk_host + 2 + 127) / 128 * 128;
        int E = pick_update_variant(ctx->L, (ctx->t_cur + 63) / 64);
        if (E < 0) E = (ctx->t_cur + 63) / 64;      // generic run-time-width kernel
        {   // algorithmic work of this K1 launch: every entry of the active block (list mode:
            // local rows x listed columns incl. the one this pivot lists, plus the dense cost row; dense mode: the
            // whole local carry) is read and written once (16 L bytes) and costs two low products of
            // N = 2 (L + E) 32-bit limbs, N (N + 1) / 2 IMAD.WIDE each
            const double cols = ctx->list_mode ? (double)(ctx->nk_host + 1) : (double)(ctx->m + 1);
            const double entries = (double)ctx->nloc * cols + (double)(ctx->m + 1);
            const double N = 2.0 * (ctx->L + E);
            ctx->k1_cur_bytes = 16.0 * ctx->L * entries;
            ctx->k1_cur_imads = entries * N * (N + 1.0);
        }
        // graphs pay when an iteration is launch-latency bound; with a large dense block the kernels run for
        // milliseconds and eager launches on three streams overlap better (measured on config 5)
        const bool small = (double)ctx->m * ((double)ctx->nd + ctx->m) < 1.5e8;
        // row-sharded runs replay graphs too (NCCL collectives are capturable; every rank derives the same
        // key from replicated state) when RG_GRAPH_NCCL=1
        const bool graphable = ctx->use_graphs && small && (ctx->world == 1 || ctx->graph_nccl) && q < 0 &&
                               fixed_row < 0 && reselect;
        if (graphable) {
            // one CUDA graph per launch shape: limb width, E variant, carry mode, list grid bound, rule, flags
            long long key = (long long)ctx->L | ((long long)E << 8) | ((long long)(ctx->list_mode ? 1 : 0) << 16) |
                            ((long long)(want_se ? 1 : 0) << 17) | ((long long)(ctx->profile == 1 ? 1 : 0) << 18) | ((long long)(ctx->profile >= 2 ? 1 : 0) << 22) |
                            ((long long)(ctx->weighted ? 1 : 0) << 19) | ((long long)ctx->rule << 20) |
                            ((long long)(want_se && (2 * ctx->t_cur + 63) / 64 > 4 ? 1 : 0) << 23) |
                            ((long long)(ctx->list_mode ? ctx->nk_grid : 0) << 24);
            rg_context::GraphEntry* ge = nullptr;
            for (auto& g : ctx->graphs) if (g.key == key) ge = &g;
            if (!ge) {
                if (ctx->graphs.size() >= 8) { cudaGraphExecDestroy(ctx->graphs.front().exec); ctx->graphs.erase(ctx->graphs.begin()); }
                long long before = ctx->launches;
                double tc0 = now_s();
                cudaGraph_t graph = nullptr;
                CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
                ctx->capturing = true;
                int rc = enqueue_iteration(ctx, q, fixed_row, want_se, reselect, E);
                LAUNCH(k_mirror, 1, 1, ctx->sc, ctx->hm_dev, ctx->L);
                ctx->capturing = false;
                cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
                if (rc != RG_OK) return rc;
                if (ce != cudaSuccess) { ctx->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce); return RG_ERR_CUDA; }
                cudaGraphExec_t exec = nullptr;
                CK(cudaGraphInstantiate(&exec, graph, 0));
                cudaGraphDestroy(graph);
                ctx->graphs.push_back({key, exec, ctx->launches - before});
                ctx->launches = before;
                ge = &ctx->graphs.back();
                g_graph_prof[0] += now_s() - tc0; g_graph_prof[3] += 1;
            }
            double tl0 = now_s();
            CK(cudaGraphLaunch(ge->exec, ctx->stream));
            double tl1 = now_s();
            ctx->launches += ge->launches;
            CK(cudaStreamSynchronize(ctx->stream));
            g_graph_prof[1] += tl1 - tl0; g_graph_prof[2] += now_s() - tl1; g_graph_prof[4] += 1;
            RG_TRY(launch_failed(ctx));
            CK(cudaGetLastError());
            ctx->t_cur = ctx->hm->t_next;
            ctx->nk_host = ctx->hm->nk;
        } else {
            RG_TRY(enqueue_iteration(ctx, q, fixed_row, want_se, reselect, E));
            RG_TRY(sync_mirror(ctx));
        }
        if (ctx->hm->status == ST_PROMOTE) {
            RG_TRY(promote(ctx));
            if (q < 0) {
                // the selection lives in sc->q and survives; pricing data is rebuilt after the pivot
            }
            set_status(ctx, ST_RUN);
            continue;
        }
        if (ctx->hm->status == ST_FATAL) {
            if (ctx->hm->fatal == 3) { ctx->err = "pivot element is zero (invalid pivot row for this column)"; return RG_ERR_ARG; }
            ctx->err = ctx->hm->fatal == 2 ? "kernel variant too narrow for the denominator (internal error)"
                                           : "numerators exceed 16 limbs";
            return ctx->hm->fatal == 2 ? RG_ERR_STATE : RG_ERR_OVERFLOW;
        }
        if (ctx->hm->pivoted) {
            static const bool trace_bits = getenv("RG_TRACE_BITS") != nullptr;
            if (trace_bits && ctx->rank == 0) {
                static double t_last = 0;
                const double t_now = now_s();
                fprintf(stderr, "RGBITS pivot=%lld L=%d maxbits=%d bitsD=%d t=%d nk=%d predicted=%d dt_us=%.0f\n",
                        (long long)ctx->pivots, ctx->L, ctx->hm->maxbits_carry, ctx->hm->bits_D, ctx->hm->t_next,
                        ctx->hm->nk, ctx->hm->predicted, (t_now - t_last) * 1e6);
                t_last = t_now;
            }
            ctx->pivots++;
            ctx->pivots_at[width_index(ctx->L)]++;
            if (ctx->profile) {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) {
                    ctx->k1_ms[width_index(ctx->L)] += ms;
                    ctx->k1_launches[width_index(ctx->L)]++;
                    ctx->k1_bytes[width_index(ctx->L)] += ctx->k1_cur_bytes;
                    ctx->k1_imads[width_index(ctx->L)] += ctx->k1_cur_imads;
                    ctx->phase_ms[3] += ms;
                }
                (void)cudaGetLastError();
            }
            if (ctx->profile >= 2) {
                float ms = 0;
                cudaEvent_t seq[7] = {ctx->evp[0], ctx->evp[1], ctx->evp[2], ctx->ev0, ctx->ev1, ctx->evp[3], ctx->evp[4]};
                const int slot[6] = {0, 1, 2, -1, 4, 5};
                for (int k = 0; k < 6; ++k)
                    if (slot[k] >= 0 && cudaEventElapsedTime(&ms, seq[k], seq[k + 1]) == cudaSuccess)
                        ctx->phase_ms[slot[k]] += ms;
                if (want_se) {   // inside phase 4: [6] bookkeeping + side-stream wait, [7] nu / sigma dots
                    if (cudaEventElapsedTime(&ms, ctx->ev1, ctx->evp[5]) == cudaSuccess) ctx->phase_ms[6] += ms;
                    if (cudaEventElapsedTime(&ms, ctx->evp[5], ctx->evp[6]) == cudaSuccess) ctx->phase_ms[7] += ms;
                }
                (void)cudaGetLastError();
            }
            ctx->identity_carry = false;
            ctx->work_valid = want_se;
            ctx->demote_need = std::max(ctx->hm->predicted, ctx->hm->bits_D + 1);
        }
        return RG_OK;
    }
}

// ------------------------------------------------------------------------------------------------
// constructors
// ------------------------------------------------------------------------------------------------
extern "C" int rg_init_identity_basis(rg_context* ctx, const int32_t* basis, const int64_t* cost) {
    if (ctx) { ctx->demote_need = 0; ctx->kappa_valid = false; }   // the carry is rebuilt / re-priced: the last pivot's bit bound no longer covers it
    if (ctx) drop_graphs(ctx);   // state the captured launch sequence depends on may change
    if (!ctx || !ctx->carry || !basis) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const int m = ctx->m, n = ctx->n;
    for (int i = 0; i < m; ++i) {
        const int j = basis[i];
        if (j >= n || j < -m) { ctx->err = "rg_init_identity_basis: column id out of range"; return RG_ERR_ARG; }
        if (j < 0) continue;
        // a real basic column must be the unit column +e_i (so that B = I, D = 1): anything else needs rg_init_basis
        const bool unit = j >= ctx->nd && ctx->h_colptr[j + 1] - ctx->h_colptr[j] == 1 &&
                          ctx->h_rowidx[ctx->h_colptr[j]] == i && ctx->h_vals[ctx->h_colptr[j]] == 1;
        if (!unit) {
            ctx->err = "rg_init_identity_basis: basic column " + std::to_string(j) + " is not the unit column of row " +
                       std::to_string(i) + " (use rg_init_basis for a general basis)";
            return RG_ERR_ARG;
        }
    }
    CK(cudaMemcpyAsync(ctx->basis, basis, sizeof(int) * m, cudaMemcpyHostToDevice, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    if (cost) {
        CK(cudaMemcpyAsync(ctx->cost, cost, sizeof(long long) * n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } else {
        CK(cudaMemsetAsync(ctx->cost, 0, sizeof(long long) * n, ctx->stream));
    }
    CK(cudaMemsetAsync(ctx->inbasis, 0, n, ctx->stream));
    // identity carry: every column of B^-1 is trivial (active-column mode)
    RG_TRY(alloc_carry(ctx, !ctx->dense_carry_opt));
    for (;;) {
        LAUNCH(k_zero, 148 * 8, 256, ctx->carry, (size_t)ctx->L * ctx->plane);
        if (ctx->list_mode) LAUNCH(k_zero, 148 * 8, 256, ctx->pk, (size_t)ctx->L * ctx->pplane);
        const BlockView bv = block_of(ctx);
        LAUNCH(k_init_identity, cdiv(std::max(ctx->nloc, 1), 256), 256, bv.base, bv.ps, bv.stride,
               ctx->nloc, ctx->row_lo, ctx->L, ctx->rhs, ctx->list_mode ? 0 : 1);
        LAUNCH(k_init_row0, cdiv(m, 256), 256, ctx->carry, ctx->plane, m, ctx->L, ctx->basis,
               ctx->weighted ? ctx->artcost : nullptr);
        LAUNCH(k_init_scalars, 1, 1, ctx->carry, ctx->plane, m, ctx->L, ctx->rhs, ctx->basis,
               ctx->weighted ? ctx->artcost : nullptr, ctx->row_lo, ctx->nloc, ctx->rank, ctx->world, ctx->sc);
        RG_TRY(sync_mirror(ctx));
        if (ctx->hm->maxbits_carry <= 64 * ctx->L - 1) break;
        RG_TRY(promote(ctx));      // the initial objective does not fit: start at the next width
    }
    if (ctx->weighted)
        LAUNCH(k_init_rowf, cdiv(m, 256), 256, ctx->basis, ctx->wf, ctx->artf, ctx->rowf, m);
    ctx->nk_host = 1;
    LAUNCH(k_init_active, cdiv(ctx->ld, 256), 256, ctx->triv, ctx->klist, ctx->kpos, ctx->ld, m,
           ctx->list_mode ? 1 : 0);
    LAUNCH(k_set_inbasis, cdiv(m, 256), 256, ctx->inbasis, ctx->basis, m);
    ctx->identity_carry = true;
    ctx->rule_ready = false; ctx->have_column = false; ctx->selected = false;
    ctx->t_cur = 0;
    RG_TRY(sync_mirror(ctx));
    if (cost) {
        // starting directly in phase two: -pi = -c_B^T I, -obj = -c_B b
        bool any = false;
        for (int i = 0; i < m; ++i) if (basis[i] >= 0 && cost[basis[i]] != 0) any = true;
        if (any) return rg_phase_switch(ctx, cost);
    }
    return RG_OK;
}

// from_basis (carry/mod.rs:444-478) = BasisInverse::invert on the selected columns: fraction-free Gauss-Jordan
// on the device.  Unit columns are relabelled (B is unchanged), every other basic column is pivoted in with the
// engine's own rank-1 update in a row that still holds an artificial; the rows are permuted at the end so that
// row i holds basis[i], and the costs are installed like at a phase switch.
extern "C" int rg_init_basis(rg_context* ctx, const int32_t* basis, const int64_t* cost) {
    if (ctx) { ctx->demote_need = 0; ctx->kappa_valid = false; }   // the carry is rebuilt / re-priced: the last pivot's bit bound no longer covers it
    if (!ctx || !ctx->carry || !basis || !cost) return RG_ERR_ARG;
    if (ctx->world > 1) { ctx->err = "rg_init_basis: single GPU only (the final row permutation is not sharded)"; return RG_ERR_STATE; }
    const int m = ctx->m, n = ctx->n;
    std::vector<char> seen(n, 0);
    for (int i = 0; i < m; ++i) {
        if (basis[i] < 0 || basis[i] >= n || seen[basis[i]]) { ctx->err = "rg_init_basis: basis must hold m distinct provider columns"; return RG_ERR_ARG; }
        seen[basis[i]] = 1;
    }
    // all-artificial identity start, dense carry (the row permutation below breaks the implicit unit columns)
    const int keep_opt = ctx->dense_carry_opt;
    ctx->dense_carry_opt = 1;
    std::vector<int> ids(m);
    for (int i = 0; i < m; ++i) ids[i] = i - m;
    int rc = rg_init_identity_basis(ctx, ids.data(), nullptr);
    ctx->dense_carry_opt = keep_opt;
    RG_TRY(rc);
    // unit columns +e_r whose row is free: relabel only
    std::vector<int> holder(m, -1);          // row -> provider column placed there
    std::vector<int> pending;
    for (int i = 0; i < m; ++i) {
        const int j = basis[i];
        const bool unit = j >= ctx->nd && ctx->h_colptr[j + 1] - ctx->h_colptr[j] == 1 && ctx->h_vals[ctx->h_colptr[j]] == 1;
        const int r = unit ? ctx->h_rowidx[ctx->h_colptr[j]] : -1;
        if (unit && holder[r] < 0) holder[r] = j; else pending.push_back(j);
    }
    {
        std::vector<int> dev_basis(m);
        for (int r = 0; r < m; ++r) dev_basis[r] = holder[r] >= 0 ? holder[r] : r - m;
        CK(cudaMemcpyAsync(ctx->basis, dev_basis.data(), sizeof(int) * m, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        LAUNCH(k_set_inbasis, cdiv(m, 256), 256, ctx->inbasis, ctx->basis, m);
    }
    // the other columns: pivot each into a free row with a non-zero entry (its own position when possible)
    std::vector<int> want_row(n, -1);
    for (int i = 0; i < m; ++i) want_row[basis[i]] = i;
    std::vector<uint64_t> col;
    size_t stalled = 0;
    while (!pending.empty()) {
        const int j = pending.front();
        pending.erase(pending.begin());
        RG_TRY(rg_generate_column(ctx, j));
        const int LU = LU_of(ctx->L);
        col.assign((size_t)m * LU, 0);
        RG_TRY(rg_get_pivot_column(ctx, col.data()));
        auto nonzero = [&](int r) { for (int l = 0; l < LU; ++l) if (col[(size_t)r * LU + l]) return true; return false; };
        int r = -1;
        if (holder[want_row[j]] < 0 && nonzero(want_row[j])) r = want_row[j];
        for (int t = 0; t < m && r < 0; ++t) if (holder[t] < 0 && nonzero(t)) r = t;
        if (r < 0) {                      // no free row yet: try again after the others (singular if nobody moves)
            pending.push_back(j);
            if (++stalled > pending.size()) { ctx->err = "rg_init_basis: the basis columns are linearly dependent"; return RG_ERR_ARG; }
            continue;
        }
        stalled = 0;
        rg_pivot_info info;
        RG_TRY(rg_bring_into_basis(ctx, j, r, 0, &info));
        holder[r] = j;
    }
    // row i must hold basis[i]: gather the carry rows through the permutation
    std::vector<int> perm(m);
    bool identity = true;
    {
        std::vector<int> row_of(n, -1);
        for (int r = 0; r < m; ++r) row_of[holder[r]] = r;
        for (int i = 0; i < m; ++i) { perm[i] = row_of[basis[i]]; identity = identity && perm[i] == i; }
    }
    if (!identity) {
        int* dperm = nullptr;
        u64* nc = nullptr;
        CK(dev_alloc(&dperm, sizeof(int) * m, ctx->stream));
        CK(cudaMemcpyAsync(dperm, perm.data(), sizeof(int) * m, cudaMemcpyHostToDevice, ctx->stream));
        CK(dev_alloc(&nc, sizeof(u64) * ctx->L * ctx->plane, ctx->stream));
        dim3 grid(std::max(1, std::min(8, cdiv(ctx->ld, 256))), m + 1);
        LAUNCH(k_permute_rows, grid, 256, nc, ctx->carry, ctx->plane, ctx->ld, m, ctx->L, dperm);
        CK(cudaStreamSynchronize(ctx->stream));
        free_dev_on(ctx->carry, ctx->stream); ctx->carry = nc;
        free_dev_on(dperm, ctx->stream);
        CK(cudaMemcpyAsync(ctx->basis, basis, sizeof(int) * m, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (ctx->weighted) LAUNCH(k_init_rowf, cdiv(m, 256), 256, ctx->basis, ctx->wf, ctx->artf, ctx->rowf, m);
    ctx->identity_carry = false;
    ctx->have_column = false; ctx->selected = false; ctx->rule_ready = false;
    return rg_phase_switch(ctx, cost);
}

template <int L>
static int launch_phase_sums_t(rg_context* ctx) {
    constexpr int LU = L + 2;
    const ColsumGeom g = colsum_geom(ctx);
    dim3 grid(cdiv(g.ncols, 32), g.chunks);
    const BlockView bv = block_of(ctx);
    LAUNCH((k_colsum1<L, 1, LU>), grid, 128, bv.base, bv.ps, bv.stride, ctx->nloc, g.rpc, g.klist,
           ctx->svec, (size_t)ctx->ld, ctx->omega_part, g.pcols, ctx->sc);
    if (ctx->world == 1) {
        LAUNCH((k_colsum2<LU, 1>), cdiv(ctx->ld, 64), 64, ctx->omega_part, ctx->ld, 2 * g.chunks, 1,
               ctx->tmprow, ctx->sc, g.triv, ctx->svec, (size_t)ctx->ld, L, g.kpos, g.pcols, 1);
        if (ctx->list_mode)
            LAUNCH((k_colsum2_list<LU>), g.ncols, 128, ctx->omega_part, 2 * g.chunks, g.pcols, g.klist, ctx->ld, 1,
                   ctx->tmprow, ctx->sc);
        return RG_OK;
    }
    size_t words = (size_t)LU * ctx->ld;
    RG_TRY(ensure_xbuf(ctx, words, words * ctx->world));
    LAUNCH((k_colsum2<LU, 1>), cdiv(ctx->ld, 64), 64, ctx->omega_part, ctx->ld, 2 * g.chunks, 0,
           ctx->xsend, ctx->sc, g.triv, ctx->svec, (size_t)ctx->ld, L, g.kpos, g.pcols, 1);
    if (ctx->list_mode)
        LAUNCH((k_colsum2_list<LU>), g.ncols, 128, ctx->omega_part, 2 * g.chunks, g.pcols, g.klist, ctx->ld, 0,
               ctx->xsend, ctx->sc);
    RG_TRY(all_gather(ctx, ctx->xsend, ctx->xrecv, words));
    LAUNCH(k_reset_tmpbits, 1, 1, ctx->sc);
    LAUNCH((k_colsum2<LU>), cdiv(ctx->ld, 64), 64, ctx->xrecv, ctx->ld, ctx->world, 1, ctx->tmprow, ctx->sc);
    return RG_OK;
}
static int launch_phase_sums(rg_context* ctx) {
    DISPATCH_L_RET(ctx->L, launch_phase_sums_t, ctx);
}

extern "C" int rg_phase_switch(rg_context* ctx, const int64_t* cost) {
    if (ctx) { ctx->demote_need = 0; ctx->kappa_valid = false; }   // the carry is rebuilt / re-priced: the last pivot's bit bound no longer covers it
    if (ctx) drop_graphs(ctx);   // state the captured launch sequence depends on may change
    if (!ctx || !ctx->carry || !cost) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->cost, cost, sizeof(long long) * ctx->n, cudaMemcpyHostToDevice, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    set_status(ctx, ST_RUN);
    for (;;) {
        LAUNCH(k_basic_costs, cdiv(ctx->nloc + 1, 256), 256, ctx->basis, ctx->cost, ctx->nloc, ctx->row_lo,
               ctx->svec);
        LAUNCH(k_reset_tmpbits, 1, 1, ctx->sc);
        RG_TRY(launch_phase_sums(ctx));
        RG_TRY(sync_mirror(ctx));
        if (ctx->hm->maxbits_tmp > 64 * ctx->L - 1) { RG_TRY(promote(ctx)); continue; }
        break;
    }
    LAUNCH(k_store_row0, cdiv(ctx->ld, 256), 256, ctx->carry, ctx->plane, ctx->ld, ctx->L, ctx->tmprow,
           LU_of(ctx->L));
    LAUNCH(k_max_into_carrybits, 1, 1, ctx->sc);
    ctx->rule_ready = false; ctx->have_column = false; ctx->selected = false;
    RG_TRY(sync_mirror(ctx));
    return RG_OK;
}

// ------------------------------------------------------------------------------------------------
// pivot rule
// ------------------------------------------------------------------------------------------------
template <int L>
static int launch_gamma_general_t(rg_context* ctx) {
    constexpr int LG = 2 * L + 6;
    const BlockView bv = block_of(ctx);
    if (ctx->world == 1) {
        LAUNCH((k_gamma_init_general<L>), ctx->n, 128, bv.base, bv.ps, bv.stride, ctx->nloc, ctx->n,
               ctx->A.colptr, ctx->A.rowidx, ctx->A.vals, ctx->inbasis, ctx->G, 1, ctx->weighted ? ctx->wf : nullptr,
               ctx->weighted ? ctx->rowf : nullptr, triv_of(ctx), kpos_of(ctx), ctx->sc);
        return RG_OK;
    }
    size_t words = (size_t)LG * ctx->n;
    RG_TRY(ensure_xbuf(ctx, words, words * ctx->world));
    LAUNCH((k_gamma_init_general<L>), ctx->n, 128, bv.base, bv.ps, bv.stride, ctx->nloc, ctx->n,
           ctx->A.colptr, ctx->A.rowidx, ctx->A.vals, ctx->inbasis, ctx->xsend, ctx->rank == 0 ? 1 : 0,
           ctx->weighted ? ctx->wf : nullptr, ctx->weighted ? ctx->rowf : nullptr, triv_of(ctx), kpos_of(ctx), ctx->sc);
    RG_TRY(all_gather(ctx, ctx->xsend, ctx->xrecv, words));
    LAUNCH((k_colsum2<LG>), cdiv(ctx->n, 64), 64, ctx->xrecv, ctx->n, ctx->world, 0, ctx->G, ctx->sc);
    return RG_OK;
}
static int launch_gamma_general(rg_context* ctx) {
    DISPATCH_L_RET(ctx->L, launch_gamma_general_t, ctx);
}

// steepest-edge weights on a general basis, one carry row at a time: stage the row like a pivot row, dot it with
// every column (tensor-core path for the dense block), add the squares.  O(m) launches, used at a phase switch.
template <int L>
static int launch_gamma_rowwise_t(rg_context* ctx) {
    LAUNCH((k_gamma_seed<L>), cdiv(ctx->n, 128), 128, ctx->n, own_of(ctx), ctx->inbasis,
           ctx->weighted ? ctx->wf : nullptr, ctx->G, ctx->sc);
    for (int row = 0; row < ctx->m; ++row) {
        int local = (row >= ctx->row_lo && row < ctx->row_lo + ctx->nloc) ? row - ctx->row_lo + 1 : -1;
        LAUNCH(k_set_rows, 1, 1, ctx->sc, local, row + 1);
        RG_TRY(launch_copyrow(ctx));
        launch_rowdot_t<L>(ctx);
        LAUNCH((k_gamma_accum<L>), cdiv(ctx->n, 128), 128, ctx->n, own_of(ctx), ctx->inbasis, ctx->nu,
               ctx->weighted ? ctx->rowf : nullptr, row, ctx->G, ctx->sc);
    }
    return RG_OK;
}
static int launch_gamma_rowwise(rg_context* ctx) {
    DISPATCH_L_RET(ctx->L, launch_gamma_rowwise_t, ctx);
}

extern "C" int rg_rule_new(rg_context* ctx, int32_t rule) {
    if (ctx) drop_graphs(ctx);   // state the captured launch sequence depends on may change
    if (!ctx || !ctx->carry || rule < 0 || rule > 3) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    ctx->rule = rule;
    set_status(ctx, ST_RUN);
    LAUNCH(k_set_pq, 1, 1, ctx->sc, -1, -1);
    if (rule == RG_RULE_STEEPEST_EDGE) {
        if (ctx->identity_carry) {
            if (ctx->nd > 0)
                LAUNCH(k_gamma_init_identity_dense, cdiv(ctx->nd, 4), 128, ctx->nd, ctx->n, ctx->m, ctx->Acm,
                       ctx->ldc, ctx->inbasis, ctx->G, LG_of(ctx->L));
            if (ctx->n > ctx->nd)
                LAUNCH(k_gamma_init_identity, cdiv(ctx->n - ctx->nd, 256), 256, ctx->n, ctx->nd, ctx->A.colptr,
                       ctx->A.rowidx, ctx->A.vals, ctx->inbasis, ctx->weighted ? ctx->wf : nullptr,
                       ctx->weighted ? ctx->rowf : nullptr, ctx->G, LG_of(ctx->L));
        } else {
            if (ctx->nd > 0) RG_TRY(launch_gamma_rowwise(ctx));   // dense block: row-wise dots on the tensor cores
            else RG_TRY(launch_gamma_general(ctx));
        }
    }
    cudaMemsetAsync(&ctx->sc->last_selected, 0xff, sizeof(int), ctx->stream);
    ctx->rule_ready = true;
    ctx->selected = false;
    RG_TRY(sync_mirror(ctx));
    return RG_OK;
}

static int to_step_status(int dev) {
    return dev == ST_OPTIMAL ? RG_STEP_OPTIMAL : (dev == ST_UNBOUNDED ? RG_STEP_UNBOUNDED : RG_STEP_PIVOTED);
}

extern "C" int rg_select_primal_pivot_column(rg_context* ctx, int32_t* status, int32_t* q) {
    if (!ctx || !ctx->carry || !status || !q) return RG_ERR_ARG;
    if (!ctx->rule_ready) { ctx->err = "rg_rule_new has not been called for this phase"; return RG_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    set_status(ctx, ST_RUN);
    launch_price(ctx);
    RG_TRY(launch_select(ctx));
    RG_TRY(sync_mirror(ctx));
    *status = to_step_status(ctx->hm->status);
    *q = ctx->hm->q;
    ctx->selected = ctx->hm->status == ST_RUN;
    return RG_OK;
}

extern "C" int rg_generate_column(rg_context* ctx, int32_t q) {
    if (!ctx || !ctx->carry || q < 0 || q >= ctx->n) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    set_status(ctx, ST_RUN);
    LAUNCH(k_reset_iter, 1, 1, ctx->sc);
    LAUNCH(k_set_pq, 1, 1, ctx->sc, q, -1);
    launch_ftran(ctx, q);
    RG_TRY(sync_mirror(ctx));
    ctx->have_column = true;
    return RG_OK;
}

extern "C" int rg_select_primal_pivot_row(rg_context* ctx, int32_t* status, int32_t* row) {
    if (!ctx || !ctx->carry || !status || !row) return RG_ERR_ARG;
    if (!ctx->have_column) { ctx->err = "no pivot column generated"; return RG_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    set_status(ctx, ST_RUN);
    RG_TRY(launch_ratio(ctx));
    RG_TRY(sync_mirror(ctx));
    *status = ctx->hm->status == ST_UNBOUNDED ? RG_STEP_UNBOUNDED : RG_STEP_PIVOTED;
    *row = ctx->hm->p - 1;
    return RG_OK;
}

extern "C" int rg_bring_into_basis(rg_context* ctx, int32_t q, int32_t row, int32_t update_rule,
                                   rg_pivot_info* info) {
    if (!ctx || !ctx->carry || q < 0 || q >= ctx->n || row < 0 || row >= ctx->m) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    set_status(ctx, ST_RUN);
    LAUNCH(k_set_pq, 1, 1, ctx->sc, q, -1);
    bool want_se = update_rule && ctx->rule == RG_RULE_STEEPEST_EDGE && ctx->rule_ready;
    RG_TRY(do_pivot(ctx, q, row, want_se, false));
    ctx->have_column = false; ctx->selected = false;
    if (info) {
        info->status = RG_STEP_PIVOTED;
        info->entering = ctx->hm->q_done; info->row = ctx->hm->p_done - 1; info->leaving = ctx->hm->leaving_done;
    }
    return RG_OK;
}

extern "C" int rg_iterate(rg_context* ctx, int64_t max_pivots, rg_pivot_info* trace, int64_t* n_done,
                          int32_t* status) {
    if (!ctx || !ctx->carry || !n_done || !status) return RG_ERR_ARG;
    if (!ctx->rule_ready) { ctx->err = "rg_rule_new has not been called for this phase"; return RG_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    *n_done = 0;
    const bool want_se = ctx->rule == RG_RULE_STEEPEST_EDGE;
    if (!ctx->selected) {
        set_status(ctx, ST_RUN);
        launch_price(ctx);
        RG_TRY(launch_select(ctx));
        RG_TRY(sync_mirror(ctx));
        if (ctx->hm->status != ST_RUN) { *status = to_step_status(ctx->hm->status); return RG_OK; }
        ctx->selected = true;
    }
    while (*n_done < max_pivots) {
        RG_TRY(do_pivot(ctx, -1, -1, want_se, true));
        if (ctx->hm->pivoted) {
            if (trace) {
                rg_pivot_info& t = trace[*n_done];
                t.status = RG_STEP_PIVOTED; t.entering = ctx->hm->q_done; t.row = ctx->hm->p_done - 1;
                t.leaving = ctx->hm->leaving_done;
            }
            (*n_done)++;
        }
        if (ctx->hm->status != ST_RUN) {
            ctx->selected = false;
            *status = to_step_status(ctx->hm->status);
            return RG_OK;
        }
    }
    *status = RG_STEP_PIVOTED;
    return RG_OK;
}

extern "C" int rg_remove_artificial_row(rg_context* ctx, int32_t row, rg_pivot_info* info) {
    if (ctx) drop_graphs(ctx);   // state the captured launch sequence depends on may change
    if (!ctx || !ctx->carry || row < 0 || row >= ctx->m || !info) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    set_status(ctx, ST_RUN);
    LAUNCH(k_reset_iter, 1, 1, ctx->sc);
    {
        int local = (row >= ctx->row_lo && row < ctx->row_lo + ctx->nloc) ? row - ctx->row_lo + 1 : -1;
        LAUNCH(k_set_rows, 1, 1, ctx->sc, local, row + 1);
    }
    launch_price(ctx);
    RG_TRY(launch_copyrow(ctx));
    LAUNCH(k_bp_nonzero, 1, 1, ctx->rowp, (size_t)ctx->ld, ctx->L, ctx->sc);
    DISPATCH_L(ctx->L, launch_rowdot_t, ctx);
    PriceView v{ctx->kappa, LU_of(ctx->L), ctx->n, ctx->inbasis, own_of(ctx)};
    launch_argbest(ctx, 0, ctx->n, CmpArtificial{v, ctx->nu, ctx->sc}, 2);
    if (ctx->world > 1) RG_TRY(merge_columns(ctx, 1));
    RG_TRY(sync_mirror(ctx));
    int q = ctx->hm->found;
    ctx->selected = false; ctx->have_column = false;
    if (q < 0) {
        info->status = RG_STEP_OPTIMAL; info->entering = -1; info->row = row; info->leaving = 0;
        return RG_OK;
    }
    LAUNCH(k_set_pq, 1, 1, ctx->sc, q, -1);
    RG_TRY(do_pivot(ctx, q, row, false, false));
    info->status = RG_STEP_PIVOTED;
    info->entering = ctx->hm->q_done; info->row = ctx->hm->p_done - 1; info->leaving = ctx->hm->leaving_done;
    return RG_OK;
}

// ------------------------------------------------------------------------------------------------
// exports
// ------------------------------------------------------------------------------------------------
__global__ void k_gather(u64* out, const u64* base, size_t stride, size_t idx0, size_t step, int count,
                         int nl) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    for (int l = 0; l < nl; ++l) out[(size_t)i * nl + l] = base[(size_t)l * stride + idx0 + (size_t)i * step];
}

static int export_planar(rg_context* ctx, const u64* base, size_t stride, size_t idx0, size_t step,
                         int count, int nl, uint64_t* out) {
    u64* tmp = nullptr;
    CK(dev_alloc(&tmp, sizeof(u64) * (size_t)count * nl, ctx->stream));
    LAUNCH(k_gather, cdiv(count, 256), 256, tmp, base, stride, idx0, step, count, nl);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpyAsync(out, tmp, sizeof(u64) * (size_t)count * nl, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    free_dev_on(tmp, ctx->stream);
    return RG_OK;
}

extern "C" int rg_get_limbs(rg_context* ctx, int32_t* limbs) {
    if (!ctx || !limbs) return RG_ERR_ARG;
    *limbs = ctx->L;
    return RG_OK;
}
extern "C" int rg_get_denominator(rg_context* ctx, uint64_t* out) {
    if (!ctx || !ctx->carry || !out) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpyAsync(out, ctx->sc->D, sizeof(u64) * ctx->L, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    return RG_OK;
}
extern "C" int rg_get_basis(rg_context* ctx, int32_t* basis) {
    if (!ctx || !ctx->carry || !basis) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpyAsync(basis, ctx->basis, sizeof(int) * ctx->m, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    return RG_OK;
}
// export one number per constraint row (rows are block-distributed when world > 1): local gather into
// a padded block, all-gather, compact on the host
static int export_rows(rg_context* ctx, const u64* base, size_t stride, size_t idx0, size_t step, int nl,
                       uint64_t* out) {
    if (ctx->world == 1) return export_planar(ctx, base, stride, idx0, step, ctx->m, nl, out);
    const int q = (ctx->m + ctx->world - 1) / ctx->world;
    size_t words = (size_t)q * nl;
    RG_TRY(ensure_xbuf(ctx, words, words * ctx->world));
    CK(cudaMemsetAsync(ctx->xsend, 0, words * sizeof(u64), ctx->stream));
    if (ctx->nloc > 0)
        LAUNCH(k_gather, cdiv(ctx->nloc, 256), 256, ctx->xsend, base, stride, idx0, step, ctx->nloc, nl);
    RG_TRY(all_gather(ctx, ctx->xsend, ctx->xrecv, words));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<u64> host(words * ctx->world);
    CK(cudaMemcpyAsync(host.data(), ctx->xrecv, host.size() * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < ctx->world; ++r) {
        int lo = std::min(ctx->m, r * q), hi = std::min(ctx->m, (r + 1) * q);
        if (hi > lo) memcpy(out + (size_t)lo * nl, host.data() + (size_t)r * words, (size_t)(hi - lo) * nl * sizeof(u64));
    }
    return RG_OK;
}

extern "C" int rg_get_b(rg_context* ctx, uint64_t* out) {
    if (!ctx || !ctx->carry || !out) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const BlockView bv = block_of(ctx);
    return export_rows(ctx, bv.base, bv.ps, (size_t)bv.stride, (size_t)bv.stride, ctx->L, out);
}
extern "C" int rg_get_minus_objective(rg_context* ctx, uint64_t* out) {
    if (!ctx || !ctx->carry || !out) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    return export_planar(ctx, ctx->carry, ctx->plane, 0, 1, 1, ctx->L, out);
}
extern "C" int rg_get_minus_pi(rg_context* ctx, uint64_t* out) {
    if (!ctx || !ctx->carry || !out) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    return export_planar(ctx, ctx->carry, ctx->plane, 1, 1, ctx->m, ctx->L, out);
}
extern "C" int rg_get_basis_inverse_row(rg_context* ctx, int32_t row, uint64_t* out) {
    if (!ctx || !ctx->carry || !out || row < 0 || row >= ctx->m) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    // stage the row like a pivot row (owner copies, sum all-reduce replicates it), then export it
    set_status(ctx, ST_RUN);
    int local = (row >= ctx->row_lo && row < ctx->row_lo + ctx->nloc) ? row - ctx->row_lo + 1 : -1;
    LAUNCH(k_set_rows, 1, 1, ctx->sc, local, row + 1);
    RG_TRY(launch_copyrow(ctx));
    ctx->have_column = false;
    return export_planar(ctx, ctx->rowp, (size_t)ctx->ld, 1, 1, ctx->m, ctx->L, out);
}
extern "C" int rg_get_pivot_column(rg_context* ctx, uint64_t* out) {
    if (!ctx || !ctx->carry || !out) return RG_ERR_ARG;
    if (!ctx->have_column) { ctx->err = "no pivot column generated"; return RG_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    return export_rows(ctx, ctx->u, (size_t)ctx->ld, 1, 1, LU_of(ctx->L), out);
}
extern "C" int rg_get_relative_costs(rg_context* ctx, uint64_t* out) {
    if (!ctx || !ctx->carry || !out) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    set_status(ctx, ST_RUN);
    {   // export every column: price the full range on this rank
        int d0 = ctx->d0, d1 = ctx->d1, s0 = ctx->s0, s1 = ctx->s1;
        ctx->d0 = 0; ctx->d1 = ctx->nd; ctx->s0 = ctx->nd; ctx->s1 = ctx->n;
        launch_price(ctx);
        ctx->d0 = d0; ctx->d1 = d1; ctx->s0 = s0; ctx->s1 = s1;
    }
    ctx->selected = false;
    return export_planar(ctx, ctx->kappa, (size_t)ctx->n, 0, 1, ctx->n, LU_of(ctx->L), out);
}
extern "C" int rg_get_gamma(rg_context* ctx, uint64_t* out) {
    if (!ctx || !ctx->carry || !out) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    return export_planar(ctx, ctx->G, (size_t)ctx->n, 0, 1, ctx->n, LG_of(ctx->L), out);
}

// generate_element (tableau/inverse_maintenance/mod.rs:200-215; carry/basis_inverse_rows.rs:179-195): the
// single entry (B^-1 a_j)[row] = row_row(B^-1) . a_j, numerator over the current denominator.  The row is
// staged like a pivot row (replicated when row-sharded), the column scattered densely, one block reduces.
template <int L>
static void launch_element_t(rg_context* ctx, int j) {
    LAUNCH(k_scatter_col, cdiv(ctx->m, 256), 256, ctx->aq, ctx->m, ctx->nd, ctx->A.colptr, ctx->A.rowidx,
           ctx->A.vals, ctx->Acm, ctx->ldc, j, ctx->sc);
    LAUNCH(k_scatter_col2, cdiv(ctx->m, 256), 256, ctx->aq, ctx->nd, ctx->A.colptr, ctx->A.rowidx, ctx->A.vals, j,
           ctx->sc);
    LAUNCH((k_ftran_row0<L>), std::max(1, std::min(32, cdiv(ctx->m, 1024))), 256, ctx->rowp, (size_t)ctx->ld, ctx->m,
           ctx->aq, (const long long*)nullptr, j, ctx->tmprow, (size_t)ctx->ld, ctx->row0_part, ctx->sc);
}
extern "C" int rg_get_element(rg_context* ctx, int32_t row, int32_t j, uint64_t* out) {
    if (!ctx || !ctx->carry || !out || row < 0 || row >= ctx->m || j < 0 || j >= ctx->n) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    set_status(ctx, ST_RUN);
    int local = (row >= ctx->row_lo && row < ctx->row_lo + ctx->nloc) ? row - ctx->row_lo + 1 : -1;
    LAUNCH(k_set_rows, 1, 1, ctx->sc, local, row + 1);
    RG_TRY(launch_copyrow(ctx));
    DISPATCH_L(ctx->L, launch_element_t, ctx, j);
    return export_planar(ctx, ctx->tmprow, (size_t)ctx->ld, 0, 1, 1, LU_of(ctx->L), out);
}

// BasisChangeComputationInfo (tableau/mod.rs:205-234) of the last basis change, for a host-side PivotRule:
//   column      column_before_change  = B_old^-1 a_q, m entries of limbs+2 words over `denominator_before`
//   work        work_vector           = column^T B_old^-1, m entries of 2*limbs+5 words over denominator_before^2
//                                       (only computed when the steepest-edge update ran; else pass NULL)
//   row         basis_inverse_row     = row p of the NEW B^-1, m entries of limbs words over the current denominator
// Any pointer may be NULL.  Valid until the next call that generates a column or changes the basis.
extern "C" int rg_get_basis_change_info(rg_context* ctx, uint64_t* column, uint64_t* work, uint64_t* row,
                                        uint64_t* denominator_before) {
    if (!ctx || !ctx->carry) return RG_ERR_ARG;
    if (ctx->pivots == 0 || ctx->hm->p_done < 1) { ctx->err = "no basis change has been performed"; return RG_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    if (denominator_before) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaMemcpyAsync(denominator_before, ctx->sc->Dold, sizeof(u64) * ctx->L, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (column) RG_TRY(export_rows(ctx, ctx->u, (size_t)ctx->ld, 1, 1, LU_of(ctx->L), column));
    if (work) {
        if (!ctx->work_valid) { ctx->err = "the work vector is only kept when the steepest-edge update ran"; return RG_ERR_STATE; }
        RG_TRY(export_planar(ctx, ctx->omega, (size_t)ctx->ld, 1, 1, ctx->m, LW_of(ctx->L), work));
    }
    if (row) RG_TRY(rg_get_basis_inverse_row(ctx, ctx->hm->p_done - 1, row));
    return RG_OK;
}
extern "C" int rg_get_stats(rg_context* ctx, rg_stats* out) {
    if (!ctx || !out) return RG_ERR_ARG;
    memset(out, 0, sizeof(*out));
    out->pivots = ctx->pivots; out->promotions = ctx->promotions; out->limbs = ctx->L;
    out->kernel_launches = ctx->launches;
    if (ctx->hm) { out->max_bits = ctx->hm->maxbits_carry; out->denominator_bits = ctx->hm->bits_D; }
    out->reserved = ctx->list_mode ? ctx->nk_host : 0;   // non-trivial carry columns (0: dense mode)
    out->demotions = ctx->demotions;
    for (int k = 0; k < RG_NWIDTHS; ++k) {
        out->limb_widths[k] = kWidths[k];
        out->pivots_at_limbs[k] = ctx->pivots_at[k];
        out->k1_launches_at_limbs[k] = ctx->k1_launches[k];
        out->k1_ms_at_limbs[k] = ctx->k1_ms[k];
        out->k1_bytes_at_limbs[k] = ctx->k1_bytes[k];
        out->k1_imads_at_limbs[k] = ctx->k1_imads[k];
    }
    out->timer_ms = ctx->timer_ms;
    for (int k = 0; k < 8; ++k) out->phase_ms[k] = ctx->phase_ms[k];
    return RG_OK;
}

extern "C" int rg_set_profile(rg_context* ctx, int32_t on) {
    if (!ctx) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (on && !ctx->ev0) {
        // the profiling events are recycled per device for the life of the process (no create / destroy churn per
        // solve; the slow solves once attributed to it were allocator stalls, see dev_alloc)
        std::lock_guard<std::mutex> lock(g_hm_mutex);
        auto& pool = g_prof_events[ctx->device & 15];
        if (!pool.empty()) {
            ProfEvents pe = pool.back(); pool.pop_back();
            ctx->ev0 = pe.e[0]; ctx->ev1 = pe.e[1];
            for (int k = 0; k < 8; ++k) ctx->evp[k] = pe.e[2 + k];
        } else {
            CK(cudaEventCreate(&ctx->ev0)); CK(cudaEventCreate(&ctx->ev1));
            for (int k = 0; k < 8; ++k) CK(cudaEventCreate(&ctx->evp[k]));
        }
    }
    ctx->profile = on < 0 ? 0 : (on > 2 ? 2 : on);
    return RG_OK;
}
extern "C" int rg_timer_start(rg_context* ctx) {
    if (!ctx) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->evt0) { CK(cudaEventCreate(&ctx->evt0)); CK(cudaEventCreate(&ctx->evt1)); }
    CK(cudaEventRecord(ctx->evt0, ctx->stream));
    return RG_OK;
}
extern "C" int rg_timer_stop(rg_context* ctx) {
    if (!ctx || !ctx->evt0) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->evt1, ctx->stream));
    CK(cudaEventSynchronize(ctx->evt1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->evt0, ctx->evt1));
    ctx->timer_ms = ms;
    return RG_OK;
}

// ------------------------------------------------------------------------------------------------
// debug / self-test hooks (used by tests/test_gpu_bigint.py and scripts/debug_gpu.py only)
// ------------------------------------------------------------------------------------------------
extern "C" int rg_debug_scalars(rg_context* ctx, void* out, int64_t bytes) {
    if (!ctx || !out) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    size_t nbytes = std::min<size_t>((size_t)bytes, sizeof(Scalars));
    CK(cudaMemcpyAsync(out, ctx->sc, nbytes, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
    return (int)sizeof(Scalars);
}
extern "C" int rg_debug_vector(rg_context* ctx, int32_t which, uint64_t* out) {
    if (!ctx || !ctx->carry || !out) return RG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (which == 0) return export_planar(ctx, ctx->u, (size_t)ctx->ld, 0, 1, ctx->m + 1, LU_of(ctx->L), out);
    if (which == 1) return export_planar(ctx, ctx->rowp, (size_t)ctx->ld, 0, 1, ctx->m + 1, ctx->L, out);
    if (which == 2) return export_planar(ctx, ctx->omega, (size_t)ctx->ld, 0, 1, ctx->m + 1, LW_of(ctx->L), out);
    return RG_ERR_ARG;
}

// op 0: mul_lo<W>(a,b); 1: mul2_lo<W>(a,b,c,d); 2: acc(c, W limbs) += a(W-2 limbs.. see below) * s;
// 3: rt_inv_odd(a) mod 2^(64W); 4: mul_full_ct<W,W>(a,b) -> 2W limbs; 5: rt_cmp_prod sign(a*b - c*d)
template <int W>
__global__ void k_selftest(int op, const u64* a, const u64* b, const u64* c, const u64* d, long long s,
                           u64* out) {
    u64 x[W], y[W], z[W], w[W], r[2 * W];
    u64 big[5 * RG_MAXW];
    for (int l = 0; l < W; ++l) { x[l] = a[l]; y[l] = b[l]; z[l] = c[l]; w[l] = d[l]; }
    for (int l = 0; l < 2 * W; ++l) r[l] = 0;
    if (op == 0) { u64 o[W]; mul_lo<W>(o, x, y); for (int l = 0; l < W; ++l) r[l] = o[l]; }
    else if (op == 1) { u64 o[W]; mul2_lo<W>(o, x, y, z, w); for (int l = 0; l < W; ++l) r[l] = o[l]; }
    else if (op == 2) {
        u64 acc[W + 2];
        for (int l = 0; l < W + 2; ++l) acc[l] = l < W ? z[l] : ((i64)z[W - 1] < 0 ? ~0ull : 0ull);
        mac_small<W + 2, W>(acc, x, s);
        for (int l = 0; l < W + 2 && l < 2 * W; ++l) r[l] = acc[l];
    } else if (op == 3) { u64* o = big; rt_inv_odd(o, x, W, W, big + RG_MAXW); for (int l = 0; l < W; ++l) r[l] = o[l]; }
    else if (op == 4) { mul_full_ct<W, W>(r, x, y); }
    else if (op == 5) { r[0] = (u64)(i64)rt_cmp_prod(x, W, y, W, z, W, w, W, big); }
    else if (op == 6) {   // 32-bit limb path used by K1: x*y + z*w mod 2^(64W)
        u32 xx[2 * W], zz[2 * W], oo[2 * W];
        for (int l = 0; l < W; ++l) { xx[2 * l] = (u32)x[l]; xx[2 * l + 1] = (u32)(x[l] >> 32); zz[2 * l] = (u32)z[l]; zz[2 * l + 1] = (u32)(z[l] >> 32); }
        mp_mul2_lo<2 * W>(oo, xx, reinterpret_cast<const u32*>(b), zz, reinterpret_cast<const u32*>(d));
        for (int l = 0; l < W; ++l) r[l] = (u64)oo[2 * l] | ((u64)oo[2 * l + 1] << 32);
    }
    for (int l = 0; l < 2 * W; ++l) out[l] = r[l];
}


// Integer-pipe peak micro-benchmark: the IMAD.WIDE carry-chain instruction mix of the K1
// products (mp_mul_lo<32>: 528 IMAD.WIDE.U32[.X] per product) on register operands, no memory traffic, every SM
// filled; timed with CUDA events after a warm-up that lets the clocks settle.  Returns IMAD.WIDE per second.
__global__ void __launch_bounds__(256) k_imad_peak(u32* out, int iters) {
    constexpr int N = 32;
    u32 x[N], y[N], r[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { x[k] = threadIdx.x * 2654435761u + k * 40503u + blockIdx.x; y[k] = x[k] ^ (0x9e3779b9u * (k + 1)); }
    for (int it = 0; it < iters; ++it) {
        mp_mul_lo<N>(r, x, y);
#pragma unroll
        for (int k = 0; k < N; ++k) x[k] = r[k];
        mp_mul_lo<N>(r, y, x);
#pragma unroll
        for (int k = 0; k < N; ++k) y[k] = r[k];
    }
    u32 acc = 0;
#pragma unroll
    for (int k = 0; k < N; ++k) acc ^= x[k] ^ y[k];
    if (acc == 0x12345678u) out[0] = acc;     // keeps the chains alive
}
extern "C" int rg_measure_imad_peak(int32_t device, double seconds, double* imad_per_s) {
    if (!imad_per_s) return RG_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return RG_ERR_CUDA;
    u32* out = nullptr;
    if (cudaMalloc(&out, 64) != cudaSuccess) return RG_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 4, iters = 256;
    const double per_launch = (double)blocks * 256 * iters * 2 * (32.0 * 33.0 / 2.0);
    double best = 0;
    const int reps = seconds > 0 ? (int)(seconds / 0.02) + 4 : 12;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        k_imad_peak<<<blocks, 256>>>(out, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return RG_ERR_CUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms > 0) best = std::max(best, per_launch / (ms * 1e-3));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *imad_per_s = best;
    return RG_OK;
}

extern "C" int rg_selftest(int32_t op, int32_t W, const uint64_t* a, const uint64_t* b, const uint64_t* c,
                           const uint64_t* d, int64_t s, uint64_t* out /* 2W words */) {
    u64* dev = nullptr;
    if (cudaMalloc(&dev, sizeof(u64) * 6 * W) != cudaSuccess) return RG_ERR_CUDA;
    cudaMemcpy(dev, a, sizeof(u64) * W, cudaMemcpyHostToDevice);
    cudaMemcpy(dev + W, b, sizeof(u64) * W, cudaMemcpyHostToDevice);
    cudaMemcpy(dev + 2 * W, c, sizeof(u64) * W, cudaMemcpyHostToDevice);
    cudaMemcpy(dev + 3 * W, d, sizeof(u64) * W, cudaMemcpyHostToDevice);
    switch (W) {
        case 1: k_selftest<1><<<1, 1>>>(op, dev, dev + W, dev + 2 * W, dev + 3 * W, s, dev + 4 * W); break;
        case 2: k_selftest<2><<<1, 1>>>(op, dev, dev + W, dev + 2 * W, dev + 3 * W, s, dev + 4 * W); break;
        case 3: k_selftest<3><<<1, 1>>>(op, dev, dev + W, dev + 2 * W, dev + 3 * W, s, dev + 4 * W); break;
        case 4: k_selftest<4><<<1, 1>>>(op, dev, dev + W, dev + 2 * W, dev + 3 * W, s, dev + 4 * W); break;
        case 5: k_selftest<5><<<1, 1>>>(op, dev, dev + W, dev + 2 * W, dev + 3 * W, s, dev + 4 * W); break;
        case 9: k_selftest<9><<<1, 1>>>(op, dev, dev + W, dev + 2 * W, dev + 3 * W, s, dev + 4 * W); break;
        case 17: k_selftest<17><<<1, 1>>>(op, dev, dev + W, dev + 2 * W, dev + 3 * W, s, dev + 4 * W); break;
        default: cudaFree(dev); return RG_ERR_ARG;
    }
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out, dev + 4 * W, sizeof(u64) * 2 * W, cudaMemcpyDeviceToHost);
    cudaFree(dev);
    return e == cudaSuccess ? RG_OK : RG_ERR_CUDA;
}
