// splendor_b200.cu -- C-ABI entry points (include/splendor_b200.h) and the native host driver of
// the level-synchronous search.  Build: nvcc -gencode arch=compute_100a,code=sm_100a (see
// __graft_entry__.build()).  No CPU fallback: without a usable device every compute call fails.
#include "../../include/splendor_b200.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "spl_kernels.cuh"
#include "spl_m2.cuh"
#include "spl_shard.cuh"
#include "spl_realistic.cuh"
#include "spl_tables.cuh"

using namespace spl;

static std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // grow to at least `bytes` (geometric), optionally preserving the first `keep` bytes
    cudaError_t ensure(size_t bytes, size_t keep, cudaStream_t st) {
        if (bytes <= cap) return cudaSuccess;
        size_t ncap = std::max(bytes, cap + cap / 2);
        ncap = (ncap + 255) & ~(size_t)255;
        void *np = nullptr;
        cudaError_t e = cudaMalloc(&np, ncap);
        if (e != cudaSuccess && ncap > bytes) {  // retry with the exact size
            cudaGetLastError();
            ncap = (bytes + 255) & ~(size_t)255;
            e = cudaMalloc(&np, ncap);
        }
        if (e != cudaSuccess) return e;
        if (keep && p) {
            e = cudaMemcpyAsync(np, p, keep, cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { cudaFree(np); return e; }
        }
        if (p) cudaFree(p);
        p = np;
        cap = ncap;
        return cudaSuccess;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
    void swap(DevBuf &o) { std::swap(p, o.p); std::swap(cap, o.cap); }
};

struct spl_ctx {
    int device = 0;
    std::string err;
    // visited table
    uint64_t *table = nullptr;
    uint64_t cap = 0 /* slots = 3 * nb */, nb = 0 /* 64-byte buckets */, max_table_bytes = 0, occupied = 0;
    uint32_t epoch = 0;
    uint64_t chunk_parents = 0;
    bool chunk_user = false;  // chunk_parents was set by the caller (spl_config), not the default
    // card-set node table of the grouped level (spl_m2.cuh)
    uint64_t *nodes = nullptr;
    uint64_t nn = 0, node_occ = 0, max_node_bytes = 0;
    uint64_t link_budget = 0;          // device bytes a solver may hold in parent-link columns before it spills (0: no limit)
    uint64_t spilled_bytes = 0;        // link-column bytes moved to pinned host memory so far
    uint16_t *d_gemrank = nullptr;
    uint16_t *d_rankgems = nullptr;   // inverse of d_gemrank
    DevBuf brec, ntk8, boff2, run_start, run_wpre, cls_list;
    int tie_link_top = 0;  // > 0: the beam cut breaks score ties on the records' link words (arrival order)
    // constant tables
    DevTables *d_tabs = nullptr;
    uint32_t *d_takes_idx = nullptr;
    uint16_t *d_takes_edges = nullptr;
    double *d_lut = nullptr;
    ScoreLuts luts{};
    // scalars
    Counters *d_ctr = nullptr, *h_ctr = nullptr;
    SelState *d_sel = nullptr, *h_sel = nullptr;
    uint32_t *d_hist = nullptr;
    ScoreDict *d_dict = nullptr, *d_dict2 = nullptr;  // d_dict2: the merged (global) dictionary of the sharded cut
    unsigned long long *d_dest = nullptr;             // [3][MAX_RANKS]: per-destination record counts / write cursors / base pointers
    int dict_skip = 0;
    int identity = IDENT_KEY;  // visited-table identity of the speedrun solver (spl_set_identity)  // levels left before the dictionary path is tried again after a miss
    DevBuf status[3];
    // scratch
    DevBuf off, cand_slot, tmp_rec, tmp_rec2, sk, y[2], idx[2], kl[2], kh[2], matrix, matrix2, os_hist, os_status;
    DevBuf pool_front, pool_uniq, pool_grank;  // frontier buffers lent to the active solver
    DevBuf rcfg, rcand, rvmask, ridx64, rtmp;  // realistic mode scratch
    RBucket *rtable = nullptr;  // realistic mode: exact-key visited table (64-byte buckets, one state each)
    uint64_t rnb = 0, rocc = 0, rslots_hint = 0;
    std::vector<DevBuf *> pool_links;  // link columns of finished solves, reused by the next one
    cudaEvent_t ev[8]{};
    long long launches = 0, h2d_bytes = 0, d2h_bytes = 0;
    int64_t dtopk_n = 0;       // distributed top-k: elements staged by spl_dtopk_begin
    int64_t route_n = -1;      // candidates of the last spl_route_keys (its permutation lives in idx[1])
    const void *active = nullptr;  // the live spl_solver: it owns the visited set, the pool buffers and the scratch
    bool rcfg_dirty = false;       // a stage operator re-uploaded the realistic config since the live solver's upload
    bool dtopk_recs = false;
    const uint64_t *dtopk_keys = nullptr;  // caller-owned [n][2] key array of the distributed top-k
};

static int fail(spl_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}
// stage operators that rewrite the visited set / identity must not run under a live solver
#define NO_LIVE_SOLVER(c, what)                                                                                   \
    do {                                                                                                          \
        if ((c)->active) return fail(c, SPL_E_STATE, what ": a solver is open on this context (destroy it first)"); \
    } while (0)
#define CK(c, call)                                                                             \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            cudaGetLastError();                                                                 \
            return fail(c, e_ == cudaErrorMemoryAllocation ? SPL_E_NOMEM : SPL_E_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                            \
        }                                                                                       \
    } while (0)
#define CKS(c, expr)              \
    do {                          \
        int s_ = (expr);          \
        if (s_ != SPL_OK) return s_; \
    } while (0)

// every entry point: select the context's device and drop any stale non-sticky error another library left in this
// thread's runtime state (the launch checks below read cudaGetLastError, which would otherwise report it as ours)
static inline cudaError_t enter_device(spl_ctx *c) {
    const cudaError_t e = cudaSetDevice(c->device);
    if (e == cudaSuccess) cudaGetLastError();
    return e;
}
static inline unsigned nblk(int64_t n, int per = TILE) { return (unsigned)((n + per - 1) / per); }
static inline int bitlen(uint64_t x) { return x ? 64 - __builtin_clzll(x) : 0; }

extern "C" {

int32_t spl_abi_version(void) { return SPL_ABI_VERSION; }

const char *spl_last_error(const spl_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

const uint32_t *spl_deck_table(int32_t *n_cards) {
    if (n_cards) *n_cards = SPL_NUM_CARDS;
    return SPL_DECK_PACKED;
}

int32_t spl_host_takes(const uint8_t gems[5], uint8_t out[100 * 5]) {
    const HostTables &T = host_tables();
    uint32_t key = 0;
    for (int c = 0; c < NCOL; ++c) {
        if (gems[c] > 7) return SPL_E_INVALID;
        key |= (uint32_t)gems[c] << (3 * c);
    }
    const uint32_t e = T.takes_idx[key], n = e & 0xff, off = e >> 8;
    for (uint32_t i = 0; i < n; ++i)
        for (int c = 0; c < NCOL; ++c) out[i * 5 + c] = (T.takes_edges[off + i] >> (3 * c)) & 7;
    return (int32_t)n;
}

int32_t spl_host_buys(const uint8_t key[5], uint8_t out[90]) {
    const HostTables &T = host_tables();
    uint64_t lo = ~0ull, hi = ~0ull;
    for (int c = 0; c < NCOL; ++c) {
        if (key[c] > 7) return SPL_E_INVALID;
        lo &= T.dev.buy_lo[c][key[c]];
        hi &= T.dev.buy_hi[c][key[c]];
    }
    int n = 0;
    for (int i = 0; i < SPL_NUM_CARDS; ++i) {
        const int b = 15 + i;
        if (b < 64 ? (lo >> b) & 1 : (hi >> (b - 64)) & 1) out[n++] = (uint8_t)i;
    }
    return n;
}

// ------------------------------------------------------------------ context
// The table is sized in slots (3 per 64-byte bucket); slot ids are u32 (bucket << 2 | slot), so at most 2^30 buckets.
constexpr uint64_t MAX_BUCKETS = 1ull << 30;
static uint64_t buckets_for(uint64_t slots) {
    return std::min<uint64_t>(std::max<uint64_t>(slots / BUCKET_SLOTS, 16), MAX_BUCKETS);
}
static int alloc_table(spl_ctx *c, uint64_t slots, cudaStream_t st) {
    const uint64_t nb = buckets_for(slots);
    CK(c, cudaMalloc(&c->table, nb * 64));
    CK(c, cudaMemsetAsync(c->table, 0, nb * 64, st));
    c->nb = nb;
    c->cap = nb * BUCKET_SLOTS;
    return SPL_OK;
}

int32_t spl_create(const spl_config *cfg, spl_ctx **out) {
    if (!cfg || !out) return fail(nullptr, SPL_E_INVALID, "spl_create: null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, SPL_E_NODEVICE, "spl_create: no CUDA device (%s); this library has no CPU fallback",
                    cudaGetErrorString(e));
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, SPL_E_INVALID, "spl_create: bad device %d", cfg->device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10)
        return fail(nullptr, SPL_E_NODEVICE, "spl_create: device %d is sm_%d%d; this build is sm_100a only", cfg->device,
                    prop.major, prop.minor);
    spl_ctx *c = new spl_ctx();
    c->device = cfg->device;
#define CKC(call)                                                                                             \
    do {                                                                                                      \
        cudaError_t e2_ = (call);                                                                             \
        if (e2_ != cudaSuccess) {                                                                             \
            cudaGetLastError();                                                                               \
            int r_ = fail(nullptr, e2_ == cudaErrorMemoryAllocation ? SPL_E_NOMEM : SPL_E_CUDA, "%s: %s", #call, \
                          cudaGetErrorString(e2_));                                                           \
            delete c;                                                                                         \
            return r_;                                                                                        \
        }                                                                                                     \
    } while (0)
    CKC(cudaSetDevice(c->device));
    size_t free_b = 0, total_b = 0;
    CKC(cudaMemGetInfo(&free_b, &total_b));
    c->max_table_bytes = cfg->max_table_bytes ? cfg->max_table_bytes : (uint64_t)(free_b * 0.6);
    c->chunk_user = cfg->chunk_parents != 0;
    c->chunk_parents = cfg->chunk_parents ? std::min<uint64_t>(cfg->chunk_parents, 16ull << 20) : (4ull << 20);
    c->chunk_parents = std::max<uint64_t>(c->chunk_parents, TILE);
    const HostTables &T = host_tables();
    CKC(cudaMalloc(&c->d_tabs, sizeof(DevTables)));
    CKC(cudaMemcpy(c->d_tabs, &T.dev, sizeof(DevTables), cudaMemcpyHostToDevice));
    CKC(cudaMalloc(&c->d_takes_idx, T.takes_idx.size() * 4));
    CKC(cudaMemcpy(c->d_takes_idx, T.takes_idx.data(), T.takes_idx.size() * 4, cudaMemcpyHostToDevice));
    CKC(cudaMalloc(&c->d_takes_edges, T.takes_edges.size() * 2));
    CKC(cudaMemcpy(c->d_takes_edges, T.takes_edges.data(), T.takes_edges.size() * 2, cudaMemcpyHostToDevice));
    const size_t n_lut = T.lut_pts.size() + T.lut_saved.size() + T.lut_small.size();
    CKC(cudaMalloc(&c->d_lut, n_lut * 8));
    CKC(cudaMemcpy(c->d_lut, T.lut_pts.data(), T.lut_pts.size() * 8, cudaMemcpyHostToDevice));
    CKC(cudaMemcpy(c->d_lut + T.lut_pts.size(), T.lut_saved.data(), T.lut_saved.size() * 8, cudaMemcpyHostToDevice));
    CKC(cudaMemcpy(c->d_lut + T.lut_pts.size() + T.lut_saved.size(), T.lut_small.data(), T.lut_small.size() * 8,
                   cudaMemcpyHostToDevice));
    c->luts.pts = c->d_lut;
    c->luts.saved = c->d_lut + T.lut_pts.size();
    c->luts.small = c->d_lut + T.lut_pts.size() + T.lut_saved.size();
    CKC(cudaMalloc(&c->d_gemrank, T.gemrank.size() * 2));
    CKC(cudaMemcpy(c->d_gemrank, T.gemrank.data(), T.gemrank.size() * 2, cudaMemcpyHostToDevice));
    {
        std::vector<uint16_t> inv(GEM_STATES, 0);
        for (size_t g = 0; g < T.gemrank.size(); ++g)
            if (T.gemrank[g] != 0xFFFF) inv[T.gemrank[g]] = (uint16_t)g;
        CKC(cudaMalloc(&c->d_rankgems, inv.size() * 2));
        CKC(cudaMemcpy(c->d_rankgems, inv.data(), inv.size() * 2, cudaMemcpyHostToDevice));
    }
    CKC(cudaMalloc(&c->d_ctr, sizeof(Counters)));
    CKC(cudaMallocHost(&c->h_ctr, sizeof(Counters)));
    CKC(cudaMalloc(&c->d_sel, sizeof(SelState)));
    CKC(cudaMallocHost(&c->h_sel, sizeof(SelState)));
    CKC(cudaMalloc(&c->d_hist, SEL_BINS * 4));
    CKC(cudaMemset(c->d_hist, 0, SEL_BINS * 4));
    CKC(cudaMalloc(&c->d_dict, sizeof(ScoreDict)));
    for (auto &ev : c->ev) CKC(cudaEventCreate(&ev));
    CKC(cudaFuncSetAttribute(expand_kernel<MODE_PROBE, IDENT_KEY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExpandSmem2)));
    CKC(cudaFuncSetAttribute(expand_kernel<MODE_PROBE, IDENT_PYHASH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExpandSmem2)));
    CKC(cudaFuncSetAttribute(expand_kernel<MODE_LIST, IDENT_KEY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExpandSmem2)));
    CKC(cudaFuncSetAttribute(m2_buys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BuySmem)));
    CKC(cudaFuncSetAttribute(m2_group_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)offsetof(WarpSmem, bsort)));
    CKC(cudaFuncSetAttribute(m2_group_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WarpSmem)));
    CKC(cudaFuncSetAttribute(m2_group_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BigSmem)));
    CKC(cudaFuncSetAttribute(os_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)offsetof(OsSmem, stage_p)));
    CKC(cudaFuncSetAttribute(os_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OsSmem)));
    CKC(cudaFuncSetAttribute(gs_buys_route_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RouteSmem)));
    CKC(cudaMalloc(&c->d_dict2, sizeof(ScoreDict)));
    CKC(cudaMalloc(&c->d_dest, 3 * MAX_RANKS * 8));
    {
        c->max_node_bytes = cfg->max_node_bytes ? cfg->max_node_bytes : (uint64_t)(free_b * 0.5);
        uint64_t nn = cfg->node_slots ? cfg->node_slots : (1ull << 12);
        nn = std::max<uint64_t>(std::min<uint64_t>(nn, c->max_node_bytes / (NODE_WORDS * 8)), 64);
        CKC(cudaMalloc(&c->nodes, nn * NODE_WORDS * 8));
        CKC(cudaMemset(c->nodes, 0, nn * NODE_WORDS * 8));
        c->nn = nn;
    }
    c->rslots_hint = cfg->table_slots;
    uint64_t slots = cfg->table_slots ? cfg->table_slots : (1ull << 22);
    slots = std::min<uint64_t>(slots, c->max_table_bytes / 64 * BUCKET_SLOTS);
    slots = std::max<uint64_t>(slots, 1024);
    if (alloc_table(c, slots, 0) != SPL_OK) {
        g_create_error = c->err;
        delete c;
        return SPL_E_NOMEM;
    }
    CKC(cudaDeviceSynchronize());
#undef CKC
    *out = c;
    return SPL_OK;
}

int32_t spl_destroy(spl_ctx *c) {
    if (!c) return SPL_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    cudaFree(c->table); cudaFree(c->d_tabs); cudaFree(c->d_takes_idx); cudaFree(c->d_takes_edges);
    cudaFree(c->d_lut); cudaFree(c->d_ctr); cudaFreeHost(c->h_ctr); cudaFree(c->d_sel); cudaFreeHost(c->h_sel);
    cudaFree(c->rtable); cudaFree(c->d_hist); cudaFree(c->d_dict); cudaFree(c->d_dict2); cudaFree(c->d_dest); cudaFree(c->nodes); cudaFree(c->d_gemrank); cudaFree(c->d_rankgems);
    for (auto *b : c->pool_links) delete b;
    for (auto &ev : c->ev) if (ev) cudaEventDestroy(ev);
    delete c;
    return SPL_OK;
}

static int reset_visited(spl_ctx *c, cudaStream_t st) {
    CK(c, enter_device(c));
    CK(c, cudaMemsetAsync(c->table, 0, c->nb * 64, st));
    if (c->node_occ) CK(c, cudaMemsetAsync(c->nodes, 0, c->nn * NODE_WORDS * 8, st));
    c->occupied = 0;
    c->node_occ = 0;
    c->epoch = 0;
    return SPL_OK;
}

int32_t spl_reset_visited(spl_ctx *c, void *stream) {
    if (!c) return SPL_E_INVALID;
    NO_LIVE_SOLVER(c, "spl_reset_visited");
    return reset_visited(c, (cudaStream_t)stream);
}

int32_t spl_set_identity(spl_ctx *c, int32_t identity) {
    if (!c) return SPL_E_INVALID;
    if (identity != SPL_IDENT_KEY && identity != SPL_IDENT_PYHASH) return fail(c, SPL_E_INVALID, "spl_set_identity: unknown identity %d", identity);
    NO_LIVE_SOLVER(c, "spl_set_identity");
    c->identity = identity == SPL_IDENT_PYHASH ? IDENT_PYHASH : IDENT_KEY;
    return SPL_OK;
}

int32_t spl_set_link_budget(spl_ctx *c, uint64_t device_bytes) {
    if (!c) return SPL_E_INVALID;
    NO_LIVE_SOLVER(c, "spl_set_link_budget");
    c->link_budget = device_bytes;
    return SPL_OK;
}

int32_t spl_spilled_bytes(spl_ctx *c, int64_t *n_host) {
    if (!c || !n_host) return SPL_E_INVALID;
    *n_host = (int64_t)c->spilled_bytes;
    return SPL_OK;
}

int32_t spl_pyhash(spl_ctx *c, const spl_key *keys, int64_t n, uint64_t *out, void *stream) {
    if (!c || n < 0 || (n && (!keys || !out))) return fail(c, SPL_E_INVALID, "spl_pyhash: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    if (n == 0) return SPL_OK;
    pyhash_kernel<<<nblk(n), TILE, 0, st>>>(keys, n, out);
    ++c->launches;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

int32_t spl_visited_count(spl_ctx *c, int64_t *n_host) {
    if (!c || !n_host) return SPL_E_INVALID;
    *n_host = (int64_t)c->occupied;
    return SPL_OK;
}

int32_t spl_launch_count(spl_ctx *c, int64_t *n_host) {
    if (!c || !n_host) return SPL_E_INVALID;
    *n_host = c->launches;
    return SPL_OK;
}

int32_t spl_transfer_bytes(spl_ctx *c, int64_t *h2d_host, int64_t *d2h_host) {
    if (!c || !h2d_host || !d2h_host) return SPL_E_INVALID;
    *h2d_host = c->h2d_bytes;
    *d2h_host = c->d2h_bytes;
    return SPL_OK;
}

// ------------------------------------------------------------------ internal helpers
static int zero_ctr(spl_ctx *c, cudaStream_t st) {
    Counters z;
    memset(&z, 0, sizeof z);
    z.sk_min = ~0ull;
    z.goal_rank = 0x7fffffffffffffffll;
    z.key_and[0] = z.key_and[1] = ~0ull;
    *c->h_ctr = z;
    CK(c, cudaMemcpyAsync(c->d_ctr, c->h_ctr, sizeof z, cudaMemcpyHostToDevice, st));
    c->h2d_bytes += sizeof z;
    return SPL_OK;
}
static int read_ctr(spl_ctx *c, cudaStream_t st) {
    CK(c, cudaMemcpyAsync(c->h_ctr, c->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    c->d2h_bytes += sizeof(Counters);
    return SPL_OK;
}
static int reset_ticket(spl_ctx *c, int id, cudaStream_t st) {
    CK(c, cudaMemsetAsync(&c->d_ctr->ticket[id], 0, 4, st));
    return SPL_OK;
}
static int prep_status(spl_ctx *c, int which, size_t ntiles, cudaStream_t st) {
    CK(c, c->status[which].ensure((ntiles + 1) * 8, 0, st));
    CK(c, cudaMemsetAsync(c->status[which].p, 0, (ntiles + 1) * 8, st));
    return SPL_OK;
}

// grow the visited table so that `need` more inserts keep the load factor <= 0.7 (if memory allows)
static int ensure_table(spl_ctx *c, uint64_t need, cudaStream_t st) {
    while ((double)(c->occupied + need) > 0.7 * (double)c->cap) {
        uint64_t nnb = std::min<uint64_t>(c->nb * 2, MAX_BUCKETS);
        if (nnb * 64 > c->max_table_bytes) nnb = c->max_table_bytes / 64;
        if (nnb <= c->nb + c->nb / 8) break;  // cannot grow meaningfully
        uint64_t *nt = nullptr;
        cudaError_t e = cudaMalloc(&nt, nnb * 64);
        if (e != cudaSuccess) { cudaGetLastError(); break; }
        CK(c, cudaMemsetAsync(nt, 0, nnb * 64, st));
        rehash_kernel<<<nblk((int64_t)c->cap), TILE, 0, st>>>(c->table, c->nb, nt, nnb, c->d_ctr);
        ++c->launches;
        CK(c, cudaGetLastError());
        unsigned int rehash_err = 0;  // read the flag now: callers zero the counters before their own launches
        CK(c, cudaMemcpyAsync(&rehash_err, &c->d_ctr->error, 4, cudaMemcpyDeviceToHost, st));
        CK(c, cudaStreamSynchronize(st));
        if (rehash_err) {
            cudaFree(nt);
            return fail(c, SPL_E_TABLE_FULL, "rehash into %llu buckets overflowed a probe sequence", (unsigned long long)nnb);
        }
        cudaFree(c->table);
        c->table = nt;
        c->nb = nnb;
        c->cap = nnb * BUCKET_SLOTS;
    }
    if (c->occupied + need / 8 > c->cap - c->cap / 16)
        return fail(c, SPL_E_TABLE_FULL, "visited table full: %llu occupied + %llu candidates vs %llu slots (max_table_bytes=%llu)",
                    (unsigned long long)c->occupied, (unsigned long long)need, (unsigned long long)c->cap,
                    (unsigned long long)c->max_table_bytes);
    return SPL_OK;
}

static int next_epoch(spl_ctx *c, uint64_t &tag) {
    if (c->epoch >= TAG_MAX) return fail(c, SPL_E_STATE, "epoch tag exhausted");
    tag = ++c->epoch;
    return SPL_OK;
}

// count + offsets for parents front[0..n): leaves total in h_ctr->total_cands (synchronises)
static int run_count(spl_ctx *c, const Rec *front, int64_t n, cudaStream_t st) {
    const unsigned nt = nblk(n);
    CK(c, c->off.ensure((size_t)n * 4 + 4, 0, st));
    CKS(c, prep_status(c, 0, nt, st));
    CKS(c, reset_ticket(c, 0, st));
    count_scan_kernel<<<nt, TILE, 0, st>>>(front, n, c->d_tabs, c->d_takes_idx, c->off.as<uint32_t>(),
                                            c->status[0].as<uint64_t>(), c->d_ctr, 0);
    ++c->launches;
    CK(c, cudaGetLastError());
    return read_ctr(c, st);
}

// det policy: split the score ties at the threshold (d_sel->prefix) by key, larger first
static int select_ties_by_key(spl_ctx *c, const uint64_t *sk, const uint64_t *kb, int ks, int64_t n, int64_t k,
                              uint64_t sk_min, cudaStream_t st) {
    const unsigned grid = std::min<unsigned>(nblk(n), 148 * 8);
    const int lt = c->tie_link_top;
    if (lt) {
        // arrival-order ties: collect the tie words (2^lt - 1 - link) of the threshold score once, then select among them
        const int64_t ntie = (int64_t)c->h_sel->tie_count;
        CK(c, c->rtmp.ensure((size_t)ntie * 8 + 8, 0, st));
        CK(c, cudaMemsetAsync(&c->d_ctr->n_ties, 0, 8, st));
        tie_collect_kernel<<<grid, TILE, 0, st>>>(sk, kb, ks, n, sk_min, c->d_sel, lt, c->rtmp.as<uint64_t>(), c->d_ctr);
        ++c->launches;
        const unsigned tgrid = std::min<unsigned>(nblk(ntie), 148 * 8);
        int top = lt, first = 1;
        while (top > 0) {
            const int bits = std::min(SEL_BITS, top), shift = top - bits;
            sel_hist_kernel<3><<<tgrid, TILE, 0, st>>>(nullptr, c->rtmp.as<uint64_t>(), 1, ntie, 0, shift, bits, first, c->d_sel, c->d_hist);
            sel_pick_kernel<<<1, 1024, 0, st>>>(c->d_hist, 2, shift, first, 0, (uint64_t)k, c->d_sel);
            c->launches += 2;
            first = 0;
            top = shift;
        }
    } else
    for (int word = 1; word <= 2; ++word) {
        int top = word == 1 ? 41 : 64, first = 1;
        while (top > 0) {
            const int bits = std::min(SEL_BITS, top), shift = top - bits;
            if (word == 1) sel_hist_kernel<1><<<grid, TILE, 0, st>>>(sk, kb, ks, n, sk_min, shift, bits, first, c->d_sel, c->d_hist);
            else sel_hist_kernel<2><<<grid, TILE, 0, st>>>(sk, kb, ks, n, sk_min, shift, bits, first, c->d_sel, c->d_hist);
            sel_pick_kernel<<<1, 1024, 0, st>>>(c->d_hist, word, shift, first, 0, (uint64_t)k, c->d_sel);
            c->launches += 2;
            first = 0;
            top = shift;
        }
    }
    CK(c, cudaGetLastError());
    c->h_sel->tie_count = ~0ull;  // marks "threshold key valid"
    return SPL_OK;
}

// radix select of the k-th largest element under (score desc[, key desc]); leaves the thresholds and
// the tie quota in d_sel / h_sel (score in x = sk - sk_min space).
static int run_select(spl_ctx *c, const uint64_t *sk, const uint64_t *kb, int ks, int64_t n, int64_t k, uint64_t sk_min,
                      uint64_t sk_max, int det, cudaStream_t st) {
    const int nbits = bitlen(sk_max - sk_min);
    const unsigned grid = std::min<unsigned>(nblk(n), 148 * 8);
    memset(c->h_sel, 0, sizeof(SelState));
    c->h_sel->k_rem = (unsigned long long)k;
    c->h_sel->tie_count = (unsigned long long)n;  // nbits == 0: every score equal
    c->h_sel->rank_t = ~0ull;
    CK(c, cudaMemcpyAsync(c->d_sel, c->h_sel, sizeof(SelState), cudaMemcpyHostToDevice, st));
    int top = nbits, first = 1, init_k = 1;
    while (top > 0) {
        const int bits = std::min(SEL_BITS, top), shift = top - bits;
        sel_hist_kernel<0><<<grid, TILE, 0, st>>>(sk, kb, ks, n, sk_min, shift, bits, first, c->d_sel, c->d_hist);
        sel_pick_kernel<<<1, 1024, 0, st>>>(c->d_hist, 0, shift, first, init_k, (uint64_t)k, c->d_sel);
        c->launches += 2;
        first = init_k = 0;
        top = shift;
    }
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(c->h_sel, c->d_sel, sizeof(SelState), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    c->d2h_bytes += sizeof(SelState);
    if (det && c->h_sel->k_rem < c->h_sel->tie_count) CKS(c, select_ties_by_key(c, sk, kb, ks, n, k, sk_min, st));
    return SPL_OK;
}

// Dictionary select (few distinct scores): count the distinct scores, rank them, find the k-th largest
// element.  Leaves d_sel / h_sel as run_select does plus rank_t; *used = 0 when the level has more
// than DICT_MAX distinct scores (nothing decided: the caller falls back to the radix select).
static int run_dict_select(spl_ctx *c, const uint64_t *sk, const uint64_t *kb, int ks, int64_t n, int64_t k,
                           uint64_t sk_min, int det, int *used, cudaStream_t st, int keep_all = -1) {
    if (keep_all < 0) keep_all = k >= n;
    *used = 0;
    if (c->dict_skip > 0) { --c->dict_skip; return SPL_OK; }
    const unsigned grid = std::min<unsigned>(nblk(n), 148 * 8);
    CK(c, cudaMemsetAsync(c->d_dict->key, 0xFF, sizeof(c->d_dict->key), st));
    CK(c, cudaMemsetAsync(c->d_dict->cnt, 0, sizeof(ScoreDict) - sizeof(c->d_dict->key), st));
    memset(c->h_sel, 0, sizeof(SelState));
    c->h_sel->rank_t = ~0ull;
    CK(c, cudaMemcpyAsync(c->d_sel, c->h_sel, sizeof(SelState), cudaMemcpyHostToDevice, st));
    dict_build_kernel<<<grid, TILE, 0, st>>>(sk, n, c->d_dict);
    dict_rank_kernel<<<1, 1024, 0, st>>>(c->d_dict, (uint64_t)k, sk_min, c->d_sel);
    c->launches += 2;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(c->h_sel, c->d_sel, sizeof(SelState), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    c->d2h_bytes += sizeof(SelState);
    if (c->h_sel->rank_t == ~0ull) {
        c->dict_skip = 3;
        return SPL_OK;
    }
    *used = 1;
    if (det && !keep_all && c->h_sel->k_rem < c->h_sel->tie_count) CKS(c, select_ties_by_key(c, sk, kb, ks, n, k, sk_min, st));
    return SPL_OK;
}

static int onesweep_sort(spl_ctx *c, uint64_t *const k[2], uint32_t *const pl[2], int64_t n, int lo_bit, int nbits, int *cur, cudaStream_t st);

// one stable LSD pass over `kept` elements: digit = (dig >> shift) & 255
static int sort_pass(spl_ctx *c, int wide, int cur, const uint64_t *dig, int64_t kept, int shift, unsigned nt,
                     size_t msz, cudaStream_t st) {
    const unsigned st_tiles = nblk((int64_t)msz, TILE * SCAN_ITEMS);
    sort_hist_kernel<<<nt, TILE, 0, st>>>(dig, kept, shift, c->matrix.as<uint32_t>(), nt);
    CKS(c, prep_status(c, 1, st_tiles, st));
    CKS(c, reset_ticket(c, 2, st));
    scan_u32_kernel<<<st_tiles, TILE, 0, st>>>(c->matrix.as<uint32_t>(), c->matrix2.as<uint32_t>(), (int64_t)msz,
                                                c->status[1].as<uint64_t>(), c->d_ctr, 2);
    if (wide)
        sort_scatter_kernel<true><<<nt, TILE, 0, st>>>(dig, c->y[cur].as<uint64_t>(), c->idx[cur].as<uint32_t>(), kept, shift,
                                                        c->matrix2.as<uint32_t>(), nt, c->y[cur ^ 1].as<uint64_t>(),
                                                        c->idx[cur ^ 1].as<uint32_t>(), c->kl[cur].as<uint64_t>(),
                                                        c->kh[cur].as<uint64_t>(), c->kl[cur ^ 1].as<uint64_t>(),
                                                        c->kh[cur ^ 1].as<uint64_t>());
    else
        sort_scatter_kernel<false><<<nt, TILE, 0, st>>>(dig, c->y[cur].as<uint64_t>(), c->idx[cur].as<uint32_t>(), kept, shift,
                                                         c->matrix2.as<uint32_t>(), nt, c->y[cur ^ 1].as<uint64_t>(),
                                                         c->idx[cur ^ 1].as<uint32_t>(), nullptr, nullptr, nullptr, nullptr);
    c->launches += 3;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

// beam cut + rank sort.  stable: arrival-order cut, stable descending sort by score.  det: cut and sort
// by (score desc, key desc).  On return idx[*which] holds the kept source indices in rank order.
static int do_cut_sort(spl_ctx *c, const uint64_t *sk, const uint64_t *kb, int ks, int64_t n, int64_t cap_kept, int keep_all,
                       int all_ties, uint64_t sk_min, uint64_t sk_max, int det, int use_dict, int *which, int64_t *kept_out,
                       cudaStream_t st);

// n = elements of sk / recs; n_valid of them are states (the rest are unused slots, sk == 0)
static int run_cut_sort(spl_ctx *c, const uint64_t *sk, const Rec *recs, int64_t n, int64_t k, uint64_t sk_min,
                        uint64_t sk_max, int det, int *which, int64_t *kept_out, cudaStream_t st, int64_t n_valid = -1) {
    if (n_valid < 0) n_valid = n;
    const int keep_all = n_valid <= k;
    const uint64_t *kb = reinterpret_cast<const uint64_t *>(recs);
    const int ks = 4;
    int use_dict = 0;
    CKS(c, run_dict_select(c, sk, kb, ks, n, std::min(n_valid, k), sk_min, det, &use_dict, st, keep_all));
    if (!keep_all && !use_dict) CKS(c, run_select(c, sk, kb, ks, n, k, sk_min, sk_max, det, st));
    const int all_ties = keep_all || c->h_sel->tie_count != ~0ull;
    CKS(c, do_cut_sort(c, sk, kb, ks, n, std::min(n_valid, k), keep_all, all_ties, sk_min, sk_max, det, use_dict, which, kept_out, st));
    if (*kept_out != std::min(n_valid, k))
        return fail(c, SPL_E_CUDA, "internal: cut kept %lld states, expected %lld", (long long)*kept_out, (long long)std::min(n_valid, k));
    return SPL_OK;
}

// cut by the thresholds currently in d_sel / h_sel (x-space score threshold `prefix`, arrival quota
// `k_rem` for the stable policy, key threshold khi/klo for det), then rank-sort the survivors.
static int do_cut_sort(spl_ctx *c, const uint64_t *sk, const uint64_t *kb, int ks, int64_t n, int64_t cap_kept, int keep_all,
                       int all_ties, uint64_t sk_min, uint64_t sk_max, int det, int use_dict, int *which, int64_t *kept_out,
                       cudaStream_t st) {
    int64_t kept = cap_kept;
    const uint64_t T = keep_all ? 0 : c->h_sel->prefix;
    for (int b = 0; b < 2; ++b) {
        CK(c, c->y[b].ensure((size_t)kept * 8 + 8, 0, st));
        CK(c, c->idx[b].ensure((size_t)kept * 4 + 4, 0, st));
        if (det) {
            CK(c, c->kl[b].ensure((size_t)kept * 8 + 8, 0, st));
            CK(c, c->kh[b].ensure((size_t)kept * 8 + 8, 0, st));
        }
    }
    const int lt0 = det == 1 ? c->tie_link_top : 0;
    const unsigned ct = nblk(n, TILE * ((lt0 && use_dict) ? CUTP_ITEMS : CUT_ITEMS));
    CKS(c, prep_status(c, 1, ct, st));
    CKS(c, prep_status(c, 2, ct, st));
    CKS(c, reset_ticket(c, 1, st));
    uint64_t vary_lo = 0, vary_hi = 0;
    const int lt = det == 1 ? c->tie_link_top : 0;
    const int pack = lt && use_dict;  // one sort word: score rank << lt | link
    if (pack) {
        CKS(c, zero_ctr(c, st));
        cut_pack_kernel<<<ct, TILE, 0, st>>>(sk, kb, ks, n, sk_min, keep_all, all_ties, c->d_sel, c->d_dict, lt, c->y[0].as<uint64_t>(),
                                              c->idx[0].as<uint32_t>(), c->status[2].as<uint64_t>(), c->d_ctr, 1);
        ++c->launches;
        CK(c, cudaGetLastError());
    } else if (det) {
        CKS(c, zero_ctr(c, st));
        if (use_dict)
            cut_det_kernel<true><<<ct, TILE, 0, st>>>(sk, kb, ks, n, sk_min, sk_max, keep_all, all_ties, c->d_sel, c->d_dict,
                                                       c->y[0].as<uint64_t>(), c->kl[0].as<uint64_t>(), c->kh[0].as<uint64_t>(),
                                                       c->idx[0].as<uint32_t>(), c->status[2].as<uint64_t>(), c->d_ctr, 1, lt, pack);
        else
            cut_det_kernel<false><<<ct, TILE, 0, st>>>(sk, kb, ks, n, sk_min, sk_max, keep_all, all_ties, c->d_sel, nullptr,
                                                        c->y[0].as<uint64_t>(), c->kl[0].as<uint64_t>(), c->kh[0].as<uint64_t>(),
                                                        c->idx[0].as<uint32_t>(), c->status[2].as<uint64_t>(), c->d_ctr, 1, lt, 0);
        ++c->launches;
        CK(c, cudaGetLastError());
        if (det == 1 && !pack) {  // which key bits vary among the survivors (constant digits need no sort pass)
            CKS(c, read_ctr(c, st));
            vary_lo = c->h_ctr->key_or[0] ^ c->h_ctr->key_and[0];
            vary_hi = (c->h_ctr->key_or[1] ^ c->h_ctr->key_and[1]) & HI_KEY_MASK;
        }
    } else {
        if (use_dict)
            cut_kernel<true><<<ct, TILE, 0, st>>>(sk, n, sk_min, sk_max, keep_all, c->d_sel, c->d_dict, c->y[0].as<uint64_t>(),
                                                   c->idx[0].as<uint32_t>(), c->status[1].as<uint64_t>(),
                                                   c->status[2].as<uint64_t>(), c->d_ctr, 1);
        else
            cut_kernel<false><<<ct, TILE, 0, st>>>(sk, n, sk_min, sk_max, keep_all, c->d_sel, nullptr, c->y[0].as<uint64_t>(),
                                                    c->idx[0].as<uint32_t>(), c->status[1].as<uint64_t>(),
                                                    c->status[2].as<uint64_t>(), c->d_ctr, 1);
        ++c->launches;
        CK(c, cudaGetLastError());
    }
    {   // survivors actually written, from the inclusive prefixes published by the last cut tile
        uint64_t last = 0, last_tie = 0;
        CK(c, cudaMemcpyAsync(&last, c->status[2].as<uint64_t>() + (ct - 1), 8, cudaMemcpyDeviceToHost, st));
        if (!det && !keep_all)
            CK(c, cudaMemcpyAsync(&last_tie, c->status[1].as<uint64_t>() + (ct - 1), 8, cudaMemcpyDeviceToHost, st));
        CK(c, cudaStreamSynchronize(st));
        c->d2h_bytes += 16;
        constexpr uint64_t VAL = (1ull << 62) - 1;
        kept = (int64_t)(last & VAL);  // det: kept states; stable: states above the threshold
        if (!det && !keep_all) kept += (int64_t)std::min<uint64_t>(last_tie & VAL, c->h_sel->k_rem);
        if (kept > cap_kept) return fail(c, SPL_E_CUDA, "internal: cut kept %lld > capacity %lld", (long long)kept, (long long)cap_kept);
    }
    // y = sk_max - sk in [0, sk_max - sk_min - T], or (dictionary) the score's descending rank in [0, rank_t]
    const int nbits = (use_dict ? bitlen(c->h_sel->rank_t) : bitlen((sk_max - sk_min) - T)) + (pack ? lt : 0);
    int cur = 0;
    const unsigned nt = nblk(kept, SORT_TILE);
    if (kept > 1 && (nbits > 0 || vary_lo || vary_hi)) {
        const size_t msz = (size_t)SORT_BINS * nt;
        CK(c, c->matrix.ensure(msz * 4, 0, st));
        CK(c, c->matrix2.ensure(msz * 4, 0, st));
        if (det == 1 && !pack) {  // least significant first: key.lo, key.hi, then the score
            for (int shift = 0; shift < 64; shift += SORT_BITS)
                if ((vary_lo >> shift) & 0xff) { CKS(c, sort_pass(c, 1, cur, c->kl[cur].as<uint64_t>(), kept, shift, nt, msz, st)); cur ^= 1; }
            for (int shift = 0; shift < 41; shift += SORT_BITS)
                if ((vary_hi >> shift) & 0xff) { CKS(c, sort_pass(c, 1, cur, c->kh[cur].as<uint64_t>(), kept, shift, nt, msz, st)); cur ^= 1; }
        }
        if (!(det == 1 && !pack) && kept < OS_MAX_ITEMS) {  // (sort word, slot) pairs: one-sweep passes of 10 bits
            uint64_t *const yk[2] = {c->y[0].as<uint64_t>(), c->y[1].as<uint64_t>()};
            uint32_t *const yi[2] = {c->idx[0].as<uint32_t>(), c->idx[1].as<uint32_t>()};
            CKS(c, onesweep_sort(c, yk, yi, kept, 0, nbits, &cur, st));
        } else {
            for (int shift = 0; shift < nbits; shift += SORT_BITS) {
                CKS(c, sort_pass(c, det == 1 && !pack, cur, c->y[cur].as<uint64_t>(), kept, shift, nt, msz, st));
                cur ^= 1;
            }
        }
    }
    if (det == 2 && kept > 0) {  // keys were in rank order already: the stable score sort kept them so; rebuild the key words
        key_words_kernel<<<nblk(kept), TILE, 0, st>>>(c->idx[cur].as<uint32_t>(), kept, kb, ks, c->kl[cur].as<uint64_t>(),
                                                       c->kh[cur].as<uint64_t>());
        ++c->launches;
        CK(c, cudaGetLastError());
    }
    *which = cur;
    *kept_out = kept;
    return SPL_OK;
}

// ------------------------------------------------------------------ stage operators
int32_t spl_expand(spl_ctx *c, const spl_key *keys, const uint64_t *aux, int64_t n, spl_key *ck, uint64_t *ca,
                   uint64_t *cl, int64_t cap, int64_t *n_out, void *stream) {
    if (!c || !n_out || n < 0 || n > (16ll << 20)) return fail(c, SPL_E_INVALID, "spl_expand: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    *n_out = 0;
    if (n == 0) return SPL_OK;
    CK(c, c->tmp_rec.ensure((size_t)n * 32, 0, st));
    pack_rec_kernel<<<nblk(n), TILE, 0, st>>>(keys, aux, n, c->tmp_rec.as<Rec>());
    ++c->launches;
    CKS(c, zero_ctr(c, st));
    CKS(c, run_count(c, c->tmp_rec.as<Rec>(), n, st));
    const int64_t total = (int64_t)c->h_ctr->total_cands;
    *n_out = total;
    if (total > cap) return fail(c, SPL_E_CAPACITY, "spl_expand: %lld successors, capacity %lld", (long long)total, (long long)cap);
    if (total == 0) return SPL_OK;
    CK(c, c->tmp_rec2.ensure((size_t)total * 32, 0, st));
    expand_kernel<MODE_LIST, IDENT_KEY><<<nblk(n), TILE, sizeof(ExpandSmem2), st>>>(
        c->tmp_rec.as<Rec>(), n, c->d_tabs, c->d_takes_idx, c->d_takes_edges, c->off.as<uint32_t>(), (uint32_t)total,
        nullptr, 0, 0, nullptr, c->tmp_rec2.as<Rec>(), 0, c->d_ctr);
    unpack_rec_kernel<<<nblk(total), TILE, 0, st>>>(c->tmp_rec2.as<Rec>(), total, ck, ca, cl);
    c->launches += 2;
    CK(c, cudaGetLastError());
    CK(c, cudaStreamSynchronize(st));
    return SPL_OK;
}

int32_t spl_dedup(spl_ctx *c, const spl_key *ck, const uint64_t *ca, int64_t n, spl_key *uk, uint64_t *ua, int64_t *usrc,
                  int64_t *n_out, void *stream) {
    if (!c || !n_out || n < 0 || n >= (1ll << 32)) return fail(c, SPL_E_INVALID, "spl_dedup: bad arguments");
    NO_LIVE_SOLVER(c, "spl_dedup");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    *n_out = 0;
    if (n == 0) return SPL_OK;
    CKS(c, zero_ctr(c, st));
    CKS(c, ensure_table(c, (uint64_t)n, st));
    uint64_t tag;
    CKS(c, next_epoch(c, tag));
    CK(c, c->cand_slot.ensure((size_t)n * 4, 0, st));
    probe_list_kernel<IDENT_KEY><<<nblk(n), TILE, 0, st>>>(ck, n, c->table, c->nb, tag, c->cand_slot.as<uint32_t>(), c->d_ctr);
    ++c->launches;
    CK(c, cudaGetLastError());
    CKS(c, read_ctr(c, st));
    if (c->h_ctr->error) return fail(c, SPL_E_TABLE_FULL, "spl_dedup: probe overflow (table full)");
    const int64_t n_new = (int64_t)c->h_ctr->n_new;
    c->occupied += n_new;
    *n_out = n_new;
    if (n_new == 0) return SPL_OK;
    CK(c, c->tmp_rec.ensure((size_t)n_new * 32, 0, st));
    const unsigned nt = nblk(n, TILE * 32);
    CKS(c, prep_status(c, 0, nt, st));
    CKS(c, reset_ticket(c, 0, st));
    resolve_kernel<SRC_LIST, false><<<nt, TILE, sizeof(ResolveSmem), st>>>(
        nullptr, 0, c->d_tabs, c->d_takes_idx, c->d_takes_edges, nullptr, (uint32_t)n, c->table,
        c->cand_slot.as<uint32_t>(), ck, ca, 0, 0, c->tmp_rec.as<Rec>(), nullptr, usrc, 0, 0, c->luts,
        c->status[0].as<uint64_t>(), c->d_ctr, 0);
    unpack_rec_kernel<<<nblk(n_new), TILE, 0, st>>>(c->tmp_rec.as<Rec>(), n_new, uk, ua, nullptr);
    c->launches += 2;
    CK(c, cudaGetLastError());
    CK(c, cudaStreamSynchronize(st));
    return SPL_OK;
}

int32_t spl_score(spl_ctx *c, int32_t heuristic, int32_t noise, const spl_key *keys, const uint64_t *aux, int64_t n,
                  double *scores, void *stream) {
    if (!c || n < 0) return fail(c, SPL_E_INVALID, "spl_score: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    if (n == 0) return SPL_OK;
    score_kernel<<<nblk(n), TILE, 0, st>>>(keys, aux, n, heuristic, noise, c->luts, scores);
    ++c->launches;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

int32_t spl_topk(spl_ctx *c, const double *scores, const spl_key *keys, int64_t n, int64_t k, int32_t tie_policy,
                 int64_t *out_idx, int64_t *n_out, void *stream) {
    if (!c || !n_out || n < 0 || n >= (1ll << 32) || k < 0) return fail(c, SPL_E_INVALID, "spl_topk: bad arguments");
    if (tie_policy != SPL_TIE_STABLE && tie_policy != SPL_TIE_KEY) return fail(c, SPL_E_INVALID, "spl_topk: unknown tie policy %d", tie_policy);
    if (tie_policy == SPL_TIE_KEY && !keys) return fail(c, SPL_E_INVALID, "spl_topk: SPL_TIE_KEY needs keys");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    *n_out = 0;
    if (n == 0 || k == 0) return SPL_OK;
    CKS(c, zero_ctr(c, st));
    CK(c, c->sk.ensure((size_t)n * 8, 0, st));
    flip_scores_kernel<<<nblk(n), TILE, 0, st>>>(scores, n, c->sk.as<uint64_t>(), c->d_ctr);
    ++c->launches;
    CKS(c, read_ctr(c, st));
    int which = 0;
    int64_t kept = 0;
    const uint64_t smin = c->h_ctr->sk_min, smax = c->h_ctr->sk_max;
    const Rec *recs = nullptr;
    if (tie_policy == SPL_TIE_KEY) {
        CK(c, c->tmp_rec.ensure((size_t)n * 32, 0, st));
        pack_rec_kernel<<<nblk(n), TILE, 0, st>>>(keys, nullptr, n, c->tmp_rec.as<Rec>());
        ++c->launches;
        recs = c->tmp_rec.as<Rec>();
    }
    CKS(c, run_cut_sort(c, c->sk.as<uint64_t>(), recs, n, k, smin, smax, tie_policy == SPL_TIE_KEY, &which, &kept, st));
    idx_widen_kernel<<<nblk(kept), TILE, 0, st>>>(c->idx[which].as<uint32_t>(), kept, out_idx);
    ++c->launches;
    CK(c, cudaGetLastError());
    CK(c, cudaStreamSynchronize(st));
    *n_out = kept;
    return SPL_OK;
}

// ---- multi-GPU building blocks ----------------------------------------------------------------
int32_t spl_owner_partition(spl_ctx *c, const spl_key *keys, int64_t n, int32_t n_ranks, int64_t *perm, int64_t *counts_host,
                            void *stream) {
    if (!c || !counts_host || n < 0 || n >= (1ll << 32) || n_ranks < 1 || n_ranks > 256)
        return fail(c, SPL_E_INVALID, "spl_owner_partition: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    for (int g = 0; g < n_ranks; ++g) counts_host[g] = 0;
    if (n == 0) return SPL_OK;
    for (int b = 0; b < 2; ++b) {
        CK(c, c->y[b].ensure((size_t)n * 8 + 8, 0, st));
        CK(c, c->idx[b].ensure((size_t)n * 4 + 4, 0, st));
    }
    owner_kernel<<<nblk(n), TILE, 0, st>>>(keys, n, (uint32_t)n_ranks, c->y[0].as<uint64_t>(), c->idx[0].as<uint32_t>());
    ++c->launches;
    const unsigned nt = nblk(n, SORT_TILE);
    const size_t msz = (size_t)SORT_BINS * nt;
    CK(c, c->matrix.ensure(msz * 4, 0, st));
    CK(c, c->matrix2.ensure(msz * 4, 0, st));
    CKS(c, sort_pass(c, 0, 0, c->y[0].as<uint64_t>(), n, 0, nt, msz, st));  // one stable pass: digit = owner
    idx_widen_kernel<<<nblk(n), TILE, 0, st>>>(c->idx[1].as<uint32_t>(), n, perm);
    ++c->launches;
    // per-owner counts = differences of the digit-major scanned histogram at tile 0
    std::vector<uint32_t> base(n_ranks + 1);
    for (int g = 0; g <= n_ranks; ++g)
        CK(c, cudaMemcpyAsync(&base[g], c->matrix2.as<uint32_t>() + (size_t)g * nt, 4, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    c->d2h_bytes += 4 * (n_ranks + 1);
    for (int g = 0; g < n_ranks; ++g) counts_host[g] = (g + 1 < SORT_BINS ? base[g + 1] : (uint32_t)n) - base[g];
    return SPL_OK;
}

// ---- row-based variants used by the sharded driver: rows are 32-byte records {lo, hi, aux, link}
int32_t spl_expand_rows(spl_ctx *c, const void *front_rows, int64_t n, int64_t rank_base, void *out_rows, int64_t cap,
                        int64_t *n_out, void *stream) {
    if (!c || !n_out || n < 0 || n > (16ll << 20)) return fail(c, SPL_E_INVALID, "spl_expand_rows: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    *n_out = 0;
    if (n == 0) return SPL_OK;
    CKS(c, zero_ctr(c, st));
    CKS(c, run_count(c, reinterpret_cast<const Rec *>(front_rows), n, st));
    const int64_t total = (int64_t)c->h_ctr->total_cands;
    *n_out = total;
    if (total > cap) return fail(c, SPL_E_CAPACITY, "spl_expand_rows: %lld successors, capacity %lld", (long long)total, (long long)cap);
    if (total == 0) return SPL_OK;
    expand_kernel<MODE_LIST, IDENT_KEY><<<nblk(n), TILE, sizeof(ExpandSmem2), st>>>(
        reinterpret_cast<const Rec *>(front_rows), n, c->d_tabs, c->d_takes_idx, c->d_takes_edges, c->off.as<uint32_t>(),
        (uint32_t)total, nullptr, 0, 0, nullptr, reinterpret_cast<Rec *>(out_rows), rank_base, c->d_ctr);
    ++c->launches;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

int32_t spl_route_keys(spl_ctx *c, const void *cand_rows, int64_t n, int32_t n_ranks, spl_key *send_keys, int64_t *counts_host,
                       void *stream) {
    if (!c || !counts_host || n < 0 || n >= (1ll << 32) || n_ranks < 1 || n_ranks > 255)
        return fail(c, SPL_E_INVALID, "spl_route_keys: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    for (int g = 0; g < n_ranks; ++g) counts_host[g] = 0;
    c->route_n = n;
    if (n == 0) return SPL_OK;
    for (int b = 0; b < 2; ++b) {
        CK(c, c->y[b].ensure((size_t)n * 8 + 8, 0, st));
        CK(c, c->idx[b].ensure((size_t)n * 4 + 4, 0, st));
    }
    const Rec *rows = reinterpret_cast<const Rec *>(cand_rows);
    if (n_ranks <= 32) {  // two kernels: per-tile owner histogram -> scan -> stable scatter of the keys
        const unsigned nt = nblk(n, PART_TILE);
        const size_t msz = (size_t)n_ranks * nt;
        CK(c, c->matrix.ensure(msz * 4 + 4, 0, st));
        CK(c, c->matrix2.ensure(msz * 4 + 4, 0, st));
        owner_hist_kernel<<<nt, TILE, 0, st>>>(rows, n, (uint32_t)n_ranks, c->matrix.as<uint32_t>(), nt);
        const unsigned st_tiles = nblk((int64_t)msz, TILE * SCAN_ITEMS);
        CKS(c, prep_status(c, 1, st_tiles, st));
        CKS(c, reset_ticket(c, 2, st));
        scan_u32_kernel<<<st_tiles, TILE, 0, st>>>(c->matrix.as<uint32_t>(), c->matrix2.as<uint32_t>(), (int64_t)msz,
                                                    c->status[1].as<uint64_t>(), c->d_ctr, 2);
        owner_scatter_kernel<<<nt, TILE, 0, st>>>(rows, n, (uint32_t)n_ranks, c->matrix2.as<uint32_t>(), nt, send_keys,
                                                   c->idx[1].as<uint32_t>());
        c->launches += 3;
        CK(c, cudaGetLastError());
        std::vector<uint32_t> base(n_ranks + 1);
        for (int g = 0; g < n_ranks; ++g)
            CK(c, cudaMemcpyAsync(&base[g], c->matrix2.as<uint32_t>() + (size_t)g * nt, 4, cudaMemcpyDeviceToHost, st));
        CK(c, cudaStreamSynchronize(st));
        base[n_ranks] = (uint32_t)n;
        c->d2h_bytes += 4 * n_ranks;
        for (int g = 0; g < n_ranks; ++g) counts_host[g] = base[g + 1] - base[g];
        return SPL_OK;
    }
    owner_rows_kernel<<<nblk(n), TILE, 0, st>>>(rows, n, (uint32_t)n_ranks, c->y[0].as<uint64_t>(), c->idx[0].as<uint32_t>());
    ++c->launches;
    const unsigned nt = nblk(n, SORT_TILE);
    const size_t msz = (size_t)SORT_BINS * nt;
    CK(c, c->matrix.ensure(msz * 4, 0, st));
    CK(c, c->matrix2.ensure(msz * 4, 0, st));
    CKS(c, sort_pass(c, 0, 0, c->y[0].as<uint64_t>(), n, 0, nt, msz, st));  // stable: digit = owner; perm stays in idx[1]
    gather_keys_kernel<<<nblk(n), TILE, 0, st>>>(rows, c->idx[1].as<uint32_t>(), n, send_keys);
    ++c->launches;
    std::vector<uint32_t> base(n_ranks + 1);
    for (int g = 0; g <= n_ranks; ++g)
        CK(c, cudaMemcpyAsync(&base[g], c->matrix2.as<uint32_t>() + (size_t)g * nt, 4, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    c->d2h_bytes += 4 * (n_ranks + 1);
    for (int g = 0; g < n_ranks; ++g) counts_host[g] = base[g + 1] - base[g];
    return SPL_OK;
}

int32_t spl_dedup_flags(spl_ctx *c, const spl_key *keys, int64_t n, uint8_t *flags, void *stream) {
    if (!c || n < 0 || n >= (1ll << 32)) return fail(c, SPL_E_INVALID, "spl_dedup_flags: bad arguments");
    NO_LIVE_SOLVER(c, "spl_dedup_flags");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    if (n == 0) return SPL_OK;
    CKS(c, zero_ctr(c, st));
    CKS(c, ensure_table(c, (uint64_t)n, st));
    uint64_t tag;
    CKS(c, next_epoch(c, tag));
    CK(c, c->cand_slot.ensure((size_t)n * 4, 0, st));
    probe_list_kernel<IDENT_KEY><<<nblk(n), TILE, 0, st>>>(keys, n, c->table, c->nb, tag, c->cand_slot.as<uint32_t>(), c->d_ctr);
    win_flags_kernel<<<nblk(n), TILE, 0, st>>>(c->cand_slot.as<uint32_t>(), c->table, n, flags);
    c->launches += 2;
    CK(c, cudaGetLastError());
    CKS(c, read_ctr(c, st));
    if (c->h_ctr->error) return fail(c, SPL_E_TABLE_FULL, "spl_dedup_flags: probe overflow (table full)");
    c->occupied += c->h_ctr->n_new;
    return SPL_OK;
}

int32_t spl_compact_winners(spl_ctx *c, const void *cand_rows, int64_t n, const uint8_t *flags_partitioned, void *out_rows,
                            int64_t *n_out, void *stream) {
    if (!c || !n_out || n < 0 || n != c->route_n) return fail(c, SPL_E_STATE, "spl_compact_winners: must follow spl_route_keys on the same candidates");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    *n_out = 0;
    if (n == 0) return SPL_OK;
    CK(c, c->rtmp.ensure((size_t)n + 8, 0, st));
    scatter_flags_kernel<<<nblk(n), TILE, 0, st>>>(flags_partitioned, c->idx[1].as<uint32_t>(), n, c->rtmp.as<uint8_t>());
    const unsigned nt = nblk(n, TILE * 8);
    CKS(c, zero_ctr(c, st));
    CKS(c, prep_status(c, 0, nt, st));
    compact_rows_kernel<<<nt, TILE, 0, st>>>(reinterpret_cast<const Rec *>(cand_rows), c->rtmp.as<uint8_t>(), n,
                                              reinterpret_cast<Rec *>(out_rows), c->status[0].as<uint64_t>(), c->d_ctr, 0);
    c->launches += 2;
    CK(c, cudaGetLastError());
    CKS(c, read_ctr(c, st));
    *n_out = (int64_t)c->h_ctr->n_emitted;
    return SPL_OK;
}

int32_t spl_score_rows(spl_ctx *c, int32_t heuristic, int32_t noise, const void *rows, int64_t n, double *scores,
                       const uint8_t *draws, void *stream) {
    if (!c || n < 0 || (noise == SPL_NOISE_EXTERNAL && !draws)) return fail(c, SPL_E_INVALID, "spl_score_rows: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    if (n == 0) return SPL_OK;
    score_rows_kernel<<<nblk(n), TILE, 0, st>>>(reinterpret_cast<const Rec *>(rows), n, heuristic, noise, c->luts, scores, draws);
    ++c->launches;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

int32_t spl_move_rows(spl_ctx *c, const void *rows, const int64_t *idx, int64_t n, void *out_rows, int32_t scatter, void *stream) {
    if (!c || n < 0) return fail(c, SPL_E_INVALID, "spl_move_rows: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    if (n == 0) return SPL_OK;
    move_rows_kernel<<<nblk(n), TILE, 0, st>>>(reinterpret_cast<const Rec *>(rows), idx, n, reinterpret_cast<Rec *>(out_rows), scatter);
    ++c->launches;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

int32_t spl_dtopk_begin(spl_ctx *c, const double *scores, const spl_key *keys, int64_t n, uint64_t *sk_min_host,
                        uint64_t *sk_max_host, void *stream) {
    if (!c || !sk_min_host || !sk_max_host || n < 0 || n >= (1ll << 32)) return fail(c, SPL_E_INVALID, "spl_dtopk_begin: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    c->dtopk_n = n;
    c->dtopk_recs = false;
    *sk_min_host = ~0ull;
    *sk_max_host = 0;
    if (n == 0) return SPL_OK;
    CKS(c, zero_ctr(c, st));
    CK(c, c->sk.ensure((size_t)n * 8, 0, st));
    flip_scores_kernel<<<nblk(n), TILE, 0, st>>>(scores, n, c->sk.as<uint64_t>(), c->d_ctr);
    ++c->launches;
    if (keys) {  // the caller's key array must stay alive until spl_dtopk_cut
        c->dtopk_keys = reinterpret_cast<const uint64_t *>(keys);
        c->dtopk_recs = true;
    }
    CKS(c, read_ctr(c, st));
    *sk_min_host = c->h_ctr->sk_min;
    *sk_max_host = c->h_ctr->sk_max;
    return SPL_OK;
}

int32_t spl_dtopk_hist(spl_ctx *c, int32_t word, int32_t shift, int32_t bits, int32_t first, uint64_t sk_min_global,
                       uint32_t **hist_dev_out, void *stream) {
    if (!c || !hist_dev_out || word < 0 || word > 2 || bits < 1 || bits > SEL_BITS) return fail(c, SPL_E_INVALID, "spl_dtopk_hist: bad arguments");
    if (word > 0 && !c->dtopk_recs && c->dtopk_n) return fail(c, SPL_E_STATE, "spl_dtopk_hist: key passes need keys in spl_dtopk_begin");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    const int64_t n = c->dtopk_n;
    *hist_dev_out = c->d_hist;
    if (n == 0) return SPL_OK;
    const unsigned grid = std::min<unsigned>(nblk(n), 148 * 8);
    const uint64_t *sk = c->sk.as<uint64_t>();
    const uint64_t *kb = c->dtopk_keys;
    if (word == 0) sel_hist_kernel<0><<<grid, TILE, 0, st>>>(sk, kb, 2, n, sk_min_global, shift, bits, first, c->d_sel, c->d_hist);
    else if (word == 1) sel_hist_kernel<1><<<grid, TILE, 0, st>>>(sk, kb, 2, n, sk_min_global, shift, bits, first, c->d_sel, c->d_hist);
    else sel_hist_kernel<2><<<grid, TILE, 0, st>>>(sk, kb, 2, n, sk_min_global, shift, bits, first, c->d_sel, c->d_hist);
    ++c->launches;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

int32_t spl_dtopk_pick(spl_ctx *c, int32_t word, int32_t shift, int32_t first, int32_t init_k, int64_t k, void *stream) {
    if (!c) return SPL_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    sel_pick_kernel<<<1, 1024, 0, st>>>(c->d_hist, word, shift, first, init_k, (uint64_t)k, c->d_sel);
    ++c->launches;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

int32_t spl_dtopk_get(spl_ctx *c, uint64_t state_host[6], void *stream) {
    if (!c || !state_host) return SPL_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    CK(c, cudaMemcpyAsync(c->h_sel, c->d_sel, sizeof(SelState), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    c->d2h_bytes += sizeof(SelState);
    memcpy(state_host, c->h_sel, 6 * sizeof(uint64_t));  // the six fields of the ABI; rank_t is internal
    return SPL_OK;
}

int32_t spl_dtopk_set(spl_ctx *c, const uint64_t state_host[6], void *stream) {
    if (!c || !state_host) return SPL_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    memcpy(c->h_sel, state_host, 6 * sizeof(uint64_t));
    c->h_sel->rank_t = ~0ull;
    CK(c, cudaMemcpyAsync(c->d_sel, c->h_sel, sizeof(SelState), cudaMemcpyHostToDevice, st));
    CK(c, cudaStreamSynchronize(st));
    c->h2d_bytes += sizeof(SelState);
    return SPL_OK;
}

int32_t spl_dtopk_cut(spl_ctx *c, int32_t tie_policy, int32_t keep_all, int32_t all_ties, uint64_t sk_min_global,
                      uint64_t sk_max_global, int64_t *out_idx, uint64_t *out_y, uint64_t *out_klo, uint64_t *out_khi,
                      int64_t *kept_host, void *stream) {
    if (!c || !kept_host) return SPL_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    const int64_t n = c->dtopk_n;
    *kept_host = 0;
    if (n == 0) return SPL_OK;
    const int det = tie_policy == SPL_TIE_KEY ? 1 : tie_policy == SPL_TIE_KEY_ORDERED ? 2 : 0;
    if (det && !c->dtopk_recs) return fail(c, SPL_E_STATE, "spl_dtopk_cut: det policy needs keys in spl_dtopk_begin");
    int which = 0;
    int64_t kept = 0;
    CKS(c, do_cut_sort(c, c->sk.as<uint64_t>(), c->dtopk_keys, 2, n, n, keep_all, all_ties, sk_min_global, sk_max_global,
                       det, 0, &which, &kept, st));
    if (kept) {
        idx_widen_kernel<<<nblk(kept), TILE, 0, st>>>(c->idx[which].as<uint32_t>(), kept, out_idx);
        ++c->launches;
        if (out_y) CK(c, cudaMemcpyAsync(out_y, c->y[which].p, (size_t)kept * 8, cudaMemcpyDeviceToDevice, st));
        if (det && out_klo) CK(c, cudaMemcpyAsync(out_klo, c->kl[which].p, (size_t)kept * 8, cudaMemcpyDeviceToDevice, st));
        if (det && out_khi) CK(c, cudaMemcpyAsync(out_khi, c->kh[which].p, (size_t)kept * 8, cudaMemcpyDeviceToDevice, st));
        CK(c, cudaGetLastError());
        CK(c, cudaStreamSynchronize(st));
    }
    *kept_host = kept;
    return SPL_OK;
}

int32_t spl_count_less(spl_ctx *c, int32_t words, int32_t inclusive, const uint64_t *ay, const uint64_t *akl,
                       const uint64_t *akh, int64_t na, const uint64_t *by, const uint64_t *bkl, const uint64_t *bkh,
                       int64_t nb, int64_t *out, int32_t accumulate, int32_t sorted_a, void *stream) {
    if (!c || (words != 1 && words != 3) || na < 0 || nb < 0) return fail(c, SPL_E_INVALID, "spl_count_less: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    if (na == 0) return SPL_OK;
    count_less_kernel<<<nblk(na), TILE, 0, st>>>(words, inclusive, ay, akl, akh, na, by, bkl, bkh, nb, out, accumulate, sorted_a);
    ++c->launches;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ fused solver
// A pooled link column for `bytes`: the smallest pooled buffer that holds them, else the largest (it is regrown with
// 1/16 headroom -- queue lengths wobble by a fraction of a percent from level to level and from rank to rank, and
// every regrowth is a cudaMalloc plus a device-synchronising cudaFree in the middle of a level).
static DevBuf *take_link_buf(spl_ctx *c, size_t bytes) {
    if (c->pool_links.empty()) return new DevBuf();
    size_t best = c->pool_links.size(), big = 0;
    for (size_t i = 0; i < c->pool_links.size(); ++i) {
        const size_t cap = c->pool_links[i]->cap;
        if (cap >= bytes && (best == c->pool_links.size() || cap < c->pool_links[best]->cap)) best = i;
        if (cap > c->pool_links[big]->cap) big = i;
    }
    const size_t pick = best != c->pool_links.size() ? best : big;
    DevBuf *b = c->pool_links[pick];
    c->pool_links.erase(c->pool_links.begin() + (long)pick);
    return b;
}
static size_t link_bytes(int64_t n) { return (size_t)n * 8 + (size_t)n / 2; }

// Parent-link columns (src/solver.py:459-464 walks them back from the goal): one per level, 8 B per queue entry.  Past
// spl_set_link_budget's device budget the oldest columns move to pinned host memory; the path walk
// reads a spilled column in place.
struct LinkCols {
    std::vector<DevBuf *> dev;      // per level (pool-owned buffers)
    std::vector<uint64_t *> host;   // per level: pinned copy once spilled, else nullptr
    std::vector<int64_t> n;         // per level: entries
    size_t next_spill = 0;          // levels below this one are on the host
    ~LinkCols() { for (uint64_t *h : host) if (h) cudaFreeHost(h); }
    void push(DevBuf *b, int64_t cnt) { dev.push_back(b); host.push_back(nullptr); n.push_back(cnt); }
    // one entry, wherever the column lives
    cudaError_t read(int level, int64_t i, uint64_t *out) const {
        if (host[level]) { *out = host[level][i]; return cudaSuccess; }
        return cudaMemcpy(out, dev[level]->as<uint64_t>() + i, 8, cudaMemcpyDeviceToHost);
    }
    // spill oldest-first until the device-resident columns fit `budget` bytes (the newest column stays on the device)
    cudaError_t spill(uint64_t budget, uint64_t *moved, cudaStream_t st) {
        if (!budget) return cudaSuccess;
        uint64_t resident = 0;
        for (size_t l = next_spill; l < dev.size(); ++l) resident += (uint64_t)n[l] * 8;
        while (resident > budget && next_spill + 1 < dev.size()) {
            const size_t l = next_spill++;
            const size_t bytes = (size_t)n[l] * 8;
            if (!bytes) continue;
            cudaError_t e = cudaMallocHost(reinterpret_cast<void **>(&host[l]), bytes);
            if (e != cudaSuccess) { host[l] = nullptr; return e; }
            e = cudaMemcpyAsync(host[l], dev[l]->p, bytes, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
            dev[l]->release();  // the memory goes back to the device, not to the column pool
            resident -= bytes;
            *moved += bytes;
        }
        return cudaSuccess;
    }
};

struct spl_solver {
    spl_ctx *c = nullptr;
    int goal = 15, use_h = 0, heuristic = 0, tie = 0, noise = 0, keep_links = 1;
    bool realistic = false;
    bool grouped = false;  // beam search on the card-set-grouped level (spl_m2.cuh); otherwise the key-table level
    spl_rconfig rcfg{};
    // external-noise mode: the level is split in two calls (expand+dedup | score+cut)
    bool pending = false;
    int64_t pend_n = 0, pend_uniq = 0;
    spl_level_info pend_info{};
    int64_t beam = 300000;
    DevBuf front, uniq;
    int64_t n_front = 0;
    int level = 0;
    bool ended = false;
    int64_t goal_rank = -1;
    LinkCols links;                // per level: link column of the queue
    ~spl_solver() {
        if (c->active == this) c->active = nullptr;
        c->pool_front.swap(front);
        c->pool_uniq.swap(uniq);
        for (auto *b : links.dev) c->pool_links.push_back(b);
    }
};

static int save_links(spl_solver *s, cudaStream_t st) {
    spl_ctx *c = s->c;
    const bool keep = s->keep_links && s->n_front > 0;
    DevBuf *b = take_link_buf(c, keep ? (size_t)s->n_front * 8 : 0);
    s->links.push(b, s->keep_links ? s->n_front : 0);
    if (!keep) return SPL_OK;
    if (b->cap < (size_t)s->n_front * 8) CK(c, b->ensure(link_bytes(s->n_front), 0, st));
    if (s->realistic)
        r_links_kernel<<<nblk(s->n_front), TILE, 0, st>>>(s->front.as<RRec>(), s->n_front, b->as<uint64_t>());
    else
        unpack_rec_kernel<<<nblk(s->n_front), TILE, 0, st>>>(s->front.as<Rec>(), s->n_front, nullptr, nullptr, b->as<uint64_t>());
    ++c->launches;
    CK(c, cudaGetLastError());
    uint64_t moved = 0;
    CK(c, s->links.spill(c->link_budget, &moved, st));
    c->d2h_bytes += moved;
    c->spilled_bytes += moved;
    return SPL_OK;
}

// ------------------------------------------------------------------ card-set-grouped level (spl_m2.cuh)
// grow the node table so that `need` more card sets keep its load factor <= 0.6 (if memory allows)
static int ensure_nodes(spl_ctx *c, uint64_t need, cudaStream_t st) {
    while ((double)(c->node_occ + need) > 0.6 * (double)c->nn) {
        uint64_t nnn = c->nn * 2;
        if (nnn * NODE_WORDS * 8 > c->max_node_bytes) nnn = c->max_node_bytes / (NODE_WORDS * 8);
        if (nnn <= c->nn + c->nn / 8) break;
        uint64_t *nt = nullptr;
        cudaError_t e = cudaMalloc(&nt, nnn * NODE_WORDS * 8);
        if (e != cudaSuccess) { cudaGetLastError(); break; }
        CK(c, cudaMemsetAsync(nt, 0, nnn * NODE_WORDS * 8, st));
        m2_rehash_kernel<<<nblk((int64_t)c->nn * 32), TILE, 0, st>>>(c->nodes, c->nn, nt, nnn, c->d_ctr);
        ++c->launches;
        CK(c, cudaGetLastError());
        unsigned int err = 0;
        CK(c, cudaMemcpyAsync(&err, &c->d_ctr->error, 4, cudaMemcpyDeviceToHost, st));
        CK(c, cudaStreamSynchronize(st));
        if (err) { cudaFree(nt); return fail(c, SPL_E_TABLE_FULL, "node rehash into %llu slots overflowed a probe sequence", (unsigned long long)nnn); }
        cudaFree(c->nodes);
        c->nodes = nt;
        c->nn = nnn;
    }
    if (c->node_occ + need > c->nn - c->nn / 16)
        return fail(c, SPL_E_TABLE_FULL, "card-set table full: %llu nodes + %llu new vs %llu slots (max_node_bytes=%llu)",
                    (unsigned long long)c->node_occ, (unsigned long long)need, (unsigned long long)c->nn,
                    (unsigned long long)c->max_node_bytes);
    return SPL_OK;
}

// stable LSD sort of packed 64-bit items on bits [lo_bit, 64): result in c->y[*cur]
// One-sweep stable LSD sort (spl_m2.cuh 2b) of k[*cur][0..n) by bits lo_bit .. lo_bit + nbits - 1, ping-ponging between
// k[0] / k[1]; with a payload column (pl != nullptr) the pairs move together.  n < OS_MAX_ITEMS.
static int onesweep_sort(spl_ctx *c, uint64_t *const k[2], uint32_t *const pl[2], int64_t n, int lo_bit, int nbits, int *cur, cudaStream_t st) {
    const int passes = (nbits + OS_BITS - 1) / OS_BITS;
    if (n <= 1 || passes <= 0) return SPL_OK;
    if (passes > OS_MAX_PASSES || n >= OS_MAX_ITEMS) return fail(c, SPL_E_INVALID, "internal: one-sweep sort of %lld values, %d bits", (long long)n, nbits);
    const unsigned snt = nblk(n, SORT_TILE);
    const size_t status_bytes = (size_t)snt * OS_BINS * 4;
    CK(c, c->os_hist.ensure(OS_MAX_PASSES * OS_BINS * 4, 0, st));
    CK(c, c->os_status.ensure(status_bytes, 0, st));
    CK(c, cudaMemsetAsync(c->os_hist.p, 0, (size_t)passes * OS_BINS * 4, st));
    os_hist_kernel<<<std::min<unsigned>(148 * 8, nblk(n)), TILE, (size_t)passes * OS_BINS * 4, st>>>(k[*cur], n, lo_bit, passes, c->os_hist.as<uint32_t>());
    os_base_kernel<<<passes, OS_BINS, 0, st>>>(c->os_hist.as<uint32_t>());
    c->launches += 2;
    for (int p = 0; p < passes; ++p) {
        CK(c, cudaMemsetAsync(c->os_status.p, 0, status_bytes, st));
        CKS(c, reset_ticket(c, 2, st));
        if (pl)
            os_scatter_kernel<true><<<snt, TILE, sizeof(OsSmem), st>>>(k[*cur], pl[*cur], n, lo_bit + p * OS_BITS, c->os_hist.as<uint32_t>() + p * OS_BINS,
                                                                       c->os_status.as<uint32_t>(), c->d_ctr, 2, k[*cur ^ 1], pl[*cur ^ 1]);
        else
            os_scatter_kernel<false><<<snt, TILE, offsetof(OsSmem, stage_p), st>>>(k[*cur], nullptr, n, lo_bit + p * OS_BITS,
                                                                                   c->os_hist.as<uint32_t>() + p * OS_BINS,
                                                                                   c->os_status.as<uint32_t>(), c->d_ctr, 2, k[*cur ^ 1], nullptr);
        ++c->launches;
        CK(c, cudaGetLastError());
        *cur ^= 1;
    }
    return SPL_OK;
}

// Stable sort of the round's packed items by their key bits ITEM_KEY_LO..63: three one-sweep passes of 10-bit digits
// (rounds below 2^30 items), else 8-bit LSD passes with per-pass histogram + scan.
static int sort_items(spl_ctx *c, int64_t n_items, int *cur, cudaStream_t st) {
    const unsigned snt = nblk(n_items, SORT_TILE);
    static const bool no_onesweep = getenv("SPL_NO_ONESWEEP") != nullptr;
    if (n_items < OS_MAX_ITEMS && !no_onesweep) {
        uint64_t *const k[2] = {c->y[0].as<uint64_t>(), c->y[1].as<uint64_t>()};
        return onesweep_sort(c, k, nullptr, n_items, ITEM_KEY_LO, 64 - ITEM_KEY_LO, cur, st);
    }
    const int lo_bit = ITEM_KEY_LO;
    const size_t msz = (size_t)SORT_BINS * snt;
    CK(c, c->matrix.ensure(msz * 4, 0, st));
    CK(c, c->matrix2.ensure(msz * 4, 0, st));
    const unsigned st_tiles = nblk((int64_t)msz, TILE * SCAN_ITEMS);
    if (n_items > 1)
        for (int shift = lo_bit; shift < 64; shift += SORT_BITS) {
            sort_hist_kernel<<<snt, TILE, 0, st>>>(c->y[*cur].as<uint64_t>(), n_items, shift, c->matrix.as<uint32_t>(), snt);
            CKS(c, prep_status(c, 1, st_tiles, st));
            CKS(c, reset_ticket(c, 2, st));
            scan_u32_kernel<<<st_tiles, TILE, 0, st>>>(c->matrix.as<uint32_t>(), c->matrix2.as<uint32_t>(), (int64_t)msz,
                                                        c->status[1].as<uint64_t>(), c->d_ctr, 2);
            psort_scatter_kernel<<<snt, TILE, 0, st>>>(c->y[*cur].as<uint64_t>(), n_items, shift, c->matrix2.as<uint32_t>(), snt,
                                                        c->y[*cur ^ 1].as<uint64_t>());
            c->launches += 3;
            CK(c, cudaGetLastError());
            *cur ^= 1;
        }
    return SPL_OK;
}

// expand + dedup (+ score) of the whole queue `front[0..n)` in rounds of parents; winners are appended to
// s->uniq / c->sk in no particular order (their link words carry the arrival order)
static int grouped_expand(spl_solver *s, const Rec *front, int64_t n, int64_t *n_uniq_out, int64_t *n_slots_out, int64_t *generated_out,
                          uint64_t *sk_min_out, uint64_t *sk_max_out, float ms[6], cudaStream_t st) {
    spl_ctx *c = s->c;
    const int64_t chunk = c->chunk_user ? (int64_t)c->chunk_parents : (16ll << 20);
    int64_t n_uniq = 0, n_slots = 0, generated = 0;
    uint64_t sk_min = ~0ull, sk_max = 0;
    for (int64_t p0 = 0; p0 < n; p0 += chunk) {
        const int64_t np = std::min<int64_t>(chunk, n - p0);
        const unsigned nt = nblk(np);
        // ---- 1. fan-out: buys offsets, take counts, sort keys of the parents
        CKS(c, zero_ctr(c, st));
        CK(c, cudaEventRecord(c->ev[0], st));
        CK(c, c->boff2.ensure((size_t)np * 4 + 4, 0, st));
        CK(c, c->ntk8.ensure((size_t)np + 8, 0, st));
        // items of the round = parents + buy records (packed: hash half << 32 | item id); the array is sized for the
        // parents first and grown (contents kept) once the buys are counted
        CK(c, c->y[0].ensure((size_t)np * 8 + 8, 0, st));
        CKS(c, prep_status(c, 0, nt, st));
        m2_count_kernel<<<nt, TILE, 0, st>>>(front + p0, np, c->d_tabs, c->d_takes_idx, c->boff2.as<uint32_t>(),
                                              c->y[0].as<uint64_t>(), c->ntk8.as<uint8_t>(), c->status[0].as<uint64_t>(), c->d_ctr, 0);
        ++c->launches;
        CK(c, cudaGetLastError());
        CKS(c, read_ctr(c, st));
        const uint64_t n_takes = c->h_ctr->total_cands, n_buys = c->h_ctr->n_buys, total = n_takes + n_buys;
        generated += (int64_t)total;
        if (total == 0) continue;
        const int64_t n_items = np + (int64_t)n_buys;
        if (total >= 0xFFFFFFFFull || (uint64_t)n_items >= 0xFFFFFFFFull)
            return fail(c, SPL_E_INVALID, "round produced %llu candidates (>= 2^32): lower spl_config.chunk_parents", (unsigned long long)total);
        CK(c, c->y[0].ensure((size_t)n_items * 8 + 8, (size_t)np * 8, st));
        CK(c, c->y[1].ensure((size_t)n_items * 8 + 8, 0, st));
        CK(c, c->brec.ensure((size_t)n_buys * 32 + 32, 0, st));
        if (n_buys) {
            m2_buys_kernel<<<nt, TILE, sizeof(BuySmem), st>>>(front + p0, np, c->d_tabs, c->d_takes_idx, c->boff2.as<uint32_t>(), p0,
                                                               c->brec.as<Rec>(), c->y[0].as<uint64_t>());
            ++c->launches;
            CK(c, cudaGetLastError());
        }
        CK(c, cudaEventRecord(c->ev[7], st));
        // ---- 2. stable LSD sort of the items by the high half of their card-set hash
        int cur = 0;
        CKS(c, sort_items(c, n_items, &cur, st));
        // ---- 3. runs of equal key + candidate-weight prefix
        const unsigned rt = nblk(n_items, TILE * RUN_ITEMS);
        CK(c, c->run_start.ensure((size_t)n_items * 4 + 8, 0, st));
        CK(c, c->run_wpre.ensure((size_t)n_items * 4 + 8, 0, st));
        CKS(c, prep_status(c, 1, rt, st));
        CKS(c, prep_status(c, 2, rt, st));
        CKS(c, reset_ticket(c, 1, st));
        m2_runs_kernel<<<rt, TILE, 0, st>>>(c->y[cur].as<uint64_t>(), n_items, (uint32_t)np,
                                             c->ntk8.as<uint8_t>(), c->run_start.as<uint32_t>(), c->run_wpre.as<uint32_t>(),
                                             c->status[1].as<uint64_t>(), c->status[2].as<uint64_t>(), c->d_ctr, 1);
        ++c->launches;
        CK(c, cudaGetLastError());
        CK(c, cudaEventRecord(c->ev[1], st));
        CKS(c, read_ctr(c, st));
        const uint64_t n_runs = c->h_ctr->n_runs;
        // every run holds at least one card set; more than one only when two sets share the 30 sorted hash bits
        CKS(c, ensure_nodes(c, n_runs + n_runs / 8 + 64, st));
        CK(c, c->cls_list.ensure((size_t)n_runs * 8 + 16, 0, st));
        CK(c, s->uniq.ensure((size_t)(n_slots + (int64_t)total) * 32, (size_t)n_slots * 32, st));
        CK(c, c->sk.ensure((size_t)(n_slots + (int64_t)total) * 8, (size_t)n_slots * 8, st));
        // ---- 4. per-run dedup: warp kernel, then the CTA kernel over the runs it queued
        GroupArgs A;
        A.front = front + p0; A.brec = c->brec.as<Rec>(); A.iv = c->y[cur].as<uint64_t>();
        A.run_start = c->run_start.as<uint32_t>(); A.run_wpre = c->run_wpre.as<uint32_t>();
        A.np = (uint32_t)np; A.rank_base = p0; A.grank = nullptr; A.unordered = 0; A.warp_max = BIG_W; A.tabs = c->d_tabs; A.takes_idx = c->d_takes_idx; A.takes_edges = c->d_takes_edges;
        A.gemrank = c->d_gemrank; A.rankgems = c->d_rankgems; A.nodes = c->nodes; A.nn = c->nn; A.out = s->uniq.as<Rec>();
        A.out_sk = c->sk.as<uint64_t>(); A.out_base = (uint64_t)n_slots;
        for (int k = 0; k < NUM_CLS; ++k) A.cls_list[k] = nullptr;
        A.cls_list[CLS_WARP] = c->cls_list.as<uint32_t>();
        A.cls_list[CLS_CTA] = c->cls_list.as<uint32_t>() + n_runs;
        A.h = s->heuristic; A.noise_mode = s->noise; A.L = c->luts; A.ctr = c->d_ctr;
        CK(c, cudaEventRecord(c->ev[2], st));
        // dispatch + thread kernel over all runs, then the warp kernel and the CTA kernel over the runs it queued
        m2_group_tiny_kernel<false><<<nblk((int64_t)n_runs), TILE, 0, st>>>(A, (uint32_t)n_runs);
        CK(c, cudaEventRecord(c->ev[4], st));
        CKS(c, reset_ticket(c, 2, st));  // the warp kernel draws its runs from ticket 2
        m2_group_warp_kernel<false><<<148 * SPL_WARP_CTAS, TILE, offsetof(WarpSmem, bsort), st>>>(A);
        CK(c, cudaEventRecord(c->ev[5], st));
        CK(c, cudaEventRecord(c->ev[6], st));
        m2_group_big_kernel<<<148 * 5, TILE, sizeof(BigSmem), st>>>(A);
        c->launches += 3;
        CK(c, cudaGetLastError());
        CK(c, cudaEventRecord(c->ev[3], st));
        CKS(c, read_ctr(c, st));
        if (c->h_ctr->error)
            return fail(c, c->h_ctr->error == 3 ? SPL_E_CUDA : SPL_E_TABLE_FULL,
                        c->h_ctr->error == 3 ? "internal: more than %d card sets share one sort key (level %d)" : "card-set table full during level %d (code %d)",
                        c->h_ctr->error == 3 ? BIG_DONE : s->level, c->h_ctr->error == 3 ? s->level : (int)c->h_ctr->error);
        const int64_t n_new = (int64_t)c->h_ctr->n_emitted;
        c->node_occ += c->h_ctr->n_new_nodes;
        c->occupied += n_new;
        if (n_new) { sk_min = std::min<uint64_t>(sk_min, c->h_ctr->sk_min); sk_max = std::max<uint64_t>(sk_max, c->h_ctr->sk_max); }
        n_uniq += n_new;
        n_slots += n_new;  // dense output
        float t;
        cudaEventElapsedTime(&t, c->ev[0], c->ev[7]); ms[0] += t;   // fan-out + buy records
        cudaEventElapsedTime(&t, c->ev[7], c->ev[1]); ms[2] += t;   // item sort + runs (grouping)
        cudaEventElapsedTime(&t, c->ev[2], c->ev[3]); ms[1] += t;   // per-run dedup + emit + score (the dominant stage)
        cudaEventElapsedTime(&t, c->ev[4], c->ev[5]); ms[5] += t;   // ... of which the warp kernel
        if (getenv("SPL_DEBUG")) {
            float tt = 0, ts = 0, tm = 0, tb = 0, tc = 0, tso = 0;
            cudaEventElapsedTime(&tt, c->ev[2], c->ev[4]);
            cudaEventElapsedTime(&ts, c->ev[4], c->ev[5]);
            cudaEventElapsedTime(&tm, c->ev[5], c->ev[6]);
            cudaEventElapsedTime(&tb, c->ev[6], c->ev[3]);
            cudaEventElapsedTime(&tc, c->ev[0], c->ev[7]);
            cudaEventElapsedTime(&tso, c->ev[7], c->ev[1]);
            fprintf(stderr, "[grouped] L%d p0=%lld np=%lld takes=%llu buys=%llu runs=%llu cls=%u/%u/%u new_nodes=%u winners=%lld | count+buys %.2f sort+runs %.2f thread %.2f warp %.2f (-) %.2f cta %.2f ms | nodes %llu/%llu\n",
                    s->level, (long long)p0, (long long)np, (unsigned long long)n_takes, (unsigned long long)n_buys,
                    (unsigned long long)n_runs, (unsigned)(n_runs - c->h_ctr->n_cls[CLS_WARP] - c->h_ctr->n_cls[CLS_CTA]), c->h_ctr->n_cls[CLS_WARP], c->h_ctr->n_cls[CLS_CTA], c->h_ctr->n_new_nodes,
                    (long long)n_new, tc, tso, tt, ts, tm, tb, (unsigned long long)c->node_occ, (unsigned long long)c->nn);
        }
    }
    *n_uniq_out = n_uniq;
    *n_slots_out = n_slots;
    *generated_out = generated;
    *sk_min_out = sk_min;
    *sk_max_out = sk_max;
    return SPL_OK;
}

// ------------------------------------------------------------------ sharded grouped level (spl_shard.cuh)
// One spl_gsolver per rank.  The host (Python, torch.distributed) issues the collectives between these calls:
//   level:  spl_gs_goal -> all-reduce MIN
//           per round: spl_gs_round_begin (-> per-destination record counts) -> all-gather of the counts ->
//                      spl_gs_round_buys (records into the send buffer) -> all-to-all -> spl_gs_round_group
//           cut:    spl_gs_dict -> all-gather -> spl_gs_threshold [-> spl_gs_tie_begin, (spl_gs_tie_hist -> all-reduce
//                      -> spl_gs_tie_pick) x passes] -> spl_gs_cut (local survivors, sorted by their sort word)
//           ranks:  sample sort of the sort words across ranks (spl_gs_partition / spl_gs_rank_sort + all-to-all) ->
//                   spl_gs_adopt (next queue = survivors with their global ranks)
struct spl_gsolver {
    spl_ctx *c = nullptr;
    int rank = 0, world = 1, goal = 15, heuristic = 0, noise = 0, keep_links = 1;
    int64_t beam = 0;
    DevBuf front, grank, uniq;
    int64_t n_local = 0, n_global = 1;
    int level = 0;
    // round
    int64_t r_p0 = 0, r_np = 0;
    uint64_t r_takes = 0, r_buys = 0;
    int64_t counts[MAX_RANKS]{};
    // level accumulators
    int64_t n_uniq = 0, generated = 0;
    float ms[4] = {0, 0, 0, 0};  // CUDA-event time of: item sort + runs, thread kernel, table warp kernel, CTA kernel
    // cut
    int lt = 0, cut_cur = 0;
    int64_t kept_local = 0;
    LinkCols link_cols, rank_cols;  // per level: links / global ranks of the local queue (path reconstruction)
    ~spl_gsolver() {
        if (c->active == this) c->active = nullptr;
        c->pool_front.swap(front);  // keep the big buffers for the next solve on this context
        c->pool_uniq.swap(uniq);
        c->pool_grank.swap(grank);
        for (auto *b : link_cols.dev) c->pool_links.push_back(b);
        for (auto *b : rank_cols.dev) c->pool_links.push_back(b);
    }
};

static int gs_save_links(spl_gsolver *s, cudaStream_t st) {
    spl_ctx *c = s->c;
    const bool keep = s->keep_links && s->n_local > 0;
    const size_t need = keep ? (size_t)s->n_local * 8 : 0;
    DevBuf *lb = take_link_buf(c, need), *rb = take_link_buf(c, need);  // columns of earlier solves on this context
    s->link_cols.push(lb, s->keep_links ? s->n_local : 0);
    s->rank_cols.push(rb, s->keep_links ? s->n_local : 0);
    if (!keep) return SPL_OK;
    if (lb->cap < need) CK(c, lb->ensure(link_bytes(s->n_local), 0, st));
    if (rb->cap < need) CK(c, rb->ensure(link_bytes(s->n_local), 0, st));
    unpack_rec_kernel<<<nblk(s->n_local), TILE, 0, st>>>(s->front.as<Rec>(), s->n_local, nullptr, nullptr, lb->as<uint64_t>());
    ++c->launches;
    CK(c, cudaMemcpyAsync(rb->p, s->grank.p, (size_t)s->n_local * 8, cudaMemcpyDeviceToDevice, st));
    CK(c, cudaGetLastError());
    uint64_t moved = 0;  // the two columns share the budget
    CK(c, s->link_cols.spill(c->link_budget / 2, &moved, st));
    CK(c, s->rank_cols.spill(c->link_budget / 2, &moved, st));
    c->d2h_bytes += moved;
    c->spilled_bytes += moved;
    return SPL_OK;
}

extern "C" {

int32_t spl_gs_create(spl_ctx *c, int32_t rank, int32_t world, const spl_key *root_key, uint64_t root_aux, int32_t goal,
                      int32_t heuristic, int64_t beam, int32_t noise, int32_t keep_links, spl_gsolver **out) {
    if (!c || !root_key || !out || world < 1 || world > MAX_RANKS || rank < 0 || rank >= world || beam < 1)
        return fail(c, SPL_E_INVALID, "spl_gs_create: bad arguments (1 <= world <= %d)", MAX_RANKS);
    if (noise != SPL_NOISE_CONST && noise != SPL_NOISE_HASH) return fail(c, SPL_E_INVALID, "spl_gs_create: noise must be const or hash");
    NO_LIVE_SOLVER(c, "spl_gs_create");
    CK(c, enter_device(c));
    cudaStream_t st = 0;
    CKS(c, reset_visited(c, st));
    spl_gsolver *s = new spl_gsolver();
    s->c = c; s->rank = rank; s->world = world; s->goal = goal; s->heuristic = heuristic; s->beam = beam; s->noise = noise;
    s->keep_links = keep_links;
    s->front.swap(c->pool_front);
    s->uniq.swap(c->pool_uniq);
    s->grank.swap(c->pool_grank);
    std::sort(c->pool_links.begin(), c->pool_links.end(), [](DevBuf *a, DevBuf *b) { return a->cap > b->cap; });
    uint64_t m0, m1;
    mask_words(root_key->lo, root_key->hi & HI_KEY_MASK, m0, m1);
    int rc = SPL_OK;
    if ((int)owner_of_mask(m0, m1, (uint32_t)world) == rank) {  // the root lives on the rank that owns its card set
        Rec r{root_key->lo, root_key->hi & HI_KEY_MASK, root_aux, ~0ull};
        uint64_t zero = 0;
        cudaError_t e = s->front.ensure(32, 0, st);
        if (e == cudaSuccess) e = s->grank.ensure(8, 0, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s->front.p, &r, 32, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s->grank.p, &zero, 8, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        c->h2d_bytes += 40;
        if (e != cudaSuccess) rc = fail(c, SPL_E_CUDA, "root upload: %s", cudaGetErrorString(e));
        if (rc == SPL_OK) rc = zero_ctr(c, st);
        if (rc == SPL_OK) {
            m2_root_kernel<<<1, 1, 0, st>>>(c->nodes, c->nn, r.lo, r.hi, c->d_gemrank, c->d_ctr);
            ++c->launches;
            c->node_occ = 1;
            c->occupied = 1;
        }
        s->n_local = 1;
    }
    if (rc == SPL_OK) rc = gs_save_links(s, st);
    if (rc == SPL_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = fail(c, SPL_E_CUDA, "spl_gs_create: sync failed");
    if (rc != SPL_OK) { delete s; return rc; }
    c->active = s;
    *out = s;
    return SPL_OK;
}

int32_t spl_gs_destroy(spl_gsolver *s) {
    if (s) { cudaSetDevice(s->c->device); cudaDeviceSynchronize(); delete s; }
    return SPL_OK;
}

// first local queue state with pts >= goal -> its GLOBAL rank (INT64_MAX: none); also the local queue length
int32_t spl_gs_goal(spl_gsolver *s, int64_t *rank_host, int64_t *n_local_host, void *stream) {
    if (!s || !rank_host || !n_local_host) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    *rank_host = 0x7fffffffffffffffll;
    *n_local_host = s->n_local;
    s->n_uniq = 0;
    s->generated = 0;
    s->ms[0] = s->ms[1] = s->ms[2] = s->ms[3] = 0;
    if (s->n_local == 0) return SPL_OK;
    CKS(c, zero_ctr(c, st));
    goal_kernel<<<nblk(s->n_local), TILE, 0, st>>>(s->front.as<Rec>(), s->n_local, s->goal, c->d_ctr);
    ++c->launches;
    CK(c, cudaGetLastError());
    CKS(c, read_ctr(c, st));
    if (c->h_ctr->goal_rank != 0x7fffffffffffffffll) {
        uint64_t g = 0;
        CK(c, cudaMemcpyAsync(&g, s->grank.as<uint64_t>() + c->h_ctr->goal_rank, 8, cudaMemcpyDeviceToHost, st));
        CK(c, cudaStreamSynchronize(st));
        c->d2h_bytes += 8;
        *rank_host = (int64_t)g;
    }
    return SPL_OK;
}

// first local index whose global rank is >= r (the local queue is ascending in global rank)
static int gs_lower(spl_gsolver *s, int64_t r, int64_t *idx, cudaStream_t st) {
    spl_ctx *c = s->c;
    int64_t lo = 0, hi = s->n_local;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        uint64_t v = 0;
        CK(c, cudaMemcpyAsync(&v, s->grank.as<uint64_t>() + mid, 8, cudaMemcpyDeviceToHost, st));
        CK(c, cudaStreamSynchronize(st));
        c->d2h_bytes += 8;
        if ((int64_t)v < r) lo = mid + 1; else hi = mid;
    }
    *idx = lo;
    return SPL_OK;
}

// round = the local parents whose global rank is in [rank_lo, rank_hi): fan-out, sort keys, and the number of buy
// records this rank will send to every rank (counts_host[world])
int32_t spl_gs_round_begin(spl_gsolver *s, int64_t rank_lo, int64_t rank_hi, int64_t *counts_host, int64_t *n_parents_host,
                           void *stream) {
    if (!s || !counts_host || !n_parents_host) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    int64_t p0 = 0, p1 = s->n_local;
    if (rank_lo > 0) CKS(c, gs_lower(s, rank_lo, &p0, st));
    if (rank_hi < s->n_global) CKS(c, gs_lower(s, rank_hi, &p1, st));
    const int64_t np = p1 - p0;
    s->r_p0 = p0; s->r_np = np; s->r_takes = s->r_buys = 0;
    for (int g = 0; g < s->world; ++g) counts_host[g] = s->counts[g] = 0;
    *n_parents_host = np;
    if (np == 0) return SPL_OK;
    if (np >= (1ll << 31)) return fail(c, SPL_E_INVALID, "round of %lld parents: use smaller rank windows", (long long)np);
    CKS(c, zero_ctr(c, st));
    CK(c, cudaMemsetAsync(c->d_dest, 0, 2 * MAX_RANKS * 8, st));
    CK(c, c->ntk8.ensure((size_t)np + 8, 0, st));
    CK(c, c->y[0].ensure((size_t)np * 8 + 8, 0, st));
    gs_count_kernel<<<nblk(np), TILE, 0, st>>>(s->front.as<Rec>() + p0, np, c->d_tabs, c->d_takes_idx, c->y[0].as<uint64_t>(),
                                                c->ntk8.as<uint8_t>(), (uint32_t)s->world, c->d_dest, c->d_ctr);
    ++c->launches;
    CK(c, cudaGetLastError());
    unsigned long long h_dest[MAX_RANKS];
    CK(c, cudaMemcpyAsync(h_dest, c->d_dest, MAX_RANKS * 8, cudaMemcpyDeviceToHost, st));
    CKS(c, read_ctr(c, st));
    c->d2h_bytes += MAX_RANKS * 8;
    s->r_takes = c->h_ctr->total_cands;
    s->r_buys = c->h_ctr->n_buys;
    uint64_t sum = 0;
    for (int g = 0; g < s->world; ++g) { counts_host[g] = s->counts[g] = (int64_t)h_dest[g]; sum += h_dest[g]; }
    if (sum != s->r_buys) return fail(c, SPL_E_CUDA, "internal: destination counts %llu != buys %llu", (unsigned long long)sum, (unsigned long long)s->r_buys);
    s->generated += (int64_t)(s->r_takes + s->r_buys);
    return SPL_OK;
}

// the round's buy records into send_dev: destination d owns [sum(counts[0..d)), +counts[d])
// the round's buy records, destination d's at dst[d] + (0 .. counts[d]) in no particular order
static int gs_route(spl_gsolver *s, void *const *dst, cudaStream_t st) {
    spl_ctx *c = s->c;
    unsigned long long h[2 * MAX_RANKS] = {0};  // cursors (relative to each destination's base), then the bases
    for (int g = 0; g < s->world; ++g) h[MAX_RANKS + g] = (unsigned long long)(uintptr_t)dst[g];
    CK(c, cudaMemcpyAsync(c->d_dest + MAX_RANKS, h, sizeof(h), cudaMemcpyHostToDevice, st));
    CK(c, cudaStreamSynchronize(st));  // h is a stack array
    c->h2d_bytes += sizeof(h);
    gs_buys_route_kernel<<<nblk(s->r_np), TILE, sizeof(RouteSmem), st>>>(
        s->front.as<Rec>() + s->r_p0, s->r_np, s->grank.as<uint64_t>() + s->r_p0, c->d_tabs, c->d_takes_idx, (uint32_t)s->world,
        reinterpret_cast<Rec *const *>(c->d_dest + 2 * MAX_RANKS), c->d_dest + MAX_RANKS);
    ++c->launches;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

int32_t spl_gs_round_buys(spl_gsolver *s, void *send_dev, void *stream) {
    if (!s) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    if (s->r_np == 0 || s->r_buys == 0) return SPL_OK;
    if (!send_dev) return fail(c, SPL_E_INVALID, "spl_gs_round_buys: null send buffer");
    void *dst[MAX_RANKS] = {nullptr};
    int64_t at = 0;
    for (int g = 0; g < s->world; ++g) { dst[g] = reinterpret_cast<Rec *>(send_dev) + at; at += s->counts[g]; }
    return gs_route(s, dst, st);
}

int32_t spl_gs_round_buys_peer(spl_gsolver *s, void *const *recv_dev_of_rank, const int64_t *offset_at_rank, void *stream) {
    if (!s || !recv_dev_of_rank || !offset_at_rank) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    if (s->r_np == 0 || s->r_buys == 0) return SPL_OK;
    void *dst[MAX_RANKS] = {nullptr};
    for (int g = 0; g < s->world; ++g) {
        if (s->counts[g] && (!recv_dev_of_rank[g] || offset_at_rank[g] < 0)) return fail(c, SPL_E_INVALID, "spl_gs_round_buys_peer: no buffer for rank %d", g);
        dst[g] = reinterpret_cast<Rec *>(recv_dev_of_rank[g]) + offset_at_rank[g];
    }
    return gs_route(s, dst, st);
}

// ---- device buffers other processes of this node can map (CUDA IPC): the receive side of spl_gs_round_buys_peer
int32_t spl_ipc_alloc(spl_ctx *c, uint64_t bytes, void **ptr_out, uint8_t handle_out[64]) {
    if (!c || !ptr_out || !handle_out || !bytes) return SPL_E_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI carries CUDA IPC handles as 64 bytes");
    CK(c, enter_device(c));
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(c, SPL_E_NOMEM, "spl_ipc_alloc: %llu bytes: %s", (unsigned long long)bytes, cudaGetErrorString(e)); }
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); cudaGetLastError(); return fail(c, SPL_E_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    memcpy(handle_out, &h, 64);
    *ptr_out = p;
    return SPL_OK;
}
int32_t spl_ipc_open(spl_ctx *c, const uint8_t handle[64], void **ptr_out) {
    if (!c || !handle || !ptr_out) return SPL_E_INVALID;
    CK(c, enter_device(c));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    cudaError_t e = cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(c, SPL_E_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); }
    return SPL_OK;
}
int32_t spl_ipc_close(spl_ctx *c, void *ptr) {
    if (!c || !ptr) return SPL_E_INVALID;
    CK(c, enter_device(c));
    CK(c, cudaIpcCloseMemHandle(ptr));
    return SPL_OK;
}
int32_t spl_ipc_free(spl_ctx *c, void *ptr) {
    if (!c || !ptr) return SPL_E_INVALID;
    CK(c, enter_device(c));
    CK(c, cudaDeviceSynchronize());
    CK(c, cudaFree(ptr));
    return SPL_OK;
}

// owner side: the round's local parents + the n_recv records received for this rank -> sort by card set, per-run
// dedup against this rank's nodes, winners (+ scores) appended to the level's list
int32_t spl_gs_round_group(spl_gsolver *s, const void *recv_dev, int64_t n_recv, int64_t *n_new_host, void *stream) {
    if (!s || !n_new_host || n_recv < 0) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    *n_new_host = 0;
    const int64_t np = s->r_np, n_items = np + n_recv;
    const uint64_t total = s->r_takes + (uint64_t)n_recv;
    if (total == 0) return SPL_OK;
    if (total >= 0xFFFFFFFFull || (uint64_t)n_items >= 0xFFFFFFFFull)
        return fail(c, SPL_E_INVALID, "round holds %llu candidates (>= 2^32): use smaller rank windows", (unsigned long long)total);
    CK(c, c->y[0].ensure((size_t)n_items * 8 + 8, (size_t)np * 8, st));
    CK(c, c->y[1].ensure((size_t)n_items * 8 + 8, 0, st));
    if (n_recv) {
        gs_recv_items_kernel<<<nblk(n_recv), TILE, 0, st>>>(reinterpret_cast<const Rec *>(recv_dev), n_recv, (uint32_t)np, c->y[0].as<uint64_t>());
        ++c->launches;
    }
    const bool dbg = getenv("SPL_DEBUG") != nullptr;
    CK(c, cudaEventRecord(c->ev[0], st));
    int cur = 0;
    CKS(c, sort_items(c, n_items, &cur, st));
    const unsigned rt = nblk(n_items, TILE * RUN_ITEMS);
    CK(c, c->run_start.ensure((size_t)n_items * 4 + 8, 0, st));
    CK(c, c->run_wpre.ensure((size_t)n_items * 4 + 8, 0, st));
    CKS(c, zero_ctr(c, st));
    CKS(c, prep_status(c, 1, rt, st));
    CKS(c, prep_status(c, 2, rt, st));
    m2_runs_kernel<<<rt, TILE, 0, st>>>(c->y[cur].as<uint64_t>(), n_items, (uint32_t)np, c->ntk8.as<uint8_t>(), c->run_start.as<uint32_t>(),
                                         c->run_wpre.as<uint32_t>(), c->status[1].as<uint64_t>(), c->status[2].as<uint64_t>(), c->d_ctr, 1);
    ++c->launches;
    CK(c, cudaGetLastError());
    CKS(c, read_ctr(c, st));
    const uint64_t n_runs = c->h_ctr->n_runs;
    CKS(c, ensure_nodes(c, n_runs + n_runs / 8 + 64, st));
    CK(c, c->cls_list.ensure((size_t)n_runs * 8 + 16, 0, st));
    CK(c, s->uniq.ensure((size_t)(s->n_uniq + (int64_t)total) * 32, (size_t)s->n_uniq * 32, st));
    CK(c, c->sk.ensure((size_t)(s->n_uniq + (int64_t)total) * 8, (size_t)s->n_uniq * 8, st));
    GroupArgs A;
    A.front = s->front.as<Rec>() + s->r_p0; A.brec = reinterpret_cast<const Rec *>(recv_dev); A.iv = c->y[cur].as<uint64_t>();
    A.run_start = c->run_start.as<uint32_t>(); A.run_wpre = c->run_wpre.as<uint32_t>();
    A.np = (uint32_t)np; A.rank_base = 0; A.grank = s->grank.as<uint64_t>() + s->r_p0; A.unordered = 1;
    A.warp_max = std::min<uint32_t>(BIG_W, BSORT_MAX);  // the warp kernel sorts at most BSORT_MAX records of a run itself
    A.tabs = c->d_tabs; A.takes_idx = c->d_takes_idx; A.takes_edges = c->d_takes_edges; A.gemrank = c->d_gemrank; A.rankgems = c->d_rankgems;
    A.nodes = c->nodes; A.nn = c->nn; A.out = s->uniq.as<Rec>(); A.out_sk = c->sk.as<uint64_t>(); A.out_base = (uint64_t)s->n_uniq;
    for (int k = 0; k < NUM_CLS; ++k) A.cls_list[k] = nullptr;
    A.cls_list[CLS_WARP] = c->cls_list.as<uint32_t>();
    A.cls_list[CLS_CTA] = c->cls_list.as<uint32_t>() + n_runs;
    A.h = s->heuristic; A.noise_mode = s->noise; A.L = c->luts; A.ctr = c->d_ctr;
    CK(c, cudaEventRecord(c->ev[1], st));
    m2_group_tiny_kernel<true><<<nblk((int64_t)n_runs), TILE, 0, st>>>(A, (uint32_t)n_runs);
    CK(c, cudaEventRecord(c->ev[2], st));
    CKS(c, reset_ticket(c, 2, st));
    m2_group_warp_kernel<true><<<148 * SPL_WARP_CTAS, TILE, sizeof(WarpSmem), st>>>(A);
    CK(c, cudaEventRecord(c->ev[3], st));
    m2_group_big_kernel<<<148 * 5, TILE, sizeof(BigSmem), st>>>(A);
    CK(c, cudaEventRecord(c->ev[4], st));
    c->launches += 3;
    CK(c, cudaGetLastError());
    CKS(c, read_ctr(c, st));
    float t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    cudaEventElapsedTime(&t1, c->ev[0], c->ev[1]); cudaEventElapsedTime(&t2, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&t3, c->ev[2], c->ev[3]); cudaEventElapsedTime(&t4, c->ev[3], c->ev[4]);
    s->ms[0] += t1; s->ms[1] += t2; s->ms[2] += t3; s->ms[3] += t4;
    if (dbg) {
        fprintf(stderr, "[gs r%d] L%d np=%lld recv=%lld takes=%llu runs=%llu warp=%u cta=%u winners=%llu | sort+runs %.2f tiny %.2f warp %.2f cta %.2f ms\n",
                s->rank, s->level, (long long)np, (long long)n_recv, (unsigned long long)s->r_takes, (unsigned long long)n_runs,
                c->h_ctr->n_cls[CLS_WARP], c->h_ctr->n_cls[CLS_CTA], (unsigned long long)c->h_ctr->n_emitted, t1, t2, t3, t4);
    }
    if (c->h_ctr->error)
        return fail(c, c->h_ctr->error == 3 ? SPL_E_CUDA : SPL_E_TABLE_FULL, "sharded level %d: device error code %u (2 = card-set table full, 3 = too many card sets under one sort key)",
                    s->level, c->h_ctr->error);
    const int64_t n_new = (int64_t)c->h_ctr->n_emitted;
    c->node_occ += c->h_ctr->n_new_nodes;
    c->occupied += n_new;
    s->n_uniq += n_new;
    *n_new_host = n_new;
    return SPL_OK;
}

int32_t spl_gs_stage_ms(spl_gsolver *s, float ms_host[4]) {
    if (!s || !ms_host) return SPL_E_INVALID;
    for (int i = 0; i < 4; ++i) ms_host[i] = s->ms[i];
    return SPL_OK;
}

int32_t spl_gs_counters(spl_gsolver *s, int64_t *n_uniq_host, int64_t *generated_host, int64_t *visited_host) {
    if (!s) return SPL_E_INVALID;
    if (n_uniq_host) *n_uniq_host = s->n_uniq;
    if (generated_host) *generated_host = s->generated;
    if (visited_host) *visited_host = (int64_t)s->c->occupied;
    return SPL_OK;
}

// local dictionary of the level's scores (distinct score -> count); *dict_dev points at sizeof(ScoreDict) = *bytes
int32_t spl_gs_dict(spl_gsolver *s, void **dict_dev, int64_t *bytes, void *stream) {
    if (!s || !dict_dev || !bytes) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    CK(c, cudaMemsetAsync(c->d_dict->key, 0xFF, sizeof(c->d_dict->key), st));
    CK(c, cudaMemsetAsync(c->d_dict->cnt, 0, sizeof(ScoreDict) - sizeof(c->d_dict->key), st));
    if (s->n_uniq) {
        dict_build_kernel<<<std::min<unsigned>(nblk(s->n_uniq), 148 * 8), TILE, 0, st>>>(c->sk.as<uint64_t>(), s->n_uniq, c->d_dict);
        ++c->launches;
        CK(c, cudaGetLastError());
    }
    *dict_dev = c->d_dict;
    *bytes = (int64_t)sizeof(ScoreDict);
    return SPL_OK;
}

// all ranks' dictionaries (all-gathered, `world` x sizeof(ScoreDict)) -> global threshold for the k best of
// n_uniq_global states.  *need_ties_host = 1: the threshold score has more states than fit (split by arrival order)
int32_t spl_gs_threshold(spl_gsolver *s, const void *dicts_dev, int64_t k, int64_t n_uniq_global, int32_t *need_ties_host,
                         void *stream) {
    if (!s || !dicts_dev || !need_ties_host || k < 1) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    CK(c, cudaMemsetAsync(c->d_dict2->key, 0xFF, sizeof(c->d_dict2->key), st));
    CK(c, cudaMemsetAsync(c->d_dict2->cnt, 0, sizeof(ScoreDict) - sizeof(c->d_dict2->key), st));
    memset(c->h_sel, 0, sizeof(SelState));
    c->h_sel->rank_t = ~0ull;
    CK(c, cudaMemcpyAsync(c->d_sel, c->h_sel, sizeof(SelState), cudaMemcpyHostToDevice, st));
    gs_dict_merge_kernel<<<nblk((int64_t)s->world * DICT_CAP), TILE, 0, st>>>(reinterpret_cast<const ScoreDict *>(dicts_dev), s->world, c->d_dict2);
    dict_rank_kernel<<<1, 1024, 0, st>>>(c->d_dict2, (uint64_t)std::min(k, n_uniq_global), 0, c->d_sel);
    c->launches += 2;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(c->h_sel, c->d_sel, sizeof(SelState), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    c->d2h_bytes += sizeof(SelState);
    if (c->h_sel->rank_t == ~0ull)
        return fail(c, SPL_E_CAPACITY, "the level has more than %d distinct scores: the dictionary cut does not apply", DICT_MAX);
    s->lt = 8 + bitlen((uint64_t)s->n_global);
    *need_ties_host = (n_uniq_global > k && c->h_sel->k_rem < c->h_sel->tie_count) ? 1 : 0;
    if (!*need_ties_host) c->h_sel->tie_count = 0;  // all ties are kept
    return SPL_OK;
}

// ties of the threshold score: collect their arrival words locally, then radix-select the quota-th one globally --
// the host all-reduces *hist_dev (2048 x uint32) between spl_gs_tie_hist and spl_gs_tie_pick; passes: shift = lt - 11,
// lt - 22, ... (bits = min(11, remaining))
int32_t spl_gs_tie_begin(spl_gsolver *s, int32_t *link_bits_host, void *stream) {
    if (!s || !link_bits_host) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    *link_bits_host = s->lt;
    CK(c, c->rtmp.ensure((size_t)std::max<int64_t>(s->n_uniq, 1) * 8 + 8, 0, st));
    CK(c, cudaMemsetAsync(&c->d_ctr->n_ties, 0, 8, st));
    if (s->n_uniq) {
        tie_collect_kernel<<<std::min<unsigned>(nblk(s->n_uniq), 148 * 8), TILE, 0, st>>>(
            c->sk.as<uint64_t>(), reinterpret_cast<const uint64_t *>(s->uniq.p), 4, s->n_uniq, 0, c->d_sel, s->lt, c->rtmp.as<uint64_t>(), c->d_ctr);
        ++c->launches;
        CK(c, cudaGetLastError());
    }
    CKS(c, read_ctr(c, st));
    return SPL_OK;
}
int32_t spl_gs_tie_hist(spl_gsolver *s, int32_t shift, int32_t bits, int32_t first, uint32_t **hist_dev, void *stream) {
    if (!s || !hist_dev) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    *hist_dev = c->d_hist;
    const int64_t ntie = (int64_t)c->h_ctr->n_ties;
    if (ntie) {
        sel_hist_kernel<3><<<std::min<unsigned>(nblk(ntie), 148 * 8), TILE, 0, st>>>(nullptr, c->rtmp.as<uint64_t>(), 1, ntie, 0, shift, bits, first,
                                                                                    c->d_sel, c->d_hist);
        ++c->launches;
        CK(c, cudaGetLastError());
    }
    return SPL_OK;
}
int32_t spl_gs_tie_pick(spl_gsolver *s, int32_t shift, int32_t first, void *stream) {
    if (!s) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    sel_pick_kernel<<<1, 1024, 0, st>>>(c->d_hist, 2, shift, first, 0, 0, c->d_sel);
    ++c->launches;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

// local cut by the global threshold, then local sort by the sort word (score rank << link bits | link: ascending ==
// better first, and a total order over all ranks).  *y_sorted_dev: the kept_local sort words, ascending.
int32_t spl_gs_cut(spl_gsolver *s, int32_t have_tie_threshold, int64_t *kept_local_host, const uint64_t **y_sorted_dev, int32_t *y_bits_host,
                   void *stream) {
    if (!s || !kept_local_host || !y_sorted_dev || !y_bits_host) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    const int64_t n = s->n_uniq;
    const int nbits = s->lt + bitlen(c->h_sel->rank_t);
    *y_bits_host = nbits;
    *kept_local_host = s->kept_local = 0;
    *y_sorted_dev = nullptr;
    s->cut_cur = 0;
    if (n == 0) return SPL_OK;
    for (int b = 0; b < 2; ++b) {
        CK(c, c->y[b].ensure((size_t)n * 8 + 8, 0, st));
        CK(c, c->idx[b].ensure((size_t)n * 4 + 4, 0, st));
    }
    const unsigned ct = nblk(n, TILE * CUTP_ITEMS);
    CKS(c, prep_status(c, 2, ct, st));
    CKS(c, zero_ctr(c, st));
    cut_pack_kernel<<<ct, TILE, 0, st>>>(c->sk.as<uint64_t>(), reinterpret_cast<const uint64_t *>(s->uniq.p), 4, n, 0, 0, have_tie_threshold ? 0 : 1,
                                          c->d_sel, c->d_dict2, s->lt, c->y[0].as<uint64_t>(), c->idx[0].as<uint32_t>(), c->status[2].as<uint64_t>(),
                                          c->d_ctr, 1);
    ++c->launches;
    CK(c, cudaGetLastError());
    uint64_t last = 0;
    CK(c, cudaMemcpyAsync(&last, c->status[2].as<uint64_t>() + (ct - 1), 8, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    c->d2h_bytes += 8;
    const int64_t kept = (int64_t)(last & ((1ull << 62) - 1));
    int cur = 0;
    if (kept > 1) {
        const unsigned nt = nblk(kept, SORT_TILE);
        const size_t msz = (size_t)SORT_BINS * nt;
        CK(c, c->matrix.ensure(msz * 4, 0, st));
        CK(c, c->matrix2.ensure(msz * 4, 0, st));
        if (kept < OS_MAX_ITEMS) {
            uint64_t *const yk[2] = {c->y[0].as<uint64_t>(), c->y[1].as<uint64_t>()};
            uint32_t *const yi[2] = {c->idx[0].as<uint32_t>(), c->idx[1].as<uint32_t>()};
            CKS(c, onesweep_sort(c, yk, yi, kept, 0, nbits, &cur, st));
        } else {
            for (int shift = 0; shift < nbits; shift += SORT_BITS) {
                CKS(c, sort_pass(c, 0, cur, c->y[cur].as<uint64_t>(), kept, shift, nt, msz, st));
                cur ^= 1;
            }
        }
    }
    s->cut_cur = cur;
    *kept_local_host = s->kept_local = kept;
    *y_sorted_dev = c->y[cur].as<uint64_t>();
    return SPL_OK;
}

// lower bounds of n_probes ascending probes in the local sorted sort words (sample-sort partition)
int32_t spl_gs_partition(spl_gsolver *s, const uint64_t *probes_host, int32_t n_probes, int64_t *bounds_host, void *stream) {
    if (!s || n_probes < 0 || n_probes > 4 * MAX_RANKS) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    if (n_probes == 0) return SPL_OK;
    CK(c, c->rtmp.ensure(16 * 4 * MAX_RANKS, 0, st));
    uint64_t *dp = c->rtmp.as<uint64_t>();
    int64_t *dout = reinterpret_cast<int64_t *>(dp + 4 * MAX_RANKS);
    CK(c, cudaMemcpyAsync(dp, probes_host, (size_t)n_probes * 8, cudaMemcpyHostToDevice, st));
    gs_lower_bound_kernel<<<1, 64, 0, st>>>(c->y[s->cut_cur].as<uint64_t>(), s->kept_local, dp, n_probes, dout);
    ++c->launches;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(bounds_host, dout, (size_t)n_probes * 8, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    return SPL_OK;
}

// receiver side of the sample sort: n sort words (any order) -> ranks_out[i] = base + (number of words smaller than
// word i); the words of all ranks are pairwise distinct (they contain the arrival index)
int32_t spl_gs_rank_sort(spl_gsolver *s, const uint64_t *y_dev, int64_t n, int32_t y_bits, int64_t base, int64_t *ranks_out_dev, void *stream) {
    if (!s || n < 0) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    if (n == 0) return SPL_OK;
    // y[] / idx[] of the context hold the local survivors (spl_gs_cut): work in separate buffers
    CK(c, c->kl[0].ensure((size_t)n * 8 + 8, 0, st));
    CK(c, c->kl[1].ensure((size_t)n * 8 + 8, 0, st));
    CK(c, c->kh[0].ensure((size_t)n * 4 + 8, 0, st));
    CK(c, c->kh[1].ensure((size_t)n * 4 + 8, 0, st));
    CK(c, cudaMemcpyAsync(c->kl[0].p, y_dev, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    gs_iota_kernel<<<nblk(n), TILE, 0, st>>>(c->kh[0].as<uint32_t>(), n);
    ++c->launches;
    int cur = 0;
    if (n >= OS_MAX_ITEMS) return fail(c, SPL_E_INVALID, "spl_gs_rank_sort: %lld sort words (limit 2^30 per rank)", (long long)n);
    {
        uint64_t *const k[2] = {c->kl[0].as<uint64_t>(), c->kl[1].as<uint64_t>()};
        uint32_t *const pl[2] = {c->kh[0].as<uint32_t>(), c->kh[1].as<uint32_t>()};
        CKS(c, onesweep_sort(c, k, pl, n, 0, y_bits, &cur, st));
    }
    gs_scatter_ranks_kernel<<<nblk(n), TILE, 0, st>>>(c->kh[cur].as<uint32_t>(), n, base, ranks_out_dev);
    ++c->launches;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

// next queue of this rank = its survivors in sort-word order (== ascending global rank) with the global ranks the
// sample sort assigned (granks_dev[i] for the i-th local survivor); n_global_next = queue length over all ranks
int32_t spl_gs_adopt(spl_gsolver *s, const int64_t *granks_dev, int64_t n_global_next, void *stream) {
    if (!s) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    const int64_t kept = s->kept_local;
    const bool dbg = getenv("SPL_DEBUG") != nullptr;
    if (dbg) CK(c, cudaEventRecord(c->ev[0], st));
    if (kept) {
        if (!granks_dev) return fail(c, SPL_E_INVALID, "spl_gs_adopt: null ranks");
        CK(c, s->front.ensure((size_t)kept * 32, 0, st));
        CK(c, s->grank.ensure((size_t)kept * 8, 0, st));
        gather_rec_kernel<<<nblk(kept), TILE, 0, st>>>(s->uniq.as<Rec>(), c->idx[s->cut_cur].as<uint32_t>(), kept, s->front.as<Rec>());
        ++c->launches;
        CK(c, cudaMemcpyAsync(s->grank.p, granks_dev, (size_t)kept * 8, cudaMemcpyDeviceToDevice, st));
        CK(c, cudaGetLastError());
    }
    s->n_local = kept;
    s->n_global = n_global_next;
    s->level += 1;
    if (dbg) CK(c, cudaEventRecord(c->ev[1], st));
    CKS(c, gs_save_links(s, st));
    if (dbg) CK(c, cudaEventRecord(c->ev[2], st));
    CK(c, cudaStreamSynchronize(st));
    if (dbg) {
        float t1 = 0, t2 = 0;
        cudaEventElapsedTime(&t1, c->ev[0], c->ev[1]); cudaEventElapsedTime(&t2, c->ev[1], c->ev[2]);
        fprintf(stderr, "[gs r%d] adopt L%d kept=%lld gather %.2f links %.2f ms\n", s->rank, s->level, (long long)kept, t1, t2);
    }
    return SPL_OK;
}

// device view of the local queue: records (32 B each) and their global ranks
int32_t spl_gs_frontier(spl_gsolver *s, const void **recs_dev, const uint64_t **granks_dev, int64_t *n_local_host) {
    if (!s || !n_local_host) return SPL_E_INVALID;
    if (recs_dev) *recs_dev = s->front.p;
    if (granks_dev) *granks_dev = s->grank.as<uint64_t>();
    *n_local_host = s->n_local;
    return SPL_OK;
}

// link (parent rank << 8 | ordinal) of the state with global rank `grank` in the queue of `level`, if this rank holds it
int32_t spl_gs_link_at(spl_gsolver *s, int32_t level, int64_t grank, int32_t *found_host, uint64_t *link_host) {
    if (!s || !found_host || !link_host || level < 0 || level >= (int)s->rank_cols.dev.size()) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    if (!s->keep_links) return fail(c, SPL_E_STATE, "spl_gs_link_at: solver was created with keep_links = 0");
    CK(c, enter_device(c));
    *found_host = 0;
    *link_host = 0;
    const int64_t n = s->rank_cols.n[level];
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        uint64_t v = 0;
        CK(c, s->rank_cols.read(level, mid, &v));
        if ((int64_t)v < grank) lo = mid + 1; else hi = mid;
    }
    if (lo < n) {
        uint64_t v = 0;
        CK(c, s->rank_cols.read(level, lo, &v));
        if ((int64_t)v == grank) {
            CK(c, s->link_cols.read(level, lo, link_host));
            *found_host = 1;
        }
    }
    return SPL_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ second half of a speedrun level
// beam cut (src/solver.py:452-456) or plain BFS hand-over, then bookkeeping
// n_uniq states in the first n_slots elements of s->uniq / c->sk (n_slots > n_uniq: the grouped level leaves unused slots)
static int speedrun_cut(spl_solver *s, spl_level_info *info, int64_t n, int64_t n_uniq, uint64_t sk_min, uint64_t sk_max,
                        cudaStream_t st, int64_t n_slots = -1) {
    spl_ctx *c = s->c;
    if (n_slots < 0) n_slots = n_uniq;
    int64_t kept = n_uniq;
    if (s->use_h && n_uniq > 0) {
        int which = 0;
        CK(c, cudaEventRecord(c->ev[4], st));
        // grouped level: the winners are unordered, so `stable` ties are broken on the link words (arrival order)
        const bool link_ties = s->grouped && s->tie == SPL_TIE_STABLE;
        c->tie_link_top = link_ties ? 8 + bitlen((uint64_t)n) : 0;
        const int rc_cut = run_cut_sort(c, c->sk.as<uint64_t>(), s->uniq.as<Rec>(), n_slots, s->beam, sk_min, sk_max,
                                        s->tie == SPL_TIE_KEY || link_ties, &which, &kept, st, n_uniq);
        c->tie_link_top = 0;
        CKS(c, rc_cut);
        CK(c, cudaEventRecord(c->ev[5], st));
        CK(c, s->front.ensure((size_t)kept * 32, 0, st));
        gather_rec_kernel<<<nblk(kept), TILE, 0, st>>>(s->uniq.as<Rec>(), c->idx[which].as<uint32_t>(), kept, s->front.as<Rec>());
        ++c->launches;
        CK(c, cudaGetLastError());
        CK(c, cudaEventRecord(c->ev[6], st));
        CK(c, cudaStreamSynchronize(st));
        float t;
        cudaEventElapsedTime(&t, c->ev[4], c->ev[5]);
        info->ms_select = t;  // select + cut + sort
        cudaEventElapsedTime(&t, c->ev[5], c->ev[6]);
        info->ms_sort = t;    // gather into rank order
    } else if (kept > 0) {
        s->front.swap(s->uniq);
    }
    info->kept = kept;
    info->visited = (int64_t)c->occupied;
    info->table_slots = s->grouped ? c->nn : c->cap;
    if (kept == 0) {  // frontier exhausted: `puzzle` is the last dequeued state (src/solver.py:438,459);
        s->ended = true;  // s->front still holds the level that was just expanded
        s->goal_rank = n - 1;
        info->ended = 1;
        return SPL_OK;
    }
    s->n_front = kept;
    s->level += 1;
    CKS(c, save_links(s, st));
    CK(c, cudaStreamSynchronize(st));
    return SPL_OK;
}

// ------------------------------------------------------------------ realistic mode host side
static int upload_rconfig(spl_ctx *c, const spl_rconfig *cfg, cudaStream_t st) {
    if (!cfg || cfg->num_players < 2 || cfg->num_players > 4 || cfg->gems_per_color < 1 || cfg->gems_per_color > 7)
        return fail(c, SPL_E_INVALID, "realistic config: num_players must be 2..4 and gems_per_color 1..7");
    const HostTables &T = host_tables();
    RConfigDev h;
    memset(&h, 0, sizeof h);
    h.P = cfg->num_players; h.target = cfg->target_points; h.gpc = cfg->gems_per_color; h.noise = cfg->noise;
    for (int t = 0; t < 3; ++t) {
        if (cfg->deck_len[t] < 0 || cfg->deck_len[t] > 40) return fail(c, SPL_E_INVALID, "realistic config: bad deck length");
        h.deck_len[t] = cfg->deck_len[t];
        memcpy(h.deck[t], cfg->deck[t], 40);
        for (int i = 0; i < cfg->deck_len[t]; ++i)
            if (cfg->deck[t][i] < SPL_NUM_CARDS) h.pos_of[cfg->deck[t][i]] = (uint8_t)i;
    }
    for (int i = 0; i < SPL_NUM_CARDS; ++i) {
        const int t = T.card_pt[i] == 0 ? 0 : T.card_pt[i] <= 2 ? 1 : 2;  // tiers by points, src/solver.py:102-104
        if (i < 64) { h.col_lo[T.card_bonus[i]] |= 1ull << i; h.pt_lo[T.card_pt[i]] |= 1ull << i; h.tier_lo[t] |= 1ull << i; }
        else { h.col_hi[T.card_bonus[i]] |= 1u << (i - 64); h.pt_hi[T.card_pt[i]] |= 1u << (i - 64); h.tier_hi[t] |= 1u << (i - 64); }
        h.card[i] = T.dev.card[i];
    }
    CK(c, c->rcfg.ensure(sizeof h, 0, st));
    c->rcfg_dirty = true;
    CK(c, cudaMemcpyAsync(c->rcfg.p, &h, sizeof h, cudaMemcpyHostToDevice, st));
    CK(c, cudaStreamSynchronize(st));  // h is a stack object
    c->h2d_bytes += sizeof h;
    return SPL_OK;
}

// realistic mode's visited table: allocate / grow so that `need` more states keep the load factor <= 0.6
static int ensure_rtable(spl_ctx *c, uint64_t need, cudaStream_t st) {
    if (!c->rtable) {
        // first use: the caller's table_slots sizes THIS table (a realistic search does not use the speedrun tables:
        // give their memory back first if they were sized large)
        if (c->occupied == 0 && c->nb > (1ull << 16)) {
            cudaFree(c->table);
            c->table = nullptr;
            CKS(c, alloc_table(c, 1ull << 12, st));
        }
        uint64_t nb = std::max<uint64_t>(c->rslots_hint ? c->rslots_hint : (1ull << 16), 1024);
        nb = std::min<uint64_t>(nb, c->max_table_bytes / 64);
        CK(c, cudaMalloc(&c->rtable, nb * 64));
        CK(c, cudaMemsetAsync(c->rtable, 0, nb * 64, st));
        c->rnb = nb;
        c->rocc = 0;
    }
    while ((double)(c->rocc + need) > 0.6 * (double)c->rnb) {
        uint64_t nnb = c->rnb * 2;
        if (nnb * 64 > c->max_table_bytes) nnb = c->max_table_bytes / 64;
        if (nnb <= c->rnb + c->rnb / 8) break;
        RBucket *nt = nullptr;
        cudaError_t e = cudaMalloc(&nt, nnb * 64);
        if (e != cudaSuccess) { cudaGetLastError(); break; }
        CK(c, cudaMemsetAsync(nt, 0, nnb * 64, st));
        r_rehash_kernel<<<nblk((int64_t)c->rnb), TILE, 0, st>>>(c->rtable, c->rnb, nt, nnb, c->d_ctr);
        ++c->launches;
        CK(c, cudaGetLastError());
        unsigned int err = 0;
        CK(c, cudaMemcpyAsync(&err, &c->d_ctr->error, 4, cudaMemcpyDeviceToHost, st));
        CK(c, cudaStreamSynchronize(st));
        if (err) { cudaFree(nt); return fail(c, SPL_E_TABLE_FULL, "realistic table rehash into %llu buckets overflowed a probe sequence", (unsigned long long)nnb); }
        cudaFree(c->rtable);
        c->rtable = nt;
        c->rnb = nnb;
    }
    if (c->rocc + need / 4 > c->rnb - c->rnb / 16)
        return fail(c, SPL_E_TABLE_FULL, "realistic visited table full: %llu states + %llu candidates vs %llu buckets (max_table_bytes=%llu)",
                    (unsigned long long)c->rocc, (unsigned long long)need, (unsigned long long)c->rnb, (unsigned long long)c->max_table_bytes);
    return SPL_OK;
}

// count + materialise the successors of front[0..n) into c->rcand
static int r_expand_all(spl_ctx *c, const RRec *front, int64_t n, int64_t rank_base, bool keys, int64_t *total_out,
                        cudaStream_t st) {
    const unsigned nt = nblk(n);
    CK(c, c->off.ensure((size_t)n * 4 + 4, 0, st));
    CK(c, c->rvmask.ensure((size_t)n * 4 + 4, 0, st));
    CKS(c, zero_ctr(c, st));
    CKS(c, prep_status(c, 0, nt, st));
    r_count_scan_kernel<<<nt, TILE, 0, st>>>(front, n, c->rcfg.as<RConfigDev>(), c->off.as<uint32_t>(), c->rvmask.as<uint32_t>(),
                                              c->status[0].as<uint64_t>(), c->d_ctr, 0);
    ++c->launches;
    CK(c, cudaGetLastError());
    CKS(c, read_ctr(c, st));
    const int64_t total = (int64_t)c->h_ctr->total_cands;
    *total_out = total;
    if (total == 0) return SPL_OK;
    CK(c, c->rcand.ensure((size_t)total * 96, 0, st));
    (void)keys;
    r_expand_kernel<<<nt, TILE, 0, st>>>(front, n, c->rcfg.as<RConfigDev>(), c->off.as<uint32_t>(), c->rvmask.as<uint32_t>(),
                                          rank_base, c->rcand.as<RRec>());
    ++c->launches;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

static int realistic_cut(spl_solver *s, spl_level_info *info, int64_t n, int64_t n_new, const uint8_t *draws, cudaStream_t st);

static int rsolver_step(spl_solver *s, spl_level_info *info, cudaStream_t st) {
    spl_ctx *c = s->c;
    memset(info, 0, sizeof *info);
    info->level = s->level;
    info->frontier = s->n_front;
    info->goal_rank = -1;
    const int64_t n = s->n_front;
    const RRec *front = s->front.as<RRec>();
    if (c->rcfg_dirty) {  // spl_rscore / spl_rexpand / spl_rmaxpts ran with their own GameConfig in between
        CKS(c, upload_rconfig(c, &s->rcfg, st));
        c->rcfg_dirty = false;
    }
    // game over on dequeue (src/solver.py:827-829)
    CKS(c, zero_ctr(c, st));
    r_goal_kernel<<<nblk(n), TILE, 0, st>>>(front, n, c->rcfg.as<RConfigDev>(), c->d_ctr);
    ++c->launches;
    CK(c, cudaGetLastError());
    CKS(c, read_ctr(c, st));
    if (c->h_ctr->goal_rank != 0x7fffffffffffffffll) {
        s->ended = true;
        s->goal_rank = c->h_ctr->goal_rank;
        info->ended = 1;
        info->goal_rank = s->goal_rank;
        info->visited = (int64_t)c->occupied;
        info->table_slots = c->rnb;
        return SPL_OK;
    }
    int64_t total = 0, n_new = 0;
    CK(c, cudaEventRecord(c->ev[0], st));
    CKS(c, r_expand_all(c, front, n, 0, true, &total, st));
    CK(c, cudaEventRecord(c->ev[1], st));
    info->expanded = n;
    info->generated = total;
    if (total >= 0xFFFFFFFFll) return fail(c, SPL_E_INVALID, "level produced %lld candidates (>= 2^32)", (long long)total);
    if (total) {  // first-arrival dedup on the exact identity key (:840-843)
        CKS(c, zero_ctr(c, st));
        CKS(c, ensure_rtable(c, (uint64_t)total, st));
        uint64_t tag;
        CKS(c, next_epoch(c, tag));
        CK(c, c->cand_slot.ensure((size_t)total * 4, 0, st));
        CKS(c, zero_ctr(c, st));
        r_probe_kernel<<<nblk(total), TILE, 0, st>>>(c->rcand.as<RRec>(), total, c->rcfg.as<RConfigDev>(), c->rtable, c->rnb, tag,
                                                      c->cand_slot.as<uint32_t>(), c->d_ctr);
        ++c->launches;
        CK(c, cudaGetLastError());
        CKS(c, read_ctr(c, st));
        if (c->h_ctr->error == 4) return fail(c, SPL_E_INVALID, "a player saved 1024 gems or more: outside the packed identity key");
        if (c->h_ctr->error) return fail(c, SPL_E_TABLE_FULL, "visited table full during level %d", s->level);
        n_new = (int64_t)c->h_ctr->n_new;
        c->rocc += n_new;
        c->occupied += n_new;
    }
    info->unique = n_new;
    if (n_new) {
        CK(c, c->ridx64.ensure((size_t)n_new * 8, 0, st));
        const unsigned nt = nblk(total, TILE * 8);
        CKS(c, prep_status(c, 0, nt, st));
        CKS(c, reset_ticket(c, 0, st));
        r_winners_kernel<<<nt, TILE, 0, st>>>(c->cand_slot.as<uint32_t>(), c->rtable, total, c->ridx64.as<int64_t>(),
                                               c->status[0].as<uint64_t>(), c->d_ctr, 0);
        CK(c, s->uniq.ensure((size_t)n_new * 96, 0, st));
        r_gather_kernel<int64_t><<<nblk(n_new), TILE, 0, st>>>(c->rcand.as<RRec>(), c->ridx64.as<int64_t>(), n_new, s->uniq.as<RRec>());
        c->launches += 2;
        CK(c, cudaGetLastError());
        if (s->rcfg.noise == SPL_NOISE_EXTERNAL) {  // wait for the host's randint draws
            CK(c, cudaStreamSynchronize(st));
            s->pending = true;
            s->pend_n = n;
            s->pend_uniq = n_new;
            info->kept = -1;
            info->visited = (int64_t)c->occupied;
            info->table_slots = c->rnb;
            s->pend_info = *info;
            return SPL_OK;
        }
    }
    return realistic_cut(s, info, n, n_new, nullptr, st);
}

static int realistic_cut(spl_solver *s, spl_level_info *info, int64_t n, int64_t n_new, const uint8_t *draws, cudaStream_t st) {
    spl_ctx *c = s->c;
    int64_t kept = 0;
    if (n_new) {
        // score + beam cut, ties by arrival order (:846)
        CK(c, cudaEventRecord(c->ev[2], st));
        CKS(c, zero_ctr(c, st));
        CK(c, c->sk.ensure((size_t)n_new * 8, 0, st));
        r_score_kernel<<<nblk(n_new), TILE, 0, st>>>(s->uniq.as<RRec>(), n_new, c->rcfg.as<RConfigDev>(), c->luts,
                                                      c->sk.as<uint64_t>(), nullptr, c->d_ctr, draws);
        ++c->launches;
        CK(c, cudaGetLastError());
        CKS(c, read_ctr(c, st));
        const uint64_t smin = c->h_ctr->sk_min, smax = c->h_ctr->sk_max;
        int which = 0;
        CKS(c, run_cut_sort(c, c->sk.as<uint64_t>(), nullptr, n_new, s->beam, smin, smax, 0, &which, &kept, st));
        if (s->level < 1000) {  // at the turn limit the queue just expanded stays in place (see below)
            CK(c, s->front.ensure((size_t)kept * 96, 0, st));
            r_gather_kernel<uint32_t><<<nblk(kept), TILE, 0, st>>>(s->uniq.as<RRec>(), c->idx[which].as<uint32_t>(), kept, s->front.as<RRec>());
            ++c->launches;
        }
        CK(c, cudaGetLastError());
        CK(c, cudaEventRecord(c->ev[3], st));
        CK(c, cudaStreamSynchronize(st));
        float t;
        cudaEventElapsedTime(&t, c->ev[2], c->ev[3]);
        info->ms_select = t;
    }
    info->kept = kept;
    info->visited = (int64_t)c->occupied;
    info->table_slots = c->rnb;
    // queue exhausted, or the turn limit: the reference leaves its loop after the iteration with turn == 1000
    // (src/solver.py:846-852), so `puzzle` is the last state dequeued from THAT queue and the path has 1001 states
    if (kept == 0 || s->level >= 1000) {
        s->ended = true;
        s->goal_rank = n - 1;
        info->ended = 1;
        return SPL_OK;
    }
    s->n_front = kept;
    s->level += 1;
    CKS(c, save_links(s, st));
    CK(c, cudaStreamSynchronize(st));
    return SPL_OK;
}

extern "C" {

int32_t spl_solver_create(spl_ctx *c, const spl_key *root_key, uint64_t root_aux, int32_t goal, int32_t use_h,
                          int32_t heuristic, int64_t beam, int32_t tie, int32_t noise, int32_t keep_links,
                          spl_solver **out) {
    if (!c || !root_key || !out) return fail(c, SPL_E_INVALID, "spl_solver_create: null argument");
    if (use_h && beam < 1) return fail(c, SPL_E_INVALID, "spl_solver_create: beam_width must be >= 1");
    if (use_h && tie != SPL_TIE_STABLE && tie != SPL_TIE_KEY) return fail(c, SPL_E_INVALID, "spl_solver_create: unknown tie policy %d", tie);
    if (noise < 0 || noise > SPL_NOISE_EXTERNAL) return fail(c, SPL_E_INVALID, "spl_solver_create: unknown noise policy %d", noise);
    NO_LIVE_SOLVER(c, "spl_solver_create");
    CK(c, enter_device(c));
    cudaStream_t st = 0;
    spl_solver *s = new spl_solver();
    s->c = c; s->goal = goal; s->use_h = use_h; s->heuristic = heuristic; s->beam = beam; s->tie = tie;
    s->noise = noise; s->keep_links = keep_links;
    // beam search with on-device scoring runs on the card-set-grouped level; exhaustive BFS (queue order == arrival
    // order), the reference's hash identity and externally drawn noise (arrival-ordered draws) on the key table
    s->grouped = use_h && noise != SPL_NOISE_EXTERNAL && c->identity == IDENT_KEY && !getenv("SPL_NO_GROUPED");
    s->front.swap(c->pool_front);
    s->uniq.swap(c->pool_uniq);
    std::sort(c->pool_links.begin(), c->pool_links.end(), [](DevBuf *a, DevBuf *b) { return a->cap > b->cap; });
    int rc = reset_visited(c, st);
    if (rc == SPL_OK) {
        Rec r{root_key->lo, root_key->hi & HI_KEY_MASK, root_aux, ~0ull};
        cudaError_t e = s->front.ensure(32, 0, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s->front.p, &r, 32, cudaMemcpyHostToDevice, st);
        c->h2d_bytes += 32;
        if (e != cudaSuccess) rc = fail(c, SPL_E_CUDA, "root upload: %s", cudaGetErrorString(e));
    }
    if (rc == SPL_OK) rc = zero_ctr(c, st);
    if (rc == SPL_OK) {  // trail = {self: None}
        uint64_t tag;
        rc = next_epoch(c, tag);
        if (rc == SPL_OK) {
            cudaError_t e = c->cand_slot.ensure(4, 0, st);
            if (e != cudaSuccess) rc = fail(c, SPL_E_NOMEM, "cand_slot alloc");
            else if (s->grouped) {
                m2_root_kernel<<<1, 1, 0, st>>>(c->nodes, c->nn, root_key->lo, root_key->hi & HI_KEY_MASK, c->d_gemrank, c->d_ctr);
                ++c->launches;
                c->occupied = 1;
                c->node_occ = 1;
            } else {
                if (c->identity == IDENT_PYHASH)
                    probe_list_kernel<IDENT_PYHASH><<<1, TILE, 0, st>>>(reinterpret_cast<const spl_key *>(s->front.p), 1, c->table,
                                                                        c->nb, tag, c->cand_slot.as<uint32_t>(), c->d_ctr);
                else
                    probe_list_kernel<IDENT_KEY><<<1, TILE, 0, st>>>(reinterpret_cast<const spl_key *>(s->front.p), 1, c->table,
                                                                     c->nb, tag, c->cand_slot.as<uint32_t>(), c->d_ctr);
                ++c->launches;
                c->occupied = 1;
            }
        }
    }
    s->n_front = 1;
    if (rc == SPL_OK) rc = save_links(s, st);
    if (rc == SPL_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = fail(c, SPL_E_CUDA, "solver create sync failed");
    if (rc != SPL_OK) { delete s; return rc; }
    c->active = s;
    *out = s;
    return SPL_OK;
}

int32_t spl_solver_destroy(spl_solver *s) {
    if (s) { cudaSetDevice(s->c->device); cudaDeviceSynchronize(); delete s; }
    return SPL_OK;
}

int32_t spl_solver_step(spl_solver *s, spl_level_info *info, void *stream) {
    if (!s || !info) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    if (s->ended) return fail(c, SPL_E_STATE, "spl_solver_step: the search has already ended");
    if (s->pending) return fail(c, SPL_E_STATE, "spl_solver_step: the previous level still waits for spl_solver_cut");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    if (s->realistic) return rsolver_step(s, info, st);
    memset(info, 0, sizeof *info);
    info->level = s->level;
    info->frontier = s->n_front;
    info->goal_rank = -1;
    const int64_t n = s->n_front;
    const Rec *front = s->front.as<Rec>();
    float ms[6] = {0, 0, 0, 0, 0, 0};  // count, expand, resolve, select, sort, (grouped level) warp kernel
    // ---- goal test on the queue (src/solver.py:443-445): the first state in queue order with
    // pts >= goal ends the search; the states before it would be expanded and discarded.
    CKS(c, zero_ctr(c, st));
    goal_kernel<<<nblk(n), TILE, 0, st>>>(front, n, s->goal, c->d_ctr);
    ++c->launches;
    CK(c, cudaGetLastError());
    CKS(c, read_ctr(c, st));
    if (c->h_ctr->goal_rank != 0x7fffffffffffffffll) {
        s->ended = true;
        s->goal_rank = c->h_ctr->goal_rank;
        info->ended = 1;
        info->goal_rank = s->goal_rank;
        info->visited = (int64_t)c->occupied;
        info->table_slots = s->grouped ? c->nn : c->cap;
        return SPL_OK;
    }
    // ---- expand + dedup by chunks of parents
    int64_t n_uniq = 0, generated = 0;
    uint64_t sk_min = ~0ull, sk_max = 0;
    if (s->grouped) {
        int64_t n_slots = 0;
        CKS(c, grouped_expand(s, front, n, &n_uniq, &n_slots, &generated, &sk_min, &sk_max, ms, st));
        info->expanded = n;
        info->generated = generated;
        info->unique = n_uniq;
        info->ms_count = ms[0]; info->ms_expand = ms[1]; info->ms_resolve = ms[2]; info->ms_warp = ms[5];
        return speedrun_cut(s, info, n, n_uniq, sk_min, sk_max, st, n_slots);
    }
    for (int64_t p0 = 0; p0 < n; p0 += (int64_t)c->chunk_parents) {
        const int64_t np = std::min<int64_t>((int64_t)c->chunk_parents, n - p0);
        const unsigned nt = nblk(np);
        CKS(c, zero_ctr(c, st));
        CK(c, cudaEventRecord(c->ev[0], st));
        CKS(c, run_count(c, front + p0, np, st));
        CK(c, cudaEventRecord(c->ev[7], st));
        const uint64_t total = c->h_ctr->total_cands;
        generated += (int64_t)total;
        if (total == 0) continue;
        if (total >= 0xFFFFFFFFull) return fail(c, SPL_E_INVALID, "chunk produced %llu candidates (>= 2^32)", (unsigned long long)total);
        CKS(c, ensure_table(c, total, st));
        uint64_t tag;
        CKS(c, next_epoch(c, tag));
        CK(c, c->cand_slot.ensure((size_t)total * 4, 0, st));
        CK(c, cudaEventRecord(c->ev[7], st));
        if (c->identity == IDENT_PYHASH)
            expand_kernel<MODE_PROBE, IDENT_PYHASH><<<nt, TILE, sizeof(ExpandSmem2), st>>>(
                front + p0, np, c->d_tabs, c->d_takes_idx, c->d_takes_edges, c->off.as<uint32_t>(), (uint32_t)total,
                c->table, c->nb, tag, c->cand_slot.as<uint32_t>(), nullptr, p0, c->d_ctr);
        else
            expand_kernel<MODE_PROBE, IDENT_KEY><<<nt, TILE, sizeof(ExpandSmem2), st>>>(
                front + p0, np, c->d_tabs, c->d_takes_idx, c->d_takes_edges, c->off.as<uint32_t>(), (uint32_t)total,
                c->table, c->nb, tag, c->cand_slot.as<uint32_t>(), nullptr, p0, c->d_ctr);
        ++c->launches;
        CK(c, cudaGetLastError());
        CK(c, cudaEventRecord(c->ev[1], st));
        CKS(c, read_ctr(c, st));
        if (c->h_ctr->error) return fail(c, SPL_E_TABLE_FULL, "visited table full during level %d (slots=%llu, occupied=%llu)",
                                         s->level, (unsigned long long)c->cap, (unsigned long long)c->occupied);
        const int64_t n_new = (int64_t)c->h_ctr->n_new;
        c->occupied += n_new;
        if (n_new) {
            CK(c, s->uniq.ensure((size_t)(n_uniq + n_new) * 32, (size_t)n_uniq * 32, st));
            if (s->use_h) CK(c, c->sk.ensure((size_t)(n_uniq + n_new) * 8, (size_t)n_uniq * 8, st));
            CKS(c, prep_status(c, 0, nt, st));
            CKS(c, reset_ticket(c, 0, st));
            CK(c, cudaEventRecord(c->ev[2], st));
            if (s->use_h && s->noise != SPL_NOISE_EXTERNAL)
                resolve_kernel<SRC_PARENT, true><<<nt, TILE, sizeof(ResolveSmem), st>>>(
                    front + p0, np, c->d_tabs, c->d_takes_idx, c->d_takes_edges, c->off.as<uint32_t>(), (uint32_t)total,
                    c->table, c->cand_slot.as<uint32_t>(), nullptr, nullptr, p0, (uint64_t)n_uniq, s->uniq.as<Rec>(),
                    c->sk.as<uint64_t>(), nullptr, s->heuristic, s->noise, c->luts, c->status[0].as<uint64_t>(), c->d_ctr, 0);
            else
                resolve_kernel<SRC_PARENT, false><<<nt, TILE, sizeof(ResolveSmem), st>>>(
                    front + p0, np, c->d_tabs, c->d_takes_idx, c->d_takes_edges, c->off.as<uint32_t>(), (uint32_t)total,
                    c->table, c->cand_slot.as<uint32_t>(), nullptr, nullptr, p0, (uint64_t)n_uniq, s->uniq.as<Rec>(),
                    nullptr, nullptr, s->heuristic, s->noise, c->luts, c->status[0].as<uint64_t>(), c->d_ctr, 0);
            ++c->launches;
            CK(c, cudaGetLastError());
            CK(c, cudaEventRecord(c->ev[3], st));
            CKS(c, read_ctr(c, st));
            if ((int64_t)c->h_ctr->n_emitted != n_uniq + n_new)
                return fail(c, SPL_E_CUDA, "internal: emitted %llu winners, expected %lld", (unsigned long long)c->h_ctr->n_emitted,
                            (long long)(n_uniq + n_new));
            if (s->use_h && s->noise != SPL_NOISE_EXTERNAL) { sk_min = std::min<uint64_t>(sk_min, c->h_ctr->sk_min); sk_max = std::max<uint64_t>(sk_max, c->h_ctr->sk_max); }
            n_uniq += n_new;
            float t;
            cudaEventElapsedTime(&t, c->ev[2], c->ev[3]);
            ms[2] += t;
        }
        float t;
        cudaEventElapsedTime(&t, c->ev[0], c->ev[7]);
        ms[0] += t;
        cudaEventElapsedTime(&t, c->ev[7], c->ev[1]);
        ms[1] += t;
    }
    info->expanded = n;
    info->generated = generated;
    info->unique = n_uniq;
    info->ms_count = ms[0]; info->ms_expand = ms[1]; info->ms_resolve = ms[2];
    if (s->use_h && s->noise == SPL_NOISE_EXTERNAL && n_uniq > 0) {  // wait for the host's randint draws
        s->pending = true;
        s->pend_n = n;
        s->pend_uniq = n_uniq;
        info->kept = -1;
        info->visited = (int64_t)c->occupied;
        info->table_slots = c->cap;
        s->pend_info = *info;
        return SPL_OK;
    }
    return speedrun_cut(s, info, n, n_uniq, sk_min, sk_max, st);
}

int32_t spl_rexpand(spl_ctx *c, const spl_rconfig *cfg, const void *recs, int64_t n, void *out_recs, int64_t cap,
                    int64_t *n_out, void *stream) {
    if (!c || !n_out || n < 0 || n > (16ll << 20)) return fail(c, SPL_E_INVALID, "spl_rexpand: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    *n_out = 0;
    if (n == 0) return SPL_OK;
    CKS(c, upload_rconfig(c, cfg, st));
    int64_t total = 0;
    CKS(c, r_expand_all(c, reinterpret_cast<const RRec *>(recs), n, 0, false, &total, st));
    *n_out = total;
    if (total > cap) return fail(c, SPL_E_CAPACITY, "spl_rexpand: %lld successors, capacity %lld", (long long)total, (long long)cap);
    if (total) CK(c, cudaMemcpyAsync(out_recs, c->rcand.p, (size_t)total * 96, cudaMemcpyDeviceToDevice, st));
    CK(c, cudaStreamSynchronize(st));
    return SPL_OK;
}

int32_t spl_rscore(spl_ctx *c, const spl_rconfig *cfg, const void *recs, int64_t n, double *scores, void *stream) {
    if (!c || n < 0) return fail(c, SPL_E_INVALID, "spl_rscore: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    if (n == 0) return SPL_OK;
    CKS(c, upload_rconfig(c, cfg, st));
    r_score_kernel<<<nblk(n), TILE, 0, st>>>(reinterpret_cast<const RRec *>(recs), n, c->rcfg.as<RConfigDev>(), c->luts,
                                              nullptr, scores, c->d_ctr);
    ++c->launches;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

int32_t spl_rpack(spl_ctx *c, const spl_rconfig *cfg, const void *recs, int64_t n, uint64_t *keys_out, void *stream) {
    if (!c || n < 0) return fail(c, SPL_E_INVALID, "spl_rpack: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    if (n == 0) return SPL_OK;
    CKS(c, upload_rconfig(c, cfg, st));
    r_pack_kernel<<<nblk(n), TILE, 0, st>>>(reinterpret_cast<const RRec *>(recs), n, c->rcfg.as<RConfigDev>(), keys_out);
    ++c->launches;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

int32_t spl_rmaxpts(spl_ctx *c, const spl_rconfig *cfg, const void *recs, int64_t n, uint8_t *out, void *stream) {
    if (!c || n < 0) return fail(c, SPL_E_INVALID, "spl_rmaxpts: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    if (n == 0) return SPL_OK;
    CKS(c, upload_rconfig(c, cfg, st));
    r_maxpts_kernel<<<nblk(n), TILE, 0, st>>>(reinterpret_cast<const RRec *>(recs), n, c->rcfg.as<RConfigDev>(), out);
    ++c->launches;
    CK(c, cudaGetLastError());
    return SPL_OK;
}

int32_t spl_rsolver_create(spl_ctx *c, const spl_rconfig *cfg, const void *root_rec_host, int64_t beam, int32_t keep_links,
                           spl_solver **out) {
    if (!c || !cfg || !root_rec_host || !out) return fail(c, SPL_E_INVALID, "spl_rsolver_create: null argument");
    if (beam < 1) return fail(c, SPL_E_INVALID, "spl_rsolver_create: beam_width must be >= 1");
    NO_LIVE_SOLVER(c, "spl_rsolver_create");
    CK(c, enter_device(c));
    cudaStream_t st = 0;
    CKS(c, upload_rconfig(c, cfg, st));
    spl_solver *s = new spl_solver();
    s->c = c; s->realistic = true; s->rcfg = *cfg; s->use_h = 1; s->beam = beam; s->keep_links = keep_links;
    s->goal = cfg->target_points;
    s->front.swap(c->pool_front);
    s->uniq.swap(c->pool_uniq);
    c->rcfg_dirty = false;
    int rc = reset_visited(c, st);
    if (rc == SPL_OK) {
        cudaError_t e = s->front.ensure(96, 0, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s->front.p, root_rec_host, 96, cudaMemcpyHostToDevice, st);
        c->h2d_bytes += 96;
        if (e != cudaSuccess) rc = fail(c, SPL_E_CUDA, "root upload: %s", cudaGetErrorString(e));
    }
    if (rc == SPL_OK) {  // trail = {self: None}: register the root's identity key
        s->n_front = 1;
        cudaError_t e = c->rcand.ensure(96, 0, st);
        if (e == cudaSuccess) e = c->cand_slot.ensure(4, 0, st);
        if (e != cudaSuccess) rc = fail(c, SPL_E_NOMEM, "realistic scratch alloc");
    }
    if (rc == SPL_OK) rc = zero_ctr(c, st);
    if (rc == SPL_OK) rc = ensure_rtable(c, 1, st);
    if (rc == SPL_OK && c->rocc) {  // forget the previous search
        if (cudaMemsetAsync(c->rtable, 0, c->rnb * 64, st) != cudaSuccess) rc = fail(c, SPL_E_CUDA, "realistic table reset failed");
        c->rocc = 0;
    }
    if (rc == SPL_OK) {
        uint64_t tag;
        rc = next_epoch(c, tag);
        if (rc == SPL_OK) {
            r_probe_kernel<<<1, TILE, 0, st>>>(s->front.as<RRec>(), 1, c->rcfg.as<RConfigDev>(), c->rtable, c->rnb, tag,
                                                c->cand_slot.as<uint32_t>(), c->d_ctr);
            ++c->launches;
            c->rocc = 1;
            c->occupied = 1;
        }
    }
    if (rc == SPL_OK) rc = save_links(s, st);
    if (rc == SPL_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = fail(c, SPL_E_CUDA, "rsolver create sync failed");
    if (rc != SPL_OK) { delete s; return rc; }
    c->active = s;
    *out = s;
    return SPL_OK;
}

int32_t spl_rsolver_frontier(spl_solver *s, const void **recs, int64_t *n) {
    if (!s || !recs || !n || !s->realistic) return SPL_E_INVALID;
    *recs = s->front.p;
    *n = s->n_front;
    return SPL_OK;
}

int32_t spl_solver_cut(spl_solver *s, const uint8_t *draws, int64_t n_draws, spl_level_info *info, void *stream) {
    if (!s || !info) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    if (!s->pending) return fail(c, SPL_E_STATE, "spl_solver_cut: no level is waiting for draws");
    if (!draws || n_draws != s->pend_uniq) return fail(c, SPL_E_INVALID, "spl_solver_cut: need exactly %lld draws", (long long)s->pend_uniq);
    cudaStream_t st = (cudaStream_t)stream;
    CK(c, enter_device(c));
    *info = s->pend_info;
    s->pending = false;
    if (s->realistic) return realistic_cut(s, info, s->pend_n, s->pend_uniq, draws, st);
    CKS(c, zero_ctr(c, st));
    CK(c, c->sk.ensure((size_t)s->pend_uniq * 8, 0, st));
    score_ext_kernel<<<nblk(s->pend_uniq), TILE, 0, st>>>(s->uniq.as<Rec>(), draws, s->pend_uniq, s->heuristic, c->luts,
                                                            c->sk.as<uint64_t>(), c->d_ctr);
    ++c->launches;
    CK(c, cudaGetLastError());
    CKS(c, read_ctr(c, st));
    return speedrun_cut(s, info, s->pend_n, s->pend_uniq, c->h_ctr->sk_min, c->h_ctr->sk_max, st);
}

int32_t spl_solver_frontier(spl_solver *s, const spl_key **keys, const uint64_t **aux, const uint64_t **link, int64_t *n) {
    // AoS view: the three pointers address fields of 32-byte records (stride 32 bytes).
    if (!s || !n) return SPL_E_INVALID;
    const char *b = reinterpret_cast<const char *>(s->front.p);
    if (keys) *keys = reinterpret_cast<const spl_key *>(b);
    if (aux) *aux = reinterpret_cast<const uint64_t *>(b + 16);
    if (link) *link = reinterpret_cast<const uint64_t *>(b + 24);
    *n = s->n_front;
    return SPL_OK;
}

int32_t spl_solver_path(spl_solver *s, int64_t *ranks, int32_t *ordinals, int32_t cap, int32_t *n_moves) {
    if (!s || !ranks || !ordinals || !n_moves) return SPL_E_INVALID;
    spl_ctx *c = s->c;
    if (!s->ended) return fail(c, SPL_E_STATE, "spl_solver_path: the search has not ended");
    if (!s->keep_links) return fail(c, SPL_E_STATE, "spl_solver_path: solver was created with keep_links = 0");
    const int L = s->level;  // level index of the final state
    if (L + 1 > cap) return fail(c, SPL_E_CAPACITY, "spl_solver_path: need %d entries", L + 1);
    CK(c, enter_device(c));
    int64_t r = s->goal_rank;
    for (int l = L; l >= 0; --l) {
        ranks[l] = r;
        if (l > 0) {
            uint64_t link = 0;
            CK(c, s->links.read(l, r, &link));
            if (!s->links.host[l]) c->d2h_bytes += 8;
            ordinals[l - 1] = (int32_t)(link & 0xff);
            r = (int64_t)(link >> 8);
        }
    }
    *n_moves = L;
    return SPL_OK;
}

}  // extern "C"
