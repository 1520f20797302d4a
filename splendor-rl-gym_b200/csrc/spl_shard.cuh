// spl_shard.cuh -- kernels of the SHARDED card-set-grouped level (one process per GPU).
//
// Ownership is by CARD SET: rank = f(hash of the 90-bit card mask).  A queue state lives on the rank that owns its
// cards, so its gem-take successors (same cards; ~2/3 of all candidates) are deduplicated on the generating GPU by
// the on-chip walk of spl_m2.cuh and never leave it; only card buys (new card set) are routed, as 32-byte records,
// straight into per-destination ranges of a send buffer by the kernel that generates them.  The owner sorts its
// parents and the records it received by card set and runs the same per-run dedup.
//
// Records received from several ranks are NOT in global arrival order (GroupArgs::unordered): the thread kernel
// compares the arrival indices (link = global parent rank << 8 | ordinal) of equal gem hands, the warp kernel first
// sorts the (link, index) pairs of a run's records in shared memory and then walks them as on one GPU, and the CTA
// kernel takes the minimum arrival index per gem hand anyway.
#pragma once
#include "spl_m2.cuh"

namespace spl {

constexpr int MAX_RANKS = 16;

// owner rank of a card set: low half of its hash (the sort key and the node slot use the high half)
__device__ __host__ __forceinline__ uint32_t owner_of_mask(uint64_t m0, uint64_t m1, uint32_t world) {
    return (uint32_t)(((mask_hash(m0, m1) & 0xFFFFFFFFull) * world) >> 32);
}

// ------------------------------------------------------------------ fan-out + per-destination buy counts
// as m2_count_kernel, plus dest_cnt[d] += card buys of the round whose new card set is owned by rank d
__global__ void __launch_bounds__(TILE) gs_count_kernel(const Rec *__restrict__ front, int64_t np,
                                                        const DevTables *__restrict__ tabs,
                                                        const uint32_t *__restrict__ takes_idx, uint64_t *__restrict__ iv,
                                                        uint8_t *__restrict__ ntk8, uint32_t world,
                                                        unsigned long long *__restrict__ dest_cnt, Counters *ctr) {
    __shared__ SmemTabs s;
    __shared__ uint32_t warp_sums[TILE / 32 + 1];
    __shared__ uint32_t s_dest[MAX_RANKS];
    load_tabs(s, tabs);
    if (threadIdx.x < MAX_RANKS) s_dest[threadIdx.x] = 0;
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * TILE + threadIdx.x;
    uint32_t nb = 0, ntk = 0;
    uint64_t pk0 = 0, pk1 = 0;
    if (p < np) {
        Rec r;
        ld_rec(front + p, r);
        uint64_t bl, bh, m0, m1;
        uint32_t tk;
        derive_parent(s, takes_idx, r.lo, r.hi, r.aux, bl, bh, nb, tk);
        ntk = tk & 0xff;
        mask_words(r.lo, r.hi, m0, m1);
        iv[p] = (mask_hash(m0, m1) & 0xFFFFFFFF00000000ull) | (uint64_t)p;
        ntk8[p] = (uint8_t)ntk;
        // key bit 15 + i <-> card i <-> mask bit i; per-thread counts packed 8 bits per destination (<= 90 buys)
        for (uint64_t m = bl >> 15; m; m &= m - 1) {
            const uint32_t d = owner_of_mask(m0 | (m & (~m + 1)), m1, world);
            if (d < 8) pk0 += 1ull << (8 * d); else pk1 += 1ull << (8 * (d - 8));
        }
        for (uint64_t m = bh; m; m &= m - 1) {
            const int i = 49 + __ffsll((long long)m) - 1;  // key.hi bit b <-> card 49 + b
            const uint64_t c0 = i < 64 ? m0 | (1ull << i) : m0, c1 = i < 64 ? m1 : m1 | (1ull << (i - 64));
            const uint32_t d = owner_of_mask(c0, c1, world);
            if (d < 8) pk0 += 1ull << (8 * d); else pk1 += 1ull << (8 * (d - 8));
        }
    }
    for (uint32_t d = 0; d < world; ++d) {  // warp sums, then one shared-memory add per warp and destination
        uint32_t v = (uint32_t)((d < 8 ? pk0 >> (8 * d) : pk1 >> (8 * (d - 8))) & 0xff);
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_dest[d], v);
    }
    __syncthreads();
    uint32_t total_tk, total_b;
    block_excl_scan(ntk, warp_sums, total_tk);
    block_excl_scan(nb, warp_sums, total_b);
    if (threadIdx.x == 0) {
        if (total_tk) atomicAdd(&ctr->total_cands, (unsigned long long)total_tk);
        if (total_b) atomicAdd(&ctr->n_buys, (unsigned long long)total_b);
    }
    if (threadIdx.x < world && s_dest[threadIdx.x]) atomicAdd(&dest_cnt[threadIdx.x], (unsigned long long)s_dest[threadIdx.x]);
}

// card buys of the round's parents as records, written into the send range of the rank that owns the new card set
// (cursor[d] = next free record of destination d; order inside a range is arbitrary)
struct RouteSmem {
    SmemTabs tabs;
    uint64_t lo[TILE], hi[TILE], aux[TILE];
    uint32_t prefb[TILE + 1];
    uint16_t blist[BUY_WIN];
    uint32_t where[BUY_WIN];  // destination << 24 | index among this window's records for that destination
    uint16_t order[BUY_WIN];  // window records grouped by destination: consecutive threads store consecutive slots
    uint32_t dstart[MAX_RANKS + 1];
    uint32_t warp_sums[TILE / 32 + 1];
    uint32_t cnt[MAX_RANKS];
    unsigned long long base[MAX_RANKS];
    Rec *dst[MAX_RANKS];
};
__global__ void __launch_bounds__(TILE) gs_buys_route_kernel(const Rec *__restrict__ front, int64_t np,
                                                             const uint64_t *__restrict__ grank,
                                                             const DevTables *__restrict__ tabs,
                                                             const uint32_t *__restrict__ takes_idx, uint32_t world,
                                                             Rec *const *__restrict__ dst, unsigned long long *cursor) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RouteSmem &S = *reinterpret_cast<RouteSmem *>(smem_raw);
    const unsigned tid = threadIdx.x;
    load_tabs(S.tabs, tabs);
    if (tid < world) S.dst[tid] = dst[tid];
    __syncthreads();
    const int64_t p0 = (int64_t)blockIdx.x * TILE, p = p0 + tid;
    uint64_t bm_lo = 0, bm_hi = 0;
    uint32_t nb = 0, tk = 0;
    if (p < np) {
        Rec r;
        ld_rec(front + p, r);
        derive_parent(S.tabs, takes_idx, r.lo, r.hi, r.aux, bm_lo, bm_hi, nb, tk);
        S.lo[tid] = r.lo; S.hi[tid] = r.hi; S.aux[tid] = r.aux;
    }
    uint32_t total_buys;
    const uint32_t prefb = block_excl_scan(nb, S.warp_sums, total_buys);
    S.prefb[tid] = prefb;
    if (tid == 0) S.prefb[TILE] = total_buys;
    for (uint32_t w0 = 0; w0 < total_buys; w0 += BUY_WIN) {
        __syncthreads();
        if (tid < MAX_RANKS) S.cnt[tid] = 0;
        {
            uint32_t q = prefb;
            uint64_t m = bm_lo;
            while (m) {
                const int pos = __ffsll((long long)m) - 1;
                m &= m - 1;
                if (q >= w0 && q < w0 + BUY_WIN) S.blist[q - w0] = (uint16_t)(tid << 7 | pos);
                ++q;
            }
            m = bm_hi;
            while (m) {
                const int pos = 64 + __ffsll((long long)m) - 1;
                m &= m - 1;
                if (q >= w0 && q < w0 + BUY_WIN) S.blist[q - w0] = (uint16_t)(tid << 7 | pos);
                ++q;
            }
        }
        __syncthreads();
        const uint32_t nbw = min((uint32_t)BUY_WIN, total_buys - w0);
        for (uint32_t i = tid; i < nbw; i += TILE) {  // destination of every record of the window
            const uint32_t ent = S.blist[i], j = ent >> 7;
            const int pos = ent & 127;
            uint64_t klo = S.lo[j], khi = S.hi[j], m0, m1;
            if (pos < 64) klo |= 1ull << pos; else khi |= 1ull << (pos - 64);
            mask_words(klo, khi, m0, m1);
            const uint32_t d = owner_of_mask(m0, m1, world);
            S.where[i] = d << 24 | atomicAdd(&S.cnt[d], 1u);
        }
        __syncthreads();
        if (tid < world && S.cnt[tid]) S.base[tid] = atomicAdd(&cursor[tid], (unsigned long long)S.cnt[tid]);
        if (tid == 0) {
            uint32_t run = 0;
            for (uint32_t d = 0; d < world; ++d) { S.dstart[d] = run; run += S.cnt[d]; }
        }
        __syncthreads();
        for (uint32_t i = tid; i < nbw; i += TILE) S.order[S.dstart[S.where[i] >> 24] + (S.where[i] & 0xffffffu)] = (uint16_t)i;
        __syncthreads();
        // one destination after the other, slot after slot: a warp stores 1 KB of consecutive records (over NVLink when
        // the destination is a peer)
        for (uint32_t t = tid; t < nbw; t += TILE) {
            const uint32_t i = S.order[t];
            const uint32_t ent = S.blist[i], j = ent >> 7;
            const int pos = ent & 127;
            const uint64_t lo = S.lo[j], aux = S.aux[j];
            uint32_t saved, cd;
            const uint32_t ng = buy_gems(S.tabs, lo, aux, pos, saved, cd);
            uint64_t klo = (lo & ~GEM_MASK) | ng, khi = S.hi[j];
            if (pos < 64) klo |= 1ull << pos; else khi |= 1ull << (pos - 64);
            const uint32_t ord = w0 + i - S.prefb[j];
            const uint64_t caux = aux + saved + ((uint64_t)((cd >> 15) & 7) << 16) + (1ull << (24 + 5 * ((cd >> 18) & 7)));
            Rec r{klo, khi, caux, (grank[p0 + j] << 8) | ord};
            // dst[d]: this rank's range of destination d's records -- a slice of the local send buffer (NCCL exchange) or
            // of rank d's receive buffer itself, mapped over NVLink (the store IS the transfer)
            st_rec(S.dst[S.where[i] >> 24] + S.base[S.where[i] >> 24] + (S.where[i] & 0xffffffu), r);
        }
    }
}

// items of the records received in a round: hash half << 32 | (np + index)
__global__ void __launch_bounds__(TILE) gs_recv_items_kernel(const Rec *__restrict__ recv, int64_t n, uint32_t np,
                                                             uint64_t *__restrict__ iv) {
    const int64_t j = (int64_t)blockIdx.x * TILE + threadIdx.x;
    if (j >= n) return;
    uint64_t lo, hi, m0, m1;
    ld_cg_u64x2(reinterpret_cast<const uint64_t *>(recv + j), lo, hi);
    mask_words(lo, hi, m0, m1);
    iv[np + j] = (mask_hash(m0, m1) & 0xFFFFFFFF00000000ull) | (uint64_t)(np + j);
}

// ------------------------------------------------------------------ dictionary merge (global beam threshold)
// every rank's score dictionary (all-gathered) -> one dictionary, identical on all ranks
__global__ void __launch_bounds__(TILE) gs_dict_merge_kernel(const ScoreDict *__restrict__ all, int world, ScoreDict *merged) {
    const int i = blockIdx.x * TILE + threadIdx.x;  // slot of one source dictionary
    if (i >= world * DICT_CAP) return;
    const ScoreDict &d = all[i / DICT_CAP];
    const int s = i % DICT_CAP;
    if (d.over && s == 0) atomicExch(&merged->over, 1u);
    if (d.key[s] != DICT_EMPTY && d.cnt[s]) dict_add(merged, d.key[s], d.cnt[s]);
}

// lower bounds of a few probes in an ascending u64 array (sample-sort partition of the survivors' sort words)
__global__ void gs_lower_bound_kernel(const uint64_t *__restrict__ a, int64_t n, const uint64_t *__restrict__ probes, int np,
                                      int64_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const uint64_t v = probes[i];
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    out[i] = lo;
}
// out[pos[i]] = base + i : global ranks of the received sort words, back in receive order
__global__ void __launch_bounds__(TILE) gs_scatter_ranks_kernel(const uint32_t *__restrict__ pos, int64_t n, int64_t base,
                                                                int64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * TILE + threadIdx.x;
    if (i < n) out[pos[i]] = base + i;
}
__global__ void __launch_bounds__(TILE) gs_iota_kernel(uint32_t *__restrict__ idx, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * TILE + threadIdx.x;
    if (i < n) idx[i] = (uint32_t)i;
}
// first index with pts >= goal -> its global rank
__global__ void gs_rank_at_kernel(const uint64_t *__restrict__ grank, int64_t idx, int64_t *out) { *out = (int64_t)grank[idx]; }

}  // namespace spl
