// spl_m2.cuh -- card-set-grouped frontier expansion (the beam-search level of the speedrun solver).
//
// The visited set of State.solve (`trail`, src/solver.py:426, :447-450) is keyed by (cards, gems).  A level
// of the search touches few distinct card sets (about one per ten queue states, measured) and many gem
// hands per card set: a gem take keeps the card set, so the ~16 take successors of a parent and of every
// other parent with the same cards fall into ONE 2898-entry space (five gem counts 0..7, at most 10 in
// hand, src/gems.py:9,40-51).  This path therefore deduplicates per card set instead of per key:
//
//   visited set  = open-addressed table of 384-byte NODES, one per card set that was ever generated:
//                  {90-bit card mask, 2898-bit bitmap over the gem hands}  (`trail` membership = one bit)
//   one level    = 1. m2_count / m2_buys : per-parent fan-out; card buys are written out as 32-byte
//                     records (they change the card set); gem takes are NOT materialised
//                  2. LSD radix sort of {parents, buy records} by the hash of their (target) card set
//                  3. m2_runs : boundaries of equal-hash runs + candidate-weight prefix
//                  4. m2_group_small / m2_group_big : one warp / one CTA per run loads the node, replays
//                     the takes of the run's parents and the run's buy records against an ON-CHIP table
//                     (min arrival index per gem hand, src/solver.py:447-450 first arrival wins), emits the
//                     winners with their scores and sets their bits.  No global atomics per candidate,
//                     no second visit to the table.
// Sorting by the hash that also selects the node makes the node accesses of a level monotone in memory.
// Winners are emitted in no particular order; link = parent_rank << 8 | ordinal carries the arrival
// order, and the beam cut breaks score ties on it (SPL_TIE_STABLE) or on the key (SPL_TIE_KEY).
#pragma once
#include "spl_kernels.cuh"

namespace spl {

constexpr int GEM_STATES = 2898;       // gem hands with every colour <= 7 and at most 10 gems
constexpr int NODE_BM_WORDS = 46;      // 2944 bits >= GEM_STATES
constexpr int NODE_WORDS = 48;         // 384 B = 6 DRAM bursts: {m0, m1 | OCC, bitmap[46]}
constexpr uint64_t NODE_OCC = 1ull << 63;
constexpr int SMALL_ITEMS = 32;        // runs handled by one warp: at most 32 items and SMALL_W candidates
constexpr int SMALL_W = 128;
constexpr int SM_TBL = 256;            // per-warp dedup slots (> SMALL_W, power of two)
constexpr int SM_STAGE = 48;           // per-warp winner staging (flushed above 16 entries)
constexpr int M2_WARPS = TILE / 32;
constexpr int BIG_DONE = 16;           // distinct card sets one equal-hash run of the CTA kernel may hold

// card mask of a key as two words (90 bits)
__device__ __host__ __forceinline__ void mask_words(uint64_t lo, uint64_t hi, uint64_t &m0, uint64_t &m1) {
    m0 = (lo >> 15) | (hi << 49);
    m1 = (hi & HI_KEY_MASK) >> 15;
}
__device__ __host__ __forceinline__ uint64_t mask_hash(uint64_t m0, uint64_t m1) { return hash_key(m0, m1); }

// ------------------------------------------------------------------ node table
// find the node of card set (m0, m1) or claim an empty one (128-bit CAS on the header).  A card set is
// looked up by exactly one warp / CTA per round (all its candidates sit in one run), so nothing races
// on a node's bitmap; claims of the same empty slot by different card sets are settled by the CAS.
__device__ __forceinline__ uint64_t node_find_or_create(uint64_t *__restrict__ nodes, uint64_t nn, uint64_t m0, uint64_t m1,
                                                        bool &fresh, unsigned int *error) {
    const uint64_t want1 = m1 | NODE_OCC;
    uint64_t i = __umul64hi(mask_hash(m0, m1), nn);
    fresh = false;
    for (int probes = 0; probes < MAX_PROBE; ++probes) {
        uint64_t *N = nodes + i * NODE_WORDS;
        uint64_t a, b;
        ld_cg_u64x2(N, a, b);
        if (b == 0) {  // empty (an occupied header always carries NODE_OCC)
            cas128(N, 0, 0, m0, want1, a, b);
            if ((a | b) == 0) { fresh = true; return i; }
        }
        if (a == m0 && b == want1) return i;
        if (++i == nn) i = 0;
    }
    atomicExch(error, 2u);
    return 0;
}
__device__ __forceinline__ bool node_bit(const uint64_t *N, uint32_t r) {
    uint64_t w;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(w) : "l"(N + 2 + (r >> 6)));
    return (w >> (r & 63)) & 1;
}

// trail = {root: None} (src/solver.py:426)
__global__ void m2_root_kernel(uint64_t *nodes, uint64_t nn, uint64_t lo, uint64_t hi, const uint16_t *__restrict__ gemrank,
                               Counters *ctr) {
    uint64_t m0, m1;
    mask_words(lo, hi, m0, m1);
    bool fresh;
    const uint64_t i = node_find_or_create(nodes, nn, m0, m1, fresh, &ctr->error);
    const uint32_t r = gemrank[lo & GEM_MASK];
    atomicOr(reinterpret_cast<unsigned long long *>(nodes + i * NODE_WORDS + 2 + (r >> 6)), 1ull << (r & 63));
    if (fresh) atomicAdd(&ctr->n_new_nodes, 1u);
}

// move every node of an old table into a larger one: one warp per old slot
__global__ void __launch_bounds__(TILE) m2_rehash_kernel(const uint64_t *__restrict__ old_nodes, uint64_t old_nn,
                                                         uint64_t *__restrict__ nodes, uint64_t nn, Counters *ctr) {
    const uint64_t s = ((uint64_t)blockIdx.x * TILE + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    if (s >= old_nn) return;
    const uint64_t *O = old_nodes + s * NODE_WORDS;
    const uint64_t m1 = O[1];
    if (m1 == 0) return;
    uint64_t i = 0;
    if (lane == 0) {
        bool fresh;
        i = node_find_or_create(nodes, nn, O[0], m1 & ~NODE_OCC, fresh, &ctr->error);
    }
    i = __shfl_sync(0xffffffffu, i, 0);
    for (int w = lane; w < NODE_BM_WORDS; w += 32) nodes[i * NODE_WORDS + 2 + w] = O[2 + w];
}

// ------------------------------------------------------------------ 1. fan-out of the round's parents
// per parent: buys offset (exclusive scan), number of takes, sort key = high half of the card-set hash
__global__ void __launch_bounds__(TILE) m2_count_kernel(const Rec *__restrict__ front, int64_t np,
                                                        const DevTables *__restrict__ tabs,
                                                        const uint32_t *__restrict__ takes_idx,
                                                        uint32_t *__restrict__ boff, uint64_t *__restrict__ ik,
                                                        uint32_t *__restrict__ iidx, uint8_t *__restrict__ ntk8,
                                                        uint64_t *status, Counters *ctr, int ticket_id) {
    __shared__ SmemTabs s;
    __shared__ uint32_t warp_sums[TILE / 32 + 1];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_base;
    load_tabs(s, tabs);
    if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->ticket[ticket_id], 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t p = (int64_t)tile * TILE + threadIdx.x;
    uint32_t nb = 0, ntk = 0;
    if (p < np) {
        Rec r;
        ld_rec(front + p, r);
        uint64_t bl, bh, m0, m1;
        uint32_t tk;
        derive_parent(s, takes_idx, r.lo, r.hi, r.aux, bl, bh, nb, tk);
        ntk = tk & 0xff;
        mask_words(r.lo, r.hi, m0, m1);
        ik[p] = mask_hash(m0, m1) >> 32;
        iidx[p] = (uint32_t)p;
        ntk8[p] = (uint8_t)ntk;
    }
    uint32_t total, total_tk;
    block_excl_scan(ntk, warp_sums, total_tk);
    const uint32_t excl = block_excl_scan(nb, warp_sums, total);
    if (threadIdx.x < 32) {
        const uint64_t e = lookback_exclusive(status, tile, total, 0);
        if (threadIdx.x == 0) {
            s_base = e;
            if (total_tk) atomicAdd(&ctr->total_cands, (unsigned long long)total_tk);  // takes of the round
            if ((int64_t)(tile + 1) * TILE >= np) ctr->n_buys = e + total;
        }
    }
    __syncthreads();
    if (p < np) boff[p] = (uint32_t)(s_base + excl);
}

// card buys of the round's parents (src/solver.py:369-374, State.buy_card :338-355) as 32-byte records
// {child key, child aux, link = parent rank << 8 | ordinal}; also their items (sort key + item id)
struct BuySmem {
    SmemTabs tabs;
    uint64_t lo[TILE], hi[TILE], aux[TILE];
    uint32_t prefb[TILE + 1];
    uint16_t blist[BUY_WIN];
    uint32_t warp_sums[TILE / 32 + 1];
};
__global__ void __launch_bounds__(TILE) m2_buys_kernel(const Rec *__restrict__ front, int64_t np,
                                                       const DevTables *__restrict__ tabs,
                                                       const uint32_t *__restrict__ takes_idx,
                                                       const uint32_t *__restrict__ boff, int64_t rank_base,
                                                       Rec *__restrict__ brec, uint64_t *__restrict__ ik,
                                                       uint32_t *__restrict__ iidx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BuySmem &S = *reinterpret_cast<BuySmem *>(smem_raw);
    const unsigned tid = threadIdx.x;
    load_tabs(S.tabs, tabs);
    __syncthreads();
    const int64_t p0 = (int64_t)blockIdx.x * TILE, p = p0 + tid;
    const uint32_t b0 = boff[p0];
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
        for (uint32_t i = tid; i < nbw; i += TILE) {
            const uint32_t ent = S.blist[i], j = ent >> 7;
            const int pos = ent & 127;
            const uint64_t lo = S.lo[j], aux = S.aux[j];
            uint32_t saved, cd;
            const uint32_t ng = buy_gems(S.tabs, lo, aux, pos, saved, cd);
            uint64_t klo = (lo & ~GEM_MASK) | ng, khi = S.hi[j];
            if (pos < 64) klo |= 1ull << pos; else khi |= 1ull << (pos - 64);
            const uint32_t ord = w0 + i - S.prefb[j];
            const uint64_t caux = aux + saved + ((uint64_t)((cd >> 15) & 7) << 16) + (1ull << (24 + 5 * ((cd >> 18) & 7)));
            const uint64_t g = (uint64_t)b0 + w0 + i;  // == boff[parent] + ord
            Rec r{klo, khi, caux, ((uint64_t)(rank_base + p0 + j) << 8) | ord};
            st_rec(brec + g, r);
            uint64_t m0, m1;
            mask_words(klo, khi, m0, m1);
            ik[np + g] = mask_hash(m0, m1) >> 32;
            iidx[np + g] = (uint32_t)(np + g);
        }
    }
}

// ------------------------------------------------------------------ 3. runs of equal sort key
// run_start[r] = first sorted item of run r, run_wpre[r] = candidates (takes of parents + buy records)
// in the runs before r.  Two look-back chains (run count, weight) walked by two warps at once.
constexpr int RUN_ITEMS = 8;
__global__ void __launch_bounds__(TILE) m2_runs_kernel(const uint64_t *__restrict__ ik, const uint32_t *__restrict__ iidx,
                                                       int64_t n_items, uint32_t np, const uint8_t *__restrict__ ntk8,
                                                       uint32_t *__restrict__ run_start, uint32_t *__restrict__ run_wpre,
                                                       uint64_t *status_f, uint64_t *status_w, Counters *ctr, int ticket_id) {
    __shared__ uint32_t warp_sums[TILE / 32 + 1];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_fbase, s_wbase;
    if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->ticket[ticket_id], 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t b0 = ((int64_t)tile * TILE + threadIdx.x) * RUN_ITEMS;
    uint32_t flags = 0, w[RUN_ITEMS], fsum = 0, wsum = 0;
    uint64_t prev = (b0 > 0 && b0 - 1 < n_items) ? ik[b0 - 1] : 0;
#pragma unroll
    for (int q = 0; q < RUN_ITEMS; ++q) {
        const int64_t i = b0 + q;
        w[q] = 0;
        if (i < n_items) {
            const uint64_t k = ik[i];
            if (i == 0 || k != prev) { flags |= 1u << q; ++fsum; }
            prev = k;
            const uint32_t id = iidx[i];
            w[q] = id < np ? ntk8[id] : 1u;
            wsum += w[q];
        }
    }
    // tile totals: runs <= 2048 (12 bits), weight <= 2048 * 100 (18 bits): one packed scan
    uint32_t tot;
    const uint32_t ex = block_excl_scan(wsum << 12 | fsum, warp_sums, tot);
    if (threadIdx.x < 32) {
        const uint64_t e = lookback_exclusive(status_f, tile, tot & 0xfffu, 0);
        if (threadIdx.x == 0) s_fbase = e;
    } else if (threadIdx.x < 64) {
        const uint64_t e = lookback_exclusive(status_w, tile, tot >> 12, 0);
        if (threadIdx.x == 32) s_wbase = e;
    }
    __syncthreads();
    uint64_t r = s_fbase + (ex & 0xfffu), wp = s_wbase + (ex >> 12);
#pragma unroll
    for (int q = 0; q < RUN_ITEMS; ++q) {
        if (flags >> q & 1) {
            run_start[r] = (uint32_t)(b0 + q);
            run_wpre[r] = (uint32_t)wp;
            ++r;
        }
        wp += w[q];
    }
    if (b0 <= n_items - 1 && n_items - 1 < b0 + RUN_ITEMS) {  // the thread that owns the last item: sentinel
        run_start[r] = (uint32_t)n_items;
        run_wpre[r] = (uint32_t)wp;
        ctr->n_runs = (unsigned int)r;
    }
}

// ------------------------------------------------------------------ 4. per-run dedup
struct GroupArgs {
    const Rec *front;              // the round's parents (rank order)
    const Rec *brec;               // the round's buy records
    const uint64_t *ik;            // sorted keys (unused by the kernels, kept for debugging)
    const uint32_t *iidx;          // sorted item ids: < np parent, else np + buy index
    const uint32_t *run_start, *run_wpre;
    uint32_t np;
    int64_t rank_base;             // global rank of parent 0 of the round
    const DevTables *tabs;
    const uint32_t *takes_idx;
    const uint16_t *takes_edges;
    const uint16_t *gemrank;
    uint64_t *nodes;
    uint64_t nn;
    Rec *out;                      // winners (any order) ...
    uint64_t *out_sk;              // ... and their order-preserving score keys (null: no scoring)
    uint64_t out_base;
    uint32_t *big_list;
    int h, noise_mode;
    ScoreLuts L;
    Counters *ctr;
};

struct SmallSmem {
    SmemTabs tabs;
    uint64_t tbl[M2_WARPS][SM_TBL];
    uint64_t st[5][M2_WARPS][SM_STAGE];  // lo, hi, aux, link, sk
};

__device__ __forceinline__ uint32_t sm_slot(uint32_t g) { return (g * 0x9E3779B1u) >> 24; }
// entry = (gems + 1) << 48 | t : same gems share the high part, so atomicMin keeps the first arrival
__device__ __forceinline__ void sm_insert(uint64_t *tbl, uint32_t g, uint64_t t) {
    const uint64_t mine = ((uint64_t)(g + 1) << 48) | t;
    uint32_t s = sm_slot(g);
    for (;;) {
        uint64_t cur = tbl[s];
        if (cur == 0) {
            cur = atomicCAS(reinterpret_cast<unsigned long long *>(&tbl[s]), 0ull, (unsigned long long)mine);
            if (cur == 0) return;
        }
        if ((cur >> 48) == (mine >> 48)) {
            atomicMin(reinterpret_cast<unsigned long long *>(&tbl[s]), (unsigned long long)mine);
            return;
        }
        s = (s + 1) & (SM_TBL - 1);
    }
}
__device__ __forceinline__ bool sm_is_first(const uint64_t *tbl, uint32_t g, uint64_t t) {
    const uint64_t mine = ((uint64_t)(g + 1) << 48) | t;
    uint32_t s = sm_slot(g);
    for (;;) {
        const uint64_t cur = tbl[s];
        if (cur == 0) return false;
        if ((cur >> 48) == (mine >> 48)) return cur == mine;
        s = (s + 1) & (SM_TBL - 1);
    }
}

// One warp per run of at most SMALL_ITEMS items / SMALL_W candidates (the vast majority of runs); larger
// runs are queued for the CTA kernel.  Persistent grid: warp w takes runs w, w + W, ... so that at any
// moment the grid works on a window of consecutive runs == consecutive nodes.
__global__ void __launch_bounds__(TILE, 4) m2_group_small_kernel(GroupArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmallSmem &S = *reinterpret_cast<SmallSmem *>(smem_raw);
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    load_tabs(S.tabs, A.tabs);
    __syncthreads();
    uint64_t *tbl = S.tbl[w];
    const uint32_t R = A.ctr->n_runs;
    uint32_t cnt = 0, n_fresh = 0;  // staged winners (warp-uniform), nodes created (lane 0)
    uint64_t kmin = ~0ull, kmax = 0;
    auto flush = [&]() {
        uint64_t base = 0;
        if (lane == 0) base = atomicAdd(&A.ctr->n_emitted, (unsigned long long)cnt);
        base = __shfl_sync(0xffffffffu, base, 0) + A.out_base;
        for (uint32_t i = lane; i < cnt; i += 32) {
            Rec r{S.st[0][w][i], S.st[1][w][i], S.st[2][w][i], S.st[3][w][i]};
            st_rec(A.out + base + i, r);
            if (A.out_sk) A.out_sk[base + i] = S.st[4][w][i];
        }
        cnt = 0;
        __syncwarp();
    };
    for (uint32_t r = blockIdx.x * M2_WARPS + w; r < R; r += gridDim.x * M2_WARPS) {
        const uint32_t s = A.run_start[r], e = A.run_start[r + 1];
        const uint32_t wgt = A.run_wpre[r + 1] - A.run_wpre[r];
        if (e - s > SMALL_ITEMS || wgt > SMALL_W) {
            if (lane == 0) A.big_list[atomicAdd(&A.ctr->n_big, 1u)] = r;
            continue;
        }
        // ---- one item per lane
        const bool valid = s + lane < e;
        Rec it{0, 0, 0, 0};
        bool isP = false;
        uint32_t nb = 0, tk = 0;
        uint64_t grank = 0, m0 = 0, m1 = 0;
        if (valid) {
            const uint32_t id = A.iidx[s + lane];
            isP = id < A.np;
            if (isP) {
                ld_rec(A.front + id, it);
                uint64_t bl, bh;
                derive_parent(S.tabs, A.takes_idx, it.lo, it.hi, it.aux, bl, bh, nb, tk);
                grank = (uint64_t)(A.rank_base + id);
            } else {
                ld_rec(A.brec + (id - A.np), it);
            }
            mask_words(it.lo, it.hi, m0, m1);
        }
        unsigned pending = __ballot_sync(0xffffffffu, valid);
        while (pending) {  // one iteration per distinct card set of the run (almost always one)
            const int lead = __ffs(pending) - 1;
            const uint64_t M0 = __shfl_sync(0xffffffffu, m0, lead), M1 = __shfl_sync(0xffffffffu, m1, lead);
            const bool mine = valid && m0 == M0 && m1 == M1;
            const unsigned grp = __ballot_sync(0xffffffffu, mine);
            pending &= ~grp;
            uint64_t node = 0;
            int fresh_i = 0;
            if (lane == 0) {
                bool fresh;
                node = node_find_or_create(A.nodes, A.nn, M0, M1, fresh, &A.ctr->error);
                fresh_i = fresh;
                n_fresh += fresh;
            }
            node = __shfl_sync(0xffffffffu, node, 0);
            const bool fresh = __shfl_sync(0xffffffffu, fresh_i, 0);
            uint64_t *N = A.nodes + node * NODE_WORDS;
#pragma unroll
            for (int i = 0; i < SM_TBL / 32; ++i) tbl[i * 32 + lane] = 0;
            __syncwarp();
            const unsigned pm = __ballot_sync(0xffffffffu, mine && isP);
            for (int pass = 0; pass < 2; ++pass) {
                // buy records of this card set: one candidate per lane
                {
                    const bool act = mine && !isP;
                    const uint32_t g = (uint32_t)(it.lo & GEM_MASK);
                    bool win = false;
                    uint32_t rk = 0;
                    if (act) {
                        rk = __ldg(A.gemrank + g);
                        if (pass == 0) { if (fresh || !node_bit(N, rk)) sm_insert(tbl, g, it.link); }
                        else win = sm_is_first(tbl, g, it.link);
                    }
                    if (pass == 1) {
                        const unsigned wb = __ballot_sync(0xffffffffu, win);
                        if (win) {
                            const uint32_t at = cnt + __popc(wb & ((1u << lane) - 1));
                            S.st[0][w][at] = it.lo; S.st[1][w][at] = it.hi; S.st[2][w][at] = it.aux; S.st[3][w][at] = it.link;
                            if (A.out_sk) {
                                const uint64_t k = flip_f64((uint64_t)__double_as_longlong(
                                    score_state(A.h, A.noise_mode, it.lo, it.hi & HI_KEY_MASK, it.aux, A.L)));
                                S.st[4][w][at] = k;
                                kmin = min(kmin, k); kmax = max(kmax, k);
                            }
                            atomicOr(reinterpret_cast<unsigned long long *>(N + 2 + (rk >> 6)), 1ull << (rk & 63));
                        }
                        cnt += __popc(wb);
                        __syncwarp();
                        if (cnt > SM_STAGE - 32) flush();
                    }
                }
                // gem takes of this card set's parents (src/solver.py:381-388), 32 table edges at a time
                for (unsigned rest = pm; rest; rest &= rest - 1) {
                    const int P = __ffs(rest) - 1;
                    const uint64_t plo = __shfl_sync(0xffffffffu, it.lo, P), phi = __shfl_sync(0xffffffffu, it.hi, P);
                    const uint64_t paux = __shfl_sync(0xffffffffu, it.aux, P), pgr = __shfl_sync(0xffffffffu, grank, P);
                    const uint32_t ptk = __shfl_sync(0xffffffffu, tk, P), pnb = __shfl_sync(0xffffffffu, nb, P);
                    const uint32_t ntk = ptk & 0xff;
                    for (uint32_t q0 = 0; q0 < ntk; q0 += 32) {
                        const uint32_t q = q0 + lane;
                        const bool act = q < ntk;
                        uint32_t g = 0, rk = 0;
                        uint64_t t = 0;
                        bool win = false;
                        if (act) {
                            g = __ldg(A.takes_edges + (ptk >> 8) + q);
                            rk = __ldg(A.gemrank + g);
                            t = (pgr << 8) | (pnb + q);
                            if (pass == 0) { if (fresh || !node_bit(N, rk)) sm_insert(tbl, g, t); }
                            else win = sm_is_first(tbl, g, t);
                        }
                        if (pass == 1) {
                            const unsigned wb = __ballot_sync(0xffffffffu, win);
                            if (win) {
                                const uint32_t at = cnt + __popc(wb & ((1u << lane) - 1));
                                const uint64_t clo = (plo & ~GEM_MASK) | g;
                                S.st[0][w][at] = clo; S.st[1][w][at] = phi; S.st[2][w][at] = paux; S.st[3][w][at] = t;
                                if (A.out_sk) {
                                    const uint64_t k = flip_f64((uint64_t)__double_as_longlong(
                                        score_state(A.h, A.noise_mode, clo, phi & HI_KEY_MASK, paux, A.L)));
                                    S.st[4][w][at] = k;
                                    kmin = min(kmin, k); kmax = max(kmax, k);
                                }
                                atomicOr(reinterpret_cast<unsigned long long *>(N + 2 + (rk >> 6)), 1ull << (rk & 63));
                            }
                            cnt += __popc(wb);
                            __syncwarp();
                            if (cnt > SM_STAGE - 32) flush();
                        }
                    }
                }
                __syncwarp();  // pass 0 inserts (and bitmap reads) complete before pass 1 reads the table (and sets bits)
            }
        }
    }
    if (cnt) flush();
    if (lane == 0 && n_fresh) atomicAdd(&A.ctr->n_new_nodes, n_fresh);
    if (A.out_sk) {
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, d));
            kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));
        }
        if (lane == 0 && kmin <= kmax) {
            atomicMin(&A.ctr->sk_min, (unsigned long long)kmin);
            atomicMax(&A.ctr->sk_max, (unsigned long long)kmax);
        }
    }
}

// ---- CTA kernel: runs with many items / candidates.  Direct table: min arrival index per gem hand.
struct BigSmem {
    SmemTabs tabs;
    uint64_t tbl[GEM_STATES + 6];
    uint64_t bm[NODE_BM_WORDS];
    uint64_t done0[BIG_DONE], done1[BIG_DONE];
    uint32_t warp_sums[TILE / 32 + 1];
    uint32_t job, next, fresh, n_done;
    uint64_t node, base, m0, m1;
};

// enumerate the candidates of the items of run [s, e) whose card set is (M0, M1): f(gems, rank, t, lo, hi, aux)
template <class F>
__device__ __forceinline__ void big_enumerate(const GroupArgs &A, BigSmem &S, uint32_t s, uint32_t e, uint64_t M0, uint64_t M1,
                                              bool note_others, F &&f) {
    for (uint32_t i = s + threadIdx.x; i < e; i += TILE) {
        const uint32_t id = A.iidx[i];
        Rec it;
        const bool isP = id < A.np;
        if (isP) ld_rec(A.front + id, it); else ld_rec(A.brec + (id - A.np), it);
        uint64_t m0, m1;
        mask_words(it.lo, it.hi, m0, m1);
        if (m0 != M0 || m1 != M1) {
            if (note_others) {  // another card set with the same 32-bit sort key: remember the first one not done yet
                bool done = false;
                for (uint32_t d = 0; d < S.n_done; ++d) done |= (S.done0[d] == m0 && S.done1[d] == m1);
                if (!done) atomicMin(&S.next, i);
            }
            continue;
        }
        if (!isP) {
            const uint32_t g = (uint32_t)(it.lo & GEM_MASK);
            f(g, (uint32_t)__ldg(A.gemrank + g), it.link, it.lo, it.hi, it.aux);
        } else {
            uint64_t bl, bh;
            uint32_t nb, tk;
            derive_parent(S.tabs, A.takes_idx, it.lo, it.hi, it.aux, bl, bh, nb, tk);
            const uint64_t tb = ((uint64_t)(A.rank_base + id) << 8) | nb;
            const uint32_t ntk = tk & 0xff;
            for (uint32_t q = 0; q < ntk; ++q) {
                const uint32_t g = __ldg(A.takes_edges + (tk >> 8) + q);
                f(g, (uint32_t)__ldg(A.gemrank + g), tb + q, (it.lo & ~GEM_MASK) | g, it.hi, it.aux);
            }
        }
    }
}

__global__ void __launch_bounds__(TILE) m2_group_big_kernel(GroupArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BigSmem &S = *reinterpret_cast<BigSmem *>(smem_raw);
    const unsigned tid = threadIdx.x, lane = tid & 31;
    load_tabs(S.tabs, A.tabs);
    const uint32_t n_big = A.ctr->n_big;
    uint32_t n_fresh = 0;
    uint64_t kmin = ~0ull, kmax = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) S.job = atomicAdd(&A.ctr->ticket[3], 1u);
        __syncthreads();
        if (S.job >= n_big) break;
        const uint32_t r = A.big_list[S.job];
        const uint32_t s = A.run_start[r], e = A.run_start[r + 1];
        if (tid == 0) { S.n_done = 0; S.next = s; }
        __syncthreads();
        while (S.next != 0xFFFFFFFFu) {  // one iteration per distinct card set of the run
            const uint32_t lead = S.next;
            __syncthreads();
            if (tid == 0) {
                const uint32_t id = A.iidx[lead];
                const Rec *src = id < A.np ? A.front + id : A.brec + (id - A.np);
                mask_words(src->lo, src->hi, S.m0, S.m1);
                bool fresh;
                S.node = node_find_or_create(A.nodes, A.nn, S.m0, S.m1, fresh, &A.ctr->error);
                S.fresh = fresh;
                n_fresh += fresh;
                S.next = 0xFFFFFFFFu;
            }
            __syncthreads();
            const uint64_t M0 = S.m0, M1 = S.m1;
            uint64_t *N = A.nodes + S.node * NODE_WORDS;
            for (int i = tid; i < GEM_STATES; i += TILE) S.tbl[i] = ~0ull;
            if (tid < NODE_BM_WORDS) S.bm[tid] = S.fresh ? 0ull : N[2 + tid];
            __syncthreads();
            // pass 0: first arrival per gem hand among the candidates not in the visited set
            big_enumerate(A, S, s, e, M0, M1, true, [&](uint32_t, uint32_t rk, uint64_t t, uint64_t, uint64_t, uint64_t) {
                if (!((S.bm[rk >> 6] >> (rk & 63)) & 1)) atomicMin(reinterpret_cast<unsigned long long *>(&S.tbl[rk]), (unsigned long long)t);
            });
            __syncthreads();
            // pass 1a: winners per thread -> output offsets
            uint32_t mywins = 0;
            big_enumerate(A, S, s, e, M0, M1, false, [&](uint32_t, uint32_t rk, uint64_t t, uint64_t, uint64_t, uint64_t) {
                mywins += S.tbl[rk] == t;
            });
            uint32_t tot;
            const uint32_t ex = block_excl_scan(mywins, S.warp_sums, tot);
            if (tid == 0) S.base = A.out_base + atomicAdd(&A.ctr->n_emitted, (unsigned long long)tot);
            __syncthreads();
            // pass 1b: emit
            uint64_t pos = S.base + ex;
            big_enumerate(A, S, s, e, M0, M1, false, [&](uint32_t, uint32_t rk, uint64_t t, uint64_t lo, uint64_t hi, uint64_t aux) {
                if (S.tbl[rk] != t) return;
                Rec o{lo, hi, aux, t};
                st_rec(A.out + pos, o);
                if (A.out_sk) {
                    const uint64_t k = flip_f64((uint64_t)__double_as_longlong(score_state(A.h, A.noise_mode, lo, hi & HI_KEY_MASK, aux, A.L)));
                    A.out_sk[pos] = k;
                    kmin = min(kmin, k); kmax = max(kmax, k);
                }
                ++pos;
                atomicOr(reinterpret_cast<unsigned long long *>(&S.bm[rk >> 6]), 1ull << (rk & 63));
            });
            __syncthreads();
            if (tid < NODE_BM_WORDS) N[2 + tid] = S.bm[tid];
            if (tid == 0) {
                if (S.n_done < BIG_DONE) { S.done0[S.n_done] = M0; S.done1[S.n_done] = M1; ++S.n_done; }
                else if (S.next != 0xFFFFFFFFu) { atomicExch(&A.ctr->error, 3u); S.next = 0xFFFFFFFFu; }
            }
            __syncthreads();
        }
    }
    if (tid == 0 && n_fresh) atomicAdd(&A.ctr->n_new_nodes, n_fresh);
    if (A.out_sk) {
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, d));
            kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));
        }
        if (lane == 0 && kmin <= kmax) {
            atomicMin(&A.ctr->sk_min, (unsigned long long)kmin);
            atomicMax(&A.ctr->sk_max, (unsigned long long)kmax);
        }
    }
}

}  // namespace spl
