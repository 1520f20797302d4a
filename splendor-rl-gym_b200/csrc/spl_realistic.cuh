// spl_realistic.cuh -- realistic multi-player mode (reference: src/solver.py:25-200, 471-860).
//
// State record RRec (96 B = three 32-byte sectors):
//   p[4]    per player: 90-bit card mask (mlo: cards 0..63, mhi: 64..89), gems (3 bits/colour), saved
//   vis[12] visible card per market slot, tier-major, in slot order (255 = empty, empties trail)
//   cur     current player;   link = parent_rank << 8 | ordinal
// Everything else the reference keeps in MultiPlayerState is derived on the fly:
//   bonus[c] / pts  : popcounts of the card mask against colour / point masks
//   gem pool        : gems_per_color - sum of the players' gems (conservation, :600-604, :685)
//   deck position   : 4 + #cards of the tier owned by anyone            (CardMarket.buy_card :121-170)
//   is_game_over()  : players[cur].pts >= target  (final round ends when play returns to its trigger, :542-549)
// Identity (src/solver.py:495-500) = (players incl. saved, pool, visible cards in slot order, current
// player): the visited table stores a 105-bit fingerprint of exactly those bytes (the reference itself
// dedups on a 64-bit hash of them).
#pragma once
#include "spl_kernels.cuh"

namespace spl {

struct RPlayer {
    uint64_t mlo;
    uint32_t mhi;
    uint16_t gems, saved;
};
struct __align__(32) RRec {
    RPlayer p[4];
    uint8_t vis[12];
    uint8_t cur, pad[3];
    uint64_t link, spare;
};
static_assert(sizeof(RRec) == 96, "RRec layout");

struct RConfigDev {
    int32_t P, target, gpc, noise;
    int32_t deck_len[3];
    uint8_t deck[3][40];
    uint64_t col_lo[NCOL], pt_lo[6], tier_lo[3];
    uint32_t col_hi[NCOL], pt_hi[6], tier_hi[3];
    uint32_t card[SPL_NUM_CARDS];
};

__device__ __forceinline__ void ld_rrec(const RRec *p, RRec &r) {
    const ulonglong4 *s = reinterpret_cast<const ulonglong4 *>(p);
    ulonglong4 *d = reinterpret_cast<ulonglong4 *>(&r);
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
}
__device__ __forceinline__ void st_rrec(RRec *p, const RRec &r) {
    const ulonglong4 *s = reinterpret_cast<const ulonglong4 *>(&r);
    ulonglong4 *d = reinterpret_cast<ulonglong4 *>(p);
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
}

__device__ __forceinline__ void r_derive(const RConfigDev &C, const RPlayer &p, int bonus[NCOL], int &pts) {
#pragma unroll
    for (int c = 0; c < NCOL; ++c) bonus[c] = __popcll(p.mlo & C.col_lo[c]) + __popc(p.mhi & C.col_hi[c]);
    pts = 0;
#pragma unroll
    for (int v = 1; v <= 5; ++v) pts += v * (__popcll(p.mlo & C.pt_lo[v]) + __popc(p.mhi & C.pt_hi[v]));
}
__device__ __forceinline__ int r_pts(const RConfigDev &C, const RPlayer &p) {
    int pts = 0;
#pragma unroll
    for (int v = 1; v <= 5; ++v) pts += v * (__popcll(p.mlo & C.pt_lo[v]) + __popc(p.mhi & C.pt_hi[v]));
    return pts;
}
// PlayerState.can_afford (src/solver.py:192-200): gems + bonus >= cost per colour, no clamp
__device__ __forceinline__ bool r_can_afford(const RConfigDev &C, uint32_t gems, const int bonus[NCOL], int card) {
    const uint32_t cd = C.card[card];
    bool ok = true;
#pragma unroll
    for (int c = 0; c < NCOL; ++c) ok &= (int)((gems >> (3 * c)) & 7) + bonus[c] >= (int)((cd >> (3 * c)) & 7);
    return ok;
}

// 27 successor slots in the reference's order (:573-748): 12 buys in market slot order, the 10
// colour triples in combinations() order, the 5 double takes.  Returns the valid-slot mask.
__device__ __forceinline__ uint32_t r_valid_mask(const RConfigDev &C, const RRec &s, const int bonus[NCOL]) {
    const RPlayer &me = s.p[s.cur];
    int pool[NCOL], hand = 0;
#pragma unroll
    for (int c = 0; c < NCOL; ++c) {
        int tot = 0;
        for (int q = 0; q < C.P; ++q) tot += (s.p[q].gems >> (3 * c)) & 7;
        pool[c] = C.gpc - tot;
        hand += (me.gems >> (3 * c)) & 7;
    }
    uint32_t m = 0;
    for (int k = 0; k < 12; ++k)
        if (s.vis[k] != 255 && r_can_afford(C, me.gems, bonus, s.vis[k])) m |= 1u << k;
    if (hand + 3 <= 10) {
        int k = 12;
        for (int a = 0; a < NCOL; ++a)
            for (int b = a + 1; b < NCOL; ++b)
                for (int c = b + 1; c < NCOL; ++c, ++k)
                    if (pool[a] > 0 && pool[b] > 0 && pool[c] > 0) m |= 1u << k;
    }
    if (hand + 2 <= 10)
        for (int c = 0; c < NCOL; ++c)
            if (pool[c] >= 4) m |= 1u << (22 + c);
    return m;
}

__device__ __forceinline__ void r_make_child(const RConfigDev &C, const RRec &s, const int bonus[NCOL], int slot, RRec &o) {
    o = s;
    const int cur = s.cur;
    RPlayer &me = o.p[cur];
    o.cur = (uint8_t)((cur + 1) % C.P);
    if (slot < 12) {  // buy (:574-632)
        const int card = s.vis[slot];
        const uint32_t cd = C.card[card];
        uint32_t ng = 0, saved = 0;
#pragma unroll
        for (int c = 0; c < NCOL; ++c) {
            const int cost = (cd >> (3 * c)) & 7;
            const int pay = max(cost - bonus[c], 0);
            saved += cost - pay;
            ng |= (uint32_t)max((int)((me.gems >> (3 * c)) & 7) - pay, 0) << (3 * c);
        }
        me.gems = (uint16_t)ng;
        me.saved = (uint16_t)(me.saved + saved);
        if (card < 64) me.mlo |= 1ull << card; else me.mhi |= 1u << (card - 64);
        const int t = slot >> 2;
        int bought = 0;
        for (int q = 0; q < C.P; ++q) bought += __popcll(s.p[q].mlo & C.tier_lo[t]) + __popc(s.p[q].mhi & C.tier_hi[t]);
        const int pos = 4 + bought;
        for (int k = slot & 3; k < 3; ++k) o.vis[t * 4 + k] = s.vis[t * 4 + k + 1];
        o.vis[t * 4 + 3] = 255;
        if (pos < C.deck_len[t]) {
            int e = 0;
            while (e < 4 && o.vis[t * 4 + e] != 255) ++e;
            o.vis[t * 4 + e] = C.deck[t][pos];
        }
    } else if (slot < 22) {  // three different colours (:671-707)
        int k = 12;
        for (int a = 0; a < NCOL; ++a)
            for (int b = a + 1; b < NCOL; ++b)
                for (int c = b + 1; c < NCOL; ++c, ++k)
                    if (k == slot) me.gems = (uint16_t)(me.gems + (1 << (3 * a)) + (1 << (3 * b)) + (1 << (3 * c)));
    } else {  // two of one colour (:710-748)
        me.gems = (uint16_t)(me.gems + (2 << (3 * (slot - 22))));
    }
}

// 105-bit fingerprint of the identity bytes (players[0..P), vis, cur)
__device__ __forceinline__ void r_fingerprint(const RConfigDev &C, const RRec &s, uint64_t &lo, uint64_t &hi) {
    uint64_t a = 0x243F6A8885A308D3ull, b = 0x13198A2E03707344ull;
    for (int q = 0; q < C.P; ++q) {
        const uint64_t w0 = s.p[q].mlo, w1 = (uint64_t)s.p[q].mhi | (uint64_t)s.p[q].gems << 32 | (uint64_t)s.p[q].saved << 48;
        a = mix64(a ^ w0, w1);
        b = mix64(b + w1 * 0x9E3779B97F4A7C15ull, w0 ^ 0xA4093822299F31D0ull);
    }
    uint64_t v0 = 0, v1 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v0 |= (uint64_t)s.vis[k] << (8 * k);
#pragma unroll
    for (int k = 8; k < 12; ++k) v1 |= (uint64_t)s.vis[k] << (8 * (k - 8));
    v1 |= (uint64_t)s.cur << 32;
    a = mix64(a ^ v0, v1);
    b = mix64(b ^ v1, v0 * 0xC2B2AE3D27D4EB4Full);
    lo = a;
    hi = b & HI_KEY_MASK;
}

// multi_competitive_heuristic (src/solver.py:778-812), bit-exact: LUT pows, separately rounded ops
__device__ __forceinline__ double r_score(const RConfigDev &C, const RRec &s, const ScoreLuts &L, int r_ext = 50) {
    const int prev = (s.cur + C.P - 1) % C.P;
    const RPlayer &me = s.p[prev], &opp = s.p[s.cur];
    int bm[NCOL], bo[NCOL], pm, po, rm = 0, ro = 0, nnz = 0, am = 0, ao = 0;
    r_derive(C, me, bm, pm);
    r_derive(C, opp, bo, po);
#pragma unroll
    for (int c = 0; c < NCOL; ++c) {
        rm += ((me.gems >> (3 * c)) & 7) + 2 * bm[c];
        ro += ((opp.gems >> (3 * c)) & 7) + 2 * bo[c];
        nnz += bm[c] > 0;
    }
    for (int k = 0; k < 12; ++k)
        if (s.vis[k] != 255) {
            am += r_can_afford(C, me.gems, bm, s.vis[k]);
            ao += r_can_afford(C, opp.gems, bo, s.vis[k]);
        }
    // point_diff = d**2.5 if me.pts > opp.pts else -(d**2.5)   (-0.0 when equal)
    const double pw = __ldg(L.pts + 0 * 256 + abs(pm - po));
    const double pd = pm > po ? pw : -pw;
    double acc = __dmul_rn(pd, 100.0);
    // resource_diff * 20: float when positive, the int 0 otherwise (x + 0 == x + 0.0)
    acc = __dadd_rn(acc, rm > ro ? __dmul_rn(__ldg(L.small + 3 * 512 + (rm - ro)), 20.0) : 0.0);
    acc = __dadd_rn(acc, (double)((am - ao) * 5));
    acc = __dadd_rn(acc, __dmul_rn(__ldg(L.small + 3 * 512 + nnz), 3.0));
    int r = C.noise == 2 ? r_ext : 50;
    if (C.noise == 1) {
        uint64_t h = 0;
        for (int q = 0; q < C.P; ++q)
            h = mix64(h ^ s.p[q].mlo, (uint64_t)s.p[q].mhi | (uint64_t)s.p[q].gems << 32 | (uint64_t)s.p[q].saved << 48);
        r = 1 + (int)(h % 100ull);
    }
    return __dadd_rn(acc, __dmul_rn((double)r, 0.01));
}

// ------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(TILE) r_goal_kernel(const RRec *__restrict__ front, int64_t n,
                                                      const RConfigDev *__restrict__ cfg, Counters *ctr) {
    const int64_t i = (int64_t)blockIdx.x * TILE + threadIdx.x;
    long long r = 0x7fffffffffffffffll;
    if (i < n) {
        const RPlayer p = front[i].p[front[i].cur];
        if (r_pts(*cfg, p) >= cfg->target) r = i;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) r = min(r, __shfl_xor_sync(0xffffffffu, r, d));
    if ((threadIdx.x & 31) == 0 && r != 0x7fffffffffffffffll) atomicMin(&ctr->goal_rank, r);
}

__global__ void __launch_bounds__(TILE) r_count_scan_kernel(const RRec *__restrict__ front, int64_t n,
                                                            const RConfigDev *__restrict__ cfg, uint32_t *__restrict__ off,
                                                            uint32_t *__restrict__ vmask, uint64_t *status, Counters *ctr,
                                                            int ticket_id) {
    __shared__ uint32_t warp_sums[TILE / 32 + 1];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_base;
    if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->ticket[ticket_id], 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t p = (int64_t)tile * TILE + threadIdx.x;
    uint32_t cnt = 0, m = 0;
    if (p < n) {
        RRec s;
        ld_rrec(front + p, s);
        int bonus[NCOL], pts;
        r_derive(*cfg, s.p[s.cur], bonus, pts);
        m = r_valid_mask(*cfg, s, bonus);
        cnt = __popc(m);
        vmask[p] = m;
    }
    uint32_t total;
    const uint32_t excl = block_excl_scan(cnt, warp_sums, total);
    if (threadIdx.x < 32) {
        const uint64_t e = lookback_exclusive(status, tile, total, 0);
        if (threadIdx.x == 0) {
            s_base = e;
            if ((int64_t)(tile + 1) * TILE >= n) ctr->total_cands = e + total;
        }
    }
    __syncthreads();
    if (p < n) off[p] = (uint32_t)(s_base + excl);
}

// one thread per parent: write its successors (record + fingerprint) at off[parent] + ordinal
__global__ void __launch_bounds__(TILE) r_expand_kernel(const RRec *__restrict__ front, int64_t n,
                                                        const RConfigDev *__restrict__ cfg,
                                                        const uint32_t *__restrict__ off,
                                                        const uint32_t *__restrict__ vmask, int64_t rank_base,
                                                        RRec *__restrict__ cand, spl_key *__restrict__ cand_key) {
    const int64_t p = (int64_t)blockIdx.x * TILE + threadIdx.x;
    if (p >= n) return;
    RRec s;
    ld_rrec(front + p, s);
    int bonus[NCOL], pts;
    r_derive(*cfg, s.p[s.cur], bonus, pts);
    uint32_t m = vmask[p], ord = 0;
    uint64_t t = off[p];
    while (m) {
        const int slot = __ffs(m) - 1;
        m &= m - 1;
        RRec o;
        r_make_child(*cfg, s, bonus, slot, o);
        o.link = ((uint64_t)(rank_base + p) << 8) | ord;
        o.spare = 0;
        st_rrec(cand + t, o);
        if (cand_key) {
            uint64_t lo, hi;
            r_fingerprint(*cfg, o, lo, hi);
            cand_key[t].lo = lo;
            cand_key[t].hi = hi;
        }
        ++t;
        ++ord;
    }
}

template <typename IDX>
__global__ void __launch_bounds__(TILE) r_gather_kernel(const RRec *__restrict__ src, const IDX *__restrict__ idx, int64_t n,
                                                        RRec *__restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * TILE + threadIdx.x;
    if (i < n) {
        RRec r;
        ld_rrec(src + idx[i], r);
        st_rrec(dst + i, r);
    }
}

__global__ void __launch_bounds__(TILE) r_score_kernel(const RRec *__restrict__ recs, int64_t n,
                                                       const RConfigDev *__restrict__ cfg, ScoreLuts L,
                                                       uint64_t *__restrict__ sk, double *__restrict__ raw, Counters *ctr,
                                                       const uint8_t *__restrict__ draws = nullptr) {
    const int64_t i = (int64_t)blockIdx.x * TILE + threadIdx.x;
    uint64_t kmin = ~0ull, kmax = 0;
    if (i < n) {
        RRec s;
        ld_rrec(recs + i, s);
        const double sc = r_score(*cfg, s, L, draws ? draws[i] : 50);
        if (raw) raw[i] = sc;
        if (sk) {
            const uint64_t k = flip_f64((uint64_t)__double_as_longlong(sc));
            sk[i] = k;
            kmin = kmax = k;
        }
    }
    if (sk) {
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, d));
            kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));
        }
        if ((threadIdx.x & 31) == 0 && kmin <= kmax) {
            atomicMin(&ctr->sk_min, (unsigned long long)kmin);
            atomicMax(&ctr->sk_max, (unsigned long long)kmax);
        }
    }
}

// max over players of pts, per state (the reference's progress line, src/solver.py:832-836)
__global__ void __launch_bounds__(TILE) r_maxpts_kernel(const RRec *__restrict__ recs, int64_t n,
                                                        const RConfigDev *__restrict__ cfg, uint8_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * TILE + threadIdx.x;
    if (i < n) {
        int m = 0;
        for (int q = 0; q < cfg->P; ++q) m = max(m, r_pts(*cfg, recs[i].p[q]));
        out[i] = (uint8_t)min(m, 255);
    }
}

__global__ void r_root_key_kernel(const RRec *root, const RConfigDev *cfg, spl_key *key) {
    RRec s;
    ld_rrec(root, s);
    uint64_t lo, hi;
    r_fingerprint(*cfg, s, lo, hi);
    key->lo = lo;
    key->hi = hi;
}

__global__ void __launch_bounds__(TILE) r_links_kernel(const RRec *__restrict__ recs, int64_t n, uint64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * TILE + threadIdx.x;
    if (i < n) out[i] = recs[i].link;
}

}  // namespace spl
