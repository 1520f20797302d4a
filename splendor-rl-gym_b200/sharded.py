"""Multi-GPU level-synchronous solver: the frontier is sharded across ranks (one process per GPU),
candidates are routed to the rank that owns their key (hash of the canonical key), and every
level is bit-identical to the single-GPU solver -- same states, same first-arrival `saved` and
parent links, same rank order after the beam cut.

Per level (G ranks; the queue is block-distributed by global rank: rank g holds a contiguous slice):
  1. goal test: local first rank with pts >= goal, all-reduce MIN                 (src/solver.py:443)
  2. expand local parents (spl_expand); global arrival index t = exscan(counts) + local index
  3. stable partition by owner (spl_owner_partition) and all-to-all of the 16-byte KEYS only;
     the receive buffer, concatenated by source rank, is already in global arrival order
  4. owner: first-arrival dedup against its slice of the visited set (spl_dedup)   (:447-450)
  5. one winner byte per candidate travels back (reverse all-to-all); the SOURCE rank, which still
     holds key/aux/link of its candidates in arrival order, compacts its winners -> the next
     queue is again block-distributed in global arrival order (pure BFS stops here)
  6. beam: scores (spl_score); global radix select = the single-GPU passes with the 2048-bin
     histogram all-reduced between spl_dtopk_hist and spl_dtopk_pick; arrival-order tie quota
     split over ranks by an exclusive scan of local tie counts (stable) or key threshold (det)
  7. local cut + local rank sort (spl_dtopk_cut); global rank of every survivor = local index +
     counts against the other ranks' sorted sort-words (all-gather + spl_count_less); survivors
     are sent to the rank that owns their global-rank block (all-to-all)          (:452-456)

The compute primitives come from a `backend` object (CudaBackend below wraps the C ABI); the
collectives go through torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ._lib import check, lib
from .engine import NOISE_IDS, TIE_IDS, Engine, _DevArray, heuristic_id

import os
import time

SEL_BITS = 11
I64_MAX = (1 << 63) - 1
TIMING = bool(os.environ.get('SPL_TIMING'))
PHASES = {}


def _tick(name, t0):
    """phase timer (SPL_TIMING=1): synchronises the device, so only for diagnosis"""
    if not TIMING:
        return 0.0
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    PHASES[name] = PHASES.get(name, 0.0) + (t1 - t0)
    return t1


def _bitlen(x: int) -> int:
    return int(x).bit_length()


class Comm:
    """Thin wrapper over torch.distributed (world size 1 works without a process group)."""

    def __init__(self, device):
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.rank = dist.get_rank() if self.on else 0
        self.world = dist.get_world_size() if self.on else 1
        self.device = device

    def gather_ints(self, *vals):
        """all_gather of a few python ints -> int64 array [world, len(vals)]"""
        t = torch.tensor(vals, dtype=torch.int64, device=self.device).reshape(1, -1)
        if not self.on:
            return t.cpu().numpy()
        out = torch.empty((self.world, t.shape[1]), dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(out, t) if self.device.type == 'cuda' else dist.all_gather(list(out.unbind(0)), t[0])
        return out.cpu().numpy()

    def all_reduce(self, t, op):
        if self.on:
            dist.all_reduce(t, op=op)
        return t

    def all_to_all_rows(self, send, send_counts, recv_counts):
        """variable all-to-all of rows (dim 0); send is grouped by destination rank"""
        if not self.on:
            return send
        out = torch.empty((int(sum(recv_counts)),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(out, send.contiguous(), [int(x) for x in recv_counts], [int(x) for x in send_counts])
        return out

    def all_gather_v(self, t, counts):
        """all_gather of 1-D tensors of different lengths -> list of tensors"""
        if not self.on:
            return [t]
        m = int(max(counts)) if len(counts) else 0
        pad = torch.zeros(m, dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        outs = [torch.empty(m, dtype=t.dtype, device=t.device) for _ in range(self.world)]
        dist.all_gather(outs, pad)
        return [o[:int(c)] for o, c in zip(outs, counts)]


class CudaBackend:
    """Compute primitives of one rank, all through the C ABI of libsplendor_b200.so."""

    def __init__(self, eng: Engine):
        self.eng = eng
        self.device = eng.tdev

    def reset_visited(self):
        self.eng.reset_visited()

    def visited_count(self):
        return self.eng.visited_count()

    def first_goal(self, front, goal):
        pts = (front[:, 2] >> 16) & 0xff
        hit = torch.nonzero(pts >= goal)
        return int(hit[0]) if hit.numel() else -1

    def expand_rows(self, front, rank_base):
        """successors of the local queue slice as rows [m, 4]; link = (rank_base + parent) << 8 | ordinal"""
        n = front.shape[0]
        cap = max(64, n * 36)
        while True:
            out = torch.empty((cap, 4), dtype=torch.int64, device=self.device)
            m = C.c_int64()
            rc = lib.spl_expand_rows(self.eng._h, front.data_ptr(), n, rank_base, out.data_ptr(), cap, C.byref(m), self.eng._stream())
            if rc == -6:
                cap = m.value
                continue
            check(rc, self.eng._h)
            return out[:m.value]

    def route_keys(self, cand, world):
        """keys of the candidates grouped by owner rank (arrival order inside a group) + per-owner counts"""
        m = cand.shape[0]
        send = torch.empty((max(m, 1), 2), dtype=torch.int64, device=self.device)
        counts = (C.c_int64 * world)()
        check(lib.spl_route_keys(self.eng._h, cand.data_ptr(), m, world, send.data_ptr(), counts, self.eng._stream()), self.eng._h)
        return send[:m], np.array(counts[:], dtype=np.int64)

    def dedup_flags(self, keys):
        """owner side: one winner byte per received key (first arrival of a never-seen key)"""
        n = keys.shape[0]
        flags = torch.empty(max(n, 1), dtype=torch.uint8, device=self.device)
        check(lib.spl_dedup_flags(self.eng._h, keys.data_ptr(), n, flags.data_ptr(), self.eng._stream()), self.eng._h)
        return flags[:n]

    def compact_winners(self, cand, flags_send_order):
        """source side: the winning rows in arrival order"""
        m = cand.shape[0]
        out = torch.empty((max(m, 1), 4), dtype=torch.int64, device=self.device)
        k = C.c_int64()
        check(lib.spl_compact_winners(self.eng._h, cand.data_ptr(), m, flags_send_order.data_ptr(), out.data_ptr(), C.byref(k),
                                      self.eng._stream()), self.eng._h)
        return out[:k.value]

    def owner_partition(self, keys, world):
        n = keys.shape[0]
        perm = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        counts = (C.c_int64 * world)()
        check(lib.spl_owner_partition(self.eng._h, keys.data_ptr(), n, world, perm.data_ptr(), counts, self.eng._stream()),
              self.eng._h)
        return perm[:n], np.array(counts[:], dtype=np.int64)

    def dedup(self, keys):
        aux = torch.zeros(keys.shape[0], dtype=torch.int64, device=self.device)
        _, _, src = self.eng.dedup(keys.contiguous(), aux)
        return src

    def score(self, heuristic, noise, rows):
        n = rows.shape[0]
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        check(lib.spl_score_rows(self.eng._h, heuristic_id(heuristic), NOISE_IDS[noise], rows.data_ptr(), n, out.data_ptr(),
                                 self.eng._stream()), self.eng._h)
        return out

    def move_rows(self, rows, idx, n_out, scatter):
        """gather (out[i] = rows[idx[i]]) or scatter (out[idx[i]] = rows[i]) of 32-byte rows"""
        out = torch.empty((max(n_out, 1), 4), dtype=torch.int64, device=self.device)
        check(lib.spl_move_rows(self.eng._h, rows.data_ptr(), idx.data_ptr(), idx.shape[0], out.data_ptr(), int(scatter),
                                self.eng._stream()), self.eng._h)
        return out[:n_out]

    # ---- distributed top-k passes
    def dtopk_begin(self, scores, keys):
        a, b = C.c_uint64(), C.c_uint64()
        check(lib.spl_dtopk_begin(self.eng._h, scores.data_ptr(), keys.data_ptr() if keys is not None else None,
                                  scores.shape[0], C.byref(a), C.byref(b), self.eng._stream()), self.eng._h)
        return a.value, b.value

    def dtopk_hist(self, word, shift, bits, first, smin):
        p = C.c_void_p()
        check(lib.spl_dtopk_hist(self.eng._h, word, shift, bits, int(first), smin, C.byref(p), self.eng._stream()), self.eng._h)
        if getattr(self, '_hist_ptr', None) != p.value:  # the histogram lives at a fixed address inside the context
            self._hist_ptr = p.value
            self._hist = torch.as_tensor(_DevArray(p.value, (1 << SEL_BITS,), '<i4'), device=self.device)
        return self._hist

    def dtopk_pick(self, word, shift, first, init_k, k):
        check(lib.spl_dtopk_pick(self.eng._h, word, shift, int(first), int(init_k), k, self.eng._stream()), self.eng._h)

    def dtopk_get(self):
        st = (C.c_uint64 * 6)()
        check(lib.spl_dtopk_get(self.eng._h, st, self.eng._stream()), self.eng._h)
        return list(st)

    def dtopk_set(self, state):
        check(lib.spl_dtopk_set(self.eng._h, (C.c_uint64 * 6)(*state), self.eng._stream()), self.eng._h)

    def dtopk_cut(self, tie, keep_all, all_ties, smin, smax, n):
        det = tie == 'det'
        idx = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        y = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        kl = torch.empty(max(n, 1) if det else 1, dtype=torch.int64, device=self.device)
        kh = torch.empty(max(n, 1) if det else 1, dtype=torch.int64, device=self.device)
        kept = C.c_int64()
        check(lib.spl_dtopk_cut(self.eng._h, TIE_IDS[tie], int(keep_all), int(all_ties), smin, smax, idx.data_ptr(),
                                y.data_ptr(), kl.data_ptr() if det else None, kh.data_ptr() if det else None,
                                C.byref(kept), self.eng._stream()), self.eng._h)
        m = kept.value
        return idx[:m], y[:m], (kl[:m] if det else None), (kh[:m] if det else None)

    def count_less(self, words, inclusive, a, b, out, accumulate):
        (ay, akl, akh), (by, bkl, bkh) = a, b
        check(lib.spl_count_less(self.eng._h, words, int(inclusive), ay.data_ptr(),
                                 akl.data_ptr() if words == 3 else None, akh.data_ptr() if words == 3 else None,
                                 ay.shape[0], by.data_ptr(), bkl.data_ptr() if words == 3 else None,
                                 bkh.data_ptr() if words == 3 else None, by.shape[0], out.data_ptr(), int(accumulate),
                                 self.eng._stream()), self.eng._h)


class ShardedSolver:
    """State.solve (src/solver.py:390-464) over a frontier sharded across `comm.world` ranks."""

    def __init__(self, backend, comm: Comm, root_key: int, root_aux: int, goal_pts: int, use_heuristic: bool,
                 heuristic: str, beam_width: int, tie: str = 'stable', noise: str = 'const'):
        self.b, self.comm = backend, comm
        self.goal, self.use_h, self.h, self.beam, self.tie, self.noise = goal_pts, use_heuristic, heuristic, beam_width, tie, noise
        dev = backend.device
        m64 = (1 << 64) - 1

        def s64(x):
            x &= m64
            return x - (1 << 64) if x >> 63 else x
        root = torch.tensor([[s64(root_key), s64(root_key >> 64), s64(root_aux), -1]], dtype=torch.int64, device=dev)
        backend.reset_visited()  # trail = {}
        # the root is held by rank 0; its key is registered in the visited set of its owner
        self.front = root if comm.rank == 0 else root[:0]
        send, counts = backend.route_keys(root, comm.world)
        if counts[comm.rank]:
            backend.dedup_flags(send)
        self.level = 0
        self.ended = False
        self.goal_rank = -1
        self.infos = []
        self.links = []  # per level: (base, local link column)
        self._save_links()

    def _sizes(self, n):
        sizes = self.comm.gather_ints(n)[:, 0]
        return sizes, int(sizes[:self.comm.rank].sum()), int(sizes.sum())

    def _save_links(self):
        _, base, _ = self._sizes(self.front.shape[0])
        self.links.append((base, self.front[:, 3].clone()))

    # ------------------------------------------------------------------ one `while queue` iteration
    def step(self) -> dict:
        b, comm, dev = self.b, self.comm, self.b.device
        G, me = comm.world, comm.rank
        front = self.front
        n_local = front.shape[0]
        _, base, n_total = self._sizes(n_local)
        info = dict(level=self.level, ended=0, frontier=n_total, expanded=0, generated=0, unique=0, kept=0, goal_rank=-1)
        # 1. goal test
        g = b.first_goal(front, self.goal) if n_local else -1
        gt = torch.tensor([base + g if g >= 0 else I64_MAX], dtype=torch.int64, device=dev)
        comm.all_reduce(gt, dist.ReduceOp.MIN)
        if int(gt) != I64_MAX:
            self.ended, self.goal_rank = True, int(gt)
            info.update(ended=1, goal_rank=self.goal_rank)
            self.infos.append(info)
            return info
        t0 = _tick('goal', time.perf_counter() if TIMING else 0.0)
        # 2. expand
        cand = b.expand_rows(front, base) if n_local else torch.empty((0, 4), dtype=torch.int64, device=dev)
        m = cand.shape[0]
        t0 = _tick('expand', t0)
        # 3. route keys to their owners
        send_keys, counts = b.route_keys(cand, G)
        all_counts = comm.gather_ints(*counts.tolist())  # [src, dst]
        recv_counts = all_counts[:, me]
        t0 = _tick('partition', t0)
        recv_keys = comm.all_to_all_rows(send_keys, counts, recv_counts)
        t0 = _tick('a2a_keys', t0)
        # 4. first-arrival dedup at the owner
        flags_recv = b.dedup_flags(recv_keys)
        t0 = _tick('dedup', t0)
        # 5. winner bytes back to the source; compaction in arrival order
        flags_back = comm.all_to_all_rows(flags_recv, recv_counts, counts)
        winners = b.compact_winners(cand, flags_back)
        t0 = _tick('flags_compact', t0)
        u_local = winners.shape[0]
        tot = comm.gather_ints(m, u_local)
        info.update(expanded=n_total, generated=int(tot[:, 0].sum()), unique=int(tot[:, 1].sum()))
        u_total = info['unique']
        if self.use_h and u_total:
            winners = self._beam_cut(winners, tot[:, 1], u_total)
        t0 = _tick('beam_cut', t0)
        self.front = winners
        kv = self.comm.gather_ints(winners.shape[0], b.visited_count())
        kept_total = int(kv[:, 0].sum())
        info['kept'] = kept_total
        info['visited'] = int(kv[:, 1].sum())
        self.infos.append(info)
        if kept_total == 0:  # frontier exhausted: `puzzle` stays the last dequeued state
            self.ended, self.goal_rank = True, n_total - 1
            info['ended'] = 1
            return info
        self.level += 1
        self._save_links()
        return info

    # ------------------------------------------------------------------ global beam cut + rank order
    def _beam_cut(self, winners, u_all, u_total):
        b, comm, dev = self.b, self.comm, self.b.device
        G, me = comm.world, comm.rank
        K = self.beam
        det = self.tie == 'det'
        u_local = winners.shape[0]
        scores = b.score(self.h, self.noise, winners) if u_local else torch.empty(0, dtype=torch.float64, device=dev)
        keys = winners[:, :2].contiguous() if det else None
        smin, smax = b.dtopk_begin(scores, keys)
        # global score-key range (unsigned 64-bit min / max, exchanged as 32-bit halves)
        allr = comm.gather_ints(smin >> 32, smin & 0xffffffff, smax >> 32, smax & 0xffffffff)
        mins = [(int(r[0]) << 32) | int(r[1]) for r, u in zip(allr, u_all) if u]
        maxs = [(int(r[2]) << 32) | int(r[3]) for r, u in zip(allr, u_all) if u]
        smin, smax = min(mins), max(maxs)
        keep_all = u_total <= K
        all_ties = True
        if not keep_all:
            nbits = _bitlen(smax - smin)
            state = [0, K, 0, u_total, 0, 0]
            b.dtopk_set(state)
            top, first, init_k = nbits, True, True
            local_last, shift_last, bits_last = None, 0, 0
            while top > 0:
                bits = min(SEL_BITS, top)
                shift = top - bits
                hist = b.dtopk_hist(0, shift, bits, first, smin)
                if shift == 0 and not det:  # last score pass: remember the local tie counts
                    local_last, shift_last, bits_last = hist.clone(), shift, bits
                comm.all_reduce(hist, dist.ReduceOp.SUM)
                b.dtopk_pick(0, shift, first, init_k, K)
                first = init_k = False
                top = shift
            state = b.dtopk_get()
            T, quota, tie_count = state[0], state[1], state[3]
            if det:
                if quota < tie_count:
                    all_ties = False
                    for word, wbits in ((1, 41), (2, 64)):
                        top, first = wbits, True
                        while top > 0:
                            bits = min(SEL_BITS, top)
                            shift = top - bits
                            hist = b.dtopk_hist(word, shift, bits, first, smin)
                            comm.all_reduce(hist, dist.ReduceOp.SUM)
                            b.dtopk_pick(word, shift, first, False, K)
                            first = False
                            top = shift
            else:
                all_ties = False
                my_ties = u_local if nbits == 0 else int(local_last[(T >> shift_last) & ((1 << bits_last) - 1)])
                ties = comm.gather_ints(my_ties)[:, 0]
                before = int(ties[:me].sum())
                state[1] = max(0, min(my_ties, quota - before))  # arrival order == (rank, local arrival)
                b.dtopk_set(state)
        idx, y, kl, kh = b.dtopk_cut(self.tie, keep_all, all_ties, smin, smax, u_local)
        # global rank of every local survivor
        k_local = idx.shape[0]
        k_all = comm.gather_ints(k_local)[:, 0]
        k_total = int(k_all.sum())
        grank = torch.arange(k_local, dtype=torch.int64, device=dev)
        if G > 1:
            ys = comm.all_gather_v(y, k_all)
            kls = comm.all_gather_v(kl, k_all) if det else [None] * G
            khs = comm.all_gather_v(kh, k_all) if det else [None] * G
            words = 3 if det else 1
            for g in range(G):
                if g == me or k_all[g] == 0 or k_local == 0:
                    continue
                # ties across ranks (stable policy only) go to the lower rank: it arrived first
                b.count_less(words, g < me, (y, kl, kh), (ys[g], kls[g], khs[g]), grank, True)
        kept_rows = b.move_rows(winners, idx, k_local, False)  # local survivors in local rank order
        if G == 1:
            return kept_rows  # grank == arange: already the new queue
        # block distribution of the new queue by global rank
        chunk = -(-k_total // G)
        dest = grank // chunk if k_local else grank
        send_counts = torch.bincount(dest, minlength=G).cpu().numpy() if k_local else np.zeros(G, np.int64)
        all_counts = comm.gather_ints(*send_counts.tolist())
        recv_counts = all_counts[:, me]
        got = comm.all_to_all_rows(kept_rows, send_counts, recv_counts)
        got_rank = comm.all_to_all_rows(grank, send_counts, recv_counts)
        return b.move_rows(got, got_rank - me * chunk, got.shape[0], True)

    # ------------------------------------------------------------------ driver helpers
    def run(self, max_levels=None):
        while not self.ended:
            self.step()
            if max_levels is not None and len(self.infos) >= max_levels:
                break
        return self.infos

    def path(self):
        """(ranks, ordinals) of the winning line, walking the distributed link columns (src/solver.py:459-464)."""
        dev = self.b.device
        L = self.level
        r = self.goal_rank
        ranks, ords = [0] * (L + 1), [0] * L
        for lv in range(L, -1, -1):
            ranks[lv] = r
            if lv > 0:
                base, col = self.links[lv]
                t = torch.zeros(1, dtype=torch.int64, device=dev)
                if base <= r < base + col.shape[0]:
                    t[0] = col[r - base]
                self.comm.all_reduce(t, dist.ReduceOp.SUM)
                link = int(t)
                ords[lv - 1] = link & 0xff
                r = link >> 8
        return ranks, ords

    def gather_frontier(self):
        """the whole current queue on every rank, in global rank order (tests / small cases only)"""
        sizes, _, _ = self._sizes(self.front.shape[0])
        cols = [torch.cat(self.comm.all_gather_v(self.front[:, c].contiguous(), sizes)) for c in range(4)]
        return torch.stack(cols, dim=1)
