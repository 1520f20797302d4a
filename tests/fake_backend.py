"""CPU stand-in for sharded.CudaBackend, used ONLY by the gloo tests of the multi-rank host logic.
Every primitive is restated with numpy / the CPU oracle (test infrastructure); the ownership hash is
deliberately different from the device one -- results must not depend on who owns which key."""
import numpy as np
import torch

import oracle

M64 = (1 << 64) - 1
SEL_BITS = 11


def _u(t):
    return t.cpu().numpy().view(np.uint64)


def _flip(scores):
    b = scores.view(np.uint64).copy()
    b[b == np.uint64(1 << 63)] = 0
    neg = (b >> np.uint64(63)).astype(bool)
    return np.where(neg, ~b, b | np.uint64(1 << 63))


class FakeBackend:
    def __init__(self):
        self.device = torch.device('cpu')
        self.visited = set()
        self.sel = [0, 0, 0, 0, 0, 0]
        self.hist = torch.zeros(1 << SEL_BITS, dtype=torch.int32)

    def reset_visited(self):
        self.visited = set()

    def visited_count(self):
        return len(self.visited)

    def first_goal(self, front, goal):
        pts = (_u(front)[:, 2] >> np.uint64(16)) & np.uint64(0xff)
        hit = np.nonzero(pts >= goal)[0]
        return int(hit[0]) if len(hit) else -1

    def expand(self, front):
        f = _u(front)
        out = []
        for r in range(f.shape[0]):
            rec = np.zeros(1, oracle.STATE_DTYPE)
            rec['lo'], rec['hi'], rec['aux'] = f[r, 0], f[r, 1], f[r, 2]
            kids = oracle.expand(rec)
            for o in range(len(kids)):
                out.append((kids['lo'][o], kids['hi'][o], kids['aux'][o], np.uint64(r << 8 | o)))
        a = np.array(out, dtype=np.uint64).reshape(-1, 4)
        return torch.from_numpy(a.view(np.int64).copy())

    def expand_rows(self, front, rank_base):
        out = self.expand(front)
        if out.shape[0]:
            out[:, 3] += rank_base << 8
        return out

    def ident_rows(self, cand):
        f = _u(cand)
        rec = np.zeros(f.shape[0], oracle.STATE_DTYPE)
        rec['lo'], rec['hi'] = f[:, 0], f[:, 1]
        out = np.zeros((f.shape[0], 4), dtype=np.uint64)
        out[:, 0] = oracle.pyhash(rec)
        return torch.from_numpy(out.view(np.int64).copy())

    def route_keys(self, cand, world):
        perm, counts = self.owner_partition(cand[:, :2], world)
        self._perm = perm
        return cand[perm][:, :2].contiguous(), counts

    def dedup_flags(self, keys):
        flags = torch.zeros(keys.shape[0], dtype=torch.uint8)
        src = self.dedup(keys)
        if len(src):
            flags[src] = 1
        return flags

    def compact_winners(self, cand, flags_send_order):
        flags = torch.zeros(cand.shape[0], dtype=torch.uint8)
        flags[self._perm] = flags_send_order
        return cand[torch.nonzero(flags).flatten()]

    def scatter_into(self, dst, rows, pos):
        dst[pos] = rows

    def move_rows(self, rows, idx, n_out, scatter):
        out = torch.empty((n_out, 4), dtype=torch.int64)
        if scatter:
            out[idx] = rows
        else:
            out = rows[idx]
        return out

    def owner_partition(self, keys, world):
        k = _u(keys)
        own = ((k[:, 0] * np.uint64(0x9E3779B97F4A7C15) + k[:, 1]) >> np.uint64(40)) % np.uint64(world) if len(k) else np.zeros(0, np.uint64)
        perm = np.argsort(own, kind='stable')
        return torch.from_numpy(perm.astype(np.int64)), np.bincount(own.astype(np.int64), minlength=world).astype(np.int64)

    def dedup(self, keys):
        k = _u(keys)
        src = []
        for i in range(k.shape[0]):
            key = (int(k[i, 0]), int(k[i, 1]))
            if key not in self.visited:
                self.visited.add(key)
                src.append(i)
        return torch.tensor(src, dtype=torch.int64)

    def score(self, heuristic, noise, rows, draws=None):
        r = _u(rows)
        recs = np.zeros(r.shape[0], oracle.STATE_DTYPE)
        recs['lo'], recs['hi'], recs['aux'] = r[:, 0], r[:, 1], r[:, 2]
        if noise == 'mt':
            raise NotImplementedError('mt noise is exercised on the GPU box (tools/mt_multi_check.py)')
        return torch.from_numpy(oracle.score(recs, heuristic, noise))

    # ---- distributed top-k passes (numpy restatement of csrc sel_hist / sel_pick / cut kernels)
    def dtopk_begin(self, scores, keys):
        self.sk = _flip(scores.numpy())
        self.keys = _u(keys) if keys is not None else None
        if len(self.sk) == 0:
            return M64, 0
        return int(self.sk.min()), int(self.sk.max())

    def _word(self, word, smin):
        x = self.sk - np.uint64(smin)
        if word == 0:
            return x, np.ones(len(x), bool)
        m = x == np.uint64(self.sel[0])
        if word == 1:
            return self.keys[:, 1], m
        return self.keys[:, 0], m & (self.keys[:, 1] == np.uint64(self.sel[4]))

    def dtopk_hist(self, word, shift, bits, first, smin):
        self.hist.zero_()
        if len(self.sk):
            x, m = self._word(word, smin)
            prefix = 0 if first else self.sel[{0: 0, 1: 4, 2: 5}[word]]
            hs = shift + bits
            if not first and hs < 64:
                m = m & ((x >> np.uint64(hs)) == np.uint64(prefix >> hs))
            d = ((x[m] >> np.uint64(shift)) & np.uint64((1 << bits) - 1)).astype(np.int64)
            self.hist += torch.from_numpy(np.bincount(d, minlength=1 << SEL_BITS).astype(np.int32))
        return self.hist

    def dtopk_pick(self, word, shift, first, init_k, k):
        h = self.hist.numpy().astype(np.int64)
        slot = {0: 0, 1: 4, 2: 5}[word]
        k_rem = k if init_k else self.sel[1]
        c_gt = 0 if init_k else self.sel[2]
        prefix = 0 if first else self.sel[slot]
        above = 0
        for b in range((1 << SEL_BITS) - 1, -1, -1):
            if above + h[b] >= k_rem:
                self.sel[slot] = prefix | (b << shift)
                self.sel[1] = k_rem - above
                self.sel[2] = c_gt + above
                self.sel[3] = int(h[b])
                break
            above += int(h[b])
        self.hist.zero_()

    def dtopk_get(self):
        return list(self.sel)

    def dtopk_set(self, state):
        self.sel = [int(v) for v in state]

    def dtopk_cut(self, tie, keep_all, all_ties, smin, smax, n):
        if n == 0:
            e = torch.empty(0, dtype=torch.int64)
            return e, e, (e if tie != 'stable' else None), (e if tie != 'stable' else None)
        x = self.sk - np.uint64(smin)
        T = np.uint64(self.sel[0])
        if keep_all:
            keep = np.ones(n, bool)
        elif tie in ('det', 'det_ordered'):
            keep = x > T
            t = x == T
            if all_ties:
                keep |= t
            else:
                hi, lo = self.keys[:, 1], self.keys[:, 0]
                keep |= t & ((hi > np.uint64(self.sel[4])) | ((hi == np.uint64(self.sel[4])) & (lo >= np.uint64(self.sel[5]))))
        else:
            keep = x > T
            ties = np.nonzero(x == T)[0]
            keep[ties[:self.sel[1]]] = True
        idx = np.nonzero(keep)[0]
        y = np.uint64(smax - smin) - x[idx]
        if tie in ('det', 'det_ordered'):
            kl, kh = ~self.keys[idx, 0], ~self.keys[idx, 1] & np.uint64((1 << 41) - 1)
            order = np.lexsort((kl, kh, y))
            if tie == 'det_ordered':
                # SPL_TIE_KEY_ORDERED: the device sorts by score only (stable) and relies on the caller's promise that
                # the keys already fall with the index -- check the promise, then do what the device does
                stable = np.argsort(y, kind='stable')
                assert np.array_equal(order, stable), 'det_ordered: local keys are not in descending order by index'
                order = stable
            f = lambda a: torch.from_numpy(a[order].view(np.int64).copy())  # noqa: E731
            return torch.from_numpy(idx[order].astype(np.int64)), f(y), f(kl), f(kh)
        order = np.argsort(y, kind='stable')
        return torch.from_numpy(idx[order].astype(np.int64)), torch.from_numpy(y[order].view(np.int64).copy()), None, None

    def count_less(self, words, inclusive, a, b, out, accumulate, sorted_a=False):
        ay, by = _u(a[0]), _u(b[0])
        if words == 1:
            c = np.searchsorted(by, ay, side='right' if inclusive else 'left')
        else:
            akl, akh, bkl, bkh = _u(a[1]), _u(a[2]), _u(b[1]), _u(b[2])
            bt = list(zip(by.tolist(), bkh.tolist(), bkl.tolist()))
            import bisect
            c = np.array([(bisect.bisect_right if inclusive else bisect.bisect_left)(bt, t)
                          for t in zip(ay.tolist(), akh.tolist(), akl.tolist())], dtype=np.int64)
        res = torch.from_numpy(np.asarray(c, dtype=np.int64))
        if accumulate:
            out += res
        else:
            out.copy_(res)
