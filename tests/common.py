"""Shared helpers of the parity tests: digests over packed state records.

`level_digest` recomputes, from packed (lo, hi, aux[, link]) arrays, exactly the digest that
tests/golden/make_golden.py computed from the reference's own State objects.
"""
import hashlib

import numpy as np

M64 = (1 << 64) - 1


def level_digest(lo, hi, aux, link=None):
    lo = np.ascontiguousarray(lo).astype(np.uint64)
    hi = np.ascontiguousarray(hi).astype(np.uint64)
    aux = np.ascontiguousarray(aux).astype(np.uint64)
    n = len(lo)
    saved = (aux & np.uint64(0xffff)).astype(np.uint16)
    pts = ((aux >> np.uint64(16)) & np.uint64(0xff)).astype(np.int64)
    rec = np.zeros(n, dtype=np.dtype([('lo', '<u8'), ('hi', '<u8'), ('saved', '<u2')]))
    rec['lo'], rec['hi'], rec['saved'] = lo, hi, saved
    d = dict(n=n,
             sum_lo=int(np.add.reduce(lo, dtype=np.uint64)) if n else 0,
             sum_hi=int(np.add.reduce(hi, dtype=np.uint64)) if n else 0,
             xor_lo=int(np.bitwise_xor.reduce(lo)) if n else 0,
             xor_hi=int(np.bitwise_xor.reduce(hi)) if n else 0,
             sum_saved=int(saved.astype(np.int64).sum()), sum_pts=int(pts.sum()),
             sha=hashlib.sha256(rec.tobytes()).hexdigest())
    if link is not None:
        link = np.ascontiguousarray(link).astype(np.uint64)
        d['sum_parent_rank'] = int((link >> np.uint64(8)).astype(object).sum()) if n else 0
        d['sum_ordinal'] = int((link & np.uint64(0xff)).astype(np.int64).sum()) if n else 0
    return d


def assert_digest(got, want, what=''):
    for k, v in want.items():
        assert got[k] == v, f'{what}: digest field {k}: got {got[k]}, want {v}'


def key_str(lo, hi):
    return str(int(lo) | int(hi) << 64)


# ---------------------------------------------------------------- realistic mode digests
def _deck_pts():
    import json
    from pathlib import Path
    deck = json.load(open(Path(__file__).resolve().parent / 'golden' / 'tables.json'))['deck']
    return np.array([c['pt'] for c in deck], dtype=np.int64)


def rlevel_digest(recs, num_players, gems_per_color, with_links=True):
    """Same digest as tests/golden/make_golden.py::rlevel_digest, from packed 96-byte records."""
    pts_tab = _deck_pts()
    n = len(recs)
    h = hashlib.sha256()
    raw = np.ascontiguousarray(recs).view(np.uint8).reshape(n, 96) if n else np.zeros((0, 96), np.uint8)
    ident = np.concatenate([raw[:, :16 * num_players], raw[:, 64:77]], axis=1)
    h.update(ident.tobytes())
    sum_saved = int(recs['p']['saved'][:, :num_players].astype(np.int64).sum()) if n else 0
    sum_pts = sum_pool = 0
    if n:
        for q in range(num_players):
            mlo = recs['p']['mlo'][:, q]
            mhi = recs['p']['mhi'][:, q].astype(np.uint64)
            for c in range(90):
                bit = ((mlo >> np.uint64(c)) & np.uint64(1)) if c < 64 else ((mhi >> np.uint64(c - 64)) & np.uint64(1))
                sum_pts += int(bit.sum()) * int(pts_tab[c])
        gems = recs['p']['gems'][:, :num_players].astype(np.int64)
        held = sum(int(((gems >> (3 * k)) & 7).sum()) for k in range(5))
        sum_pool = 5 * gems_per_color * n - held
    d = dict(n=n, sha=h.hexdigest(), sum_saved=sum_saved, sum_pts=sum_pts, sum_pool=sum_pool)
    if with_links:
        link = recs['link'].astype(np.uint64)
        d['sum_parent_rank'] = int((link >> np.uint64(8)).astype(object).sum()) if n else 0
        d['sum_ordinal'] = int((link & np.uint64(0xff)).astype(np.int64).sum()) if n else 0
    return d
