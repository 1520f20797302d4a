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
