"""Possible-buys table (public names of the reference's src/buys.py).

Built natively (csrc/spl_tables.cuh) in separable per-colour mask form; `possible_buys()`
materialises the reference's dict view of it.  The reference's pickle cache
(src/buys.py:20-36) is I/O outside the hot path and is not reproduced.
"""
import ctypes as C
from functools import cache

from ._lib import lib
from .gems import Gems, all_gem_sets

Buys = dict[Gems, tuple[int, ...]]


def buys_for(key: Gems) -> tuple[int, ...]:
    buf = (C.c_uint8 * 90)()
    n = lib.spl_host_buys((C.c_uint8 * 5)(*key), buf)
    if n < 0:
        raise ValueError(f'invalid buys key {key!r}')
    return tuple(buf[i] for i in range(n))


def possible_buys() -> Buys:
    """{clamped gems+bonus: ascending card indices with cost <= key} (src/buys.py:13-17)."""
    return {g: buys_for(g) for g in all_gem_sets}


@cache
def get_buys() -> Buys:
    return possible_buys()


def load_buys(*, update: bool = False) -> Buys:  # noqa: ARG001 - signature parity with src/buys.py:25
    return get_buys()


def store_buys(buys: Buys, path='buys.pickle'):
    """Write the table in the reference's cache format (src/buys.py:20-22): a pickled {gems: card indices} dict
    that the reference's `load_buys()` reads back unchanged."""
    import pickle
    with open(path, 'wb') as f:
        pickle.dump(buys, f, pickle.HIGHEST_PROTOCOL)


def export_buys_to_txt(path: str = 'buys.txt'):
    """Dump `key: card indices` per line, the format of the reference's debug exporter (src/buys.py:44-49)."""
    buys = get_buys()
    with open(path, mode='w', encoding='utf-8') as f:
        print('Writing buys to a text file...')
        for g in buys:
            f.write(f'{g}: {buys[g]}\n')
