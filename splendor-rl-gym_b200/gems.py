"""Gem arithmetic and the gem-take table (public names of the reference's src/gems.py).

The take table itself is built natively inside libsplendor_b200.so (csrc/spl_tables.cuh) --
the same table the expand kernel reads -- and only *viewed* here.
"""
import ctypes as C
from functools import cache
from itertools import product

from ._lib import lib
from .color import COLOR_NUM, Color

MAX_GEMS = 7
Gems = tuple[int, ...]
all_gem_sets = tuple(product(range(MAX_GEMS + 1), repeat=COLOR_NUM))


def take_gems(g: Gems):
    """All hands reachable by one take action, in the reference's order (src/gems.py:85-108)."""
    buf = (C.c_uint8 * 500)()
    n = lib.spl_host_takes((C.c_uint8 * 5)(*g), buf)
    if n < 0:
        raise ValueError(f'invalid gem tuple {g!r}')
    return tuple(tuple(buf[i * 5 + c] for c in range(5)) for i in range(n))


@cache
def get_takes() -> dict[Gems, tuple[Gems, ...]]:
    """{gems: successor gem tuples} over all 8^5 keys (src/gems.py:111-113)."""
    return {g: take_gems(g) for g in all_gem_sets}


def add(g1: Gems, g2: Gems) -> Gems:
    return tuple(a + b for a, b in zip(g1, g2))


def is_valid(g: Gems) -> bool:
    return all(0 <= x <= MAX_GEMS for x in g)


def subtract_with_bonus(gems: Gems, cost: Gems, bonus: Gems) -> tuple[Gems, int]:
    """`gems - max(cost - bonus, 0)` per colour, clamped at 0, and the gems saved (src/gems.py:116-129)."""
    out, saved = [], 0
    for g, c, b in zip(gems, cost, bonus):
        pay = max(c - b, 0)
        saved += c - pay
        out.append(max(g - pay, 0))
    return tuple(out), saved


def increase_bonus(bonus: Gems, color: Color) -> Gems:
    """bonus with one more card of `color` (src/gems.py:141-143)."""
    return tuple(b + (i == color.value) for i, b in enumerate(bonus))
