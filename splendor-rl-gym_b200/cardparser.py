"""The 90-card deck (mirrors the public names of the reference's src/cardparser.py).

The deck is not parsed from a CSV here: it comes from the packed constant table compiled
into libsplendor_b200.so (`spl_deck_table`, csrc/deck_table.h), so host and device share
one source of truth.
"""
import ctypes as C
from dataclasses import dataclass
from functools import cache, cached_property

from ._lib import lib
from .color import Color

CardIndex = int
CardIndices = tuple[CardIndex, ...]
Gems = tuple[int, ...]


@dataclass(frozen=True)
class Card:
    """cost / pt / bonus / index, as src/cardparser.py:17-22."""
    cost: Gems
    pt: int
    bonus: Color
    index: int

    @cached_property
    def str_id(self) -> str:
        """`<pt><colour letter><sorted non-zero costs>`, e.g. `2W124` (src/cardparser.py:34-45)."""
        letter = self.bonus.name[0] if self.bonus is not Color.BLACK else 'K'
        return f'{self.pt}{letter}' + ''.join(sorted(str(x) for x in self.cost if x))

    def __str__(self):
        return self.str_id

    def __hash__(self):
        return self.index

    def __eq__(self, other):
        return self.index == other.index


@cache
def get_deck() -> tuple[Card, ...]:
    n = C.c_int32()
    tab = lib.spl_deck_table(C.byref(n))
    deck = []
    for i in range(n.value):
        v = tab[i]
        deck.append(Card(cost=tuple((v >> (3 * c)) & 7 for c in range(5)), pt=(v >> 15) & 7,
                         bonus=Color((v >> 18) & 7), index=i))
    return tuple(deck)


def sort_cards(cards):
    """Sort by points, total cost, sorted cost tuple, colour (src/cardparser.py:69-76)."""
    return tuple(sorted(cards, key=lambda c: (c.pt, sum(c.cost), sorted(c.cost), c.bonus.value)))
