"""The five gem colours, in the column order of the packed records (bits 3c..3c+2 of the gems word,
bonus field c of the aux word).  Same names and values as the reference's `Color` enum (src/color.py)."""
import enum

_NAMES = ('WHITE', 'BLUE', 'GREEN', 'RED', 'BLACK')


class Color(enum.Enum):
    _ignore_ = ['_i', '_n']
    for _i, _n in enumerate(_NAMES):
        vars()[_n] = _i

    def __repr__(self):
        return f'Color.{self.name}'


COLOR_NUM = len(_NAMES)
assert [c.value for c in Color] == list(range(COLOR_NUM))
