"""The five gem colours, in the column order of the packed records (bits 3c..3c+2 of the gems word,
bonus field c of the aux word).  Same names and values as the reference's `Color` enum (src/color.py:4-15)."""
import enum


class Color(enum.Enum):
    WHITE = 0
    BLUE = 1
    GREEN = 2
    RED = 3
    BLACK = 4

    def __repr__(self):
        return f'Color.{self.name}'


COLOR_NUM = len(Color)
