"""Small helpers on the hot path's edge (bit-list conversions used by measurement read-out)."""
import math
from fractions import Fraction
from typing import Sequence

__all__ = ['bitlist_to_int', 'int_to_bitlist', 'invert_map', 'rationalize', 'symbolize']

# denominators a gate parameter is recognised with (quantumflow/utils.py:166-189)
_DENOMINATORS = frozenset([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])


def rationalize(flt: float, denominators=None) -> Fraction:
    """Fraction with a standard small denominator closest to `flt`; ValueError when there is none."""
    frac = Fraction.from_float(float(flt)).limit_denominator()
    if frac.denominator not in (_DENOMINATORS if denominators is None else denominators):
        raise ValueError('Cannot rationalize')
    return frac


def symbolize(flt: float) -> str:
    """Text of a real parameter the way the reference prints it in Quil (quantumflow/utils.py:192-208 via sympy):
    a small fraction ('3', '1/2') or a small fraction of pi ('pi/2', '3*pi/4'); ValueError otherwise."""
    try:
        frac = rationalize(flt)
        return str(frac.numerator) if frac.denominator == 1 else '{}/{}'.format(frac.numerator, frac.denominator)
    except ValueError:
        frac = rationalize(flt / math.pi)
    num, den = frac.numerator, frac.denominator
    if num == 0:
        return '0'
    head = 'pi' if num == 1 else '-pi' if num == -1 else '{}*pi'.format(num)
    return head if den == 1 else '{}/{}'.format(head, den)


def bitlist_to_int(bitlist: Sequence[int]) -> int:
    """[1, 0, 0] -> 4 (most significant bit first, the order of `State.measure()` outputs)."""
    value = 0
    for bit in bitlist:
        value = (value << 1) | int(bit)
    return value


def int_to_bitlist(x: int, pad: int = None) -> Sequence[int]:
    """4 -> [1, 0, 0]; `pad` left-fills with zeros to at least that many bits."""
    bits = [int(ch) for ch in bin(int(x))[2:]]
    if pad is not None and len(bits) < pad:
        bits = [0] * (pad - len(bits)) + bits
    return bits


def invert_map(mapping: dict, one_to_one: bool = True) -> dict:
    if one_to_one:
        return {v: k for k, v in mapping.items()}
    inv: dict = {}
    for k, v in mapping.items():
        inv.setdefault(v, set()).add(k)
    return inv
