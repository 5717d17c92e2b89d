"""Small helpers on the hot path's edge (bit-list conversions used by measurement read-out)."""
from typing import Sequence

__all__ = ['bitlist_to_int', 'int_to_bitlist', 'invert_map']


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
