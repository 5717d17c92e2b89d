"""Classical memory addresses (keys of `State.memory`). Behavioural contract: quantumflow/cbits.py:24-80 --
hashable, totally ordered by (dtype, register name, key), printable as `name[key]`."""
from functools import total_ordering
from typing import Any, Hashable

__all__ = ['Register', 'Addr']

DTYPE = {'BIT', 'REAL', 'INT', 'OCTET', 'ANY'}
DEFAULT_REGION = 'ro'


@total_ordering
class Register:
    """A named region of classical memory; indexing it yields an :class:`Addr`."""
    __slots__ = ('name', 'dtype')

    def __init__(self, name: str = DEFAULT_REGION, dtype: str = 'BIT') -> None:
        if dtype not in DTYPE:
            raise AssertionError('unknown register dtype {!r}'.format(dtype))
        self.name = name
        self.dtype = dtype

    def _key(self):
        return (self.dtype, self.name)

    def __getitem__(self, key: Hashable) -> 'Addr':
        return Addr(self, key)

    def __eq__(self, other: Any) -> bool:
        return self._key() == other._key() if isinstance(other, Register) else NotImplemented

    def __lt__(self, other: Any) -> bool:
        return self._key() < other._key() if isinstance(other, Register) else NotImplemented

    def __hash__(self) -> int:
        return hash(self._key())

    def __repr__(self) -> str:
        return 'Register({!r}, {!r})'.format(self.name, self.dtype)


@total_ordering
class Addr:
    """One cell of a :class:`Register`."""
    __slots__ = ('register', 'key')

    def __init__(self, register: Register, key: Hashable) -> None:
        self.register = register
        self.key = key

    @property
    def dtype(self) -> str:
        return self.register.dtype

    def _key(self):
        return (self.register, self.key)

    def __eq__(self, other: Any) -> bool:
        return self._key() == other._key() if isinstance(other, Addr) else NotImplemented

    def __lt__(self, other: Any) -> bool:
        return self._key() < other._key() if isinstance(other, Addr) else NotImplemented

    def __hash__(self) -> int:
        return hash(self._key())

    def __str__(self) -> str:
        return '{}[{}]'.format(self.register.name, self.key)

    def __repr__(self) -> str:
        return '{!r}[{!r}]'.format(self.register, self.key)
