"""Chunked on-disk dump of a state vector / density (checkpoint and interop for states of up to
128 GiB per GPU; the reference has nothing comparable -- its states are numpy arrays one would np.save).

File layout (little endian): a 64-byte header -- magic b'QFBSTATE', u32 version, u32 index bits, u32 rank of the
tensor (1 state, 2 density), u32 reserved, u64 number of amplitudes, padding -- followed by the amplitudes as
complex128 in flat C order, i.e. exactly the reference's `[2]*N` tensor (`ket.vec.asarray().ravel()`), so that
`numpy.fromfile(path, dtype=complex128, offset=64)` reads it anywhere. The copy goes through one pinned staging
buffer of `chunk_bytes`, so a 128 GiB state needs no second 128 GiB of host memory. A sharded state is one file
per rank (the shard in its physical layout) plus whatever the caller records of `final_phys_of`.
"""
import struct
from typing import Optional, Tuple

import numpy as np
import torch

MAGIC = b'QFBSTATE'
VERSION = 1
HEADER_BYTES = 64
_HEADER = struct.Struct('<8sIIIIQ')


def _header(nbits: int, rank: int, count: int) -> bytes:
    head = _HEADER.pack(MAGIC, VERSION, nbits, rank, 0, count)
    return head + b'\0' * (HEADER_BYTES - len(head))


def read_header(path: str) -> Tuple[int, int, int]:
    """(index bits, tensor rank, number of amplitudes) of a dump."""
    with open(path, 'rb') as f:
        raw = f.read(HEADER_BYTES)
    if len(raw) != HEADER_BYTES:
        raise ValueError('{}: not a state dump (short header)'.format(path))
    magic, version, nbits, rank, _reserved, count = _HEADER.unpack(raw[:_HEADER.size])
    if magic != MAGIC or version != VERSION:
        raise ValueError('{}: not a state dump (magic / version)'.format(path))
    if count != 1 << nbits:
        raise ValueError('{}: header is inconsistent'.format(path))
    return nbits, rank, count


def write_amplitudes(tensor: torch.Tensor, path: str, rank: int = 1, chunk_bytes: int = 1 << 28) -> None:
    """Dump a contiguous complex128 tensor of 2^n amplitudes (device or host) to `path`."""
    if tensor.dtype != torch.complex128 or not tensor.is_contiguous():
        raise TypeError('write_amplitudes: need a contiguous complex128 tensor')
    flat = tensor.reshape(-1)
    count = flat.numel()
    nbits = count.bit_length() - 1
    if count != 1 << nbits:
        raise ValueError('write_amplitudes: the number of amplitudes is not a power of two')
    chunk = max(1, min(count, int(chunk_bytes) // 16))
    staging = torch.empty(chunk, dtype=torch.complex128, pin_memory=tensor.is_cuda)
    with open(path, 'wb') as f:
        f.write(_header(nbits, rank, count))
        for start in range(0, count, chunk):
            n = min(chunk, count - start)
            staging[:n].copy_(flat[start:start + n])          # synchronous device -> pinned host copy
            f.write(staging[:n].numpy().tobytes() if n < chunk else staging.numpy().data)


def read_amplitudes(path: str, device: Optional[torch.device] = None, chunk_bytes: int = 1 << 28) -> torch.Tensor:
    """The flat complex128 tensor of a dump, assembled on `device` (default: host) chunk by chunk."""
    nbits, _rank, count = read_header(path)
    out = torch.empty(count, dtype=torch.complex128, device=device)
    chunk = max(1, min(count, int(chunk_bytes) // 16))
    staging = torch.empty(chunk, dtype=torch.complex128, pin_memory=out.is_cuda)
    view = staging.numpy()
    with open(path, 'rb') as f:
        f.seek(HEADER_BYTES)
        for start in range(0, count, chunk):
            n = min(chunk, count - start)
            got = f.readinto(memoryview(view[:n]).cast('B'))
            if got != 16 * n:
                raise ValueError('{}: truncated dump'.format(path))
            out[start:start + n].copy_(staging[:n])
    return out


def save_state(state, path: str, chunk_bytes: int = 1 << 28) -> None:
    """Dump a State (rank 1) or Density (rank 2); the qubit labels are the caller's to keep."""
    rank = getattr(state, '_RANK', 1)
    tensor = state.tensor
    write_amplitudes(tensor if tensor.is_contiguous() else tensor.contiguous(), path, rank, chunk_bytes)


def load_state(path: str, qubits=None, chunk_bytes: int = 1 << 28):
    """State / Density from a dump, resident in HBM (amplitude tensors live on the device)."""
    from . import backend as bk
    from .states import Density, State
    nbits, rank, _count = read_header(path)
    tensor = read_amplitudes(path, bk.device(), chunk_bytes).reshape([2] * nbits)
    nqubits = nbits // rank
    labels = tuple(range(nqubits)) if qubits is None else tuple(qubits)
    return (Density if rank == 2 else State)(tensor, labels)


# ---------------------------------------------------------------------------------------------------------
# pyQuil wavefunction order (quantumflow/forest/__init__.py:350-366)
# ---------------------------------------------------------------------------------------------------------
# pyQuil labels basis states backwards: its flat amplitude vector is the reference's [2]*N tensor with the axes
# reversed (`amplitudes.transpose().reshape(size)`), i.e. the flat index with its N bits in reverse order. pyquil
# itself is not needed for the layout half (and is not installed); the functions below return / accept the flat
# vector a `pyquil.Wavefunction` holds. The bit reversal runs on the device (qfb_permute_bits).

def state_to_wavefunction_amplitudes(state) -> np.ndarray:
    """Flat complex128 vector in pyQuil's order: what `forest.state_to_wavefunction(state).amplitudes` holds."""
    from . import engine
    n = state.qubit_nb
    reversed_bits = engine.permute_bits(state.tensor.reshape(-1), list(range(n - 1, -1, -1)))
    return reversed_bits.cpu().numpy().reshape(-1)


def state_from_pyquil_order(amplitudes, qubits=None):
    """Inverse of state_to_wavefunction_amplitudes: the state whose pyQuil-ordered flat vector is `amplitudes`.
    (The reference's own `wavefunction_to_state` transposes a 1-D array, which is a no-op, so it re-reads the vector
    in QuantumFlow order, forest/__init__.py:361-364; this function performs the bit reversal.)"""
    from . import engine
    from .states import State
    vec = np.asarray(amplitudes, dtype=np.complex128).reshape(-1)
    n = int(np.log2(vec.size))
    st = State(vec.reshape([2] * n), qubits)
    tensor = engine.permute_bits(st.tensor.reshape(-1), list(range(n - 1, -1, -1)))
    return State(tensor.reshape([2] * n), st.qubits)
