"""Host side of the chunked state dump (quantumflow_b200/stateio.py): file layout, chunking, error paths. The
device copies are the same torch calls with a CUDA tensor (tests/test_gpu_states.py::test_state_dump_round_trip)."""
import numpy as np
import pytest
import torch

from quantumflow_b200 import stateio


@pytest.mark.parametrize('nbits,chunk_bytes', [(0, 1 << 20), (5, 16), (10, 16 * 100), (12, 1 << 28), (11, 16 * 2048)])
def test_dump_layout_and_chunking(tmp_path, nbits, chunk_bytes):
    rng = np.random.RandomState(nbits)
    vec = rng.normal(size=1 << nbits) + 1j * rng.normal(size=1 << nbits)
    path = str(tmp_path / 'state.qfb')
    stateio.write_amplitudes(torch.from_numpy(vec.copy()), path, rank=1, chunk_bytes=chunk_bytes)
    assert stateio.read_header(path) == (nbits, 1, 1 << nbits)
    # readable without this package: flat C order complex128 behind a 64-byte header
    assert np.array_equal(np.fromfile(path, dtype=np.complex128, offset=stateio.HEADER_BYTES), vec)
    assert np.array_equal(stateio.read_amplitudes(path, chunk_bytes=chunk_bytes).numpy(), vec)
    assert np.array_equal(stateio.read_amplitudes(path, chunk_bytes=48).numpy(), vec)


def test_dump_error_paths(tmp_path):
    path = str(tmp_path / 'bad.qfb')
    with pytest.raises(TypeError):
        stateio.write_amplitudes(torch.zeros(4, dtype=torch.float64), path)
    with pytest.raises(ValueError):
        stateio.write_amplitudes(torch.zeros(6, dtype=torch.complex128), path)
    with open(path, 'wb') as f:
        f.write(b'short')
    with pytest.raises(ValueError):
        stateio.read_header(path)
    good = str(tmp_path / 'good.qfb')
    stateio.write_amplitudes(torch.ones(8, dtype=torch.complex128), good, rank=1)
    raw = open(good, 'rb').read()
    with open(path, 'wb') as f:
        f.write(raw[:-16])
    with pytest.raises(ValueError):
        stateio.read_amplitudes(path)
    with open(path, 'wb') as f:
        f.write(b'NOTSTATE' + raw[8:])
    with pytest.raises(ValueError):
        stateio.read_header(path)
