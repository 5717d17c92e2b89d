"""In-tree build of libqfb200.so (the CUDA kernels + C ABI) for sm_100a.

The shared object is written next to the sources (quantumflow_b200/csrc/libqfb200.so); it is git-ignored but
travels with the working tree. nvcc cross-compiles without a GPU.
"""
import hashlib
import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc')
LIB_PATH = os.path.join(CSRC, 'libqfb200.so')
STAMP_PATH = os.path.join(CSRC, '.libqfb200.stamp')
SOURCES = ['qfb_api.cu', 'qfb_apply.cu', 'qfb_reduce.cu', 'qfb_sweep.cu', 'qfb_planhost.cu', 'qfb_jit.cu', 'qfb_remap.cu', 'qfb_batch.cu', 'qfb_small.cu']
HEADERS = ['qfb_common.cuh', 'qfb_plan.h', 'qfb_jit.h', 'qfb_oploop.inc', os.path.join('..', '..', 'include', 'qfb200.h')]

NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-shared']


def _find_nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; libqfb200.so cannot be built')


def _source_digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), 'rb') as f:
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def library_is_current() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return False
    with open(STAMP_PATH) as f:
        return f.read().strip() == _source_digest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile libqfb200.so if the sources changed (or `force`). Returns the library path."""
    if not force and library_is_current():
        return LIB_PATH
    nvcc = _find_nvcc()
    # the PTX compiler (sweep-specialised kernels, qfb_jit.cu) is linked statically; the driver API is resolved
    # with dlopen at run time so that the library still loads where libcuda is absent
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB_PATH] + SOURCES + \
        ['-lnvptxcompiler_static', '-lpthread', '-ldl']
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    with open(STAMP_PATH, 'w') as f:
        f.write(_source_digest())
    return LIB_PATH


if __name__ == '__main__':
    import sys
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
