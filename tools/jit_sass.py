#!/usr/bin/env python
"""SASS of the sweep-specialised kernels of the benchmark plan, without a GPU: plan the W-B circuit, let the library
emit the PTX of every sweep (qfb_jit_ptx; the QFB_JIT_* environment knobs apply), assemble with ptxas for sm_100a
and summarise the tile loop: registers, spills, static instruction mix.
Usage: python tools/jit_sass.py [qubits=30] [depth=20] [seed=0] [tile_bits=11] [sweeps=0,7] [keep_dir]"""
import collections
import ctypes
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantumflow_b200 as qf                      # noqa: E402
from quantumflow_b200 import _lib, planner, workloads    # noqa: E402


def sweep_ptx(lib, blob, i):
    need, ncoef = ctypes.c_size_t(0), ctypes.c_size_t(0)
    rc = lib.qfb_jit_ptx(blob, len(blob), i, None, 0, ctypes.byref(need), ctypes.byref(ncoef))
    if rc != 0:
        return None, 0
    buf = ctypes.create_string_buffer(need.value)
    assert lib.qfb_jit_ptx(blob, len(blob), i, buf, need.value, None, None) == 0
    return buf.value.decode(), ncoef.value


def sass_summary(ptx, workdir, tag):
    p = os.path.join(workdir, tag + '.ptx')
    c = os.path.join(workdir, tag + '.cubin')
    open(p, 'w').write(ptx)
    r = subprocess.run(['ptxas', '-arch=sm_100a', '-O3', '-v', p, '-o', c], capture_output=True, text=True)
    info = r.stderr
    regs = re.search(r'Used (\d+) registers', info)
    spill = re.search(r'(\d+) bytes spill stores, (\d+) bytes spill loads', info)
    sass = subprocess.run(['cuobjdump', '-sass', c], capture_output=True, text=True).stdout
    ops = []
    for line in sass.splitlines():
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            ops.append((m.group(2), line))
    mix = collections.Counter(o.split('.')[0] for o, _ in ops)
    cbank = sum(1 for o, l in ops if o.startswith(('DFMA', 'DMUL', 'DADD')) and 'c[' in l)
    ureg = sum(1 for o, l in ops if o.startswith(('DFMA', 'DMUL', 'DADD')) and re.search(r'\bUR\d+', l))
    return dict(regs=int(regs.group(1)) if regs else -1, spill=(spill.groups() if spill else None), n=len(ops), mix=mix,
                fp64_cbank=cbank, fp64_ureg=ureg)


def main():
    a = sys.argv[1:]
    n = int(a[0]) if len(a) > 0 else 30
    depth = int(a[1]) if len(a) > 1 else 20
    seed = int(a[2]) if len(a) > 2 else 0
    tile = int(a[3]) if len(a) > 3 else 11
    which = [int(x) for x in a[4].split(',')] if len(a) > 4 else [0, 7]
    keep = a[5] if len(a) > 5 else None
    circ = workloads.wb_circuit(qf, n, depth, seed)
    bitops = [(g.matrix(), [n - 1 - circ.qubits.index(q) for q in g.qubits]) for g in circ.elements]
    segments = planner.build_segments(n, bitops, tile_bits=tile, reg_bits=4)
    lib = _lib.load()
    workdir = keep or tempfile.mkdtemp()
    os.makedirs(workdir, exist_ok=True)
    for seg in segments:
        for i in which:
            ptx, ncoef = sweep_ptx(lib, seg.blob, i)
            if ptx is None:
                continue
            s = sass_summary(ptx, workdir, 'sweep%d' % i)
            top = ' '.join('%s:%d' % kv for kv in s['mix'].most_common(14))
            print('sweep %d: coef %d regs %d spill %s SASS %d  fp64 with c[] %d, with UR %d\n   %s' % (
                i, ncoef, s['regs'], s['spill'], s['n'], s['fp64_cbank'], s['fp64_ureg'], top))


if __name__ == '__main__':
    main()
