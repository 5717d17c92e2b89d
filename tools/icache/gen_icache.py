#!/usr/bin/env python
"""Instruction-cache probe (CPU part): straight-line FP64 code of a given size inside a loop, 128 threads per CTA and
168 registers like the sweep kernels. Writes icache_<KB>.cubin for run_icache.py. Each "op" is 64 DFMA = 1 KiB of SASS."""
import random, subprocess, sys, os
R = 5; NE = 32
HERE = os.path.dirname(os.path.abspath(__file__))

def pairs_of(j):
    for p in range(NE // 2):
        e0 = ((p >> j) << (j + 1)) | (p & ((1 << j) - 1)); yield e0, e0 | (1 << j)

def gen(nops):
    L = ["""
.version 8.7
.target sm_100a
.address_size 64
.visible .entry k(.param .u64 state, .param .u32 iters, .param .f64 pc0, .param .f64 pc1)
.maxntid 128,1,1
.minnctapersm 3
{
.reg .f64 r<32>, i<32>, c<8>;
.reg .b64 a, b; .reg .b32 t, n; .reg .pred p;
ld.param.u64 a, [state];
cvta.to.global.u64 a, a;
ld.param.u32 n, [iters];
ld.param.f64 c0, [pc0];
ld.param.f64 c1, [pc1];
mov.u32 t, %tid.x;
mul.wide.u32 b, t, 16;
add.u64 a, a, b;
"""]
    for e in range(NE):
        L.append("ld.global.v2.f64 {r%d, i%d}, [a+%d];" % (e, e, e * 2048))
    L.append("LOOP:")
    rnd = random.Random(1)
    for o in range(nops):
        j = rnd.randrange(R)
        for x, y in pairs_of(j):
            L.append("fma.rn.f64 r%d, c0, r%d, r%d;" % (x, y, x))
            L.append("fma.rn.f64 i%d, c0, i%d, i%d;" % (x, y, x))
            L.append("fma.rn.f64 r%d, c1, r%d, r%d;" % (y, x, y))
            L.append("fma.rn.f64 i%d, c1, i%d, i%d;" % (y, x, y))
    L.append("sub.u32 n, n, 1; setp.ne.u32 p, n, 0; @p bra LOOP;")
    for e in range(NE):
        L.append("st.global.v2.f64 [a+%d], {r%d, i%d};" % (e * 2048, e, e))
    L.append("ret;\n}")
    return "\n".join(L)

if __name__ == '__main__':
    for kb in [int(x) for x in sys.argv[1:]] or [8, 16, 24, 32, 40, 48, 64, 96, 128]:
        ptx = os.path.join(HERE, 'icache_%d.ptx' % kb)
        open(ptx, 'w').write(gen(kb))
        subprocess.check_call(['ptxas', '-arch=sm_100a', '-O3', ptx, '-o', os.path.join(HERE, 'icache_%d.cubin' % kb)])
        os.remove(ptx)
        print('built', kb)
