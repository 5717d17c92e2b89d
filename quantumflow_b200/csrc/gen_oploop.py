#!/usr/bin/env python
"""Writes qfb_oploop.inc: the op interpreter of sweep_kernel (qfb_sweep.cu) as one PTX block.

Why PTX and why generated: the interpreter needs (a) ONE indirect jump per op -- `brx.idx` through a branch-target
table; the CUDA C++ front end lowers `switch` to a compare tree (4-5 serial ISETP+BRA levels per op, measured in
profiles/r1_sweep_v6_summary.txt) -- and (b) every amplitude component pinned to one register across all
handlers (no register-renaming copies at the merge points of the loop). The handlers are the same 2^R-amplitude
update unrolled over register bit J / register mask, so they are emitted by the loops below instead of being
typed 48 times. The output is committed next to this script; `python gen_oploop.py` regenerates it and
tests/test_abi.py checks that the committed file is current.

asm operands (see sweep_kernel), NE = 2^R amplitudes per thread: %0..%(2NE-1) amplitude components (a[e].re =
%(2e), a[e].im = %(2e+1)), then the round's running scalar phase (re, im), the shared-memory address of the next
op record (in/out) and the full index of the thread's first amplitude (tile base | thread bits | rank bits << nbits).
"""
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
R = 5
NE = 1 << R
PAIRS = [(j0, j1) for j0 in range(R) for j1 in range(j0)]
PHR, PHI, OP, TFULL = ['%%%d' % (2 * NE + i) for i in range(4)]


def handler_ids():
    text = open(os.path.join(HERE, 'qfb_plan.h')).read()
    return {m.group(1): int(m.group(2)) for m in re.finditer(r'QFB_H_(\w+) = (\d+)', text)}


def re_(e):
    return '%%%d' % (2 * e)


def im_(e):
    return '%%%d' % (2 * e + 1)


def pairs_of(j):
    for p in range(NE // 2):
        e0 = ((p >> j) << (j + 1)) | (p & ((1 << j) - 1))
        yield e0, e0 | (1 << j)


class Emit:
    def __init__(self):
        self.lines = []
        self.nlabel = 0

    def __call__(self, *lines):
        self.lines.extend(lines)

    def label(self, stem):
        self.nlabel += 1
        return '%s_%d' % (stem, self.nlabel)


def emit_on_check(E, off='TAIL'):
    """Skip the op (branch to `off`) unless all idx_cmask bits are set in the thread's full index."""
    E('and.b64 tm, cm, %s;' % TFULL, 'setp.ne.b64 poff, tm, cm;', '@poff bra %s;' % off)


def emit_rc(E):
    E('ld.shared.u8 rc, [cur+6];')          # reg_cmask byte of the op header


def emit_long_tail(E, nbytes):
    """End of a handler whose record is longer than the 32-byte stride the loop assumes: step to the real next
    record and fetch its handler id again."""
    E('add.u32 %s, cur, %d;' % (OP, nbytes), 'ld.shared.u32 hn, [%s];' % OP, 'bra TAIL;')


def emit_general_pair(E, x, y):
    xr, xi, yr, yi = re_(x), im_(x), re_(y), im_(y)
    E('mul.f64 t0, n3, %s;' % yi, 'fma.rn.f64 t0, c2, %s, t0;' % yr,      # pr = m01r y.re - m01i y.im
      'mul.f64 t1, c3, %s;' % yr, 'fma.rn.f64 t1, c2, %s, t1;' % yi,      # pi = m01r y.im + m01i y.re
      'mul.f64 t2, n5, %s;' % xi, 'fma.rn.f64 t2, c4, %s, t2;' % xr,      # qr = m10r x.re - m10i x.im
      'mul.f64 t3, c5, %s;' % xr, 'fma.rn.f64 t3, c4, %s, t3;' % xi,      # qi = m10r x.im + m10i x.re
      'fma.rn.f64 t4, c1, %s, t1;' % xr,                                  # m00i x.re + pi
      'fma.rn.f64 t5, c7, %s, t3;' % yr,                                  # m11i y.re + qi
      'fma.rn.f64 %s, %s, c0, t0;' % (xr, xr), 'fma.rn.f64 %s, n1, %s, %s;' % (xr, xi, xr),
      'fma.rn.f64 %s, %s, c0, t4;' % (xi, xi),
      'fma.rn.f64 %s, %s, c6, t2;' % (yr, yr), 'fma.rn.f64 %s, n7, %s, %s;' % (yr, yi, yr),
      'fma.rn.f64 %s, %s, c6, t5;' % (yi, yi))


def emit_load_matrix(E):
    for q in range(1, 4):          # c0, c1 come from the loop top
        E('ld.shared.v2.f64 {c%d, c%d}, [cur+%d];' % (2 * q, 2 * q + 1, 16 + 16 * q))
    E('neg.f64 n1, c1;', 'neg.f64 n3, c3;', 'neg.f64 n5, c5;', 'neg.f64 n7, c7;')


def emit_pair_guard(E, e0, j):
    """Branch around a pair whose register index does not contain the register control mask rc."""
    free = (NE - 1) & ~(1 << j) & ~e0          # register bits that are 0 in e0 (and are not the target)
    skip = E.label('SKIP')
    E('and.b32 t32, rc, %d;' % free, 'setp.ne.u32 pe, t32, 0;', '@pe bra.uni %s;' % skip)
    return skip


def gen(has_g2):
    H = handler_ids()
    E = Emit()
    targets = ['L_END'] * H['COUNT']
    body = Emit()

    def handler(hid, name):
        targets[hid] = name
        body(name + ':')

    # ---- uncontrolled dense 1-bit operator ----
    for j in range(R):
        handler(H['G1_GENERAL'] + j, 'L_G1G%d' % j)
        emit_load_matrix(body)
        for x, y in pairs_of(j):
            emit_general_pair(body, x, y)
        emit_long_tail(body, 80)
    # ---- pivoted kinds ----
    for j in range(R):
        handler(H['G1_SUMDIFF'] + j, 'L_SD%d' % j)
        # x' = x + r0 y, y' = x' + (r1 - r0) y: sums only (exact zeros under destructive interference)
        body('sub.f64 c2, c1, c0;')
        for x, y in pairs_of(j):
            body('fma.rn.f64 %s, c0, %s, %s;' % (re_(x), re_(y), re_(x)),
                 'fma.rn.f64 %s, c0, %s, %s;' % (im_(x), im_(y), im_(x)),
                 'fma.rn.f64 %s, %s, c2, %s;' % (re_(y), re_(y), re_(x)),
                 'fma.rn.f64 %s, %s, c2, %s;' % (im_(y), im_(y), im_(x)))
        body('bra TAIL;')
    for j in range(R):
        handler(H['G1_LU_R'] + j, 'L_LR%d' % j)
        # two real shears: x += a y; y += b x  (in place, no temporaries)
        for x, y in pairs_of(j):
            body('fma.rn.f64 %s, c0, %s, %s;' % (re_(x), re_(y), re_(x)),
                 'fma.rn.f64 %s, c0, %s, %s;' % (im_(x), im_(y), im_(x)),
                 'fma.rn.f64 %s, c1, %s, %s;' % (re_(y), re_(x), re_(y)),
                 'fma.rn.f64 %s, c1, %s, %s;' % (im_(y), im_(x), im_(y)))
        body('bra TAIL;')
    for j in range(R):
        handler(H['G1_LU_I'] + j, 'L_LI%d' % j)
        # two imaginary shears: x += i a y; y += i b x
        body('neg.f64 n1, c0;', 'neg.f64 n3, c1;')
        for x, y in pairs_of(j):
            body('fma.rn.f64 %s, n1, %s, %s;' % (re_(x), im_(y), re_(x)),
                 'fma.rn.f64 %s, c0, %s, %s;' % (im_(x), re_(y), im_(x)),
                 'fma.rn.f64 %s, n3, %s, %s;' % (re_(y), im_(x), re_(y)),
                 'fma.rn.f64 %s, c1, %s, %s;' % (im_(y), re_(x), im_(y)))
        body('bra TAIL;')
    # ---- controlled dense / X ----
    for j in range(R):
        handler(H['G1C_GENERAL'] + j, 'L_CG%d' % j)
        emit_on_check(body, 'L_CGT%d' % j)
        emit_rc(body)
        emit_load_matrix(body)
        for x, y in pairs_of(j):
            skip = emit_pair_guard(body, x, j)
            emit_general_pair(body, x, y)
            body(skip + ':')
        body('L_CGT%d:' % j)
        emit_long_tail(body, 80)
    for j in range(R):
        handler(H['G1C_SWAPX'] + j, 'L_CX%d' % j)
        emit_on_check(body)
        emit_rc(body)
        for x, y in pairs_of(j):
            skip = emit_pair_guard(body, x, j)
            # the swap as multiplications by the payload's 1.0 (exact copies): one FP64-pipe instruction per
            # 64-bit move instead of two 32-bit register moves, i.e. half the issue slots; the FP64 pipe has
            # room (28 % busy on the benchmark, profiles/r1_sweep_v8_summary.txt)
            body('mul.f64 t0, %s, c0;' % re_(x), 'mul.f64 t1, %s, c0;' % im_(x),
                 'mul.f64 %s, %s, c0;' % (re_(x), re_(y)), 'mul.f64 %s, %s, c0;' % (im_(x), im_(y)),
                 'mul.f64 %s, t0, c0;' % re_(y), 'mul.f64 %s, t1, c0;' % im_(y))
            body(skip + ':')
        body('bra TAIL;')
    # ---- phase terms ----
    handler(H['CPH_SCALAR'], 'L_PS')
    emit_on_check(body)
    body('neg.f64 n1, c1;',
         'mul.f64 t0, n1, %s;' % PHI, 'mul.f64 t1, c1, %s;' % PHR,
         'fma.rn.f64 %s, c0, %s, t0;' % (PHR, PHR), 'fma.rn.f64 %s, c0, %s, t1;' % (PHI, PHI), 'bra TAIL;')

    def cmul(e):
        body('mul.f64 t0, n1, %s;' % im_(e), 'mul.f64 t1, c1, %s;' % re_(e),
             'fma.rn.f64 %s, %s, c0, t0;' % (re_(e), re_(e)), 'fma.rn.f64 %s, %s, c0, t1;' % (im_(e), im_(e)))

    def neg(e):
        body('xor.b64 %s, %s, 0x8000000000000000;' % (re_(e), re_(e)),
             'xor.b64 %s, %s, 0x8000000000000000;' % (im_(e), im_(e)))

    for j in range(R):
        handler(H['CPH_REG1'] + j, 'L_P1%d' % j)
        emit_on_check(body)
        body('neg.f64 n1, c1;')
        for e in range(NE):
            if (e >> j) & 1:
                cmul(e)
        body('bra TAIL;')
    for j in range(R):
        handler(H['CPH_RSC1'] + j, 'L_R1%d' % j)
        emit_on_check(body)
        for e in range(NE):
            if (e >> j) & 1:
                body('mul.f64 %s, %s, c0;' % (re_(e), re_(e)), 'mul.f64 %s, %s, c0;' % (im_(e), im_(e)))
        body('bra TAIL;')
    for j in range(R):
        handler(H['CPH_NEG1'] + j, 'L_N1%d' % j)
        emit_on_check(body)
        for e in range(NE):
            if (e >> j) & 1:
                neg(e)
        body('bra TAIL;')
    for pi, (j0, j1) in enumerate(PAIRS):
        handler(H['CPH_NEG2'] + pi, 'L_N2%d' % pi)
        emit_on_check(body)
        mask = (1 << j0) | (1 << j1)
        for e in range(NE):
            if (e & mask) == mask:
                neg(e)
        body('bra TAIL;')
    for name, is_neg in (('CPH_REGM', False), ('CPH_NEGM', True)):
        handler(H[name], 'L_PM%d' % int(is_neg))
        emit_on_check(body)
        emit_rc(body)
        if not is_neg:
            body('neg.f64 n1, c1;')
        for e in range(1, NE):
            skip = body.label('SKIPM')
            body('and.b32 t32, rc, %d;' % ((NE - 1) & ~e), 'setp.ne.u32 pe, t32, 0;', '@pe bra.uni %s;' % skip)
            neg(e) if is_neg else cmul(e)
            body(skip + ':')
        body('bra TAIL;')
    # ---- diagonal table over the register index: a[e] *= table[e] where e & reg_cmask != 0 (flag: every e) ----
    handler(H['CPH_TABLE'], 'L_TB')
    emit_rc(body)
    body('ld.shared.u8 t32, [cur+7];', 'setp.ne.u32 poff, t32, 0;', 'selp.u32 rc, %d, rc, poff;' % (NE - 1))
    for e in range(NE):
        skip = body.label('SKIPT')
        if e == 0:
            body('@!poff bra.uni %s;' % skip)
        else:
            body('and.b32 t32, rc, %d;' % e, 'setp.eq.u32 pe, t32, 0;', '@pe bra.uni %s;' % skip)
        body('ld.shared.v2.f64 {c0, c1}, [cur+%d];' % (16 + 16 * e), 'neg.f64 n1, c1;')
        cmul(e)
        body(skip + ':')
    emit_long_tail(body, 16 + 16 * NE)
    # ---- dense 2-bit operator on register bits j0 > j1 (operator index = bit(j0) << 1 | bit(j1)) ----
    if has_g2:
        for pi, (j0, j1) in enumerate(PAIRS):
            handler(H['G2'] + pi, 'L_G2%d' % pi)
            emit_on_check(body, 'L_G2T%d' % pi)
            emit_rc(body)
            others = [b for b in range(R) if b not in (j0, j1)]
            for g in range(1 << len(others)):
                eb = sum(((g >> i) & 1) << b for i, b in enumerate(others))
                skip = body.label('SKIPG')
                free = (NE - 1) & ~(1 << j0) & ~(1 << j1) & ~eb
                body('and.b32 t32, rc, %d;' % free, 'setp.ne.u32 pe, t32, 0;', '@pe bra.uni %s;' % skip)
                ids = [eb, eb | (1 << j1), eb | (1 << j0), eb | (1 << j0) | (1 << j1)]
                for r in range(4):
                    for c in range(4):
                        body('ld.shared.v2.f64 {c0, c1}, [cur+%d];' % (16 + 16 * (4 * r + c)), 'neg.f64 n1, c1;')
                        if c == 0:
                            body('mul.f64 o%d, c0, %s;' % (2 * r, re_(ids[c])), 'mul.f64 o%d, c0, %s;' % (2 * r + 1, im_(ids[c])))
                        else:
                            body('fma.rn.f64 o%d, c0, %s, o%d;' % (2 * r, re_(ids[c]), 2 * r),
                                 'fma.rn.f64 o%d, c0, %s, o%d;' % (2 * r + 1, im_(ids[c]), 2 * r + 1))
                        body('fma.rn.f64 o%d, n1, %s, o%d;' % (2 * r, im_(ids[c]), 2 * r),
                             'fma.rn.f64 o%d, c1, %s, o%d;' % (2 * r + 1, re_(ids[c]), 2 * r + 1))
                for r in range(4):
                    body('mov.f64 %s, o%d;' % (re_(ids[r]), 2 * r), 'mov.f64 %s, o%d;' % (im_(ids[r]), 2 * r + 1))
                body(skip + ':')
            body('L_G2T%d:' % pi)
            emit_long_tail(body, 288)
    if has_g2:
        for pi, (j0, j1) in enumerate(PAIRS):
            handler(H['G2X'] + pi, 'L_GX%d' % pi)
            # real X-shaped operator: (a0, a3) and (a1, a2) mix as two real 2x2 blocks (operator index = bit(j0) << 1 |
            # bit(j1)); c0..c7 = m00 m03 m30 m33 m11 m12 m21 m22
            emit_on_check(body, 'L_GXT%d' % pi)
            emit_rc(body)
            for q in range(1, 4):
                body('ld.shared.v2.f64 {c%d, c%d}, [cur+%d];' % (2 * q, 2 * q + 1, 16 + 16 * q))
            others = [b for b in range(R) if b not in (j0, j1)]
            for g in range(1 << len(others)):
                eb = sum(((g >> i) & 1) << b for i, b in enumerate(others))
                skip = body.label('SKIPX')
                free = (NE - 1) & ~(1 << j0) & ~(1 << j1) & ~eb
                body('and.b32 t32, rc, %d;' % free, 'setp.ne.u32 pe, t32, 0;', '@pe bra.uni %s;' % skip)
                ids = [eb, eb | (1 << j1), eb | (1 << j0), eb | (1 << j0) | (1 << j1)]
                for (p, q, base) in ((ids[0], ids[3], 0), (ids[1], ids[2], 4)):
                    for comp in (re_, im_):
                        body('mul.f64 t0, c%d, %s;' % (base + 2, comp(p)),                      # m_qp * a_p
                             'mul.f64 %s, %s, c%d;' % (comp(p), comp(p), base),                 # a_p *= m_pp
                             'fma.rn.f64 %s, c%d, %s, %s;' % (comp(p), base + 1, comp(q), comp(p)),   # += m_pq a_q
                             'fma.rn.f64 %s, %s, c%d, t0;' % (comp(q), comp(q), base + 3))      # a_q = m_qq a_q + t0
                body(skip + ':')
            body('L_GXT%d:' % pi)
            emit_long_tail(body, 80)
    targets[H['END']] = 'L_END'

    E('{',
      '.reg .b32 h, hn, cur, rc, t32;',
      '.reg .b64 cm, tm;',
      '.reg .pred poff, pe;',
      '.reg .f64 c<8>, n<8>, t<6>, o<8>;',
      'ts: .branchtargets %s;' % ', '.join(targets),
      'ld.shared.u32 h, [%s];' % OP,
      'LOOP:',
      # every op record the common handlers use is 32 bytes (header + two doubles); the few long ones (dense
      # matrices) correct the pointer themselves (emit_long_tail), so the loop needs no size decode
      'mov.u32 cur, %s;' % OP,
      'add.u32 %s, %s, 32;' % (OP, OP),
      'ld.shared.u32 hn, [%s];' % OP,              # the next handler id is in flight while the handler runs
      # control mask and the first two payload doubles of THIS op: issued before the jump so that they land
      # while the branch resolves (ops without them read the next record's bytes, harmlessly)
      'ld.shared.u64 cm, [cur+8];',
      'ld.shared.v2.f64 {c0, c1}, [cur+16];',
      'brx.idx h, ts;')
    E(*body.lines)
    E('TAIL:', 'mov.u32 h, hn;', 'bra LOOP;', 'L_END:', '}')
    return E.lines


def render():
    out = ['// GENERATED by gen_oploop.py -- do not edit; see that script for the why and the operand map.']
    for has_g2, name in ((False, 'QFB_OPLOOP_PTX'), (True, 'QFB_OPLOOP_PTX_G2')):
        lines = gen(has_g2)
        out.append('#define %s \\' % name)
        for i, line in enumerate(lines):
            out.append('    "%s\\n\\t"%s' % (line, ' \\' if i + 1 < len(lines) else ''))
        out.append('')
    return '\n'.join(out) + '\n'


if __name__ == '__main__':
    path = os.path.join(HERE, 'qfb_oploop.inc')
    with open(path, 'w') as f:
        f.write(render())
    print(path)
