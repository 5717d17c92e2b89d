"""A CPU emulator for the PTX that csrc/qfb_jit.cu generates (TEST INFRASTRUCTURE ONLY).

The sweep-specialised kernels are straight-line PTX with one CTA-uniform tile loop and a closed set of ~55
instructions (integer address arithmetic, predicates, fp64 multiply / fma / select, 16-byte shared and global
accesses, cp.async copies, one warp ballot). This module executes that subset for every thread of a CTA at once
(registers are numpy arrays with one element per thread) on a numpy state vector, so that the CPU test tier can RUN
the generated code -- not only compile it -- against the oracle (tests/test_jit_emulated.py).

What it models: one CTA at a time, all threads in lock step, instruction by instruction. Asynchronous copies complete
at issue, barriers are no-ops; both are valid schedules of a race-free kernel (the hardware racecheck log is
profiles/r2_sanitizer_racecheck_jit.log), so a result that differs from the oracle is a code-generation bug.
fma.rn.f64 is computed as a * b + c with two roundings (numpy has no fused operation): results agree with the
hardware to a few ulp, far inside the 1e-10 parity bar. The tile loop must be CTA-uniform (asserted); forward
branches that only some threads take (a skipped operator block) reconverge at their target.
"""
import re

import numpy as np

GLOBAL_BASE = 1 << 44          # "device address" of the state vector

_U64 = np.uint64
_U32 = np.uint32
_MASK = {32: 0xFFFFFFFF, 64: 0xFFFFFFFFFFFFFFFF}


class Instr:
    __slots__ = ('pred', 'neg', 'op', 'args', 'line')

    def __init__(self, pred, neg, op, args, line):
        self.pred, self.neg, self.op, self.args, self.line = pred, neg, op, args, line


def _split_args(text: str):
    """'a, [b+4], {c, d}' -> ['a', '[b+4]', '{c, d}']"""
    out, depth, cur = [], 0, ''
    for ch in text:
        if ch in '[{':
            depth += 1
        elif ch in ']}':
            depth -= 1
        if ch == ',' and depth == 0:
            out.append(cur.strip())
            cur = ''
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


class Kernel:
    """Parsed PTX of one qfb_sweep kernel."""

    def __init__(self, ptx: str):
        self.instrs = []
        self.labels = {}
        self.threads = None
        self.const_bytes = 0
        body = False
        for raw in ptx.splitlines():
            line = raw.strip()
            if not line or line.startswith('//'):
                continue
            if not body:
                m = re.match(r'\.maxntid (\d+)', line)
                if m:
                    self.threads = int(m.group(1))
                m = re.match(r'\.const .*qfb_coef\[(\d+)\]', line)
                if m:
                    self.const_bytes = int(m.group(1))
                if line == '{':
                    body = True
                continue
            if line == '}':
                break
            if line.startswith('.reg'):
                continue
            if line.endswith(':'):
                self.labels[line[:-1]] = len(self.instrs)
                continue
            assert line.endswith(';'), raw
            line = line[:-1].strip()
            pred, neg = None, False
            m = re.match(r'@(!?)(%p\d+)\s+(.*)', line)
            if m:
                neg, pred, line = bool(m.group(1)), m.group(2), m.group(3)
            parts = line.split(None, 1)
            op = parts[0]
            args = _split_args(parts[1]) if len(parts) > 1 else []
            self.instrs.append(Instr(pred, neg, op, args, raw))
        assert self.threads, 'no .maxntid'


class _Cta:
    def __init__(self, kernel, state_f64, coef, params, ctaid, nctaid, smem_bytes):
        self.k = kernel
        self.T = kernel.threads
        self.g = state_f64
        self.coef = coef
        self.params = params
        self.smem = np.zeros(max(smem_bytes, 16) // 8 + 2, dtype=np.float64)
        self.regs = {}
        tid = np.arange(self.T, dtype=_U32)
        self.special = {'%tid.x': tid, '%ctaid.x': np.full(self.T, ctaid, dtype=_U32),
                        '%nctaid.x': np.full(self.T, nctaid, dtype=_U32)}
        self.executed = 0

    # ---- operands ----------------------------------------------------------------------------------
    def imm(self, text, bits):
        if text in ('qfb_smem', 'qfb_coef'):
            return (_U32 if bits == 32 else _U64)(0)
        v = int(text, 0)
        return (_U32 if bits == 32 else _U64)(v & _MASK[bits])

    def val(self, text, bits=64, kind='u'):
        """Operand as a numpy array / scalar. kind: 'u' integer of `bits`, 'f' float64, 'p' predicate."""
        if text in self.special:
            return self.special[text]
        if text.startswith('!'):                 # negated predicate operand
            return ~self.regs[text[1:]]
        if text.startswith('%'):
            v = self.regs[text]
            if kind == 'u' and v.dtype == np.float64:       # bit operations on a floating-point register
                return v.view(_U64)
            return v
        if kind == 'f':
            if text.startswith('0d') or text.startswith('0D'):
                return np.array([int(text[2:], 16)], dtype=_U64).view(np.float64)[0]
            return np.float64(float(text))
        return self.imm(text, bits)

    def addr(self, text, bits):
        """'[%r5+128]' / '[%rd7]' / '[qfb_coef+16]' -> byte addresses (array or scalar)."""
        inner = text.strip()[1:-1].strip()
        m = re.match(r'([^+]+?)\s*(?:\+\s*(-?\w+))?$', inner)
        base, off = m.group(1).strip(), m.group(2)
        b = self.val(base, bits)
        if isinstance(b, np.ndarray):
            b = b.astype(np.int64)
        else:
            b = np.int64(int(b))
        return b + (int(off, 0) if off else 0)

    def write(self, name, value, mask):
        old = self.regs.get(name)
        if name.startswith('%fd'):
            dtype = np.float64
        elif name.startswith('%rd'):
            dtype = _U64
        elif name.startswith('%r'):
            dtype = _U32
        else:
            dtype = np.bool_
        value = np.asarray(value)
        if dtype == np.float64 and value.dtype == _U64:
            value = value.view(np.float64)
        elif value.dtype != dtype:
            value = value.astype(dtype)
        if value.ndim == 0:
            value = np.full(self.T, value, dtype=dtype)
        if mask is not None:
            if old is None:
                old = np.zeros(self.T, dtype=dtype)
            value = np.where(mask, value, old)
        self.regs[name] = value

    def vec(self, text):
        return [t.strip() for t in text.strip()[1:-1].split(',')]

    # ---- execution ---------------------------------------------------------------------------------
    def run(self, max_instructions=50_000_000):
        """Threads that take a forward branch the others do not take (the generated code skips an operator block for
        the threads whose control bit is 0) wait at the branch target; the CTA reconverges there. Everything else --
        the tile loop, the prefetch guards -- must be uniform."""
        k = self.k
        pc = 0
        n = len(k.instrs)
        active = np.ones(self.T, dtype=np.bool_)
        everyone = True
        waiting = {}                 # instruction index -> threads parked there
        with np.errstate(over='ignore'):
            while pc < n:
                if waiting and pc in waiting:
                    active = active | waiting.pop(pc)
                    everyone = bool(active.all())
                ins = k.instrs[pc]
                self.executed += 1
                if self.executed > max_instructions:
                    raise RuntimeError('instruction budget exceeded (endless loop?)')
                mask = None if everyone else active
                if ins.pred is not None:
                    pm = self.regs[ins.pred]
                    if ins.neg:
                        pm = ~pm
                    mask = pm if mask is None else (pm & mask)
                op = ins.op
                if op == 'bra':
                    target = k.labels[ins.args[0]]
                    taken = active if mask is None else mask
                    if not taken.any():
                        pc += 1
                        continue
                    if np.array_equal(taken, active):             # every running thread goes
                        if target <= pc:
                            assert everyone and not waiting, 'backward branch inside a divergent region: ' + ins.line
                        pc = target
                        continue
                    assert target > pc, 'divergent backward branch: ' + ins.line
                    waiting[target] = waiting.get(target, np.zeros(self.T, dtype=np.bool_)) | taken
                    active = active & ~taken
                    everyone = False
                    pc += 1
                    continue
                if op == 'ret':
                    gone = active if mask is None else mask
                    assert np.array_equal(gone, active) and everyone and not waiting, 'divergent ret'
                    return
                assert op.split('.')[0] != 'bar' or everyone, 'barrier inside a divergent region'
                self.step(ins, op, mask)
                pc += 1
        assert not waiting

    def step(self, ins, op, mask):
        a = ins.args
        parts = op.split('.')
        head = parts[0]
        if head in ('bar', 'prefetch', 'nanosleep') or op.startswith('cp.async.commit') or \
                op.startswith('cp.async.wait'):
            return
        if head == 'mov':
            t = parts[1]
            if t == 'f64':
                self.write(a[0], self.val(a[1], kind='f'), mask)
            elif t == 'pred':
                self.write(a[0], self.val(a[1], kind='p'), mask)
            else:
                bits = 32 if t.endswith('32') else 64
                self.write(a[0], self.val(a[1], bits), mask)
            return
        if head == 'ld' and parts[1] == 'param':
            self.write(a[0], _U64(self.params[a[1].strip('[]')] & _MASK[64]), mask)
            return
        if head == 'cvta':
            self.write(a[0], self.val(a[1]), mask)
            return
        if head == 'cvt':
            src = self.val(a[1], 32 if parts[2] == 'u32' else 64)
            self.write(a[0], np.asarray(src).astype(_U32 if parts[1] == 'u32' else _U64), mask)
            return
        if head in ('and', 'or', 'xor', 'not') and parts[1] == 'pred':
            x = self.val(a[1], kind='p')
            if head == 'not':
                self.write(a[0], ~x, mask)
                return
            y = self.val(a[2], kind='p')
            self.write(a[0], (x & y) if head == 'and' else (x | y) if head == 'or' else (x ^ y), mask)
            return
        if head in ('and', 'or', 'xor', 'add', 'sub', 'shl', 'shr', 'min', 'rem') and parts[1] != 'f64' and \
                parts[-1] in ('b32', 'b64', 'u32', 'u64', 's64', 's32'):
            bits = 32 if parts[-1].endswith('32') else 64
            dt = _U32 if bits == 32 else _U64
            x = np.asarray(self.val(a[1], bits)).astype(dt)
            if head in ('shl', 'shr'):
                sh = self.val(a[2], 32)
                if not isinstance(sh, np.ndarray):
                    sh = int(sh)
                    if sh >= bits:
                        self.write(a[0], np.zeros(self.T, dtype=dt), mask)
                        return
                    sh = dt(sh)
                else:
                    sh = sh.astype(dt)
                self.write(a[0], (x << sh) if head == 'shl' else (x >> sh), mask)
                return
            y = np.asarray(self.val(a[2], bits)).astype(dt)
            if head == 'and':
                r = x & y
            elif head == 'or':
                r = x | y
            elif head == 'xor':
                r = x ^ y
            elif head == 'add':
                r = x + y
            elif head == 'sub':
                r = x - y
            elif head == 'min':
                r = np.minimum(x, y)
            else:
                r = x % y
            self.write(a[0], r, mask)
            return
        if head == 'neg' and parts[1] == 's64':
            x = np.asarray(self.val(a[1], 64)).astype(_U64)
            self.write(a[0], (~x) + _U64(1), mask)
            return
        if head == 'mul' and parts[1] == 'lo':
            x = np.asarray(self.val(a[1], 32)).astype(_U32)
            y = np.asarray(self.val(a[2], 32)).astype(_U32)
            self.write(a[0], x * y, mask)
            return
        if head == 'mul' and parts[1] == 'wide':
            x = np.asarray(self.val(a[1], 32)).astype(_U64)
            y = np.asarray(self.val(a[2], 32)).astype(_U64)
            self.write(a[0], x * y, mask)
            return
        if head == 'mad' and parts[1] == 'lo':
            x = np.asarray(self.val(a[1], 32)).astype(_U32)
            y = np.asarray(self.val(a[2], 32)).astype(_U32)
            z = np.asarray(self.val(a[3], 32)).astype(_U32)
            self.write(a[0], x * y + z, mask)
            return
        if head == 'setp':
            cmp_, t = parts[1], parts[2]
            bits = 32 if t.endswith('32') else 64
            dt = _U32 if bits == 32 else _U64
            x = np.asarray(self.val(a[1], bits)).astype(dt)
            y = np.asarray(self.val(a[2], bits)).astype(dt)
            r = {'ne': x != y, 'eq': x == y, 'lt': x < y, 'ge': x >= y, 'le': x <= y, 'gt': x > y}[cmp_]
            self.write(a[0], r, mask)
            return
        if head == 'selp':
            t = parts[1]
            p = self.val(a[3], kind='p')
            if t == 'f64':
                x, y = self.val(a[1], kind='f'), self.val(a[2], kind='f')
            else:
                bits = 32 if t.endswith('32') else 64
                x, y = self.val(a[1], bits), self.val(a[2], bits)
            self.write(a[0], np.where(p, x, y), mask)
            return
        if op == 'neg.f64':
            self.write(a[0], -self.val(a[1], kind='f'), mask)
            return
        if op == 'mul.f64':
            self.write(a[0], self.val(a[1], kind='f') * self.val(a[2], kind='f'), mask)
            return
        if op in ('add.f64', 'sub.f64'):
            x, y = self.val(a[1], kind='f'), self.val(a[2], kind='f')
            self.write(a[0], x + y if op == 'add.f64' else x - y, mask)
            return
        if op == 'fma.rn.f64':
            self.write(a[0], self.val(a[1], kind='f') * self.val(a[2], kind='f') + self.val(a[3], kind='f'), mask)
            return
        if op == 'ld.const.f64':
            idx = np.asarray(self.addr(a[1], 64))
            assert (idx % 8 == 0).all() and (idx >= 0).all() and (idx < 8 * max(1, self.coef.size)).all(), ins.line
            self.write(a[0], self.coef[idx >> 3], mask)
            return
        if op == 'ld.shared.v2.f64':
            idx = np.broadcast_to(self.addr(a[1], 32), (self.T,))
            assert (idx % 16 == 0).all(), ins.line
            d0, d1 = self.vec(a[0])
            self.write(d0, self.smem[idx >> 3], mask)
            self.write(d1, self.smem[(idx >> 3) + 1], mask)
            return
        if op == 'ld.shared.u32':
            idx = np.broadcast_to(self.addr(a[1], 32), (self.T,))
            self.write(a[0], self.smem.view(_U32)[idx >> 2], mask)
            return
        if op == 'st.shared.v2.f64':
            idx = np.broadcast_to(self.addr(a[0], 32), (self.T,))
            assert (idx % 16 == 0).all(), ins.line
            s0, s1 = self.vec(a[1])
            v0 = np.broadcast_to(self.val(s0, kind='f'), (self.T,))
            v1 = np.broadcast_to(self.val(s1, kind='f'), (self.T,))
            sel = slice(None) if mask is None else mask
            self.smem[(idx >> 3)[sel]] = v0[sel]
            self.smem[(idx >> 3)[sel] + 1] = v1[sel]
            return
        if op.startswith('cp.async.c'):
            sidx = np.broadcast_to(self.addr(a[0], 32), (self.T,))
            gidx = self.gindex(np.broadcast_to(self.addr(a[1], 64), (self.T,)), ins)
            assert int(a[2]) == 16 and (sidx % 16 == 0).all()
            sel = slice(None) if mask is None else mask
            self.smem[(sidx >> 3)[sel]] = self.g[gidx[sel]]
            self.smem[(sidx >> 3)[sel] + 1] = self.g[gidx[sel] + 1]
            return
        if op.startswith('ld.global') and op.endswith('v2.f64'):
            gidx = self.gindex(np.broadcast_to(self.addr(a[1], 64), (self.T,)), ins)
            d0, d1 = self.vec(a[0])
            self.write(d0, self.g[gidx], mask)
            self.write(d1, self.g[gidx + 1], mask)
            return
        if op.startswith('st.global') and op.endswith('v2.f64'):
            gidx = self.gindex(np.broadcast_to(self.addr(a[0], 64), (self.T,)), ins)
            s0, s1 = self.vec(a[1])
            v0 = np.broadcast_to(self.val(s0, kind='f'), (self.T,))
            v1 = np.broadcast_to(self.val(s1, kind='f'), (self.T,))
            sel = slice(None) if mask is None else mask
            self.g[gidx[sel]] = v0[sel]
            self.g[gidx[sel] + 1] = v1[sel]
            return
        if op == 'vote.sync.ballot.b32':
            p = np.broadcast_to(self.val(a[1], kind='p'), (self.T,))
            out = np.zeros(self.T, dtype=_U32)
            for w in range(0, self.T, 32):
                bits = 0
                for lane, v in enumerate(p[w:w + 32]):
                    bits |= int(bool(v)) << lane
                out[w:w + 32] = bits
            self.write(a[0], out, mask)
            return
        raise NotImplementedError('PTX emulator: ' + ins.line)

    def gindex(self, addr, ins):
        off = addr - GLOBAL_BASE
        assert (off % 16 == 0).all() and (off >= 0).all() and (off < 8 * self.g.size).all(), \
            'global access outside the state: ' + ins.line
        return off >> 3


class SparseState:
    """The memory of a state that is too large to allocate (the 30-qubit benchmark: 16 GiB): amplitudes that were
    never written read as 0. Indexed like the float64 view of the state vector (amplitude k = elements 2k, 2k+1), which
    is what the emulator's loads and stores use; `amplitude` / `set_amplitude` for the test's side."""

    def __init__(self, nbits: int):
        self.nbits = nbits
        self.size = 2 << nbits
        self.data = {}

    def __getitem__(self, idx):
        get = self.data.get
        return np.fromiter((get(int(i), 0.0) for i in np.asarray(idx).reshape(-1)), dtype=np.float64,
                           count=np.asarray(idx).size)

    def __setitem__(self, idx, values):
        for i, v in zip(np.asarray(idx).reshape(-1).tolist(), np.asarray(values).reshape(-1).tolist()):
            self.data[i] = v

    def amplitude(self, k: int) -> complex:
        return complex(self.data.get(2 * k, 0.0), self.data.get(2 * k + 1, 0.0))

    def set_amplitude(self, k: int, value: complex) -> None:
        self.data[2 * k], self.data[2 * k + 1] = float(value.real), float(value.imag)


def run_sweep(ptx: str, coef: np.ndarray, state, index_hi: int = 0, fix_value: int = 0, grid: int = 2,
              smem_bytes: int = 1 << 16, groups: int = 1, ctas=None) -> int:
    """Execute one launch of a generated sweep kernel in place on `state` (complex128 vector of 2^nbits amplitudes, or a
    SparseState) with `grid` CTAs, one after the other; `ctas`: run only these CTA indices of the grid (a few tiles of
    a state too large to walk). index_hi: the rank bits of a sharded state (kernel parameter p_hi = index_hi << nbits);
    fix_value: the fixed index bits of a slice launch (p_fix). Returns the instructions executed."""
    if isinstance(state, SparseState):
        nbits, g = state.nbits, state
    else:
        assert state.dtype == np.complex128 and state.flags['C_CONTIGUOUS']
        nbits = int(state.size).bit_length() - 1
        g = state.view(np.float64)
    kernel = Kernel(ptx)
    params = {'p_state': GLOBAL_BASE, 'p_hi': int(index_hi) << nbits, 'p_zero': 0, 'p_fix': int(fix_value)}
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    if coef.size == 0:
        coef = np.zeros(2)
    total = 0
    for cta in (range(grid) if ctas is None else ctas):
        c = _Cta(kernel, g, coef, params, int(cta), grid, smem_bytes * max(1, groups))
        c.run()
        total += c.executed
    return total
