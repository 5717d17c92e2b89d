"""Host planner for the tiled multi-gate executor (the insertion point is Circuit.run / Circuit.evolve,
reference loop quantumflow/circuits.py:87-109).

Input  : a list of bit-level operators (matrix, index-bit positions) in program order.
Output : segments, each either a binary plan (csrc/qfb_plan.h) executed by qfb_plan_launch, or a single operator
         that the executor cannot express (>2 mixing bits) and that goes through qfb_apply_dense.

Three greedy passes, all order preserving up to commutation:

1. classify   every operator becomes D (diagonal over any bits), or G (1 or 2 mixing bits + any number of
              control bits, controls peeled by classify.peel_controls). Diagonal uses and control uses of a bit
              commute with each other, so they never constrain tiling.
2. sweeps     walk the list; an operator joins the current sweep when it does not conflict with a deferred
              operator and its mixing bits fit into the tile (M bits, the L lowest index bits are always members
              so that global accesses are whole 128-byte lines). A cost cap keeps a sweep HBM-bound.
3. rounds     inside a sweep the same walk assigns operators to rounds of R=4 register bits; the first and
              last round keep the low tile bits on the lanes (coalesced LDG/STG).
"""
import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import classify

PLAN_MAGIC = 0x50424651
PLAN_VERSION = 3
REG_BITS = 4
MAX_TILE_BITS = 13
MIN_TILE_BITS = 5
MAX_HOLES = 48
MAX_SWEEP_BYTES = 40 * 1024
MAX_DIAG_BITS = 5
OP_G1, OP_G2, OP_CPH = 1, 2, 3

DEFAULT_TILE_BITS = 12
DEFAULT_LOW_BITS = 3
# cost units ~ FP64 work per amplitude relative to a dense 1-bit operator
COST = {'G1': 1.0, 'G1_cheap': 0.5, 'G1_swap': 0.2, 'G2': 2.5, 'P': 0.15}
DEFAULT_MAX_COST = 28.0


class POp:
    """A classified operator."""
    __slots__ = ('kind', 'mix', 'ctrl', 'dbits', 'mat', 'cost', 'mixset', 'diagset', 'anyset', 'gate_index')

    def __init__(self, kind, mix=(), ctrl=(), dbits=(), mat=None, cost=1.0, gate_index=-1):
        self.kind = kind
        self.mix = tuple(int(b) for b in mix)
        self.ctrl = tuple(int(b) for b in ctrl)
        self.dbits = tuple(int(b) for b in dbits)
        self.mat = mat
        self.cost = cost
        self.mixset = frozenset(self.mix)
        self.diagset = frozenset(self.ctrl) | frozenset(self.dbits)
        self.anyset = self.mixset | self.diagset
        self.gate_index = gate_index


class Fallback:
    """An operator executed by the one-gate kernels (qfb_apply_dense)."""
    __slots__ = ('mat', 'bits')

    def __init__(self, mat, bits):
        self.mat = mat
        self.bits = tuple(int(b) for b in bits)


def phase_polynomial(table: np.ndarray, k: int):
    """Diagonal operator -> list of (qubit subset mask over the k table bits, factor) with
    d[s] = prod over subsets T of s of factor[T]; bit q of the table index has weight 1 << (k-1-q).
    Requires every entry to be non-zero."""
    phi = {}
    for s in range(1 << k):
        val = complex(table[s])
        for t in range(s):
            if (t & s) == t and t in phi:
                val = val / phi[t]
        phi[s] = val
    return phi


def classify_op(mat: np.ndarray, bits: Sequence[int], gate_index: int = -1):
    """Turn (matrix, bits) into a list of POps, a Fallback, or None (identity)."""
    k = len(bits)
    mat = classify.as_matrix(mat, k)
    if classify.is_identity(mat):
        return None
    if classify.is_diagonal(mat):
        table = np.ascontiguousarray(np.diagonal(mat))
        if np.all(table != 0) and k <= MAX_DIAG_BITS:
            terms = []
            for sub, factor in phase_polynomial(table, k).items():
                if factor == 1:
                    continue
                tbits = [bits[q] for q in range(k) if (sub >> (k - 1 - q)) & 1]
                terms.append(POp('P', dbits=tbits, mat=complex(factor), cost=COST['P'], gate_index=gate_index))
            return terms or None
        if k > 2:
            return Fallback(mat, bits)
        # singular diagonal (projectors): a dense 1- or 2-bit operator
        cost = COST['G1_cheap'] if k == 1 else COST['G2'] * 0.25 + 0.3
        return [POp('G', mix=bits, ctrl=(), mat=mat, cost=cost, gate_index=gate_index)]
    controls, targets, reduced = ([], list(range(k)), mat) if k == 1 else classify.peel_controls(mat, k)
    if len(targets) > 2:
        return Fallback(mat, bits)
    cbits = [bits[q] for q in controls]
    tbits = [bits[q] for q in targets]
    if len(targets) == 1:
        kind = g1_kind(reduced, bool(cbits))
        cost = COST['G1_swap'] if kind == 3 else (COST['G1_cheap'] if kind in (1, 2, 4, 5) else COST['G1'])
        return [POp('G', mix=tbits, ctrl=cbits, mat=reduced, cost=cost, gate_index=gate_index)]
    nnz = int(np.count_nonzero(reduced))
    return [POp('G', mix=tbits, ctrl=cbits, mat=reduced, cost=COST['G2'] * max(nnz, 4) / 16.0 + 0.3,
                gate_index=gate_index)]


def g1_kind(m: np.ndarray, controlled: bool) -> int:
    """QFB_G1_* kind of a 2x2 operator; controlled operators only use SWAPX / GENERAL (fewer kernel variants)."""
    kind = classify.g1_kind(m)
    m = np.asarray(m).reshape(2, 2)
    if kind == 1 and np.all(m.real != 0) and len({abs(v) for v in m.real.reshape(-1)}) == 1:
        kind = 5   # HLIKE
    if controlled and kind not in (0, 3):
        kind = 0
    return kind


def merge_phase_terms(ops: List[POp]) -> List[POp]:
    """Multiply together phase terms with the same bit mask when no mixing operator on those bits sits between
    them (phase terms commute with each other and with controls). Terms that become 1 are dropped."""
    out: List[POp] = []
    open_terms: Dict[frozenset, int] = {}
    for op in ops:
        if op.kind == 'P':
            key = op.diagset
            pos = open_terms.get(key)
            if pos is not None:
                prev = out[pos]
                out[pos] = POp('P', dbits=prev.dbits, mat=prev.mat * op.mat, cost=prev.cost,
                               gate_index=prev.gate_index)
            else:
                open_terms[key] = len(out)
                out.append(op)
        else:
            for key in [k for k in open_terms if k & op.mixset]:
                del open_terms[key]
            out.append(op)
    return [op for op in out if not (op.kind == 'P' and op.mat == 1)]


def _conflicts(op: POp, def_any: set, def_mix: set) -> bool:
    """Does `op` fail to commute with some deferred operator?"""
    return bool(op.mixset & def_any) or bool(op.diagset & def_mix)


def swizzle_class(pos: int) -> int:
    return pos if pos < 3 else (pos - 3) % 3


class SweepPlan:
    __slots__ = ('tile', 'ops', 'rounds', 'cost')

    def __init__(self, tile: List[int], ops: List[POp]):
        self.tile = tile      # index-bit positions, ascending, length M
        self.ops = ops
        self.rounds = []      # list of (regpos[4], thrpos[M-4], [POp])
        self.cost = sum(o.cost for o in ops)


class Planner:
    def __init__(self, nbits: int, tile_bits: int = None, low_bits: int = None, max_cost: float = None):
        self.nbits = int(nbits)
        m = DEFAULT_TILE_BITS if tile_bits is None else int(tile_bits)
        m = min(m, self.nbits, MAX_TILE_BITS)
        if m < MIN_TILE_BITS:
            raise ValueError('state too small for the tiled executor (need >= {} bits)'.format(MIN_TILE_BITS))
        if self.nbits - m > MAX_HOLES:
            raise ValueError('state too large for one plan')
        self.M = m
        low = DEFAULT_LOW_BITS if low_bits is None else int(low_bits)
        self.L = max(0, min(low, m - REG_BITS))
        self.max_cost = DEFAULT_MAX_COST if max_cost is None else float(max_cost)

    # ---- pass 2: sweeps ---------------------------------------------------------------------------
    def _form_sweep(self, ops: List[POp]) -> Tuple[List[POp], List[POp], List[int]]:
        tile = set(range(self.L))
        chosen: List[POp] = []
        deferred: List[POp] = []
        def_any: set = set()
        def_mix: set = set()
        cost = 0.0
        nbytes = 80
        full = False
        for op in ops:
            ok = not full and not _conflicts(op, def_any, def_mix)
            if ok and op.kind == 'G':
                if any(b >= self.nbits for b in op.mix):
                    raise ValueError('operator mixes bit {} outside the local index; remap first'.format(
                        max(op.mix)))
                need = op.mixset - tile
                if len(tile) + len(need) > self.M:
                    ok = False
            if ok and cost + op.cost > self.max_cost and chosen:
                ok = False
            opbytes = 16 + (64 if (op.kind == 'G' and len(op.mix) == 1) else 272 if op.kind == 'G' else 16)
            if ok and nbytes + opbytes + 32 * 8 > MAX_SWEEP_BYTES:
                ok = False
                full = True
            if ok:
                chosen.append(op)
                cost += op.cost
                nbytes += opbytes
                if op.kind == 'G':
                    tile |= op.mixset
            else:
                deferred.append(op)
                def_any |= op.anyset
                def_mix |= op.mixset
        # pad the tile with the lowest free bits (locality of the strided tile accesses)
        b = 0
        while len(tile) < self.M:
            if b not in tile:
                tile.add(b)
            b += 1
        return chosen, deferred, sorted(tile)

    # ---- pass 3: rounds ---------------------------------------------------------------------------
    def _thread_order(self, regs: Sequence[int], coalesced: bool) -> List[int]:
        free = [p for p in range(self.M) if p not in regs]
        if coalesced:
            return free  # ascending: lanes cover tile positions 0..L-1 = index bits 0..L-1
        head: List[int] = []
        for cls in range(3):
            for p in free:
                if swizzle_class(p) == cls and p not in head:
                    head.append(p)
                    break
        rest = [p for p in free if p not in head]
        return head + rest

    def _form_rounds(self, sweep: SweepPlan) -> None:
        pos_of = {b: j for j, b in enumerate(sweep.tile)}
        remaining = list(sweep.ops)
        rounds: List[Tuple[List[int], List[POp]]] = []
        first = True
        while remaining:
            regs: List[int] = []
            chosen: List[POp] = []
            deferred: List[POp] = []
            def_any: set = set()
            def_mix: set = set()
            for op in remaining:
                ok = not _conflicts(op, def_any, def_mix)
                if ok and op.kind == 'G':
                    need = [pos_of[b] for b in op.mix if pos_of[b] not in regs]
                    if first and any(pos_of[b] < self.L for b in op.mix):
                        ok = False
                    elif len(regs) + len(need) > REG_BITS:
                        ok = False
                    else:
                        regs += need
                if ok:
                    chosen.append(op)
                else:
                    deferred.append(op)
                    def_any |= op.anyset
                    def_mix |= op.mixset
            if not chosen and first:
                # nothing can run with the low bits on the lanes; open a pure load round
                pass
            rounds.append((regs, chosen))
            remaining = deferred
            first = False
        if not rounds:
            rounds.append(([], []))
        # the last round stores to HBM: its register bits must avoid the low tile positions
        if any(p < self.L for p in rounds[-1][0]):
            rounds.append(([], []))
        # drop an empty first round when the sweep has another round that can serve as the load round
        if len(rounds) > 1 and not rounds[0][1] and not any(p < self.L for p in rounds[1][0]):
            rounds.pop(0)
        final = []
        nr = len(rounds)
        for r, (regs, chosen) in enumerate(rounds):
            edge = (r == 0) or (r == nr - 1)
            regs = list(regs)
            # fill up to R register bits with high free positions (never low ones on edge rounds)
            cand = [p for p in range(self.M - 1, -1, -1) if p not in regs and (p >= self.L or not edge)]
            # prefer a spread of swizzle classes so that lane bits always find three distinct classes
            while len(regs) < REG_BITS:
                counts = {c: sum(1 for p in regs if swizzle_class(p) == c) for c in range(3)}
                cand.sort(key=lambda p: (counts[swizzle_class(p)], -p))
                regs.append(cand.pop(0))
            regs.sort()
            final.append((regs, self._thread_order(regs, edge), chosen))
        sweep.rounds = final

    # ---- driver -----------------------------------------------------------------------------------
    def plan(self, pops: List[POp]) -> List[SweepPlan]:
        sweeps: List[SweepPlan] = []
        remaining = list(pops)
        while remaining:
            chosen, remaining, tile = self._form_sweep(remaining)
            if not chosen:
                raise RuntimeError('planner made no progress')
            sweep = SweepPlan(tile, merge_phase_terms(chosen))
            self._form_rounds(sweep)
            sweeps.append(sweep)
        return sweeps

    # ---- serialisation ----------------------------------------------------------------------------
    def _emit_op(self, op: POp, sweep: SweepPlan, regs: Sequence[int]) -> bytes:
        pos_of = {b: j for j, b in enumerate(sweep.tile)}
        reg_of = {p: i for i, p in enumerate(regs)}   # tile position -> register bit

        def reg_index(bit: int) -> Optional[int]:
            p = pos_of.get(bit)
            return reg_of.get(p) if p is not None else None

        if op.kind == 'P':
            reg_cmask = 0
            idx_cmask = 0
            for bit in op.dbits:
                ri = reg_index(bit)
                if ri is None:
                    idx_cmask |= 1 << bit
                else:
                    reg_cmask |= 1 << ri
            factor = complex(op.mat)
            kind = 1 if (factor == -1 and reg_cmask != 0) else 0
            payload = struct.pack('<dd', factor.real, factor.imag)
            return struct.pack('<BBBBBBHQ', OP_CPH, kind, 0, 0, reg_cmask, 0, 16 + len(payload), idx_cmask) + payload
        reg_cmask = 0
        idx_cmask = 0
        for c in op.ctrl:
            ri = reg_index(c)
            if ri is None:
                idx_cmask |= 1 << c
            else:
                reg_cmask |= 1 << ri
        if len(op.mix) == 1:
            j0 = reg_index(op.mix[0])
            mat = np.array(op.mat, dtype=np.complex128).reshape(2, 2)
            kind = g1_kind(mat, bool(op.ctrl))
            if kind == 5:   # HLIKE: out0 = h0 (x + r0 y), out1 = h1 (x + r1 y); ratios ride in the imaginary slots
                h0, h1 = mat[0, 0].real, mat[1, 0].real
                mat[0, 0] = complex(h0, mat[0, 1].real / h0)
                mat[1, 0] = complex(h1, mat[1, 1].real / h1)
            payload = np.ascontiguousarray(mat).tobytes()
            header = struct.pack('<BBBBBBHQ', OP_G1, kind, j0, 0, reg_cmask, 0, 16 + len(payload), idx_cmask)
            return header + payload
        j0, j1 = reg_index(op.mix[0]), reg_index(op.mix[1])
        mat = np.ascontiguousarray(op.mat, dtype=np.complex128).reshape(2, 2, 2, 2)
        if j0 < j1:   # kernel wants the operator's MSB qubit on the higher register bit
            mat = mat.transpose(1, 0, 3, 2)
            j0, j1 = j1, j0
        mat = np.ascontiguousarray(mat).reshape(4, 4)
        nz = 0
        for r in range(4):
            for c in range(4):
                if mat[r, c] != 0:
                    nz |= 1 << (4 * r + c)
        payload = mat.tobytes() + struct.pack('<I12x', nz)
        header = struct.pack('<BBBBBBHQ', OP_G2, 0, j0, j1, reg_cmask, 0, 16 + len(payload), idx_cmask)
        return header + payload

    def serialise(self, sweeps: List[SweepPlan]) -> bytes:
        body = b''
        for sweep in sweeps:
            rounds_blob = b''
            nops = 0
            for regs, thr, ops in sweep.rounds:
                emitted = [self._emit_op(op, sweep, regs) for op in ops]
                ops_blob = b''.join(emitted)
                nops += len(ops)
                has_scalar = int(any(e[0] == OP_CPH and e[4] == 0 for e in emitted))
                thrpad = list(thr) + [0] * (12 - len(thr))
                rounds_blob += struct.pack('<II4B12BB7x', len(ops), 32 + len(ops_blob), *regs, *thrpad,
                                           has_scalar) + ops_blob
            holes = [b for b in range(self.nbits) if b not in sweep.tile]
            gpos = list(sweep.tile) + [0] * (16 - len(sweep.tile))
            hole = holes + [0] * (MAX_HOLES - len(holes))
            size = 80 + len(rounds_blob)
            if size > MAX_SWEEP_BYTES:
                raise RuntimeError('sweep record too large ({} bytes)'.format(size))
            body += struct.pack('<IIII16B48B', size, len(sweep.rounds), nops, 0, *gpos, *hole) + rounds_blob
        total = 32 + len(body)
        header = struct.pack('<IIIIIIQ', PLAN_MAGIC, PLAN_VERSION, self.nbits, self.M, REG_BITS, len(sweeps), total)
        return header + body


class Segment:
    """One executable piece: a plan blob (kind 'plan') or a single operator (kind 'op')."""
    __slots__ = ('kind', 'blob', 'mat', 'bits', 'nsweeps', 'nops', 'nrounds', 'uploaded')

    def __init__(self, kind, blob=None, mat=None, bits=None, nsweeps=0, nops=0, nrounds=0):
        self.kind = kind
        self.blob = blob
        self.mat = mat
        self.bits = bits
        self.nsweeps = nsweeps
        self.nops = nops
        self.nrounds = nrounds
        self.uploaded = None


def build_segments(nbits: int, bitops: Sequence[Tuple[np.ndarray, Sequence[int]]], tile_bits: int = None,
                   low_bits: int = None, max_cost: float = None) -> List[Segment]:
    """Plan a list of (matrix, bits) operators for a state with `nbits` local index bits."""
    planner = Planner(nbits, tile_bits, low_bits, max_cost)
    segments: List[Segment] = []
    pending: List[POp] = []

    def flush():
        if pending:
            sweeps = planner.plan(pending)
            blob = planner.serialise(sweeps)
            segments.append(Segment('plan', blob=blob, nsweeps=len(sweeps),
                                    nops=len({op.gate_index for op in pending}),
                                    nrounds=sum(len(s.rounds) for s in sweeps)))
            pending.clear()

    for gi, (mat, bits) in enumerate(bitops):
        item = classify_op(np.asarray(mat), list(bits), gi)
        if item is None:
            continue
        if isinstance(item, Fallback):
            flush()
            segments.append(Segment('op', mat=item.mat, bits=item.bits, nsweeps=1, nops=1))
        else:
            pending.extend(item)
    flush()
    return segments


def plan_stats(segments: Sequence[Segment]) -> Dict[str, int]:
    return {'segments': len(segments), 'sweeps': sum(s.nsweeps for s in segments),
            'ops': sum(s.nops for s in segments), 'rounds': sum(s.nrounds for s in segments),
            'plan_bytes': sum(len(s.blob) for s in segments if s.blob)}
