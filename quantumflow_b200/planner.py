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
PLAN_VERSION = 2
REG_BITS = 4
MAX_TILE_BITS = 13
MIN_TILE_BITS = 5
MAX_HOLES = 48
MAX_SWEEP_BYTES = 40 * 1024
MAX_DIAG_BITS = 6
OP_G1, OP_G2, OP_D = 1, 2, 3

DEFAULT_TILE_BITS = 12
DEFAULT_LOW_BITS = 3
# cost units ~ FP64 work per amplitude relative to a dense 1-bit operator
COST = {'G1': 1.0, 'G1_cheap': 0.5, 'G1_swap': 0.15, 'G2': 2.5, 'D': 0.5}
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


def classify_op(mat: np.ndarray, bits: Sequence[int], gate_index: int = -1):
    """Turn (matrix, bits) into a POp, a Fallback, or None (identity)."""
    k = len(bits)
    mat = classify.as_matrix(mat, k)
    if classify.is_identity(mat):
        return None
    if classify.is_diagonal(mat):
        if k > MAX_DIAG_BITS:
            return Fallback(mat, bits)
        # drop bits the table does not depend on
        table = np.ascontiguousarray(np.diagonal(mat))
        keep = []
        t = table.reshape([2] * k)
        for q in range(k):
            a, b = np.take(t, 0, axis=q), np.take(t, 1, axis=q)
            if not np.array_equal(a, b):
                keep.append(q)
        if len(keep) < k:
            idx = tuple(slice(None) if q in keep else 0 for q in range(k))
            table = np.ascontiguousarray(t[idx]).reshape(-1)
            bits = [bits[q] for q in keep]
            if not keep:   # global phase
                return POp('D', dbits=(), mat=table.reshape(1), cost=COST['D'], gate_index=gate_index)
        return POp('D', dbits=bits, mat=table, cost=COST['D'], gate_index=gate_index)
    controls, targets, reduced = ([], list(range(k)), mat) if k == 1 else classify.peel_controls(mat, k)
    if len(targets) > 2:
        return Fallback(mat, bits)
    cbits = [bits[q] for q in controls]
    tbits = [bits[q] for q in targets]
    if len(targets) == 1:
        kind = classify.g1_kind(reduced)
        if classify.is_diagonal(reduced):
            # controlled phase that was not caught as fully diagonal cannot happen (diag checked first)
            pass
        cost = COST['G1_swap'] if kind == 3 else (COST['G1_cheap'] if kind in (1, 2, 4) else COST['G1'])
        return POp('G', mix=tbits, ctrl=cbits, mat=reduced, cost=cost, gate_index=gate_index)
    nnz = int(np.count_nonzero(reduced))
    return POp('G', mix=tbits, ctrl=cbits, mat=reduced, cost=COST['G2'] * max(nnz, 4) / 16.0 + 0.3,
               gate_index=gate_index)


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
            opbytes = 16 + (64 if (op.kind == 'G' and len(op.mix) == 1) else
                            272 if op.kind == 'G' else 16 + (32 << max(1, len(op.dbits))))
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
            sweep = SweepPlan(tile, chosen)
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

        if op.kind == 'D':
            nb = len(op.dbits)
            if nb == 0:   # global phase: a 1-bit table on bit 0 with equal entries
                table = np.array([op.mat[0], op.mat[0]], dtype=np.complex128)
                dbits = (0,)
                nb = 1
            else:
                table = np.asarray(op.mat, dtype=np.complex128)
                dbits = op.dbits
            pos = [0xFF] * 8
            econ = [0] * 4
            for q, bit in enumerate(dbits):
                ri = reg_index(bit)
                if ri is None:
                    pos[q] = bit
                else:
                    econ[ri] = 1 << (nb - 1 - q)
            payload = struct.pack('<8B4B4x', *pos, *econ) + table.tobytes()
            header = struct.pack('<BBBBBBHQ', OP_D, 0, 0, 0, 0, nb, 16 + len(payload), 0)
            return header + payload
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
            mat = np.ascontiguousarray(op.mat, dtype=np.complex128).reshape(2, 2)
            kind = classify.g1_kind(mat)
            payload = mat.tobytes()
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
                ops_blob = b''.join(self._emit_op(op, sweep, regs) for op in ops)
                nops += len(ops)
                thrpad = list(thr) + [0] * (12 - len(thr))
                rounds_blob += struct.pack('<II4B12B8x', len(ops), 32 + len(ops_blob), *regs, *thrpad) + ops_blob
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
            segments.append(Segment('plan', blob=blob, nsweeps=len(sweeps), nops=len(pending),
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
            pending.append(item)
    flush()
    return segments


def plan_stats(segments: Sequence[Segment]) -> Dict[str, int]:
    return {'segments': len(segments), 'sweeps': sum(s.nsweeps for s in segments),
            'ops': sum(s.nops for s in segments), 'rounds': sum(s.nrounds for s in segments),
            'plan_bytes': sum(len(s.blob) for s in segments if s.blob)}
