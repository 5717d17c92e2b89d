"""Host planner for the tiled multi-gate executor (the insertion point is Circuit.run / Circuit.evolve,
reference loop quantumflow/circuits.py:87-109).

Input  : a list of bit-level operators (matrix, index-bit positions) in program order.
Output : segments, each either a binary plan (csrc/qfb_plan.h) executed by qfb_plan_launch, or a single operator
         that the executor cannot express (>2 mixing bits) and that goes through qfb_apply_dense.

Passes, all order preserving up to commutation:

1. classify   a diagonal operator becomes phase-polynomial terms P(mask, factor): multiply by `factor` where all
              mask bits are 1 (any bits of the full index, so diagonal gates never constrain tiling; terms with
              factor 1 vanish, e.g. CZ is the single term P({a,b}, -1)). Everything else becomes G: 1 or 2
              mixing bits + any number of control bits (controls peeled by classify.peel_controls). Diagonal uses
              and control uses of a bit commute with each other.
2. sweeps     walk the list; an operator joins the current sweep when it does not conflict with a deferred
              operator and its mixing bits fit into the tile (M bits, the L lowest index bits are always members
              so that global accesses are whole 128-byte lines). A cost cap keeps a sweep close to HBM-bound.
3. rounds     inside a sweep the same walk assigns operators to rounds of R=4 register bits; the first and
              last round keep the low tile bits on the lanes (coalesced LDG/STG). Phase terms are then moved,
              inside their commutation window, to the round where they are cheapest: a term whose bits are all
              thread-level is a per-thread scalar (4 FP64 ops instead of up to 64).
4. encode     1-bit operators are divided by their (0,0) entry when that exposes a cheaper form (Hadamard:
              sums only; RX / RY: two fused multiply-adds per amplitude); the pivots of a sweep are multiplied
              into one uniform scalar that is applied once.
"""
import os
import random
import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import classify

PLAN_MAGIC = 0x50424651
PLAN_VERSION = 13
REG_BITS = 5                # register bits of the interpreter's plans; sweep-specialised kernels also take 4
HANDLER_STRIDE = 5          # handler ids are laid out for 5 register bits whatever a plan uses
TABLE_ENTRIES = 32
MAX_TILE_BITS = 13
MIN_TILE_BITS = 6
MAX_HOLES = 48
MAX_SWEEP_BYTES = 40 * 1024
MAX_DIAG_BITS = 5
# handler ids (csrc/qfb_plan.h)
(H_G1_GENERAL, H_G1_SUMDIFF, H_G1_LU_R, H_G1_LU_I, H_G1C_GENERAL, H_G1C_SWAPX, H_CPH_SCALAR, H_CPH_REG1,
 H_CPH_RSC1, H_CPH_NEG1, H_CPH_NEG2, H_CPH_REGM, H_CPH_NEGM, H_END, H_G2, H_G2X, H_CPH_TABLE) = \
    0, 5, 10, 15, 20, 25, 30, 31, 36, 41, 46, 56, 57, 58, 59, 69, 79
G2_PAIRS = [(j0, j1) for j0 in range(HANDLER_STRIDE) for j1 in range(j0)]
SWEEP_FLAG_G2, SWEEP_FLAG_STORE_SYNC, SWEEP_FLAG_STORE_PERM = 1, 2, 4


def _op_record(handler: int, reg_cmask: int, idx_cmask: int, payload: bytes = b'', flag: int = 0) -> bytes:
    size = 16 + len(payload)
    assert size % 16 == 0 and size < 65536
    return struct.pack('<IHBBQ', handler, size, reg_cmask, flag, idx_cmask) + payload


def is_scalar_term(record: bytes) -> bool:
    return struct.unpack_from('<I', record, 0)[0] == H_CPH_SCALAR


def _round_header(nops, nbytes, regs, thrpad, has_scalar, has_g2, rgb, rst) -> bytes:
    """qfb_round_header without the thread LUTs (192 bytes)."""
    pad8 = [0] * (8 - len(regs))
    return struct.pack('<II8B12BBB2x8I8q8q', nops, nbytes, *regs, *pad8, *thrpad, has_scalar, has_g2,
                       *[swz(1 << p) << 4 for p in regs], *pad8, *rgb, *pad8, *rst, *pad8)


def swz(idx: int) -> int:
    """XOR swizzle of the exchange buffer (16-byte granularity), see csrc/qfb_sweep.cu."""
    x = idx >> 3
    return idx ^ ((x ^ (x >> 3) ^ (x >> 6) ^ (x >> 9)) & 7)


SWEEP_HEADER_BYTES = 112
ROUND_HEADER_BYTES = 192 + 16 * (16 + 32)

# QFB_G1_* kinds (csrc/qfb_plan.h)
K_GENERAL, K_REAL, K_RXLIKE, K_SWAPX, K_ANTIDIAG, K_SUMDIFF, K_LU_R, K_LU_I = range(8)

DEFAULT_TILE_BITS = 12          # interpreter plans (5 register bits)
DEFAULT_TILE_BITS_SPECIALISED = 11   # plans of the sweep-specialised kernels: 4 CTAs of 128 threads per SM instead of 2 of
                                    # 256 -- more tiles in different phases per SM (146.6 against 151.3 ms on the benchmark)


def default_tile_bits(reg_bits: int) -> int:
    v = os.environ.get('QFB_TILE_BITS')          # experiments
    if v:
        return int(v)
    return DEFAULT_TILE_BITS if reg_bits == REG_BITS else DEFAULT_TILE_BITS_SPECIALISED
DEFAULT_LOW_BITS = 3
# cost units ~ FP64 work per amplitude relative to a dense 1-bit operator (16 FP64 ops per amplitude pair)
COST = {K_GENERAL: 1.0, K_REAL: 1.0, K_RXLIKE: 1.0, K_SWAPX: 0.3, K_ANTIDIAG: 1.0, K_SUMDIFF: 0.25, K_LU_R: 0.25,
        K_LU_I: 0.25, 'G2': 2.5, 'P': 0.1}
DEFAULT_MAX_COST = 28.0
# randomised variants of the greedy sweep split tried by Planner._partition (0 = plain greedy)
DEFAULT_TRIES = 24
# randomised variants of the split of a sweep into rounds, for plans of the sweep-specialised kernels (an exchange less
# is ~1.5 ms of a 30-qubit sweep: 3-round sweeps run at 0.87 of the HBM roofline, 4-round sweeps at 0.70)
ROUND_TRIES_SPECIALISED = 256
# tile refinement (Planner._refine_tile): operator lists shorter than this are not worth the search
REFINE_MIN_OPS = 24
SWEEP_WINDOW = 4096         # operators of the list that one sweep may draw from
REFINE_PASSES = 6
REFINE_TRIES = 24
LOOKAHEAD_PASSES = 3
LOOKAHEAD_TRIES = 1        # partitions formed with the two-sweep look-ahead (kept only when they need fewer sweeps)
REFINE_WINDOW = 512      # operators of the list that the search scores (a sweep rarely executes more)
# pivot on the (0,0) entry unless it is this much smaller than the largest entry
PIVOT_RATIO = 1e-3


# one operator for the native tile search (qfb_plan_refine_tile): mix mask, diag mask, cost, record bytes
OP_ROW = struct.Struct('<QQdI4x')
OP_ROW_DTYPE = np.dtype([('mix', '<u8'), ('diag', '<u8'), ('cost', '<f8'), ('bytes', '<u4'), ('pad', '<u4')])


class POp:
    """A classified operator: kind 'G' (mixing) or 'P' (phase term)."""
    __slots__ = ('kind', 'mix', 'ctrl', 'dbits', 'mat', 'cost', 'mixset', 'diagset', 'anyset', 'gate_index', 'enc',
                 'mixmask', 'diagmask', 'plan_bytes', 'terms', 'row')

    def __init__(self, kind, mix=(), ctrl=(), dbits=(), mat=None, cost=1.0, gate_index=-1, enc=None, terms=None):
        self.enc = enc          # (kind, payload, scalar) chosen by absorb_scales for an uncontrolled 1-bit operator
        self.row = None         # (nbits, packed OP_ROW) for the native tile search, filled on first use (_refine_tile)
        self.terms = terms      # kind 'T' (diagonal table over register bits): [(bits, factor)], see _form_tables
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
        # the same sets as integer masks (Planner._closure scans operator lists thousands of times)
        self.mixmask = sum(1 << b for b in self.mixset)
        self.diagmask = sum(1 << b for b in self.diagset)
        # upper bound of the operator's records in a sweep: a flipped control can split an operator in two, a
        # flipped phase term of k bits into 2^k (absorb_frame)
        self.plan_bytes = 2 * (16 + 64) if (kind == 'G' and len(self.mix) == 1) else 2 * (16 + 272) if kind == 'G' \
            else (16 + 512) if kind == 'T' else 32 * (1 << len(self.dbits))


class Fallback:
    """An operator executed by the one-gate kernels (qfb_apply_dense)."""
    __slots__ = ('mat', 'bits')

    def __init__(self, mat, bits):
        self.mat = mat
        self.bits = tuple(int(b) for b in bits)


# ---------------------------------------------------------------------------------------------------------
# pass 1: classification
# ---------------------------------------------------------------------------------------------------------

def phase_polynomial(table: np.ndarray, k: int) -> Dict[int, complex]:
    """Diagonal operator -> {subset mask over the k table bits: factor} with d[s] = prod_{T subset of s} factor[T];
    bit q of the table index has weight 1 << (k-1-q). Requires every entry to be non-zero."""
    phi: Dict[int, complex] = {}
    for s in range(1 << k):
        val = complex(table[s])
        for t in range(s):
            if (t & s) == t:
                val = val / phi[t]
        phi[s] = val
    return phi


# LDU pivots on the (0,0) entry; absorb_frame swaps the rows (a free X flip) when the (1,0) entry is larger
LU_PIVOT = 0.05
# pending scales outside [1 / SCALE_LIMIT, SCALE_LIMIT] are emitted as phase terms instead of being carried on
SCALE_LIMIT = 65536.0


def encode_g1(mat: np.ndarray, controlled: bool) -> Tuple[int, np.ndarray, Optional[complex], complex]:
    """(kind, payload doubles, scalar, pending) for a 2x2 operator: the kernel applies S (named by kind + payload)
    and the operator equals  scalar . diag(1, pending) . S.  `scalar` (None = 1) is uniform and joins the sweep
    scalar; `pending` (1 = none) is a relative scale of the bit's |1> half that absorb_frame carries forward.
    Payload: 8 doubles (row-major matrix) for GENERAL / SWAPX, 2 doubles for the structured kinds."""
    m = np.array(mat, dtype=np.complex128).reshape(2, 2)
    plain = np.ascontiguousarray(m).view(np.float64).reshape(-1).copy()
    if classify.g1_kind(m) == K_SWAPX:
        return K_SWAPX, plain, None, 1.0
    g00 = m[0, 0]
    if controlled or g00 == 0 or abs(g00) < LU_PIVOT * abs(m[1, 0]):
        return K_GENERAL, plain, None, 1.0
    # the structure is judged with the phase of the pivot divided out (e.g. i . RX-like is RX-like)
    unit = g00 / abs(g00)
    if unit.imag == 0 or unit.real == 0:
        m = m / unit                       # exact: a division by +-1 or +-i
    else:
        unit = 1.0
    base = classify.g1_kind(m)            # 0 general, 1 real, 2 rxlike, 3 swapx, 4 antidiag
    if base not in (K_REAL, K_RXLIKE):
        return K_GENERAL, plain, None, 1.0
    p = m[0, 0].real
    if base == K_REAL and m[1, 0] == m[0, 0]:
        # [[1, r0], [1, r1]] . p: sums and differences only (Hadamard: r = +-1 keeps exact zeros)
        return K_SUMDIFF, np.array([m[0, 1].real / p, m[1, 1].real / p]), complex(p * unit), 1.0
    # LDU: G = diag(p, q) . [[1, 0], [l, 1]] . [[1, u], [0, 1]],  p = g00, q = det / g00
    if base == K_REAL:
        g01, g10, g11 = m[0, 1].real, m[1, 0].real, m[1, 1].real
        cross = g10 * g01 / p
        kind = K_LU_R
    else:
        g01, g10, g11 = m[0, 1].imag, m[1, 0].imag, m[1, 1].real      # off-diagonal entries are i g01, i g10
        cross = -g10 * g01 / p                                        # (i g10)(i g01) = -g10 g01
        kind = K_LU_I
    q = g11 - cross
    if abs(q) <= 1e-9 * (abs(g11) + abs(cross)):
        return K_GENERAL, plain, None, 1.0                            # singular: projectors, full damping
    return kind, np.array([g01 / p, g10 / q]), complex(p * unit), q / p


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
        cost = COST[K_REAL] if k == 1 else COST['G2'] * 0.25 + 0.3
        return [POp('G', mix=bits, ctrl=(), mat=mat, cost=cost, gate_index=gate_index)]
    controls, targets, reduced = ([], list(range(k)), mat) if k == 1 else classify.peel_controls(mat, k)
    if len(targets) > 2:
        return Fallback(mat, bits)
    cbits = [bits[q] for q in controls]
    tbits = [bits[q] for q in targets]
    if len(targets) == 1:
        kind = encode_g1(reduced, bool(cbits))[0]
        return [POp('G', mix=tbits, ctrl=cbits, mat=reduced, cost=COST[kind], gate_index=gate_index)]
    return [POp('G', mix=tbits, ctrl=cbits, mat=reduced, cost=_g_cost(tbits, cbits, reduced), gate_index=gate_index)]


_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)


X_SHAPE = np.array([[1, 0, 0, 1], [0, 1, 1, 0], [0, 1, 1, 0], [1, 0, 0, 1]], dtype=bool)


def is_real_x_shaped(mat: np.ndarray) -> bool:
    """Non-zeros only on the two diagonals of the 4x4 operator, all real: every Pauli channel and amplitude damping
    as a superoperator on (ket bit, bra bit); invariant under the frame pass (index XOR, real column scales)."""
    m = np.asarray(mat).reshape(4, 4)
    return not np.any(m[~X_SHAPE]) and not np.any(m.imag)


def _g_cost(mix, ctrl, mat) -> float:
    if len(mix) == 1:
        return COST[encode_g1(mat, bool(ctrl))[0]]
    if is_real_x_shaped(mat):
        return 1.0 + (0.3 if ctrl else 0.0)
    return COST['G2'] * max(int(np.count_nonzero(mat)), 4) / 16.0 + 0.3


def absorb_frame(ops: List[POp], scale: Dict[int, complex]) -> Tuple[List[POp], int]:
    """Pauli-X gates and relative scales are not executed, they are tracked. While a sweep's operators are walked
    in order, the stored state phi relates to the true one by  psi = X_F . D . phi :

      F   pending bit flips: an uncontrolled X on bit b toggles bit b of F and costs nothing; the kernel applies F
          as an XOR on the addresses of the sweep's final store (every flipped bit is a mixing bit of some
          operator of the sweep, hence a tile bit);
      D   = prod_b diag(1, scale[b]): what LDU leaves behind. A real (or RX-like) 1-bit operator G is executed as
          two shears; G = p . diag(1, s) . L . U, p joins the sweep scalar and s stays pending on the bit until
          the next operator that mixes the bit absorbs it (G' = G . diag(1, s)) or it has to be emitted as a phase
          term (before a controlled target, at the end of the plan). `scale` is carried from sweep to sweep.

    Every later operator U is replaced by what must act on phi:
      phase term   commutes with D; flipped bits expand phi^[all bits 1] into 2^|flipped| terms (phi or 1/phi)
      controls     commute with D; a flipped control (control on 0) is rewritten  C0(W) = W . C1(W^-1)
      1-bit G      X^f' G X^f diag(1, s): f' = f or 1 - f (a free row swap), whichever gives the better pivot
      2-bit G      rows and columns permuted by F, scales multiplied in

    Returns the rewritten list and F."""
    flip = 0
    out: List[POp] = []

    def phase(bits, factor, gi):
        fl = [b for b in bits if (flip >> b) & 1]
        keep = [b for b in bits if not (flip >> b) & 1]
        for sub in range(1 << len(fl)):
            tb = [fl[i] for i in range(len(fl)) if (sub >> i) & 1]
            f = complex(factor) if len(tb) % 2 == 0 else 1.0 / complex(factor)
            if f != 1:
                out.append(POp('P', dbits=keep + tb, mat=f, cost=COST['P'], gate_index=gi))

    def flush_scale(b, gi):
        lam = scale.pop(b, 1.0)
        if lam != 1:
            out.append(POp('P', dbits=[b], mat=complex(lam), cost=COST['P'], gate_index=gi))   # physical frame

    def gate(mix, ctrl, mat, flipped_ctrl, gi):
        nonlocal flip
        if flipped_ctrl:
            a, rest = flipped_ctrl[0], flipped_ctrl[1:]
            gate(mix, ctrl, np.linalg.inv(mat), rest, gi)                 # C_a(W^-1), control a now on 1
            gate(mix, [c for c in ctrl if c != a], mat, rest, gi)         # W without control a
            return
        k = len(mix)
        if k == 1 and not ctrl:
            b = mix[0]
            if np.array_equal(mat, _X):
                flip ^= 1 << b
                return
            if mat[0, 0] == 0 and mat[1, 1] == 0:
                # antidiagonal = diag(m01, m10) . X
                flip ^= 1 << b
                phase([], mat[0, 1], gi)
                phase([b], mat[1, 0] / mat[0, 1], gi)
                return
            if (flip >> b) & 1:
                mat = _X @ mat @ _X
            g = mat @ np.diag([1.0, scale.pop(b, 1.0)])
            if classify.is_diagonal(g):
                # the operator was the inverse of the pending scale up to a phase: back to a phase term
                out.append(POp('P', dbits=[], mat=complex(g[0, 0]), cost=0.0, gate_index=gi))
                if g[1, 1] != g[0, 0]:
                    scale[b] = g[1, 1] / g[0, 0]
                return
            kind, payload, scalar, pending = encode_g1(g, False)
            if kind != K_SUMDIFF and abs(g[0, 0]) < 0.6 * abs(g[1, 0]):
                # swap the rows (toggle the flip): the pivot is the larger entry of the first column
                swapped = encode_g1(_X @ g, False)
                if COST[swapped[0]] <= COST[kind]:
                    flip ^= 1 << b
                    g = _X @ g
                    kind, payload, scalar, pending = swapped
            out.append(POp('G', mix=mix, mat=np.ascontiguousarray(g), cost=COST[kind], gate_index=gi,
                           enc=(kind, payload, scalar)))
            if pending != 1:
                scale[b] = pending
                if not 1.0 / SCALE_LIMIT < abs(pending) < SCALE_LIMIT:
                    flush_scale(b, gi)          # keep the stored amplitudes within a sane dynamic range
            return
        if k == 1:
            flush_scale(mix[0], gi)                 # a controlled target cannot absorb the scale
            if (flip >> mix[0]) & 1:
                mat = _X @ mat @ _X
        else:
            perm = [i ^ (2 * ((flip >> mix[0]) & 1)) ^ ((flip >> mix[1]) & 1) for i in range(4)]
            mat = mat[np.ix_(perm, perm)]
            if ctrl:
                flush_scale(mix[0], gi)
                flush_scale(mix[1], gi)
            else:
                la, lb = scale.pop(mix[0], 1.0), scale.pop(mix[1], 1.0)
                mat = mat @ np.diag([1.0, lb, la, la * lb])
        mat = np.ascontiguousarray(mat)
        if classify.is_identity(mat):
            return
        out.append(POp('G', mix=mix, ctrl=ctrl, mat=mat, cost=_g_cost(mix, ctrl, mat), gate_index=gi))

    for op in ops:
        if op.kind == 'P':
            if any((flip >> b) & 1 for b in op.dbits):
                phase(list(op.dbits), op.mat, op.gate_index)
            else:
                out.append(op)
        else:
            k = len(op.mix)
            mat = np.asarray(op.mat, dtype=np.complex128).reshape(1 << k, 1 << k)
            gate(list(op.mix), list(op.ctrl), mat, [c for c in op.ctrl if (flip >> c) & 1], op.gate_index)
    return out, flip


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


def sink_phase_terms(parts: List[Tuple[List[POp], List[int]]]) -> List[Tuple[List[POp], List[int]]]:
    """Move every phase term forward, across sweeps, to just before the next operator that mixes one of its bits
    (its anchor). Legal: a phase term commutes with everything that does not mix its bits, and the execution
    order keeps the program order of operators that share a mixed bit. Why: the anchor's target is a REGISTER bit
    in the anchor's round, so the term arrives where it costs least -- as an entry of the round's diagonal table
    (_form_tables) instead of a per-thread scalar pass over all amplitudes (128 FP64 instructions per thread and
    round) or a phase pass of its own in an earlier sweep. Terms without an anchor (nothing mixes their bits any
    more) and global phases stay where they are. A sweep whose record would outgrow MAX_SWEEP_BYTES takes no
    more terms."""
    flat: List[Tuple[int, POp]] = [(si, op) for si, (chosen, _) in enumerate(parts) for op in chosen]
    next_mix: Dict[int, int] = {}           # bit -> flat position of the next operator that mixes it
    anchor_of: Dict[int, int] = {}          # flat position of a phase term -> flat position of its anchor
    for pos in range(len(flat) - 1, -1, -1):
        op = flat[pos][1]
        if op.kind == 'P':
            hits = [next_mix[b] for b in op.dbits if b in next_mix]
            if hits:
                anchor_of[pos] = min(hits)
        else:
            for b in op.mix:
                next_mix[b] = pos
    room = [MAX_SWEEP_BYTES - 4 * ROUND_HEADER_BYTES - (SWEEP_HEADER_BYTES + 8 * ROUND_HEADER_BYTES) -
            sum(op.plan_bytes for op in chosen) for chosen, _ in parts]
    before: Dict[int, List[POp]] = {}       # flat position of an anchor -> the terms that now precede it
    moved = set()
    for pos, apos in anchor_of.items():
        si, op = flat[pos]
        sa = flat[apos][0]
        if sa == si or room[sa] < op.plan_bytes:
            continue
        room[sa] -= op.plan_bytes
        room[si] += op.plan_bytes
        before.setdefault(apos, []).append(op)
        moved.add(pos)
    if not moved:
        return parts
    out: List[Tuple[List[POp], List[int]]] = [([], tile) for _, tile in parts]
    for pos, (si, op) in enumerate(flat):
        if pos in moved:
            continue
        out[si][0].extend(before.get(pos, ()))
        out[si][0].append(op)
    return out


def _conflicts(op: POp, def_any: set, def_mix: set) -> bool:
    """Does `op` fail to commute with some deferred operator?"""
    return bool(op.mixset & def_any) or bool(op.diagset & def_mix)


def default_reg_bits(nbits: int) -> int:
    """Register bits of a plan: 5 (32 amplitudes per thread) for the interpreter; 4 for states that run the sweep-
    specialised kernels (csrc/qfb_jit.cu), whose straight-line code must fit the SM's 32 KiB instruction cache
    (profiles/r2_icache_probe.jsonl): half the amplitudes per thread is half the code per operator. The rule is
    the library's (jit_wanted, csrc/qfb_sweep.cu): QFB_JIT=0 switches the kernels off, otherwise states of at
    least QFB_JIT_MIN_BITS (24) index bits use them. QFB_REG_BITS overrides (experiments)."""
    v = os.environ.get('QFB_REG_BITS')
    if v:
        return int(v)
    jit = os.environ.get('QFB_JIT')
    if jit:
        return 4 if int(jit) != 0 else 5
    return 4 if nbits >= int(os.environ.get('QFB_JIT_MIN_BITS') or 24) else 5


def swizzle_class(pos: int) -> int:
    return pos if pos < 3 else (pos - 3) % 3


class Round:
    __slots__ = ('regs', 'thr', 'ops')

    def __init__(self, regs, thr, ops):
        self.regs = regs   # tile positions of register bits 0..3
        self.thr = thr     # tile positions of thread bits 0..M-5
        self.ops = ops     # POps in execution order


class SweepPlan:
    __slots__ = ('tile', 'ops', 'rounds', 'cost', 'store_xor', 'spos')

    def __init__(self, tile: List[int], ops: List[POp], store_xor: int = 0):
        self.tile = tile      # index-bit positions, ascending, length M
        self.ops = ops
        self.rounds: List[Round] = []
        self.cost = sum(o.cost for o in ops)
        self.store_xor = store_xor   # pending X flips, applied by the final store (absorb_flips); STORE positions
        self.spos = list(tile)       # index-bit position tile bit j is stored to (!= tile: in-place bit permutation)


# length of the random streams handed to qfb_plan_split_rounds (a split consumes one number per operator that asks
# for a new register bit, a few hundred at most; a split that runs out is redone in Python)
ROUND_STREAM_LEN = 1024
_round_streams: Dict[int, np.ndarray] = {}


def _round_stream(trial: int) -> np.ndarray:
    """The first ROUND_STREAM_LEN numbers of random.Random(trial): what Planner._split_rounds draws for that trial."""
    stream = _round_streams.get(trial)
    if stream is None:
        draw = random.Random(trial).random
        stream = np.fromiter((draw() for _ in range(ROUND_STREAM_LEN)), dtype=np.float64, count=ROUND_STREAM_LEN)
        _round_streams[trial] = stream
    return stream


class _RoundSplitter:
    """The operators of one sweep as arrays for qfb_plan_split_rounds (csrc/qfb_planhost.cu), the native statement of
    Planner._split_rounds: score() = (number of rounds, cost of the last round's operators) of one split, rounds() =
    the split itself as _split_rounds returns it. Trial 0 is the plain greedy, trial t > 0 the randomised variant
    with random.Random(t) and p_new = 0.85."""

    def __init__(self, planner: 'Planner', sweep: 'SweepPlan'):
        from . import _lib
        self.ops = sweep.ops
        self.R, self.L = planner.R, planner.L
        n = len(self.ops)
        pos_of = {b: j for j, b in enumerate(sweep.tile)}
        self.lib = _lib.load()
        try:
            self.mix = np.fromiter((op.mixmask for op in self.ops), dtype=np.uint64, count=n)
            self.diag = np.fromiter((op.diagmask for op in self.ops), dtype=np.uint64, count=n)
            self.pos = np.fromiter((sum(1 << pos_of[b] for b in op.mix) if op.kind == 'G' else 0xffffffff
                                    for op in self.ops), dtype=np.uint32, count=n)
            self.cost = np.fromiter((op.cost for op in self.ops), dtype=np.float64, count=n)
        except OverflowError:           # index bits beyond 63: the Python loop handles them
            self.lib = None

    def _call(self, trial: int, backward: bool, round_of, regs, max_rounds: int):
        import ctypes
        from . import _lib
        stream = _round_stream(trial) if trial else None
        nrounds, consumed, tail = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0.0)
        _lib.check(self.lib.qfb_plan_split_rounds(
            self.mix.ctypes.data, self.diag.ctypes.data, self.pos.ctypes.data, self.cost.ctypes.data, len(self.ops),
            self.R, self.L, stream.ctypes.data if trial else None, ROUND_STREAM_LEN if trial else 0, 0.85,
            1 if backward else 0, round_of, regs, max_rounds, ctypes.byref(nrounds), ctypes.byref(tail),
            ctypes.byref(consumed)))
        return (nrounds.value, tail.value) if consumed.value >= 0 else None

    def score(self, trial: int, backward: bool) -> Optional[Tuple[int, float]]:
        if self.lib is None:
            return None
        return self._call(trial, backward, None, None, 0)

    def rounds(self, trial: int, backward: bool) -> List[Tuple[List[int], List['POp']]]:
        n = len(self.ops)
        round_of = np.zeros(max(n, 1), dtype=np.int32)
        regs = np.zeros(n + 3, dtype=np.uint32)
        nrounds, _ = self._call(trial, backward, round_of.ctypes.data, regs.ctypes.data, n + 3)
        out: List[Tuple[List[int], List[POp]]] = [([p for p in range(32) if (int(regs[r]) >> p) & 1], [])
                                                  for r in range(nrounds)]
        for op, r in zip(self.ops, round_of.tolist()):
            out[r][1].append(op)
        return out


class Planner:
    def __init__(self, nbits: int, tile_bits: int = None, low_bits: int = None, max_cost: float = None,
                 tries: int = None, refine: bool = None, reg_bits: int = None):
        self.nbits = int(nbits)
        self.R = default_reg_bits(self.nbits) if reg_bits is None else int(reg_bits)
        if self.R not in (3, 4, 5):
            raise ValueError('reg_bits must be 3, 4 or 5')
        # Diagonal tables (and the placement of phase terms that feeds them) change WHERE a phase is multiplied in.
        # That is exact in real arithmetic but not in floating point: amplitudes that cancel to an exact 0 in the
        # reference's gate-by-gate arithmetic may end as 1e-17 (measured: tools/zero_pattern.py), and
        # np.random.multinomial consumes no random number for a zero-probability bin -- so the bit-exact sampling
        # fixtures only hold with the plain placement. Plans of the interpreter (5 register bits: states below
        # QFB_JIT_MIN_BITS, where the whole probability vector goes to the host's numpy calls) keep it; plans of
        # the sweep-specialised kernels (4 register bits: large states, sampled on the device) take the tables.
        v = os.environ.get('QFB_PLAN_TABLES')
        self.tables = (self.R != REG_BITS) if not v else v != '0'
        # latest-possible rounds (see _split_rounds): plans of the sweep-specialised kernels only, for the same reason
        v = os.environ.get('QFB_PLAN_LATE')
        self.late_rounds = (self.R != REG_BITS) if not v else v != '0'
        m = default_tile_bits(self.R) if tile_bits is None else int(tile_bits)
        m = min(m, self.nbits, MAX_TILE_BITS)
        if m < MIN_TILE_BITS:
            raise ValueError('state too small for the tiled executor (need >= {} bits)'.format(MIN_TILE_BITS))
        if self.nbits - m > MAX_HOLES:
            raise ValueError('state too large for one plan')
        self.M = m
        low = DEFAULT_LOW_BITS if low_bits is None else int(low_bits)
        if m - self.R < 3:
            self.R = REG_BITS          # tiny tiles: only the interpreter runs them
        while m - self.R > 9:
            self.R += 1                # the thread tables of a round record hold 9 thread bits
        self.L = max(0, min(low, m - self.R))
        self.max_cost = DEFAULT_MAX_COST if max_cost is None else float(max_cost)
        self.tries = DEFAULT_TRIES if tries is None else int(tries)
        self.refine = True if refine is None else bool(refine)
        # results of the tile searches, keyed by the operators scanned and the start tile: the randomised partitions
        # of one plan, and even more so the candidate schedules of a sharded circuit (sharded.schedule hands one dict
        # to all of them), repeat many searches. Only valid while the operator objects are alive (ids in the key).
        self.search_cache: Optional[Dict[tuple, int]] = None

    # ---- pass 2: sweeps ---------------------------------------------------------------------------
    def _form_sweep(self, ops: List[POp], rnd=None, p_new: float = 1.0, forbidden: frozenset = frozenset(),
                    required: frozenset = frozenset(), lookahead: bool = False
                    ) -> Tuple[List[POp], List[POp], List[int]]:
        """One sweep: the operators that join it, the deferred rest, the tile bits.

        A greedy walk in program order picks the tile (an operator joins while its mixing bits fit; with `rnd`, an
        operator that would bring a NEW bit into the tile is only admitted with probability p_new: randomised
        variants, see _partition). The walk lets the first operators it meets claim the tile, so the tile is then
        refined by local search (_refine_tile): single-bit exchanges are kept while they raise the number of
        operators the sweep executes. Operators that mix a bit of `forbidden` (the global qubits of a sharded
        state, sharded.schedule) are deferred and such bits never pad the tile. The bits of `required` are tile
        bits whatever the operators need (the bit positions a qubit remap moves: the sweep that holds them all
        can store the permutation, attach_permutation)."""
        tmask = (1 << self.L) - 1
        for b in required:
            tmask |= 1 << b
        tcount = bin(tmask).count('1')
        if tcount > self.M:
            raise ValueError('required bits do not fit a tile')
        fmask = sum(1 << b for b in forbidden)
        # Only the next SWEEP_WINDOW operators are candidates (everything behind them is deferred as it stands:
        # deferring is always legal, and a sweep holds a few hundred operators at most). Keeps the planner linear
        # in the length of the circuit instead of quadratic.
        beyond = ops[SWEEP_WINDOW:]
        ops = ops[:SWEEP_WINDOW] if beyond else ops
        cost = 0.0
        nbytes = SWEEP_HEADER_BYTES + 8 * ROUND_HEADER_BYTES
        cap = MAX_SWEEP_BYTES - 4 * ROUND_HEADER_BYTES
        da = dm = 0                    # index bits touched / mixed by the operators deferred so far (bit masks)
        started = False
        for op in ops:
            mm = op.mixmask
            ok = not ((mm & da) or (op.diagmask & dm))
            if ok and op.kind == 'G':
                if mm & fmask:
                    ok = False
                else:
                    if mm >> self.nbits:
                        raise ValueError('operator mixes bit {} outside the local index; remap first'.format(
                            max(op.mix)))
                    need = mm & ~tmask
                    if tcount + bin(need).count('1') > self.M:
                        ok = False
                    elif need and rnd is not None and started and rnd.random() > p_new:
                        ok = False
            if ok and cost + op.cost > self.max_cost and started:
                ok = False
            if ok and nbytes + op.plan_bytes > cap:
                break
            if ok:
                started = True
                cost += op.cost
                nbytes += op.plan_bytes
                if op.kind == 'G' and mm & ~tmask:
                    tmask |= mm
                    tcount = bin(tmask).count('1')
            else:
                da |= mm | op.diagmask
                dm |= mm
        # pad the tile with the lowest free bits (locality of the strided tile accesses)
        b = 0
        while tcount < self.M:
            if not ((tmask | fmask) >> b) & 1:
                tmask |= 1 << b
                tcount += 1
            b += 1
        if self.refine and len(ops) >= REFINE_MIN_OPS:
            tmask = self._refine_tile(ops, tmask, fmask, sum(1 << b for b in required), lookahead)
        chosen, deferred = self._closure(ops, tmask, fmask)
        return chosen, deferred + beyond, [b for b in range(self.nbits) if (tmask >> b) & 1]

    def _closure(self, ops: List[POp], tmask: int, fmask: int) -> Tuple[List[POp], List[POp]]:
        """(chosen, deferred): the operators a sweep over the tile `tmask` executes, in program order, and the rest.
        An operator joins when it mixes tile bits only (never a bit of `fmask`), commutes with everything deferred
        before it and fits the sweep's work and size caps."""
        allow = tmask & ~fmask
        every = (1 << self.nbits) - 1
        max_cost = self.max_cost
        cap = MAX_SWEEP_BYTES - 4 * ROUND_HEADER_BYTES
        chosen: List[POp] = []
        deferred: List[POp] = []
        da = dm = 0                    # bits touched / mixed by the deferred operators so far
        cost = 0.0
        nbytes = SWEEP_HEADER_BYTES + 8 * ROUND_HEADER_BYTES
        started = False
        blocked = len(ops)             # from here on every bit is blocked: only bit-free scalar factors still pass
        for i, op in enumerate(ops):
            mm = op.mixmask
            dd = op.diagmask
            if (mm & da) or (dd & dm) or (mm & ~allow) or (started and cost + op.cost > max_cost):
                deferred.append(op)
                da |= mm | dd
                dm |= mm
                if dm & every == every:
                    blocked = i + 1
                    break
                continue
            if nbytes + op.plan_bytes > cap:
                # the sweep record is full: everything from here on waits for the next sweep
                deferred.extend(ops[i:])
                break
            started = True
            cost += op.cost
            nbytes += op.plan_bytes
            chosen.append(op)
        for i in range(blocked, len(ops)):
            op = ops[i]
            if (op.mixmask | op.diagmask) or (started and cost + op.cost > max_cost):
                deferred.append(op)
            elif nbytes + op.plan_bytes > cap:
                deferred.extend(ops[i:])
                break
            else:
                started = True
                cost += op.cost
                nbytes += op.plan_bytes
                chosen.append(op)
        return chosen, deferred

    def _refine_tile(self, ops: List[POp], tmask: int, fmask: int, keep: int = 0, lookahead: bool = False) -> int:
        """Local search over the tile of one sweep: exchange one tile bit (never the low bits, which every sweep
        needs for whole 128-byte lines, nor the bits of `keep`) for one outside bit while that raises the number
        of operators the sweep executes; the best exchange of a pass is taken, up to REFINE_PASSES passes. The
        score is a bit-mask scan over the next REFINE_WINDOW operators, a few thousand scans per sweep: it runs
        natively (qfb_plan_refine_tile, csrc/qfb_planhost.cu; the same scan in Python, which a test compares it
        with: tests/plan_emulator.py::count_executed). On the 30-qubit benchmark the search takes the plan from
        20 sweeps (randomised greedy walks alone) to 15."""
        import ctypes
        from . import _lib
        lib = _lib.load()
        window = ops[:REFINE_WINDOW]
        n = len(window)
        key = None
        if self.search_cache is not None:
            key = (tuple(map(id, window)), tmask, fmask, keep, lookahead, self.nbits, self.L, self.M, self.max_cost)
            hit = self.search_cache.get(key)
            if hit is not None:
                return hit
        every = (1 << self.nbits) - 1
        nb = self.nbits
        # the operators as parallel arrays; every operator keeps its packed row (the same operators are scanned by
        # hundreds of searches)
        rows = []
        for op in window:
            row = op.row
            if row is None or row[0] != nb:
                row = op.row = (nb, OP_ROW.pack(op.mixmask & every, op.diagmask & every, op.cost, op.plan_bytes))
            rows.append(row[1])
        rec = np.frombuffer(b''.join(rows), dtype=OP_ROW_DTYPE)
        mix = np.ascontiguousarray(rec['mix'])
        diag = np.ascontiguousarray(rec['diag'])
        cost = np.ascontiguousarray(rec['cost'])
        nbytes = np.ascontiguousarray(rec['bytes'])
        room = MAX_SWEEP_BYTES - 4 * ROUND_HEADER_BYTES - (SWEEP_HEADER_BYTES + 8 * ROUND_HEADER_BYTES)
        out = ctypes.c_uint64(0)
        _lib.check(lib.qfb_plan_refine_tile(mix.ctypes.data, diag.ctypes.data, cost.ctypes.data, nbytes.ctypes.data, n,
                                            self.nbits, tmask, fmask, ((1 << self.L) - 1) | keep,
                                            float(self.max_cost), room, REFINE_PASSES, ctypes.byref(out), None))
        if lookahead:
            # second search, from the tile found above: score = operators of this sweep + operators of the best
            # sweep that can follow it (what a tile leaves behind matters as much as what it takes)
            _lib.check(lib.qfb_plan_refine_tile_lookahead(
                mix.ctypes.data, diag.ctypes.data, cost.ctypes.data, nbytes.ctypes.data, n, self.nbits, self.L,
                self.M, int(out.value), fmask, ((1 << self.L) - 1) | keep, float(self.max_cost), room,
                REFINE_PASSES, LOOKAHEAD_PASSES, ctypes.byref(out), None))
        if key is not None:
            self.search_cache[key] = int(out.value)
        return int(out.value)

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

    def _split_rounds(self, sweep: SweepPlan, rnd=None, p_new: float = 1.0,
                      backward: bool = False) -> List[Tuple[List[int], List[POp]]]:
        """Greedy split of a sweep's operators into rounds of R register bits; with `rnd`, an operator that needs
        a NEW register bit is only admitted with probability p_new (randomised variants, see _form_rounds).
        `backward`: the same greedy on the reversed operator list (commutation is symmetric, and so are the
        constraints of the two rounds that touch HBM), i.e. every operator goes to the LATEST round that can take
        it: the full rounds end up at the end of the sweep, where the sweep-specialised kernels hide the
        asynchronous copy of the next tile (it can only start after the last exchange)."""
        if backward:
            mirror = SweepPlan(sweep.tile, list(reversed(sweep.ops)))
            rounds = self._split_rounds(mirror, rnd, p_new)
            return [(regs, list(reversed(chosen))) for regs, chosen in reversed(rounds)]
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
                    elif len(regs) + len(need) > self.R:
                        ok = False
                    elif need and regs and rnd is not None and rnd.random() > p_new:
                        ok = False
                    else:
                        regs += need
                if ok:
                    chosen.append(op)
                else:
                    deferred.append(op)
                    def_any |= op.anyset
                    def_mix |= op.mixset
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
        return rounds

    def _form_rounds(self, sweep: SweepPlan) -> None:
        # every extra round is one more trip of the tile through shared memory: keep the split with the fewest.
        # The search scores a few hundred splits per sweep: they are computed natively (_RoundSplitter) and only the
        # one that is kept is turned into operator lists. A candidate is (rounds, cost of the last round, split or key).
        splitter = _RoundSplitter(self, sweep)

        def split(trial: int, backward: bool):
            scored = splitter.score(trial, backward)
            if scored is None:          # no native answer (masks beyond 64 bits, random stream too short)
                rounds = self._split_rounds(sweep, random.Random(trial) if trial else None, 0.85 if trial else 1.0,
                                            backward)
                return len(rounds), sum(op.cost for op in rounds[-1][1]), rounds
            return scored[0], scored[1], (trial, backward)

        best = split(0, False)
        tries = max(self.tries, ROUND_TRIES_SPECIALISED) if self.late_rounds and self.tries > 0 else self.tries
        if len(sweep.ops) >= 16:
            for trial in range(1, 1 + tries):
                if best[0] <= 2 or (trial > self.tries and best[0] <= 3):
                    break           # the long search is for sweeps that still need four rounds
                cand = split(trial, False)
                if cand[0] < best[0]:
                    best = cand
        if self.late_rounds and best[0] > 1:
            # the latest-possible split, kept when it needs no more rounds and leaves more work behind the last
            # exchange (cost of the last round's operators)
            cands = [split(0, True)]
            if len(sweep.ops) >= 16:
                for trial in range(1, 1 + tries):
                    if trial > self.tries and min(best[0], min(c[0] for c in cands)) <= 3:
                        break
                    cands.append(split(trial, True))
            for cand in cands:
                if cand[0] < best[0] or (cand[0] == best[0] and cand[1] > best[1]):
                    best = cand
        rounds = splitter.rounds(*best[2]) if isinstance(best[2], tuple) else best[2]
        final: List[Round] = []
        nr = len(rounds)
        for r, (regs, chosen) in enumerate(rounds):
            edge = (r == 0) or (r == nr - 1)   # rounds that touch HBM keep the low bits on the lanes
            regs = list(regs)
            # fill up to R register bits with high free positions (never low ones on edge rounds)
            cand = [p for p in range(self.M - 1, -1, -1) if p not in regs and (p >= self.L or not edge)]
            # prefer a spread of swizzle classes so that lane bits always find three distinct classes
            while len(regs) < self.R:
                counts = {c: sum(1 for p in regs if swizzle_class(p) == c) for c in range(3)}
                cand.sort(key=lambda p: (counts[swizzle_class(p)], -p))
                regs.append(cand.pop(0))
            regs.sort()
            final.append(Round(regs, self._thread_order(regs, edge), chosen))
        sweep.rounds = final
        self._relocate_phase_terms(sweep)
        if self.tables:
            self._form_tables(sweep)

    def _relocate_phase_terms(self, sweep: SweepPlan) -> None:
        """Move every phase term, inside its commutation window, to the round where it is cheapest."""
        pos_of = {b: j for j, b in enumerate(sweep.tile)}
        rounds = sweep.rounds
        nr = len(rounds)
        placed: List[List[POp]] = [[op for op in rd.ops if op.kind == 'G'] for rd in rounds]
        terms: List[Tuple[POp, Tuple[int, int], Tuple[int, int]]] = []
        for r, rd in enumerate(rounds):
            gcount = 0
            for op in rd.ops:
                if op.kind == 'G':
                    gcount += 1
                    continue
                # window: after the last conflicting G op before the term, before the first one after it
                lo = (-1, -1)
                for rr in range(r, -1, -1):
                    glist = placed[rr]
                    upto = gcount if rr == r else len(glist)
                    hit = [i for i in range(upto) if glist[i].mixset & op.diagset]
                    if hit:
                        lo = (rr, hit[-1])
                        break
                hi = (nr, 0)
                for rr in range(r, nr):
                    glist = placed[rr]
                    start = gcount if rr == r else 0
                    hit = [i for i in range(start, len(glist)) if glist[i].mixset & op.diagset]
                    if hit:
                        hi = (rr, hit[0])
                        break
                terms.append((op, lo, hi))

        def reg_mask(op: POp, r: int) -> int:
            mask = 0
            for b in op.dbits:
                p = pos_of.get(b)
                if p is not None and p in rounds[r].regs:
                    mask |= 1 << rounds[r].regs.index(p)
            return mask

        def legal_rounds(lo, hi):
            return range(max(lo[0], 0), min(hi[0], nr - 1) + 1)

        # the round that can host most terms as per-thread scalars becomes the sweep's scalar round
        votes = [0] * nr
        for op, lo, hi in terms:
            for r in legal_rounds(lo, hi):
                if reg_mask(op, r) == 0:
                    votes[r] += 1
        scalar_rounds = {max(range(nr), key=lambda r: (votes[r], -r))} if terms else set()
        # after[r][i]: terms that run right after G op i of round r (i = -1: at the start of the round)
        after: List[Dict[int, List[POp]]] = [dict() for _ in range(nr)]
        # a term whose bits are ALL register bits of a round joins that round's diagonal table (_form_tables):
        # almost free, and the more terms a table collects the better
        allin = [0] * nr
        for op, lo, hi in terms:
            for r in legal_rounds(lo, hi):
                if op.dbits and bin(reg_mask(op, r)).count('1') == len(op.dbits):
                    allin[r] += 1
        for op, lo, hi in terms:
            best = None
            for r in legal_rounds(lo, hi):
                mask = reg_mask(op, r)
                if mask == 0:
                    cost = 0.05 if r in scalar_rounds else 1.05
                elif self.tables and bin(mask).count('1') == len(op.dbits) and allin[r] >= 2:
                    cost = 0.04 - 0.001 * min(allin[r], 20)
                else:
                    touched = (1 << self.R) >> bin(mask).count('1')
                    cost = (0.1 if op.mat == -1 else 0.5) * touched / 8.0
                if best is None or cost < best[0]:
                    best = (cost, r, mask)
            _, r, mask = best
            if mask == 0:
                scalar_rounds.add(r)
            if r == lo[0]:
                anchor = lo[1]                     # right after the conflicting predecessor
            elif r == hi[0]:
                anchor = hi[1] - 1                 # right before the conflicting successor
            else:
                anchor = len(placed[r]) - 1        # anywhere: at the end
            after[r].setdefault(anchor, []).append(op)
        for r, rd in enumerate(rounds):
            ops: List[POp] = list(after[r].get(-1, []))
            for i, g in enumerate(placed[r]):
                ops.append(g)
                ops.extend(after[r].get(i, []))
            rd.ops = ops

    def _form_tables(self, sweep: SweepPlan) -> None:
        """Phase terms whose bits are all register bits of their round are multiplied into diagonal TABLES: one
        complex multiplication per touched amplitude for the whole group (4 FP64 instructions) instead of one
        pass per term (a 1-bit term alone touches half the amplitudes). A term may sit anywhere between the last
        operator before it and the first one after it that mix one of its bits; the fewest table positions that
        serve all terms of a round are found by stabbing these intervals (latest legal position first). A
        position that serves a single term keeps the term as it is."""
        for rd in sweep.rounds:
            regbits = {sweep.tile[p] for p in rd.regs}
            ops = rd.ops
            gidx = [i for i, op in enumerate(ops) if op.kind == 'G']
            ng = len(gidx)
            cands = []                 # (hi, lo, position in ops): complex factors
            signs = []                 # the same for factor -1 terms: sign flips cost no FP64 work on their own
            seen_g = 0
            for i, op in enumerate(ops):
                if op.kind == 'G':
                    seen_g += 1
                    continue
                if op.kind != 'P' or not op.dbits or not all(b in regbits for b in op.dbits):
                    continue
                lo = seen_g
                while lo > 0 and not (ops[gidx[lo - 1]].mixset & op.diagset):
                    lo -= 1
                hi = seen_g
                while hi < ng and not (ops[gidx[hi]].mixset & op.diagset):
                    hi += 1
                (signs if op.mat == -1 else cands).append((hi, lo, i))
            if len(cands) < 2:
                continue
            cands.sort()
            groups: Dict[int, List[int]] = {}      # gap (index of the G operator the table precedes) -> terms
            covered = set()
            for hi, lo, i in cands:
                if i in covered:
                    continue
                members = [j for (h2, l2, j) in cands if j not in covered and l2 <= hi <= h2]
                covered.update(members)
                groups[hi] = members
            tables = {gap: members for gap, members in groups.items() if len(members) >= 2}
            if not tables:
                continue
            # a sign flip whose bits the table touches anyway rides along for free
            for hi, lo, i in signs:
                for gap, members in tables.items():
                    if lo <= gap <= hi and set(ops[i].dbits) <= {b for j in members for b in ops[j].dbits}:
                        members.append(i)
                        break
            drop = {i for members in tables.values() for i in members}
            new_ops: List[POp] = []
            g = 0

            def emit(gap: int) -> None:
                members = tables.get(gap)
                if members:
                    bits = sorted({b for i in members for b in ops[i].dbits})
                    new_ops.append(POp('T', dbits=bits, cost=COST['P'], gate_index=ops[members[0]].gate_index,
                                       terms=[(ops[i].dbits, complex(ops[i].mat)) for i in members]))

            for i, op in enumerate(ops):
                if op.kind == 'G':
                    emit(g)
                    g += 1
                if i not in drop:
                    new_ops.append(op)
            emit(ng)
            rd.ops = new_ops

    # ---- driver -----------------------------------------------------------------------------------
    def attach_permutation(self, sweeps: List[SweepPlan], perm: Sequence[int]) -> None:
        """Make the plan end with the in-place bit permutation `perm` (destination bit j <- source bit perm[j],
        the convention of qfb_permute_bits) at no extra pass when the moved bits are tile bits of the last sweep:
        the sweep then stores tile bit b at position dst_of[b]. Otherwise bare sweeps are appended (one usually;
        more when the moved bits do not fit into one tile)."""
        nb = self.nbits
        assert sorted(perm) == list(range(nb))
        content = list(range(nb))               # content[pos] = source bit currently at position pos
        want = list(perm)                       # position j must end up holding source bit perm[j]

        def apply(sweep: SweepPlan, mapping: Dict[int, int]) -> None:
            # mapping: position -> new position, for the tile bits of `sweep` (identity elsewhere)
            sweep.spos = [mapping.get(b, b) for b in sweep.tile]
            sweep.store_xor = sum(1 << mapping.get(b, b) for b in range(nb) if (sweep.store_xor >> b) & 1)
            if not sweep.rounds:
                # bare sweep: a load round (lanes on index bits 0.. under tile[]) before the store round
                regs = [p for p in range(self.M - 1, -1, -1) if p >= self.L][:self.R]
                regs.sort()
                sweep.rounds.append(Round(regs, self._thread_order(regs, True), []))
            # store round: the lanes walk the tile positions that are stored to index bits 0, 1, 2 ...
            lanes = [sweep.spos.index(t) for t in range(min(self.L, self.M - self.R))]
            regs = [p for p in range(self.M - 1, -1, -1) if p not in lanes][:self.R]
            regs.sort()
            thr = lanes + [p for p in range(self.M) if p not in lanes and p not in regs]
            sweep.rounds.append(Round(regs, thr, []))

        first = True
        while content != want:
            # transpositions that each put one source bit into its final position, grouped while they fit a tile
            moved: List[int] = []
            mapping: Dict[int, int] = {}
            trial = list(content)
            last = sweeps[-1] if (first and sweeps) else None
            cap = self.M - self.L
            while trial != want:
                j = next(q for q in range(nb) if trial[q] != want[q])
                q = trial.index(want[j])
                new_moved = sorted(set(moved) | {j, q})
                if last is not None:
                    fits = all(b in last.tile for b in new_moved)
                else:
                    fits = len([b for b in new_moved if b >= self.L]) <= cap
                if not fits:
                    break
                moved = new_moved
                trial[j], trial[q] = trial[q], trial[j]
            if not moved:
                if last is not None:
                    first = False       # the last sweep cannot host it: start over with appended bare sweeps
                    continue
                raise RuntimeError('bit permutation does not fit a tile')
            # composite position map of this group: where does the content of position b end up
            for b in moved:
                mapping[b] = trial.index(content[b])
            if last is not None:
                apply(last, mapping)
            else:
                tile = set(range(self.L)) | set(moved)
                b = 0
                while len(tile) < self.M:
                    tile.add(b)
                    b += 1
                bare = SweepPlan(sorted(tile), [])
                apply(bare, mapping)
                sweeps.append(bare)
            content = trial
            first = False

    def _partition(self, pops: List[POp]) -> List[Tuple[List[POp], List[int]]]:
        """Split the operator list into sweeps. Every sweep costs one pass over the state (the dominant cost), so
        besides the plain greedy walk a few randomised variants are tried (fixed seeds: the plan is deterministic)
        and the split with the fewest sweeps wins."""
        own_cache = self.search_cache is None
        if own_cache:
            self.search_cache = {}          # `pops` keeps the operators alive while it is in use
        try:
            return self._partition_search(pops)
        finally:
            if own_cache:
                self.search_cache = None

    def _partition_search(self, pops: List[POp]) -> List[Tuple[List[POp], List[int]]]:
        best = None
        # with tile refinement a trial costs ~0.25 s on 1000 operators of 30 bits: fewer randomised variants
        tries = min(self.tries, REFINE_TRIES) if self.refine else self.tries
        for trial in range(1 + tries):
            rnd = random.Random(trial) if trial else None
            p_new = 1.0 if trial == 0 else (0.9 if trial % 3 else 0.8)
            parts: List[Tuple[List[POp], List[int]]] = []
            remaining = list(pops)
            while remaining:
                chosen, remaining, tile = self._form_sweep(remaining, rnd, p_new)
                if not chosen:
                    raise RuntimeError('planner made no progress')
                parts.append((chosen, tile))
                if best is not None and len(parts) >= len(best):
                    break
            else:
                best = parts
            if len(pops) < 64:
                break
        if self.refine and len(pops) >= 64:
            # the two-sweep look-ahead costs ~100x the plain search: a few partitions only, and only a partition
            # with FEWER sweeps replaces the one found above
            for trial in range(LOOKAHEAD_TRIES):
                rnd = random.Random(1000 + trial) if trial else None
                parts = []
                remaining = list(pops)
                while remaining and len(parts) < len(best) - 1:
                    chosen, remaining, tile = self._form_sweep(remaining, rnd, 1.0 if trial == 0 else 0.9,
                                                               lookahead=True)
                    parts.append((chosen, tile))
                if not remaining:
                    best = parts
        return best

    def plan(self, pops: List[POp], parts: List[Tuple[List[POp], List[int]]] = None) -> List[SweepPlan]:
        """Sweeps for the operator list. `parts` is a split made elsewhere -- (operators, tile bits) per sweep in
        order, covering `pops` (sharded.schedule forms the sweeps itself, and knows which tile must hold the bits
        of the next remap's permutation); without it the list is partitioned here."""
        sweeps: List[SweepPlan] = []
        scale: Dict[int, complex] = {}     # pending relative scales (absorb_frame), carried across sweeps
        if parts is None:
            parts = self._partition(pops) if pops else []
        else:
            parts = [(list(chosen), sorted(tile)) for chosen, tile in parts]
            if sum(len(chosen) for chosen, _ in parts) != len(pops):
                raise ValueError('preset sweeps do not cover the operator list')
            for chosen, tile in parts:
                if len(set(tile)) != self.M or tile[:self.L] != list(range(self.L)) or tile[-1] >= self.nbits or \
                        any(not op.mixset <= set(tile) for op in chosen):
                    raise ValueError('preset sweep does not fit its tile')
        if os.environ.get('QFB_PLAN_SINK', '1') != '0':          # experiments: 0 keeps every phase term in its sweep
            parts = [(chosen, tile) for chosen, tile in sink_phase_terms(parts) if chosen]
        for index, (chosen, tile) in enumerate(parts):
            remaining = index + 1 < len(parts)
            ops, store_xor = absorb_frame(chosen, scale)
            assert all(b in tile for b in range(self.nbits) if (store_xor >> b) & 1)
            if not remaining:
                # end of the plan: the pending scales become phase terms (they act before the final store)
                for b in sorted(scale):
                    if scale[b] != 1:
                        ops.append(POp('P', dbits=[b], mat=complex(scale[b]), cost=COST['P']))
                scale.clear()
            for b in list(scale):
                if (store_xor >> b) & 1:
                    # the final store swaps the two halves of bit b: diag(1, s) becomes diag(s, 1) = s diag(1, 1/s)
                    ops.append(POp('P', dbits=[], mat=complex(scale[b]), cost=0.0))
                    scale[b] = 1.0 / scale[b]
            sweep = SweepPlan(tile, merge_phase_terms(ops), store_xor)
            self._form_rounds(sweep)
            sweeps.append(sweep)
        return sweeps

    # ---- pass 4: encoding -------------------------------------------------------------------------
    @staticmethod
    def _emit_phase(bits: Sequence[int], factor: complex, pos_of, reg_of) -> bytes:
        reg_cmask = 0
        idx_cmask = 0
        for bit in bits:
            p = pos_of.get(bit)
            ri = reg_of.get(p) if p is not None else None
            if ri is None:
                idx_cmask |= 1 << bit
            else:
                reg_cmask |= 1 << ri
        factor = complex(factor)
        regs = [i for i in range(HANDLER_STRIDE) if (reg_cmask >> i) & 1]
        if not regs:
            handler = H_CPH_SCALAR
        elif factor == -1:
            handler = (H_CPH_NEG1 + regs[0]) if len(regs) == 1 else \
                (H_CPH_NEG2 + G2_PAIRS.index((regs[1], regs[0]))) if len(regs) == 2 else H_CPH_NEGM
        elif len(regs) == 1:
            handler = (H_CPH_RSC1 if factor.imag == 0 else H_CPH_REG1) + regs[0]    # real scale: 2 FP64 per amplitude
        else:
            handler = H_CPH_REGM
        return _op_record(handler, reg_cmask, idx_cmask, struct.pack('<dd', factor.real, factor.imag))

    @staticmethod
    def _emit_gate(op: POp, pos_of, reg_of) -> Tuple[bytes, Optional[complex]]:
        def reg_index(bit: int) -> Optional[int]:
            p = pos_of.get(bit)
            return reg_of.get(p) if p is not None else None

        reg_cmask = 0
        idx_cmask = 0
        for c in op.ctrl:
            ri = reg_index(c)
            if ri is None:
                idx_cmask |= 1 << c
            else:
                reg_cmask |= 1 << ri
        if len(op.mix) == 1:
            if op.enc is not None:
                kind, payload, pivot = op.enc
            else:
                kind, payload, pivot, pending = encode_g1(op.mat, bool(op.ctrl))
                if pending != 1:
                    raise RuntimeError('an operator with a pending scale must go through absorb_scales')
            j = reg_index(op.mix[0])
            if op.ctrl:
                if kind == K_SWAPX:
                    # payload (1.0, 0): the kernel swaps by multiplying with it (FP64-pipe moves)
                    return _op_record(H_G1C_SWAPX + j, reg_cmask, idx_cmask, struct.pack('<dd', 1.0, 0.0)), None
                return _op_record(H_G1C_GENERAL + j, reg_cmask, idx_cmask, payload.tobytes()), None
            if kind == K_SWAPX:
                raise RuntimeError('uncontrolled X must have been absorbed into the flip mask')
            handler = {K_GENERAL: H_G1_GENERAL, K_SUMDIFF: H_G1_SUMDIFF, K_LU_R: H_G1_LU_R, K_LU_I: H_G1_LU_I}[kind]
            return _op_record(handler + j, reg_cmask, idx_cmask, payload.tobytes()), pivot
        j0, j1 = reg_index(op.mix[0]), reg_index(op.mix[1])
        mat = np.ascontiguousarray(op.mat, dtype=np.complex128).reshape(2, 2, 2, 2)
        if j0 < j1:   # kernel wants the operator's MSB qubit on the higher register bit
            mat = mat.transpose(1, 0, 3, 2)
            j0, j1 = j1, j0
        mat = np.ascontiguousarray(mat).reshape(4, 4)
        if is_real_x_shaped(mat):
            m = mat.real
            payload = struct.pack('<8d', m[0, 0], m[0, 3], m[3, 0], m[3, 3], m[1, 1], m[1, 2], m[2, 1], m[2, 2])
            return _op_record(H_G2X + G2_PAIRS.index((j0, j1)), reg_cmask, idx_cmask, payload), None
        nz = 0
        for r in range(4):
            for c in range(4):
                if mat[r, c] != 0:
                    nz |= 1 << (4 * r + c)
        return _op_record(H_G2 + G2_PAIRS.index((j0, j1)), reg_cmask, idx_cmask,
                          mat.tobytes() + struct.pack('<I12x', nz)), None

    @staticmethod
    def _thread_luts(ipos: Sequence[int], thr: Sequence[int]) -> bytes:
        """lut_lo[v] deposits thread bits 0..3 of v, lut_hi[v] thread bits 4.. of (v << 4); ipos[p] = index-bit
        image of tile position p."""
        def entry(value: int, first: int) -> bytes:
            tb = tg = 0
            for t, p in enumerate(thr):
                if t >= first and t < first + (4 if first == 0 else 8) and (value >> (t - first)) & 1:
                    tb |= 1 << p
                    tg |= 1 << ipos[p]
            return struct.pack('<IIQ', swz(tb) << 4, tb, tg)

        return b''.join(entry(v, 0) for v in range(16)) + b''.join(entry(v, 4) for v in range(32))

    @staticmethod
    def _emit_table(op: POp, pos_of, reg_of, scalar: complex = 1.0) -> bytes:
        """Diagonal table over the register index: entry e = product of the factors of the terms whose bits are
        all set in e. With `scalar` (the plan's uniform factor) every entry is multiplied by it and the record
        is flagged to act on ALL amplitudes (entry 0 included)."""
        masks = []
        for bits, factor in op.terms:
            mask = 0
            for b in bits:
                mask |= 1 << reg_of[pos_of[b]]
            masks.append((mask, complex(factor)))
        union = 0
        for mask, _ in masks:
            union |= mask
        entries = np.ones(TABLE_ENTRIES, dtype=np.complex128)
        for e in range(TABLE_ENTRIES):
            for mask, factor in masks:
                if (e & mask) == mask:
                    entries[e] *= factor
        whole = scalar != 1
        if whole:
            entries *= complex(scalar)
        return _op_record(H_CPH_TABLE, union, 0, entries.tobytes(), flag=int(whole))

    def serialise(self, sweeps: List[SweepPlan]) -> bytes:
        body = b''
        carry = 1.0 + 0j       # uniform factor (pivots, global phases) of the sweeps so far, applied once
        for si, sweep in enumerate(sweeps):
            pos_of = {b: j for j, b in enumerate(sweep.tile)}
            encoded: List[Tuple[Round, List[object]]] = []
            scalar = 1.0 + 0j
            for rd in sweep.rounds:
                reg_of = {p: i for i, p in enumerate(rd.regs)}
                blobs: List[object] = []
                for op in rd.ops:
                    if op.kind == 'T':
                        blobs.append((op, reg_of))      # encoded below, once the sweep's uniform factor is known
                    elif op.kind == 'P':
                        if not op.dbits:          # global phase: joins the sweep scalar
                            scalar *= complex(op.mat)
                            continue
                        blobs.append(self._emit_phase(op.dbits, op.mat, pos_of, reg_of))
                    else:
                        blob, pivot = self._emit_gate(op, pos_of, reg_of)
                        if pivot is not None:
                            scalar *= pivot
                        blobs.append(blob)
                encoded.append((rd, blobs))
            # The uniform factor commutes with everything: it rides along from sweep to sweep and is applied once,
            # by the plan's last sweep (or earlier, when it leaves a range that is safe for the stored amplitudes'
            # exponents), inside a diagonal table if the sweep has one -- otherwise as a per-thread scalar term.
            carry *= scalar
            last = si + 1 == len(sweeps) or os.environ.get('QFB_PLAN_CARRY', '1') == '0'
            due = carry != 1 and (last or not 2.0 ** -200 < abs(carry) < 2.0 ** 200)
            if due:
                host = next(((i, k) for i, (_, bl) in enumerate(encoded) for k, b in enumerate(bl)
                             if isinstance(b, tuple)), None)
                if host is not None:
                    i, k = host
                    op, reg_of = encoded[i][1][k]
                    encoded[i][1][k] = self._emit_table(op, pos_of, reg_of, carry)
                else:
                    target = next((i for i, (_, bl) in enumerate(encoded)
                                   if any(isinstance(b, bytes) and is_scalar_term(b) for b in bl)), len(encoded) - 1)
                    encoded[target][1].append(self._emit_phase((), carry, pos_of, {}))
                carry = 1.0 + 0j
            for _, bl in encoded:
                for k, b in enumerate(bl):
                    if isinstance(b, tuple):
                        bl[k] = self._emit_table(b[0], pos_of, b[1])
            rounds_blob = b''
            nops = 0
            any_g2 = False
            for rd, blobs in encoded:
                ops_blob = b''.join(blobs) + _op_record(H_END, 0, 0)
                nops += len(blobs)
                has_scalar = int(any(is_scalar_term(b) for b in blobs))
                has_g2 = int(any(H_G2 <= struct.unpack_from('<I', b, 0)[0] < H_G2X + len(G2_PAIRS) for b in blobs))
                any_g2 = any_g2 or bool(has_g2)
                thrpad = list(rd.thr) + [0] * (12 - len(rd.thr))
                rgb = [16 << sweep.tile[p] for p in rd.regs]
                rst = [-v if (sweep.store_xor >> sweep.tile[p]) & 1 else v for v, p in zip(rgb, rd.regs)]
                rounds_blob += _round_header(len(blobs), ROUND_HEADER_BYTES + len(ops_blob), rd.regs, thrpad,
                                             has_scalar, has_g2, rgb, rst)
                rounds_blob += self._thread_luts(sweep.tile, rd.thr) + ops_blob
            perm = sweep.spos != list(sweep.tile)
            if perm:
                # header-only store record: the last round's bit assignment with the images under spos
                rd = sweep.rounds[-1]
                thrpad = list(rd.thr) + [0] * (12 - len(rd.thr))
                rgb = [16 << sweep.tile[p] for p in rd.regs]
                rst = [(-1 if (sweep.store_xor >> sweep.spos[p]) & 1 else 1) * (16 << sweep.spos[p]) for p in rd.regs]
                rounds_blob += _round_header(0, ROUND_HEADER_BYTES + 16, rd.regs, thrpad, 0, 0, rgb, rst)
                rounds_blob += self._thread_luts(sweep.spos, rd.thr) + _op_record(H_END, 0, 0)
            holes = [b for b in range(self.nbits) if b not in sweep.tile]
            gpos = list(sweep.tile) + [0] * (16 - len(sweep.tile))
            hole = holes + [0] * (MAX_HOLES - len(holes))
            size = SWEEP_HEADER_BYTES + len(rounds_blob)
            if size > MAX_SWEEP_BYTES:
                raise RuntimeError('sweep record too large ({} bytes)'.format(size))
            flags = (SWEEP_FLAG_G2 if any_g2 else 0) | (SWEEP_FLAG_STORE_PERM if perm else 0) | \
                (SWEEP_FLAG_STORE_SYNC if ((sweep.store_xor or perm) and len(sweep.rounds) == 1) else 0)
            spos = list(sweep.spos) + [0] * (16 - len(sweep.spos))
            body += struct.pack('<IIII16B48BQ8x16B', size, len(sweep.rounds), nops, flags, *gpos, *hole,
                                sweep.store_xor, *spos) + rounds_blob
        total = 32 + len(body)
        header = struct.pack('<IIIIIIQ', PLAN_MAGIC, PLAN_VERSION, self.nbits, self.M, self.R, len(sweeps), total)
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


def classify_all(bitops: Sequence[Tuple[np.ndarray, Sequence[int]]]) -> List[object]:
    """(matrix, bits) operators -> POp / Fallback items in program order (identities dropped)."""
    items: List[object] = []
    for gi, (mat, bits) in enumerate(bitops):
        item = classify_op(np.asarray(mat), list(bits), gi)
        if item is None:
            continue
        if isinstance(item, Fallback):
            items.append(item)
        else:
            items.extend(item)
    return items


def remap_item(item, phys_of: Sequence[int]):
    """The same POp / Fallback with every bit b replaced by phys_of[b]."""
    if isinstance(item, Fallback):
        return Fallback(item.mat, [phys_of[b] for b in item.bits])
    return POp(item.kind, mix=[phys_of[b] for b in item.mix], ctrl=[phys_of[b] for b in item.ctrl],
               dbits=[phys_of[b] for b in item.dbits], mat=item.mat, cost=item.cost, gate_index=item.gate_index)


def item_bitop(item) -> Tuple[np.ndarray, List[int]]:
    """(matrix, bits) of a POp / Fallback: what a reference executor (tests) applies."""
    if isinstance(item, Fallback):
        return item.mat, list(item.bits)
    if item.kind == 'P':
        k = len(item.dbits)
        diag = np.ones(1 << k, dtype=np.complex128)
        diag[-1] = item.mat
        return np.diag(diag), list(item.dbits)
    k, c = len(item.mix), len(item.ctrl)
    full = np.eye(1 << (k + c), dtype=np.complex128)
    full[-(1 << k):, -(1 << k):] = np.asarray(item.mat, dtype=np.complex128).reshape(1 << k, 1 << k)
    return full, list(item.ctrl) + list(item.mix)


def build_segments(nbits: int, bitops: Sequence[Tuple[np.ndarray, Sequence[int]]], tile_bits: int = None,
                   low_bits: int = None, max_cost: float = None, final_perm: Sequence[int] = None,
                   reg_bits: int = None) -> List[Segment]:
    """Plan a list of (matrix, bits) operators for a state with `nbits` local index bits. `final_perm` (destination
    bit j <- source bit final_perm[j]) is an in-place bit permutation executed after the last operator, fused into
    the last sweep when its tile holds the moved bits (the local half of a qubit remap, sharded.py)."""
    return build_segments_from_items(nbits, classify_all(bitops), tile_bits, low_bits, max_cost, final_perm,
                                     reg_bits=reg_bits)


def build_segments_from_items(nbits: int, items: Sequence[object], tile_bits: int = None, low_bits: int = None,
                              max_cost: float = None, final_perm: Sequence[int] = None,
                              preset: Sequence[Tuple[int, Optional[Sequence[int]]]] = None,
                              reg_bits: int = None) -> List[Segment]:
    """build_segments for operators that are already classified (POp / Fallback, see classify_all). `preset`
    fixes the split into sweeps: (number of items, tile bits) per sweep in order, (1, None) for a Fallback."""
    planner = Planner(nbits, tile_bits, low_bits, max_cost, reg_bits=reg_bits)
    segments: List[Segment] = []
    pending: List[POp] = []
    groups: List[Tuple[List[POp], List[int]]] = []      # preset sweeps, consumed from the end as items stream by
    if final_perm is not None and list(final_perm) == list(range(nbits)):
        final_perm = None
    if preset is not None:
        if sum(count for count, _ in preset) != len(items):
            raise ValueError('preset sweeps do not cover the items')
        at = 0
        for count, tile in preset:
            if tile is not None:
                groups.append((list(items[at:at + count]), list(tile)))
            at += count
        groups.reverse()

    def flush(last: bool = False):
        if pending or (last and final_perm is not None):
            parts = None
            if preset is not None:
                parts, need = [], len(pending)
                while need > 0:
                    parts.append(groups.pop())
                    need -= len(parts[-1][0])
                if need != 0:
                    raise ValueError('preset sweeps do not line up with the Fallback operators')
            sweeps = planner.plan(pending, parts) if pending else []
            if last and final_perm is not None:
                planner.attach_permutation(sweeps, final_perm)
            blob = planner.serialise(sweeps)
            segments.append(Segment('plan', blob=blob, nsweeps=len(sweeps),
                                    nops=len({op.gate_index for op in pending}),
                                    nrounds=sum(len(s.rounds) for s in sweeps)))
            pending.clear()

    for item in items:
        if isinstance(item, Fallback):
            flush()
            segments.append(Segment('op', mat=item.mat, bits=item.bits, nsweeps=1, nops=1))
        else:
            pending.append(item)
    flush(last=True)
    return segments


def plan_stats(segments: Sequence[Segment]) -> Dict[str, int]:
    return {'segments': len(segments), 'sweeps': sum(s.nsweeps for s in segments),
            'ops': sum(s.nops for s in segments), 'rounds': sum(s.nrounds for s in segments),
            'plan_bytes': sum(len(s.blob) for s in segments if s.blob)}
