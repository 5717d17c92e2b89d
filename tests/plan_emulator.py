"""CPU emulator of the binary plan format (csrc/qfb_plan.h) -- TEST INFRASTRUCTURE ONLY.

Walks a plan exactly the way sweep_kernel does (tiles, rounds, register/thread bit assignment, control masks,
diagonal tables) on a numpy vector, so that the planner and the serialiser can be verified without a GPU.
It also reports the shared-memory bank-conflict degree of every round's LDS/STS pattern under the kernel's
XOR swizzle.
"""
import struct

import numpy as np

R = 5              # register bits of the plan being parsed / executed (parse() sets it from the header: 4 or 5)
NE = 1 << R
HS = 5             # handler ids are laid out for 5 register bits
TABLE = 32         # entries of a diagonal table record


def swz(idx):
    x = idx >> 3
    return idx ^ ((x ^ (x >> 3) ^ (x >> 6) ^ (x >> 9)) & 7)


G2_PAIRS = [(j0, j1) for j0 in range(HS) for j1 in range(j0)]
(H_G1_GENERAL, H_G1_SUMDIFF, H_G1_ROT_R, H_G1_ROT_I, H_G1C_GENERAL, H_G1C_SWAPX, H_CPH_SCALAR, H_CPH_REG1,
 H_CPH_RSC1, H_CPH_NEG1, H_CPH_NEG2, H_CPH_REGM, H_CPH_NEGM, H_END, H_G2, H_G2X, H_CPH_TABLE) = \
    0, 5, 10, 15, 20, 25, 30, 31, 36, 41, 46, 56, 57, 58, 59, 69, 79
SWEEP_HEADER, ROUND_HEADER = 112, 192 + 768


def parse(blob: bytes):
    global R, NE
    magic, version, nbits, M, rbits, nsweeps, total = struct.unpack_from('<IIIIIIQ', blob, 0)
    assert magic == 0x50424651 and version == 13 and rbits in (3, 4, 5) and total == len(blob)
    R, NE = rbits, 1 << rbits
    off = 32
    sweeps = []
    for _ in range(nsweeps):
        size, nrounds, nops, flags = struct.unpack_from('<IIII', blob, off)
        gpos = list(blob[off + 16: off + 16 + M])
        hole = list(blob[off + 32: off + 32 + (nbits - M)])
        store_xor = struct.unpack_from('<Q', blob, off + 80)[0]
        spos = list(blob[off + 96: off + 96 + M])
        assert sorted(spos) == sorted(gpos), 'spos must permute the tile bits'
        perm = spos != gpos
        assert all((store_xor >> b) & 1 == 0 for b in hole), 'store_xor leaves the tile'
        assert bool(flags & 2) == ((store_xor != 0 or perm) and nrounds == 1)
        assert bool(flags & 4) == perm
        roff = off + SWEEP_HEADER
        rounds = []
        for _r in range(nrounds + int(perm)):
            ipos = spos if _r == nrounds else gpos
            rn, rbytes = struct.unpack_from('<II', blob, roff)
            regpos = list(blob[roff + 8: roff + 8 + R])
            thrpos = list(blob[roff + 16: roff + 16 + (M - R)])
            has_scalar, has_g2 = blob[roff + 28], blob[roff + 29]
            ps_b = struct.unpack_from('<8I', blob, roff + 32)
            rgb = struct.unpack_from('<8q', blob, roff + 64)
            rst = struct.unpack_from('<8q', blob, roff + 128)
            for i in range(R):
                g, sg = 16 << gpos[regpos[i]], 16 << ipos[regpos[i]]
                assert ps_b[i] == swz(1 << regpos[i]) << 4 and rgb[i] == g
                assert rst[i] == (-sg if (store_xor >> ipos[regpos[i]]) & 1 else sg)
            # thread LUTs (16 + 32 entries of <IIQ): must reproduce the deposit of the thread bits
            lut = [struct.unpack_from('<IIQ', blob, roff + 192 + 16 * i) for i in range(48)]
            for tid in range(1 << (M - R)):
                tb = tg = 0
                for t in range(M - R):
                    if (tid >> t) & 1:
                        tb |= 1 << thrpos[t]
                        tg |= 1 << ipos[thrpos[t]]
                lo, hi = lut[tid & 15], lut[16 + ((tid >> 4) & 31)]
                assert (lo[1] | hi[1], lo[2] | hi[2], lo[0] ^ hi[0]) == (tb, tg, swz(tb) << 4), 'bad thread LUT'
            ooff = roff + ROUND_HEADER
            ops = []
            for _o in range(rn + 1):
                handler, obytes, rcm, flag, icm = struct.unpack_from('<IHBBQ', blob, ooff)
                payload = blob[ooff + 16: ooff + obytes]
                ooff += obytes
                if _o == rn:
                    assert handler == H_END and obytes == 16, 'round must end with an END record'
                    break
                if handler == H_CPH_TABLE:
                    # diagonal table over the register index; flag 1 = acts on every amplitude (entry 0 included)
                    assert obytes == 16 + 16 * TABLE and icm == 0 and flag in (0, 1) and (rcm != 0 or flag == 1)
                    typ, kind, j0, j1 = 4, 'table', 0, 0
                    tbl = np.frombuffer(payload, dtype=np.complex128, count=TABLE)
                    assert rcm < NE
                    if not flag:
                        assert all(tbl[e] == tbl[e & rcm] for e in range(NE)) and tbl[0] == 1
                elif handler < H_G1C_GENERAL:
                    assert rcm == 0 and icm == 0
                    kind = ['general', 'sumdiff', 'rot_r', 'rot_i'][handler // HS]
                    assert obytes == (80 if kind == 'general' else 32)
                    typ, j0, j1 = 1, handler % HS, 0
                    assert j0 < R
                elif handler < H_CPH_SCALAR:
                    kind = 'general' if handler < H_G1C_SWAPX else 'swapx'
                    assert obytes == (80 if kind == 'general' else 32)
                    typ, j0, j1 = 1, handler % HS, 0
                    assert j0 < R and not (rcm >> j0) & 1
                elif handler < H_END:
                    typ, j0, j1 = 3, 0, 0
                    assert obytes == 32
                    kind = 'neg' if handler in range(H_CPH_NEG1, H_CPH_REGM) or handler == H_CPH_NEGM else 'mul'
                    if handler == H_CPH_SCALAR:
                        assert rcm == 0
                    elif handler < H_CPH_RSC1:
                        assert rcm == 1 << (handler - H_CPH_REG1)
                    elif handler < H_CPH_NEG1:
                        assert rcm == 1 << (handler - H_CPH_RSC1)
                        assert struct.unpack_from('<dd', payload, 0)[1] == 0.0
                    elif handler < H_CPH_NEG2:
                        assert rcm == 1 << (handler - H_CPH_NEG1)
                    elif handler < H_CPH_REGM:
                        q0, q1 = G2_PAIRS[handler - H_CPH_NEG2]
                        assert rcm == (1 << q0) | (1 << q1)
                    else:
                        assert 0 < rcm < NE
                elif handler >= H_G2X:
                    assert handler < H_G2X + len(G2_PAIRS) and obytes == 16 + 64
                    typ, kind = 2, 'xshape'
                    j0, j1 = G2_PAIRS[handler - H_G2X]
                    v = struct.unpack_from('<8d', payload, 0)
                    m = np.zeros((4, 4), dtype=np.complex128)
                    m[0, 0], m[0, 3], m[3, 0], m[3, 3], m[1, 1], m[1, 2], m[2, 1], m[2, 2] = v
                    nzm = sum(1 << (4 * r + c) for r in range(4) for c in range(4) if m[r, c] != 0)
                    payload = m.tobytes() + struct.pack('<I12x', nzm)      # what _apply_g2 reads
                else:
                    assert H_G2 <= handler < H_G2 + len(G2_PAIRS) and obytes == 16 + 272
                    typ, kind = 2, 'dense'
                    j0, j1 = G2_PAIRS[handler - H_G2]
                ops.append(dict(type=typ, kind=kind, j0=j0, j1=j1, reg_cmask=rcm, idx_cmask=icm, payload=payload,
                                flag=flag))
            assert ooff == roff + rbytes
            assert has_g2 == int(any(o['type'] == 2 for o in ops))
            if _r == nrounds:
                assert rn == 0 and (regpos, thrpos) == (rounds[-1]['regpos'], rounds[-1]['thrpos']), 'bad store record'
            else:
                rounds.append(dict(regpos=regpos, thrpos=thrpos, ops=ops, has_scalar=has_scalar))
            roff += rbytes
        assert roff == off + size
        assert bool(flags & 1) == any(o['type'] == 2 for rd in rounds for o in rd['ops'])
        sweeps.append(dict(gpos=gpos, spos=spos, hole=hole, rounds=rounds, store_xor=store_xor))
        off += size
    assert off == len(blob)
    return dict(nbits=nbits, M=M, sweeps=sweeps)


def _apply_g1(a, op):
    j, rc, kind = op['j0'], op['reg_cmask'], op['kind']
    if kind == 'general':
        m = np.frombuffer(op['payload'], dtype=np.complex128, count=4).reshape(2, 2)
    elif kind != 'swapx':
        c0, c1 = struct.unpack_from('<dd', op['payload'], 0)
    for p in range(NE // 2):
        e0 = ((p >> j) << (j + 1)) | (p & ((1 << j) - 1))
        e1 = e0 | (1 << j)
        if (e0 & rc) != rc:
            continue
        x, y = a[e0], a[e1]
        if kind == 'swapx':
            a[e0], a[e1] = y, x
        elif kind == 'sumdiff':     # pivoted: x' = x + r0 y, y' = x' + (r1 - r0) y
            a[e0] = x + c0 * y
            a[e1] = a[e0] + (c1 - c0) * y      # formed from x' like the kernel does
        elif kind == 'rot_r':       # LU_R, two shears: x += a y; y += b x
            x = x + c0 * y
            y = y + c1 * x
            a[e0], a[e1] = x, y
        elif kind == 'rot_i':       # LU_I: x += i a y; y += i b x
            x = x + 1j * c0 * y
            y = y + 1j * c1 * x
            a[e0], a[e1] = x, y
        else:
            a[e0] = m[0, 0] * x + m[0, 1] * y
            a[e1] = m[1, 0] * x + m[1, 1] * y


def _apply_g2(a, op):
    m = np.frombuffer(op['payload'], dtype=np.complex128, count=16).reshape(4, 4)
    nz = struct.unpack_from('<I', op['payload'], 256)[0]
    j0, j1, rc = op['j0'], op['j1'], op['reg_cmask']
    assert j0 > j1
    others = [b for b in range(R) if b not in (j0, j1)]
    for g in range(1 << len(others)):
        eb = sum(((g >> i) & 1) << b for i, b in enumerate(others))
        if (eb & rc) != rc:
            continue
        ids = [eb, eb | (1 << j1), eb | (1 << j0), eb | (1 << j0) | (1 << j1)]
        vin = [a[i] for i in ids]
        for r in range(4):
            acc = 0j
            for c in range(4):
                assert ((nz >> (4 * r + c)) & 1) == int(m[r, c] != 0)
                acc += m[r, c] * vin[c]
            a[ids[r]] = acc


def _apply_cph(a, op, on, scalar):
    """Returns the updated per-thread scalar."""
    factor = complex(*struct.unpack_from('<dd', op['payload'], 0))
    rc = op['reg_cmask']
    if not on:
        return scalar
    if rc == 0:
        return scalar * factor
    for e in range(NE):
        if (e & rc) == rc:
            if op['kind'] == 'neg':
                assert factor == -1
                a[e] = -a[e]
            else:
                a[e] = factor * a[e]
    return scalar


def conflict_degree(plan_round, M):
    """Worst quarter-warp bank-conflict degree (1 = conflict free) of the round's LDS/STS.128 pattern."""
    T = 1 << (M - R)
    worst = 1
    regpos, thrpos = plan_round['regpos'], plan_round['thrpos']
    for e in (0, NE - 1):
        off = 0
        for i in range(R):
            if (e >> i) & 1:
                off |= 1 << regpos[i]
        for q0 in range(0, min(T, 64), 8):
            groups = {}
            for tid in range(q0, min(q0 + 8, T)):
                tb = 0
                for t in range(M - R):
                    tb |= ((tid >> t) & 1) << thrpos[t]
                quad = swz(tb | off) & 7
                groups[quad] = groups.get(quad, 0) + 1
            worst = max(worst, max(groups.values()))
    return worst


def _run_tile(plan, sweep, tile_id: int, state, index_hi: int, check_layout: bool) -> None:
    """One tile of one sweep, in place on `state` (anything indexable by flat amplitude index: a numpy vector, or a
    dict-backed sparse memory when only a few tiles of a large state are of interest)."""
    nbits, M = plan['nbits'], plan['M']
    T = 1 << (M - R)
    gpos, spos, hole = sweep['gpos'], sweep['spos'], sweep['hole']
    nrounds = len(sweep['rounds'])
    gb = 0
    for i, h in enumerate(hole):
        gb |= ((tile_id >> i) & 1) << h
    tile = np.zeros(1 << M, dtype=np.complex128)   # "shared memory", indexed by swizzled tile index
    # asynchronous tile loader: tile-local index i <- state[gb | deposit(i through gpos)]
    for i in range(1 << M):
        g = 0
        for j in range(M):
            g |= ((i >> j) & 1) << gpos[j]
        tile[swz(i)] = state[gb | g]
    for rnd, rd in enumerate(sweep['rounds']):
        regpos, thrpos = rd['regpos'], rd['thrpos']
        assert sorted(regpos + thrpos) == list(range(M))
        if check_layout and (rnd == 0 or rnd == nrounds - 1):
            # edge rounds: lanes must walk the lowest index bits (coalesced 128-byte lines)
            nlow = min(3, M - R)
            if rnd == 0:
                assert [gpos[thrpos[t]] for t in range(nlow)] == list(range(nlow)), 'uncoalesced load round'
            if rnd == nrounds - 1:
                assert [spos[thrpos[t]] for t in range(nlow)] == list(range(nlow)), 'uncoalesced store round'
        regs_all = np.zeros((T, NE), dtype=np.complex128)
        for tid in range(T):
            tb = tg = 0
            for t in range(M - R):
                bit = (tid >> t) & 1
                tb |= bit << thrpos[t]
                tg |= bit << gpos[thrpos[t]]
            a = np.zeros(NE, dtype=np.complex128)
            for e in range(NE):
                toff = goff = 0
                for i in range(R):
                    if (e >> i) & 1:
                        toff |= 1 << regpos[i]
                        goff |= 1 << gpos[regpos[i]]
                a[e] = tile[swz(tb | toff)]
            tfull = (index_hi << nbits) | gb | tg
            scalar = 1.0 + 0j
            for op in rd['ops']:
                on = (tfull & op['idx_cmask']) == op['idx_cmask']
                if op['type'] == 4:
                    tbl = np.frombuffer(op['payload'], dtype=np.complex128, count=TABLE)
                    for e in range(NE):
                        if op['flag'] or (e & op['reg_cmask']):
                            a[e] = tbl[e] * a[e]
                elif op['type'] == 3:
                    assert op['reg_cmask'] != 0 or rd['has_scalar'] == 1
                    scalar = _apply_cph(a, op, on, scalar)
                elif on:
                    if op['type'] == 1:
                        _apply_g1(a, op)
                    else:
                        _apply_g2(a, op)
            if rd['has_scalar']:
                a = a * scalar
            else:
                assert scalar == 1
            regs_all[tid] = a
        # all threads have read the tile before anyone writes it (the kernel's barriers)
        staged = []
        for tid in range(T):
            tb = tg = 0
            for t in range(M - R):
                bit = (tid >> t) & 1
                tb |= bit << thrpos[t]
                tg |= bit << spos[thrpos[t]]       # only used by the final store: STORE positions
            for e in range(NE):
                toff = goff = 0
                for i in range(R):
                    if (e >> i) & 1:
                        toff |= 1 << regpos[i]
                        goff |= 1 << spos[regpos[i]]
                if rnd + 1 < nrounds:
                    tile[swz(tb | toff)] = regs_all[tid, e]
                else:
                    staged.append(((gb | tg | goff) ^ sweep['store_xor'], regs_all[tid, e]))
        if rnd + 1 == nrounds:
            # pending X flips: amplitude i lands at i ^ store_xor (all loads of the tile came first)
            for addr, val in staged:
                state[addr] = val


def execute(blob: bytes, state: np.ndarray, index_hi: int = 0, check_layout: bool = True) -> np.ndarray:
    plan = parse(blob)
    nbits, M = plan['nbits'], plan['M']
    assert state.size == 1 << nbits
    state = np.array(state, dtype=np.complex128).reshape(-1)
    for sweep in plan['sweeps']:
        assert sorted(sweep['gpos'] + sweep['hole']) == list(range(nbits))
        for tile_id in range(1 << (nbits - M)):
            _run_tile(plan, sweep, tile_id, state, index_hi, check_layout)
    return state


def execute_tiles(blob: bytes, sweep_index: int, tile_ids, memory, index_hi: int = 0) -> None:
    """The tiles `tile_ids` of sweep `sweep_index` only, in place on `memory` (flat amplitude index -> complex; a
    mapping that returns 0 for absent keys works): what a few CTAs of the sweep's kernel do to a state that is too
    large to emulate as a whole."""
    plan = parse(blob)
    sweep = plan['sweeps'][sweep_index]
    for tile_id in tile_ids:
        _run_tile(plan, sweep, int(tile_id), memory, index_hi, True)


# ---------------------------------------------------------------------------------------------------------
# The planner's tile score and tile search in plain Python: the statement that csrc/qfb_planhost.cu
# (qfb_plan_count_executed / qfb_plan_refine_tile) is checked against (tests/test_planner.py).
# ---------------------------------------------------------------------------------------------------------

def count_executed(recs, tmask, fmask, max_cost, room):
    """recs: (mixmask, diagmask, cost, bytes) per operator in program order. Number of operators that touch a bit
    and that a sweep over the tile `tmask` executes."""
    allow = tmask & ~fmask
    da = dm = 0
    work = 0.0
    count = 0
    for mm, dd, c, nb in recs:
        if (mm & da) or (dd & dm) or (mm & ~allow) or (count and work + c > max_cost):
            da |= mm | dd
            dm |= mm
            if not (allow & ~da):
                break
            continue
        room -= nb
        if room < 0:
            break
        work += c
        if mm | dd:
            count += 1
    return count


def refine_tile(recs, nbits, tmask, fmask, keep, max_cost, room, passes):
    """Best single-bit exchange per pass while the count rises; returns (tile mask, count)."""
    best = count_executed(recs, tmask, fmask, max_cost, room)
    for _ in range(passes):
        base = tmask
        improved = False
        for bi in range(nbits):
            if not (base >> bi) & 1 or (keep >> bi) & 1:
                continue
            without = base & ~(1 << bi)
            for bo in range(nbits):
                if (base >> bo) & 1 or (fmask >> bo) & 1:
                    continue
                n = count_executed(recs, without | (1 << bo), fmask, max_cost, room)
                if n > best:
                    best, tmask, improved = n, without | (1 << bo), True
        if not improved:
            break
    return tmask, best
