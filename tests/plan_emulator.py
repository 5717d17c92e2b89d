"""CPU emulator of the binary plan format (csrc/qfb_plan.h) -- TEST INFRASTRUCTURE ONLY.

Walks a plan exactly the way sweep_kernel does (tiles, rounds, register/thread bit assignment, control masks,
diagonal tables) on a numpy vector, so that the planner and the serialiser can be verified without a GPU.
It also reports the shared-memory bank-conflict degree of every round's LDS/STS pattern under the kernel's
XOR swizzle.
"""
import struct

import numpy as np

R = 4
NE = 1 << R


def swz(idx):
    x = idx >> 3
    return idx ^ ((x ^ (x >> 3) ^ (x >> 6) ^ (x >> 9)) & 7)


def parse(blob: bytes):
    magic, version, nbits, M, rbits, nsweeps, total = struct.unpack_from('<IIIIIIQ', blob, 0)
    assert magic == 0x50424651 and version == 6 and rbits == R and total == len(blob)
    off = 32
    sweeps = []
    for _ in range(nsweeps):
        size, nrounds, nops, _ = struct.unpack_from('<IIII', blob, off)
        gpos = list(blob[off + 16: off + 16 + M])
        hole = list(blob[off + 32: off + 32 + (nbits - M)])
        roff = off + 80
        rounds = []
        for _r in range(nrounds):
            rn, rbytes = struct.unpack_from('<II', blob, roff)
            regpos = list(blob[roff + 8: roff + 12])
            thrpos = list(blob[roff + 12: roff + 12 + (M - R)])
            has_scalar, has_g2 = blob[roff + 24], blob[roff + 25]
            # thread LUTs (16 + 32 entries of <IIQ): must reproduce the deposit of the thread bits
            lut = [struct.unpack_from('<IIQ', blob, roff + 32 + 16 * i) for i in range(48)]
            for tid in range(1 << (M - R)):
                tb = tg = 0
                for t in range(M - R):
                    if (tid >> t) & 1:
                        tb |= 1 << thrpos[t]
                        tg |= 1 << gpos[thrpos[t]]
                lo, hi = lut[tid & 15], lut[16 + ((tid >> 4) & 31)]
                assert (lo[0] | hi[0], lo[2] | hi[2]) == (tb, tg), 'bad thread LUT'
            ooff = roff + 32 + 768
            ops = []
            for _o in range(rn + 1):
                handler, rcm, size16, _p0, _p1, icm = struct.unpack_from('<BBBBIQ', blob, ooff)
                obytes = 16 * size16
                payload = blob[ooff + 16: ooff + obytes]
                ooff += obytes
                if _o == rn:
                    assert handler == 37 and obytes == 16, 'round must end with an END record'
                    break
                if handler < 20:
                    assert rcm == 0 and icm == 0
                    typ, kind, j0, j1 = 1, [0, 3, 5, 6, 7][handler // 4], handler % 4, 0
                elif handler < 28:
                    typ, kind, j0, j1 = 1, (0 if handler < 24 else 3), handler % 4, 0
                elif handler in (28, 29, 30):
                    typ, kind, j0, j1 = 3, int(handler == 30), 0, 0
                    assert (handler == 28) == (rcm == 0)
                else:
                    assert 31 <= handler < 37
                    typ, kind = 2, 0
                    j0, j1 = [(1, 0), (2, 0), (2, 1), (3, 0), (3, 1), (3, 2)][handler - 31]
                ops.append(dict(type=typ, kind=kind, j0=j0, j1=j1, reg_cmask=rcm, idx_cmask=icm, payload=payload))
            assert ooff == roff + rbytes
            assert has_g2 == int(any(o['type'] == 2 for o in ops))
            rounds.append(dict(regpos=regpos, thrpos=thrpos, ops=ops, has_scalar=has_scalar))
            roff += rbytes
        assert roff == off + size
        sweeps.append(dict(gpos=gpos, hole=hole, rounds=rounds))
        off += size
    assert off == len(blob)
    return dict(nbits=nbits, M=M, sweeps=sweeps)


def _apply_g1(a, op):
    m = np.frombuffer(op['payload'], dtype=np.complex128, count=4).reshape(2, 2)
    j, rc = op['j0'], op['reg_cmask']
    for p in range(NE // 2):
        e0 = ((p >> j) << (j + 1)) | (p & ((1 << j) - 1))
        e1 = e0 | (1 << j)
        if (e0 & rc) != rc:
            continue
        x, y = a[e0], a[e1]
        if op['kind'] == 3:      # SWAPX
            a[e0], a[e1] = y, x
        elif op['kind'] == 1:    # REAL
            a[e0] = m[0, 0].real * x + m[0, 1].real * y
            a[e1] = m[1, 0].real * x + m[1, 1].real * y
        elif op['kind'] == 2:    # RXLIKE
            a[e0] = m[0, 0].real * x + 1j * m[0, 1].imag * y
            a[e1] = 1j * m[1, 0].imag * x + m[1, 1].real * y
        elif op['kind'] == 4:    # ANTIDIAG
            a[e0] = m[0, 1] * y
            a[e1] = m[1, 0] * x
        elif op['kind'] == 5:    # SUMDIFF (pivoted): x' = x + r0 y, y' = x + r1 y
            r0, r1 = m[0, 0].real, m[0, 0].imag
            assert abs(r0) == 1 and abs(r1) == 1
            a[e0] = x + r0 * y
            a[e1] = a[e0] + (r1 - r0) * y      # formed from x' like the kernel does
        elif op['kind'] == 6:    # ROT_R (pivoted): x' = x + r y, y' = y + s x
            a[e0] = x + m[0, 0].real * y
            a[e1] = y + m[0, 0].imag * x
        elif op['kind'] == 7:    # ROT_I (pivoted): x' = x + i a y, y' = y + i b x
            a[e0] = x + 1j * m[0, 0].real * y
            a[e1] = y + 1j * m[0, 0].imag * x
        else:
            a[e0] = m[0, 0] * x + m[0, 1] * y
            a[e1] = m[1, 0] * x + m[1, 1] * y


def _apply_g2(a, op):
    m = np.frombuffer(op['payload'], dtype=np.complex128, count=16).reshape(4, 4)
    nz = struct.unpack_from('<I', op['payload'], 256)[0]
    j0, j1, rc = op['j0'], op['j1'], op['reg_cmask']
    assert j0 > j1
    others = [b for b in range(R) if b not in (j0, j1)]
    for g in range(4):
        eb = ((g & 1) << others[0]) | ((g >> 1) << others[1])
        if (eb & rc) != rc:
            continue
        ids = [eb, eb | (1 << j1), eb | (1 << j0), eb | (1 << j0) | (1 << j1)]
        vin = [a[i] for i in ids]
        for r in range(4):
            acc = 0j
            for c in range(4):
                if (nz >> (4 * r + c)) & 1:
                    acc += m[r, c] * vin[c]
                else:
                    assert m[r, c] == 0
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
            if op['kind'] == 1:
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


def execute(blob: bytes, state: np.ndarray, index_hi: int = 0, check_layout: bool = True) -> np.ndarray:
    plan = parse(blob)
    nbits, M = plan['nbits'], plan['M']
    assert state.size == 1 << nbits
    state = np.array(state, dtype=np.complex128).reshape(-1)
    T = 1 << (M - R)
    for sweep in plan['sweeps']:
        gpos, hole = sweep['gpos'], sweep['hole']
        assert sorted(gpos + hole) == list(range(nbits))
        nrounds = len(sweep['rounds'])
        for tile_id in range(1 << (nbits - M)):
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
                    assert [gpos[thrpos[t]] for t in range(nlow)] == list(range(nlow)), 'uncoalesced edge round'
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
                        if op['type'] == 3:
                            assert op['reg_cmask'] != 0 or rd['has_scalar'] == 1
                            scalar = _apply_cph(a, op, on, scalar)
                        elif on:
                            if op['type'] == 1:
                                assert op['reg_cmask'] == 0 or op['kind'] in (0, 3)
                                _apply_g1(a, op)
                            else:
                                _apply_g2(a, op)
                    if rd['has_scalar']:
                        a = a * scalar
                    else:
                        assert scalar == 1
                    regs_all[tid] = a
                # all threads have read the tile before anyone writes it (the kernel's barriers)
                for tid in range(T):
                    tb = tg = 0
                    for t in range(M - R):
                        bit = (tid >> t) & 1
                        tb |= bit << thrpos[t]
                        tg |= bit << gpos[thrpos[t]]
                    for e in range(NE):
                        toff = goff = 0
                        for i in range(R):
                            if (e >> i) & 1:
                                toff |= 1 << regpos[i]
                                goff |= 1 << gpos[regpos[i]]
                        if rnd + 1 < nrounds:
                            tile[swz(tb | toff)] = regs_all[tid, e]
                        else:
                            state[gb | tg | goff] = regs_all[tid, e]
    return state
