/*
 * qfb_plan.h -- binary layout of an execution plan for the tiled multi-gate executor (qfb_run_plan).
 *
 * A plan is a list of SWEEPS. One sweep = one kernel launch = one read + one write of the whole state
 * (32 B per amplitude). In a sweep every CTA owns tiles of 2^M amplitudes: the M "tile bits" are index bit
 * positions chosen by the planner (gpos[], ascending; the low bits are always included so that every global
 * access is a full 128-byte line). A tile is processed in ROUNDS. In a round each thread holds 2^R amplitudes
 * in registers (R = 5: 32 amplitudes = 128 registers; the cost of an op is its arithmetic plus a fixed ~100-cycle
 * dispatch bubble per SM sub-partition, profiles/README.md, so the more amplitudes an op updates per thread the
 * better): R "register bits" (regpos[], tile-bit positions) enumerate the amplitudes of one thread and
 * the remaining M-R tile bits (thrpos[]) enumerate the threads. Ops of a round are
 *
 *   G1   dense 1-bit operator on a register bit, optionally controlled
 *   G2   dense 2-bit operator on two register bits, optionally controlled
 *   CPH  one term of a diagonal operator in phase-polynomial form: multiply the amplitude by `factor` when all
 *        bits of a mask are 1. Mask bits may be ANY bit of the full index (register bits, thread bits, tile-id
 *        bits, rank bits of a sharded state), so diagonal gates never constrain tiling. A term without register
 *        bits is a per-thread scalar: it is accumulated into one running factor and applied once per round.
 *
 * Controls are split the same way: `reg_cmask` over the register index, `idx_cmask` over the full index.
 * Between rounds the tile is exchanged through shared memory (XOR-swizzled, see qfb_sweep.cu). Round 0 loads
 * from HBM, the last round stores to HBM.
 *
 * Pauli-X gates are never executed: the planner tracks them as a pending bit-flip mask of the sweep, rewrites
 * the operators that follow (operator conjugated by X on its flipped bits) and hands the mask to the kernel as
 * `store_xor`: the last round stores amplitude i to address i ^ store_xor (tile bits only, so the permutation
 * stays inside the CTA's tile and inside whole 128-byte lines).
 *
 * All records are multiples of 16 bytes; integers little endian. The Python planner
 * (quantumflow_b200/planner.py) writes this layout with struct.pack; keep the two in sync.
 */
#ifndef QFB_PLAN_H
#define QFB_PLAN_H
#include <stdint.h>

#define QFB_PLAN_MAGIC 0x50424651u /* "QFBP" */
#define QFB_PLAN_VERSION 13u
#define QFB_PLAN_REG_BITS 5
#define QFB_PLAN_MAX_TILE_BITS 13
#define QFB_PLAN_MIN_TILE_BITS 6
#define QFB_PLAN_MAX_HOLES 48
#define QFB_PLAN_MAX_SWEEP_BYTES (40 * 1024)

typedef struct {
    uint32_t magic;
    uint32_t version;
    uint32_t nbits;     /* index bits of the (local) state the plan was built for */
    uint32_t tile_bits; /* M */
    uint32_t reg_bits;  /* R (= QFB_PLAN_REG_BITS) */
    uint32_t nsweeps;
    uint64_t total_bytes;
} qfb_plan_header; /* 32 bytes */

#define QFB_SWEEP_FLAG_G2 1u          /* some round holds G2 ops (selects the kernel variant) */
#define QFB_SWEEP_FLAG_STORE_SYNC 2u  /* single-round sweep that permutes on store: barrier between loads and stores */
#define QFB_SWEEP_FLAG_STORE_PERM 4u  /* spos != gpos: a header-only "store record" follows the last round */

typedef struct {
    uint32_t bytes; /* whole sweep record including this header */
    uint32_t nrounds;
    uint32_t nops; /* informational */
    uint32_t flags;
    uint8_t gpos[16]; /* gpos[j] = index bit position of tile bit j (ascending) */
    uint8_t hole[QFB_PLAN_MAX_HOLES]; /* hole[i] = index bit position of tile-id bit i (ascending) */
    uint64_t store_xor; /* pending X flips of the sweep, in STORE bit positions (subset of the tile bits) */
    uint8_t pad[8];
    uint8_t spos[16]; /* index bit position tile bit j is STORED to: a permutation of gpos[] (= gpos[] unless
                         QFB_SWEEP_FLAG_STORE_PERM). The sweep then also performs the in-place bit permutation
                         gpos[j] -> spos[j] of the state (qubit remap of a sharded state, no extra pass). The store
                         record is a round header with nops = 0 and an END op whose regpos / thrpos equal the last
                         round's and whose rst[] / thread LUT tg are the images under spos[]. */
} qfb_sweep_header; /* 112 bytes */

typedef struct {
    uint32_t stb; /* byte offset of the thread's first amplitude in the swizzled exchange buffer: swz(tb) << 4 */
    uint32_t tb;  /* tile-local index contribution (register bits zero) */
    uint64_t tg;  /* the same bits at their index-bit positions */
} qfb_thread_lut; /* 16 bytes */

#define QFB_PLAN_LUT_LO 16 /* entries indexed by the low 4 thread bits */
#define QFB_PLAN_LUT_HI 32 /* entries indexed by the remaining (M-R-4 <= 5) thread bits */

typedef struct {
    uint32_t nops;
    uint32_t bytes;     /* whole round record: header + thread LUTs + ops */
    uint8_t regpos[8];  /* tile-bit position of register bit i, i < R */
    uint8_t thrpos[12]; /* tile-bit position of thread bit t, t < M-R */
    uint8_t has_scalar; /* 1 when the round holds CPH terms without register bits */
    uint8_t has_g2;     /* 1 when the round holds G2 ops */
    uint8_t pad[2];
    uint32_t ps_b[8];   /* swz(1 << regpos[i]) << 4: byte offset of register bit i in the exchange buffer */
    int64_t rgb[8];     /* 16 << gpos[regpos[i]]: byte distance in the state between register bit i = 0 and 1 */
    int64_t rst[8];     /* the same for the final store: negative when store_xor flips that bit */
    /* thread id -> (stb, tb, tg) in two table look-ups instead of a per-bit deposit loop:
     * x = lut_lo[tid & 15].x ^ lut_hi[tid >> 4].x (the bit sets are disjoint and swz is linear over XOR) */
    qfb_thread_lut lut_lo[QFB_PLAN_LUT_LO];
    qfb_thread_lut lut_hi[QFB_PLAN_LUT_HI];
} qfb_round_header; /* 192 + 768 bytes */

/* kinds of dense 1-bit operators: structure of the 2x2 operator, chosen by the planner to save FP64 work and
 * register copies; (x, y) = the pair of amplitudes, every update is in place.
 *   GENERAL  16 FP64 per pair
 *   SWAPX    Pauli X: swap, no arithmetic (only as a controlled operator; uncontrolled X is store_xor)
 *   SUMDIFF  Hadamard-like divided by its (0,0) entry ("pivot"; the planner multiplies the pivots of a sweep into
 *            one uniform scalar that rides on a CPH term): x' = x + r0 y, y' = x' + (r1 - r0) y with r = +-1.
 *            Sums only, so destructive interference gives exact zeros like the reference's h*x + h*y. 4 per pair.
 *   LU_R     any real 2x2 operator G (RY, X.H.X, ...) in LDU form  G = p . diag(1, s) . [[1,0],[b,1]] . [[1,a],[0,1]]:
 *            the kernel runs the two shears  x += a y; y += b x  (4 per pair, in place, no temporaries), p joins
 *            the sweep scalar, and the relative scale s stays PENDING on that bit: the planner multiplies it into
 *            the next operator that mixes the bit (or emits it as a CPH term when it must).
 *   LU_I     the same for real diagonal / imaginary off-diagonal operators (RX): x += i a y; y += i b x. */
/* Handler ids: ONE jump table (brx.idx) in the kernel's op interpreter. j = register bit (0..R-1), pair = index
 * of (j0 > j1) in (1,0) (2,0) (2,1) (3,0) (3,1) (3,2) (4,0) (4,1) (4,2) (4,3).
 *   QFB_H_G1_GENERAL + j  uncontrolled dense 1-bit operator on register bit j (16 FP64 per pair)
 *   QFB_H_G1_SUMDIFF + j  pivoted Hadamard-like
 *   QFB_H_G1_LU_R + j     two real shears (any real operator, see above)
 *   QFB_H_G1_LU_I + j     two imaginary shears (RX-like operators)
 *   QFB_H_G1C_GENERAL + j controlled dense 1-bit operator   (reg_cmask / idx_cmask)
 *   QFB_H_G1C_SWAPX + j   controlled X (CNOT, CCNOT ...)
 *   QFB_H_CPH_SCALAR      phase term without register bits: accumulates into the round's scalar
 *   QFB_H_CPH_REG1 + j    phase term on register bit j (and idx_cmask)
 *   QFB_H_CPH_RSC1 + j    the same with a real factor (a pending LDU scale that had to be emitted): 2 FP64 per amplitude
 *   QFB_H_CPH_NEG1 + j    the same with factor -1 (sign flip, no FP64 work)
 *   QFB_H_CPH_NEG2 + pair factor -1 on two register bits (CZ between register bits)
 *   QFB_H_CPH_REGM / NEGM any other register mask (reg_cmask)
 *   QFB_H_CPH_TABLE       diagonal table over the register index (reg_cmask = the register bits it depends on):
 *                         amplitude e is multiplied by table[e] where e & reg_cmask != 0 -- all phase terms of a
 *                         round whose bits are register bits, multiplied together by the planner (4 FP64 per
 *                         touched amplitude for the whole group). With flag = 1 the table also carries the plan's
 *                         uniform factor and acts on every amplitude, entry 0 included.
 *   QFB_H_END             terminates the round's op list
 *   QFB_H_G2 + pair       dense 2-bit operator on register bits (j0, j1)
 *   QFB_H_G2X + pair      real "X-shaped" 2-bit operator: non-zeros only at (0,0) (0,3) (3,0) (3,3) and (1,1) (1,2)
 *                         (2,1) (2,2) -- the superoperator of every Pauli channel and of amplitude damping on
 *                         (ket bit, bra bit): 4 FP64 per amplitude instead of 16 */
enum {
    QFB_H_G1_GENERAL = 0,
    QFB_H_G1_SUMDIFF = 5,
    QFB_H_G1_LU_R = 10,
    QFB_H_G1_LU_I = 15,
    QFB_H_G1C_GENERAL = 20,
    QFB_H_G1C_SWAPX = 25,
    QFB_H_CPH_SCALAR = 30,
    QFB_H_CPH_REG1 = 31,
    QFB_H_CPH_RSC1 = 36,
    QFB_H_CPH_NEG1 = 41,
    QFB_H_CPH_NEG2 = 46,
    QFB_H_CPH_REGM = 56,
    QFB_H_CPH_NEGM = 57,
    QFB_H_END = 58,
    QFB_H_G2 = 59,
    QFB_H_G2X = 69,
    QFB_H_CPH_TABLE = 79,
    QFB_H_COUNT = 80
};

typedef struct {
    uint32_t handler;
    uint16_t bytes;    /* whole op record including this header (multiple of 16) */
    uint8_t reg_cmask; /* control / phase mask over the register index */
    uint8_t flag;      /* QFB_H_CPH_TABLE: 1 = acts on every amplitude; 0 elsewhere */
    uint64_t idx_cmask; /* control / phase mask over thread-level bits of the FULL index (incl. rank bits) */
} qfb_op_header; /* 16 bytes */

/* payloads (follow the header)
 *   G1 GENERAL / G1C_GENERAL: double m[8]   row-major 2x2 complex (64 B); G1C_SWAPX: double (1.0, 0) (16 B)
 *   G1 SUMDIFF: double r[2] = (r0, r1); LU_R / LU_I: double (a, b)   (16 B)
 *   G2X: double m[8] = m00 m03 m30 m33 m11 m12 m21 m22 (64 B)
 *   G2 : double m[32]; uint32 nzmask; uint32 pad[3]   row-major 4x4 complex, bit (4r+c) of nzmask set when
 *                      entry (r,c) is non-zero (272 B)
 *   CPH: double factor[2]  (16 B); CPH_TABLE: double table[2^R][2]  (512 B)
 *   END: none
 */
#endif
