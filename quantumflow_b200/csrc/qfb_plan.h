/*
 * qfb_plan.h -- binary layout of an execution plan for the tiled multi-gate executor (qfb_run_plan).
 *
 * A plan is a list of SWEEPS. One sweep = one kernel launch = one read + one write of the whole state
 * (32 B per amplitude). In a sweep every CTA owns tiles of 2^M amplitudes: the M "tile bits" are index bit
 * positions chosen by the planner (gpos[], ascending; the low bits are always included so that every global
 * access is a full 128-byte line). A tile is processed in ROUNDS. In a round each thread holds 2^R amplitudes
 * in registers: R "register bits" (regpos[], tile-bit positions) enumerate the amplitudes of one thread and
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
 * All records are multiples of 16 bytes; integers little endian. The Python planner
 * (quantumflow_b200/planner.py) writes this layout with struct.pack; keep the two in sync.
 */
#ifndef QFB_PLAN_H
#define QFB_PLAN_H
#include <stdint.h>

#define QFB_PLAN_MAGIC 0x50424651u /* "QFBP" */
#define QFB_PLAN_VERSION 6u
#define QFB_PLAN_REG_BITS 4
#define QFB_PLAN_MAX_TILE_BITS 13
#define QFB_PLAN_MIN_TILE_BITS 5
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

typedef struct {
    uint32_t bytes; /* whole sweep record including this header */
    uint32_t nrounds;
    uint32_t nops; /* informational */
    uint32_t reserved;
    uint8_t gpos[16]; /* gpos[j] = index bit position of tile bit j (ascending) */
    uint8_t hole[QFB_PLAN_MAX_HOLES]; /* hole[i] = index bit position of tile-id bit i (ascending) */
} qfb_sweep_header; /* 80 bytes */

typedef struct {
    uint32_t tb;  /* tile-local index contribution (register bits zero) */
    uint32_t pad;
    uint64_t tg;  /* the same bits at their index-bit positions */
} qfb_thread_lut; /* 16 bytes */

#define QFB_PLAN_LUT_LO 16 /* entries indexed by the low 4 thread bits */
#define QFB_PLAN_LUT_HI 32 /* entries indexed by the remaining (M-R-4 <= 5) thread bits */

typedef struct {
    uint32_t nops;
    uint32_t bytes;     /* whole round record: header + thread LUTs + ops */
    uint8_t regpos[4];  /* tile-bit position of register bit i */
    uint8_t thrpos[12]; /* tile-bit position of thread bit t, t < M-R */
    uint8_t has_scalar; /* 1 when the round holds CPH terms without register bits */
    uint8_t has_g2;     /* 1 when the round holds G2 ops (selects the kernel variant for the whole plan) */
    uint8_t pad[6];
    /* thread id -> (tb, tg) in two table look-ups instead of a per-bit deposit loop:
     * tb = lut_lo[tid & 15].tb | lut_hi[tid >> 4].tb, same for tg */
    qfb_thread_lut lut_lo[QFB_PLAN_LUT_LO];
    qfb_thread_lut lut_hi[QFB_PLAN_LUT_HI];
} qfb_round_header; /* 32 + 768 bytes */

/* kinds of QFB_OP_G1: structure of the 2x2 operator, chosen by the planner to save FP64 work. The "pivoted"
 * kinds apply the operator divided by its (0,0) entry; the planner multiplies the pivots of a sweep into one
 * uniform scalar that rides on the sweep's unconditional CPH term. (x, y) = the pair of amplitudes. */
enum {
    QFB_G1_GENERAL = 0,  /* 16 FP64 per pair */
    QFB_G1_REAL = 1,     /* all entries real: 8 per pair */
    QFB_G1_RXLIKE = 2,   /* real diagonal, imaginary off-diagonal: 8 per pair */
    QFB_G1_SWAPX = 3,    /* Pauli X: swap, no arithmetic */
    QFB_G1_ANTIDIAG = 4, /* zero diagonal (Y, phased X): 8 per pair */
    QFB_G1_SUMDIFF = 5,  /* pivoted Hadamard-like: x' = x + r0 y, y' = x + r1 y with r = +-1 (m[0], m[1]); sums only,
                            so destructive interference gives exact zeros like the reference's h*x + h*y: 4 per pair */
    QFB_G1_ROT_R = 6,    /* pivoted real rotation (RY): x' = x + r y, y' = y + s x  (m[0], m[1]): 4 per pair */
    QFB_G1_ROT_I = 7     /* pivoted RX-like: x' = x + i a y, y' = y + i b x  (m[0], m[1]): 4 per pair */
};
/* Handler ids: ONE DENSE switch in the kernel's op interpreter (dense ids let ptxas emit a jump table, BRX,
 * instead of a compare tree whose serial branches cost ~15 cycles per level).
 *   QFB_H_G1_GENERAL + j  uncontrolled dense 1-bit operator on register bit j (16 FP64 per pair)
 *   QFB_H_G1_SWAPX + j    X
 *   QFB_H_G1_SUMDIFF + j  pivoted Hadamard-like
 *   QFB_H_G1_ROT_R + j    pivoted real rotation (RY)
 *   QFB_H_G1_ROT_I + j    pivoted RX-like
 *   QFB_H_G1C_GENERAL + j controlled dense 1-bit operator   (reg_cmask / idx_cmask)
 *   QFB_H_G1C_SWAPX + j   controlled X (CNOT, CCNOT ...)
 *   QFB_H_CPH_SCALAR      phase term without register bits: accumulates into the round's scalar
 *   QFB_H_CPH_REG / _NEG  phase term on the register elements selected by reg_cmask (NEG: factor -1)
 *   QFB_H_G2 + pair       dense 2-bit operator, (j0, j1) = (1,0) (2,0) (2,1) (3,0) (3,1) (3,2)
 *   QFB_H_END             terminates the round's op list */
enum {
    QFB_H_G1_GENERAL = 0,
    QFB_H_G1_SWAPX = 4,
    QFB_H_G1_SUMDIFF = 8,
    QFB_H_G1_ROT_R = 12,
    QFB_H_G1_ROT_I = 16,
    QFB_H_G1C_GENERAL = 20,
    QFB_H_G1C_SWAPX = 24,
    QFB_H_CPH_SCALAR = 28,
    QFB_H_CPH_REG = 29,
    QFB_H_CPH_NEG = 30,
    QFB_H_G2 = 31,
    QFB_H_END = 37,
    QFB_H_COUNT = 38
};

typedef struct {
    uint8_t handler;
    uint8_t reg_cmask; /* control / phase mask over the register index */
    uint8_t size16;    /* whole op record including this header, in 16-byte units */
    uint8_t pad0;
    uint32_t pad1;
    uint64_t idx_cmask; /* control / phase mask over thread-level bits of the FULL index (incl. rank bits) */
} qfb_op_header; /* 16 bytes */

/* payloads (follow the header)
 *   G1 : double m[8]   row-major 2x2 complex (64 B); pivoted kinds use m[0], m[1] as described above
 *   G2 : double m[32]; uint32 nzmask; uint32 pad[3]   row-major 4x4 complex, bit (4r+c) of nzmask set when
 *                      entry (r,c) is non-zero (272 B)
 *   CPH: double factor[2]  (16 B)
 *   END: none
 */
#endif
