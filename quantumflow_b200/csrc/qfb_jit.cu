// qfb_jit.cu -- sweep-specialised kernels: every sweep of a plan (qfb_plan.h) is emitted as straight-line PTX,
// compiled in-process for sm_100a (nvPTXCompiler, statically linked) and loaded through the driver API.
//
// Why: the interpreter of qfb_sweep.cu pays a dispatch per operator, reads payloads from
// shared memory and executes every register move; the plan of a sweep is known before the launch, so
//   * operators become straight-line code that ptxas schedules ACROSS operators (no dispatch, no op headers),
//   * coefficients are operands from the constant bank (the module's `qfb_coef`, written once per plan) --
//     the code depends on the STRUCTURE of the sweep only, so one compiled image serves every plan with the
//     same structure (a parametrised circuit in an optimisation loop compiles once),
//   * a controlled X between register bits is a renaming of registers at code-generation time: no instruction
//     (the FP64 pipe is the scarce resource: 1.83 warp instructions per clock and SM, profiles/r2_fp64_peaks.jsonl),
//   * bit positions are immediates: tile-id deposit, thread-bit deposit and exchange offsets need no tables.
// The data movement is the interpreter's: coalesced LDG.128 of round 0 (next tile prefetched into L2), XOR-swizzled
// exchange through shared memory between rounds, coalesced STG.128 (with the pending X flips / bit permutation).
//
// Entry points: jit_generate (PTX text + coefficient values of one sweep), jit_compile (PTX -> cubin, no GPU
// needed: the CPU test-suite compiles every benchmark sweep), JitSweep (loaded module) used by qfb_plan_upload.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdlib.h>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <cuda.h>
#include <nvPTXCompiler.h>
#include "qfb_common.cuh"
#include "qfb_jit.h"
#include "qfb_plan.h"

namespace qfb {

namespace {

// Register bits of the plan being generated: 5, or 4 (half the code per operator; the straight-line code of a sweep
// has to fit the SM's 32 KiB instruction cache, profiles/r2_icache_probe.jsonl). Set per generating thread.
constexpr int RMAX = QFB_PLAN_REG_BITS, NEMAX = 1 << RMAX;
constexpr int HS = 5;                  // handler ids are laid out for 5 register bits
thread_local int R = RMAX, NE = NEMAX;
const int J0[10] = {1, 2, 2, 3, 3, 3, 4, 4, 4, 4}, J1[10] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3};

uint32_t swz(uint32_t idx) {
    const uint32_t x = idx >> 3;
    return idx ^ ((x ^ (x >> 3) ^ (x >> 6) ^ (x >> 9)) & 7u);
}

struct Amp {
    int re, im;   // %fd register numbers
};

// PTX text builder with virtual-register counters. Registers are single-assignment except inside predicated
// regions (see Gen::op_*), which update the current registers in place so that both paths meet in one name.
struct Gen {
    std::string body;
    std::vector<double> coef;
    int nfd = 0, nrd = 0, nr = 0, np = 0, nlabel = 0;
    int M = 0, nbits = 0, T = 0;

    void e(const char *fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        body += buf;
        body += '\n';
    }
    int fd() { return nfd++; }
    int rd() { return nrd++; }
    int r32() { return nr++; }
    int pr() { return np++; }
    int label() { return nlabel++; }
    // a coefficient of the plan: an operand from the module's constant bank (value-independent code)
    // QFB_JIT_COEF_PIN (default 1): the constant-bank address carries a term that is zero at run time but depends on
    // the tile (coef_base, set per tile), so ptxas cannot hoist the coefficient loads out of the tile loop -- it did,
    // ran out of registers, spilled them and re-read them with LDL + R2UR (51 LDL, 54 STL, 83 R2UR per tile and thread)
    int coef_base = -1;
    int cst(double v) {
        const int reg = fd();
        if (coef_base >= 0)
            e("ld.const.f64 %%fd%d, [%%rd%d+%zu];", reg, coef_base, coef.size() * 8);
        else
            e("ld.const.f64 %%fd%d, [qfb_coef+%zu];", reg, coef.size() * 8);
        coef.push_back(v);
        return reg;
    }
    // the negative of a coefficient that is already in a register: neg.f64 folds into the consumer's operand modifier,
    // so it costs neither a constant load (LDC: an issue slot and two registers each) nor an instruction.
    // QFB_JIT_NEGFOLD=0 loads the negated value as a coefficient of its own (experiments).
    int neg_of(int reg, double v) {
        static const bool fold = [] {
            const char *e = getenv("QFB_JIT_NEGFOLD");
            return !(e && *e) || atoi(e) != 0;
        }();
        if (!fold) return cst(-v);
        const int out = fd();
        this->e("neg.f64 %%fd%d, %%fd%d;", out, reg);
        return out;
    }
};

struct RoundInfo {
    const qfb_round_header *rh;
    const uint8_t *ops;       // first op record
};

// deposit the low bits of the 32-bit register `src` (bit t -> bit pos[t]) into a fresh 64-bit register
int deposit64(Gen &g, int src32, const int *pos, int n) {
    const int acc = g.rd();
    g.e("mov.u64 %%rd%d, 0;", acc);
    const int wide = g.rd();
    g.e("cvt.u64.u32 %%rd%d, %%r%d;", wide, src32);
    int t = 0;
    while (t < n) {
        int len = 1;
        while (t + len < n && pos[t + len] == pos[t] + len) ++len;       // a run of consecutive positions
        const int tmp = g.rd();
        g.e("shr.u64 %%rd%d, %%rd%d, %d;", tmp, wide, t);
        g.e("and.b64 %%rd%d, %%rd%d, %llu;", tmp, tmp, (unsigned long long)((1ull << len) - 1));
        g.e("shl.b64 %%rd%d, %%rd%d, %d;", tmp, tmp, pos[t]);
        g.e("or.b64 %%rd%d, %%rd%d, %%rd%d;", acc, acc, tmp);
        t += len;
    }
    return acc;
}

// the same from a 64-bit source (tile id through the holes)
int deposit64_from64(Gen &g, int src64, const int *pos, int n) {
    const int acc = g.rd();
    g.e("mov.u64 %%rd%d, 0;", acc);
    int t = 0;
    while (t < n) {
        int len = 1;
        while (t + len < n && pos[t + len] == pos[t] + len) ++len;
        const int tmp = g.rd();
        g.e("shr.u64 %%rd%d, %%rd%d, %d;", tmp, src64, t);
        g.e("and.b64 %%rd%d, %%rd%d, %llu;", tmp, tmp, (unsigned long long)((1ull << len) - 1));
        g.e("shl.b64 %%rd%d, %%rd%d, %d;", tmp, tmp, pos[t]);
        g.e("or.b64 %%rd%d, %%rd%d, %%rd%d;", acc, acc, tmp);
        t += len;
    }
    return acc;
}

// swizzled byte offset of the thread's first amplitude in the exchange buffer: swz(tb) << 4 (32-bit register)
int thread_stb(Gen &g, int tid32, const uint8_t *thrpos, int nthr) {
    const int tb = g.r32();
    g.e("mov.u32 %%r%d, 0;", tb);
    for (int t = 0; t < nthr; ++t) {
        const int tmp = g.r32();
        g.e("shr.u32 %%r%d, %%r%d, %d;", tmp, tid32, t);
        g.e("and.b32 %%r%d, %%r%d, 1;", tmp, tmp);
        g.e("shl.b32 %%r%d, %%r%d, %d;", tmp, tmp, (int)thrpos[t]);
        g.e("or.b32 %%r%d, %%r%d, %%r%d;", tb, tb, tmp);
    }
    // f = (x ^ x>>3 ^ x>>6 ^ x>>9) & 7 with x = tb >> 3
    const int x = g.r32(), f = g.r32(), t1 = g.r32();
    g.e("shr.u32 %%r%d, %%r%d, 3;", x, tb);
    g.e("shr.u32 %%r%d, %%r%d, 3;", t1, x);
    g.e("xor.b32 %%r%d, %%r%d, %%r%d;", f, x, t1);
    g.e("shr.u32 %%r%d, %%r%d, 6;", t1, x);
    g.e("xor.b32 %%r%d, %%r%d, %%r%d;", f, f, t1);
    g.e("shr.u32 %%r%d, %%r%d, 9;", t1, x);
    g.e("xor.b32 %%r%d, %%r%d, %%r%d;", f, f, t1);
    g.e("and.b32 %%r%d, %%r%d, 7;", f, f);
    const int stb = g.r32();
    g.e("xor.b32 %%r%d, %%r%d, %%r%d;", stb, tb, f);
    g.e("shl.b32 %%r%d, %%r%d, 4;", stb, stb);
    return stb;
}

// exchange-buffer address of register index e for a round: (stb ^ low(e) << 4) + high(e) -- the thread bits and
// the register bits occupy different tile positions, so only the three swizzled low index bits need an XOR
struct XchgAddr {
    int base[8];      // 32-bit registers: smem + (stb ^ (k << 4)), -1 when unused
    uint32_t low[NEMAX], high[NEMAX];
};

XchgAddr exchange_addresses(Gen &g, int smem32, int stb32, const uint8_t *regpos) {
    XchgAddr x;
    for (int k = 0; k < 8; ++k) x.base[k] = -1;
    for (int e = 0; e < NE; ++e) {
        uint32_t rb = 0;
        for (int i = 0; i < R; ++i)
            if ((e >> i) & 1) rb |= 1u << regpos[i];
        const uint32_t s = swz(rb);
        x.low[e] = s & 7u;
        x.high[e] = (s & ~7u) << 4;
    }
    const int sum = g.r32();
    g.e("add.u32 %%r%d, %%r%d, %%r%d;", sum, smem32, stb32);      // smem is 128-byte aligned: + == | here
    for (int e = 0; e < NE; ++e) {
        const int k = (int)x.low[e];
        if (x.base[k] < 0) {
            if (k == 0) {
                x.base[k] = sum;
            } else {
                x.base[k] = g.r32();
                g.e("xor.b32 %%r%d, %%r%d, %d;", x.base[k], sum, k << 4);
            }
        }
    }
    return x;
}

// The same for the write side of an exchange when conditional flips are pending (RoundState::pend): in the threads
// where predicate pend[j] holds, the register with index e carries the amplitude of index e ^ (1 << j) and goes to THAT
// slot. swz is linear over GF(2), so the slot of e ^ F is slot(e) ^ slot(F): the thread's part of the address takes the
// XOR of the selected slot(1 << j); the pending bits themselves then cannot go into the immediate (their address bits
// are thread dependent), so there is one set of base registers per value v of the pending bits.
struct XchgStore {
    int reg[NEMAX];
    uint32_t imm[NEMAX];
};

XchgStore exchange_store_addresses(Gen &g, int smem32, int stb32, const uint8_t *regpos, const int *pend) {
    XchgStore x;
    auto slot = [&](int e) {
        uint32_t rb = 0;
        for (int i = 0; i < R; ++i)
            if ((e >> i) & 1) rb |= 1u << regpos[i];
        return swz(rb);
    };
    int jm = 0, t = stb32;
    for (int j = 0; j < R; ++j) {
        if (pend[j] < 0) continue;
        jm |= 1 << j;
        const int sel = g.r32(), nt = g.r32();
        g.e("selp.b32 %%r%d, %u, 0, %%p%d;", sel, slot(1 << j) << 4, pend[j]);
        g.e("xor.b32 %%r%d, %%r%d, %%r%d;", nt, t, sel);
        t = nt;
    }
    std::map<uint32_t, int> bases;       // (v, k) -> register
    for (int e = 0; e < NE; ++e) {
        const int v = e & jm, rest = e & ~jm;
        const uint32_t sr = slot(rest), k = sr & 7u;
        const uint32_t key = ((uint32_t)v << 3) | k;
        auto it = bases.find(key);
        if (it == bases.end()) {
            const uint32_t c = (slot(v) << 4) ^ (k << 4);
            const int tv = g.r32(), b = g.r32();
            if (c) {
                g.e("xor.b32 %%r%d, %%r%d, %u;", tv, t, c);
                g.e("add.u32 %%r%d, %%r%d, %%r%d;", b, smem32, tv);
            } else {
                g.e("add.u32 %%r%d, %%r%d, %%r%d;", b, smem32, t);
            }
            it = bases.emplace(key, b).first;
        }
        x.reg[e] = it->second;
        x.imm[e] = (sr & ~7u) << 4;
    }
    return x;
}

// Global address of `base + off` as an operand text: offsets are split into a part that needs a register of its own
// (bits of 2^22 bytes and above: one 64-bit add per distinct high part) and an immediate, so that 16 accesses of a
// tile need a handful of address registers instead of 16 (less register pressure, fewer integer instructions).
struct AddrSet {
    std::map<long long, int> bases;      // high part -> 64-bit register
    int root;
    explicit AddrSet(int root_reg) : root(root_reg) {}
    std::string operand(Gen &g, long long off) {
        const long long lo = ((off % (1ll << 22)) + (1ll << 22)) % (1ll << 22), hi = off - lo;
        int reg = root;
        if (hi != 0) {
            auto it = bases.find(hi);
            if (it == bases.end()) {
                reg = g.rd();
                g.e("add.s64 %%rd%d, %%rd%d, %lld;", reg, root, hi);
                bases[hi] = reg;
            } else {
                reg = it->second;
            }
        }
        char buf[64];
        if (lo)
            snprintf(buf, sizeof(buf), "[%%rd%d+%lld]", reg, lo);
        else
            snprintf(buf, sizeof(buf), "[%%rd%d]", reg);
        return buf;
    }
};

struct OpView {
    qfb_op_header h;
    const uint8_t *payload;
};

double payload_f64(const OpView &op, int i) {
    double v;
    memcpy(&v, op.payload + 8 * i, 8);
    return v;
}

// ---- operators -------------------------------------------------------------------------------------------

void pairs_of(int j, int p, int &e0, int &e1) {
    e0 = ((p >> j) << (j + 1)) | (p & ((1 << j) - 1));
    e1 = e0 | (1 << j);
}

// predicate "all idx_cmask bits are 1 in the thread's full index"; -1 when the mask is empty
int thread_predicate(Gen &g, uint64_t cm, int tfull64) {
    if (cm == 0) return -1;
    const int t = g.rd(), p = g.pr();
    g.e("and.b64 %%rd%d, %%rd%d, %llu;", t, tfull64, (unsigned long long)cm);
    g.e("setp.eq.u64 %%p%d, %%rd%d, %llu;", p, t, (unsigned long long)cm);
    return p;
}

// a *= (c + i s) into fresh registers; nc = -s
void cmul_fresh(Gen &g, Amp &a, int c, int s, int ns) {
    const int t0 = g.fd(), t1 = g.fd(), re = g.fd(), im = g.fd();
    g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", t0, ns, a.im);
    g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", t1, s, a.re);
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", re, a.re, c, t0);
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", im, a.im, c, t1);
    a.re = re;
    a.im = im;
}

// the same in place (inside a predicated region)
void cmul_inplace(Gen &g, const Amp &a, int c, int s, int ns) {
    const int t0 = g.fd(), t1 = g.fd();
    g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", t0, ns, a.im);
    g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", t1, s, a.re);
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", a.re, a.re, c, t0);
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", a.im, a.im, c, t1);
}

// dense 2x2 on (x, y): the interpreter's arithmetic (16 FP64), results in place when `inplace`
void general_pair(Gen &g, Amp &x, Amp &y, const int *c, const int *n, bool inplace) {
    const int t0 = g.fd(), t1 = g.fd(), t2 = g.fd(), t3 = g.fd(), t4 = g.fd(), t5 = g.fd();
    g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", t0, n[3], y.im);
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", t0, c[2], y.re, t0);
    g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", t1, c[3], y.re);
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", t1, c[2], y.im, t1);
    g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", t2, n[5], x.im);
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", t2, c[4], x.re, t2);
    g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", t3, c[5], x.re);
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", t3, c[4], x.im, t3);
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", t4, c[1], x.re, t1);
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", t5, c[7], y.re, t3);
    Amp nx = x, ny = y;
    if (!inplace) {
        nx.re = g.fd();
        nx.im = g.fd();
        ny.re = g.fd();
        ny.im = g.fd();
    }
    // order matters for the in-place form: x.re is rewritten before x.im reads the OLD x.re? No: x.im' needs old
    // x.im and t4 (which already holds m00i * old x.re); x.re' needs old x.re and old x.im -> write x.re' to a
    // temporary first
    const int xr = g.fd(), yr = g.fd();
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", xr, x.re, c[0], t0);
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", xr, n[1], x.im, xr);
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", nx.im, x.im, c[0], t4);
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", yr, y.re, c[6], t2);
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", yr, n[7], y.im, yr);
    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", ny.im, y.im, c[6], t5);
    g.e("mov.f64 %%fd%d, %%fd%d;", nx.re, xr);
    g.e("mov.f64 %%fd%d, %%fd%d;", ny.re, yr);
    x = nx;
    y = ny;
}

struct RoundState {
    Amp a[NEMAX];
    int tfull;       // 64-bit register: rank bits | tile bits | thread bits of the first amplitude
    int phr, phi;    // running per-thread scalar (has_scalar rounds)
    bool scalar_live;
    // Pending conditional flips: pend[j] >= 0 is a predicate register q; in the threads where q holds the amplitude
    // of register index e sits in a[e ^ (1 << j)] (a thread-controlled X on register bit j that has not been carried
    // out). A flip stays pending while no later operator of the round looks at bit j; what is still pending at the end
    // of the round goes into the ADDRESSES of the exchange (or of the final store) -- a dozen integer instructions
    // instead of 8 selects per amplitude pair.
    int pend[RMAX];
};

// QFB_JIT_DEFER (default 1): 0 = every thread-controlled X is carried out at once with selects
bool defer_flips() {
    static const bool on = [] {
        const char *e = getenv("QFB_JIT_DEFER");
        return !(e && *e) || atoi(e) != 0;
    }();
    return on;
}

// carry out the pending flip of register bit j
void flush_flip(Gen &g, RoundState &st, int j) {
    const int q = st.pend[j];
    if (q < 0) return;
    st.pend[j] = -1;
    for (int k = 0; k < NE / 2; ++k) {
        int e0, e1;
        pairs_of(j, k, e0, e1);
        Amp &x = st.a[e0], &y = st.a[e1];
        const int xr = g.fd(), xi = g.fd(), yr = g.fd(), yi = g.fd();
        g.e("selp.f64 %%fd%d, %%fd%d, %%fd%d, %%p%d;", xr, y.re, x.re, q);
        g.e("selp.f64 %%fd%d, %%fd%d, %%fd%d, %%p%d;", xi, y.im, x.im, q);
        g.e("selp.f64 %%fd%d, %%fd%d, %%fd%d, %%p%d;", yr, x.re, y.re, q);
        g.e("selp.f64 %%fd%d, %%fd%d, %%fd%d, %%p%d;", yi, x.im, y.im, q);
        x.re = xr; x.im = xi; y.re = yr; y.im = yi;
    }
}

void flush_mask(Gen &g, RoundState &st, int mask) {
    for (int j = 0; j < R; ++j)
        if ((mask >> j) & 1) flush_flip(g, st, j);
}

int pending_mask(const RoundState &st) {
    int m = 0;
    for (int j = 0; j < R; ++j)
        if (st.pend[j] >= 0) m |= 1 << j;
    return m;
}

void load_matrix(Gen &g, const OpView &op, int *c, int *n) {
    for (int i = 0; i < 8; ++i) c[i] = g.cst(payload_f64(op, i));
    for (int i = 1; i < 8; i += 2) n[i] = g.neg_of(c[i], payload_f64(op, i));
}

int emit_op(Gen &g, RoundState &st, const OpView &op, std::string &err) {
    const int hd = (int)op.h.handler;
    const int rcm = op.h.reg_cmask;
    const uint64_t icm = op.h.idx_cmask;
    if (hd >= QFB_H_G1_GENERAL && hd < QFB_H_G1C_GENERAL) {
        const int kind = hd / HS, j = hd % HS;
        flush_flip(g, st, j);
        if (kind == 0) {
            int c[8], n[8];
            load_matrix(g, op, c, n);
            for (int p = 0; p < NE / 2; ++p) {
                int e0, e1;
                pairs_of(j, p, e0, e1);
                general_pair(g, st.a[e0], st.a[e1], c, n, false);
            }
        } else if (kind == 1) {          // SUMDIFF: x' = x + r0 y ; y' = x' + (r1 - r0) y
            const double r0 = payload_f64(op, 0), r1 = payload_f64(op, 1);
            const int c0 = g.cst(r0), c2 = g.cst(r1 - r0);
            for (int p = 0; p < NE / 2; ++p) {
                int e0, e1;
                pairs_of(j, p, e0, e1);
                Amp &x = st.a[e0], &y = st.a[e1];
                const int xr = g.fd(), xi = g.fd(), yr = g.fd(), yi = g.fd();
                g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", xr, c0, y.re, x.re);
                g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", xi, c0, y.im, x.im);
                g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", yr, y.re, c2, xr);
                g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", yi, y.im, c2, xi);
                x.re = xr; x.im = xi; y.re = yr; y.im = yi;
            }
        } else if (kind == 2) {          // LU_R: x += a y ; y += b x
            const int ca = g.cst(payload_f64(op, 0)), cb = g.cst(payload_f64(op, 1));
            for (int p = 0; p < NE / 2; ++p) {
                int e0, e1;
                pairs_of(j, p, e0, e1);
                Amp &x = st.a[e0], &y = st.a[e1];
                const int xr = g.fd(), xi = g.fd(), yr = g.fd(), yi = g.fd();
                g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", xr, ca, y.re, x.re);
                g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", xi, ca, y.im, x.im);
                g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", yr, cb, xr, y.re);
                g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", yi, cb, xi, y.im);
                x.re = xr; x.im = xi; y.re = yr; y.im = yi;
            }
        } else {                         // LU_I: x += i a y ; y += i b x
            const double a = payload_f64(op, 0), b = payload_f64(op, 1);
            const int ca = g.cst(a), cb = g.cst(b), na = g.neg_of(ca, a), nb = g.neg_of(cb, b);
            for (int p = 0; p < NE / 2; ++p) {
                int e0, e1;
                pairs_of(j, p, e0, e1);
                Amp &x = st.a[e0], &y = st.a[e1];
                const int xr = g.fd(), xi = g.fd(), yr = g.fd(), yi = g.fd();
                g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", xr, na, y.im, x.re);
                g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", xi, ca, y.re, x.im);
                g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", yr, nb, xi, y.re);
                g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", yi, cb, xr, y.im);
                x.re = xr; x.im = xi; y.re = yr; y.im = yi;
            }
        }
        return QFB_OK;
    }
    if (hd >= QFB_H_G1C_GENERAL && hd < QFB_H_G1C_SWAPX) {
        const int j = hd - QFB_H_G1C_GENERAL;
        flush_mask(g, st, rcm | (1 << j));
        const int p = thread_predicate(g, icm, st.tfull);
        int c[8], n[8];
        load_matrix(g, op, c, n);
        const int skip = g.label();
        if (p >= 0) g.e("@!%%p%d bra L%d;", p, skip);
        for (int q = 0; q < NE / 2; ++q) {
            int e0, e1;
            pairs_of(j, q, e0, e1);
            if ((e0 & rcm) != rcm) continue;
            general_pair(g, st.a[e0], st.a[e1], c, n, p >= 0);
        }
        if (p >= 0) g.e("L%d:", skip);
        return QFB_OK;
    }
    if (hd >= QFB_H_G1C_SWAPX && hd < QFB_H_G1C_SWAPX + HS) {
        const int j = hd - QFB_H_G1C_SWAPX;
        flush_mask(g, st, rcm);            // a pending X on the target commutes with this X, one on a control does not
        const int p = thread_predicate(g, icm, st.tfull);
        if (p >= 0 && rcm == 0 && defer_flips()) {
            if (st.pend[j] < 0) {
                st.pend[j] = p;
            } else {
                const int both = g.pr();
                g.e("xor.pred %%p%d, %%p%d, %%p%d;", both, st.pend[j], p);
                st.pend[j] = both;
            }
            return QFB_OK;
        }
        if (p >= 0) flush_flip(g, st, j);
        for (int q = 0; q < NE / 2; ++q) {
            int e0, e1;
            pairs_of(j, q, e0, e1);
            if ((e0 & rcm) != rcm) continue;
            if (p < 0) {
                std::swap(st.a[e0], st.a[e1]);            // controls in registers only: a renaming, no instruction
            } else {
                Amp &x = st.a[e0], &y = st.a[e1];
                const int xr = g.fd(), xi = g.fd(), yr = g.fd(), yi = g.fd();
                g.e("selp.f64 %%fd%d, %%fd%d, %%fd%d, %%p%d;", xr, y.re, x.re, p);
                g.e("selp.f64 %%fd%d, %%fd%d, %%fd%d, %%p%d;", xi, y.im, x.im, p);
                g.e("selp.f64 %%fd%d, %%fd%d, %%fd%d, %%p%d;", yr, x.re, y.re, p);
                g.e("selp.f64 %%fd%d, %%fd%d, %%fd%d, %%p%d;", yi, x.im, y.im, p);
                x.re = xr; x.im = xi; y.re = yr; y.im = yi;
            }
        }
        return QFB_OK;
    }
    if (hd == QFB_H_CPH_SCALAR) {
        // ph *= (p ? factor : 1)
        const int fr = g.cst(payload_f64(op, 0)), fi = g.cst(payload_f64(op, 1));
        int sr = fr, si = fi;
        const int p = thread_predicate(g, icm, st.tfull);
        if (p >= 0) {
            sr = g.fd();
            si = g.fd();
            g.e("selp.f64 %%fd%d, %%fd%d, 0d3FF0000000000000, %%p%d;", sr, fr, p);
            g.e("selp.f64 %%fd%d, %%fd%d, 0d0000000000000000, %%p%d;", si, fi, p);
        }
        if (!st.scalar_live) {
            st.phr = sr;
            st.phi = si;
            st.scalar_live = true;
        } else {
            const int t0 = g.fd(), t1 = g.fd(), nr = g.fd(), ni = g.fd(), nsi = g.fd();
            g.e("neg.f64 %%fd%d, %%fd%d;", nsi, si);
            g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", t0, nsi, st.phi);
            g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", t1, si, st.phr);
            g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", nr, sr, st.phr, t0);
            g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", ni, sr, st.phi, t1);
            st.phr = nr;
            st.phi = ni;
        }
        return QFB_OK;
    }
    if ((hd >= QFB_H_CPH_REG1 && hd < QFB_H_CPH_NEG1) || hd == QFB_H_CPH_REGM) {
        const bool real_scale = hd >= QFB_H_CPH_RSC1 && hd < QFB_H_CPH_NEG1;
        flush_mask(g, st, rcm);
        const double fr = payload_f64(op, 0), fi = payload_f64(op, 1);
        const int c = g.cst(fr);
        const int s = real_scale ? -1 : g.cst(fi), ns = real_scale ? -1 : g.neg_of(s, fi);
        const int p = thread_predicate(g, icm, st.tfull);
        const int skip = g.label();
        if (p >= 0) g.e("@!%%p%d bra L%d;", p, skip);
        for (int e = 0; e < NE; ++e) {
            if ((e & rcm) != rcm) continue;
            if (real_scale) {
                if (p >= 0) {
                    g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", st.a[e].re, st.a[e].re, c);
                    g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", st.a[e].im, st.a[e].im, c);
                } else {
                    const int re = g.fd(), im = g.fd();
                    g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", re, st.a[e].re, c);
                    g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", im, st.a[e].im, c);
                    st.a[e].re = re;
                    st.a[e].im = im;
                }
            } else if (p >= 0) {
                cmul_inplace(g, st.a[e], c, s, ns);
            } else {
                cmul_fresh(g, st.a[e], c, s, ns);
            }
        }
        if (p >= 0) g.e("L%d:", skip);
        return QFB_OK;
    }
    if ((hd >= QFB_H_CPH_NEG1 && hd < QFB_H_CPH_REGM) || hd == QFB_H_CPH_NEGM) {
        const int p = thread_predicate(g, icm, st.tfull);
        const int pm = pending_mask(st) & rcm;
        if (pm != 0 && (rcm & (rcm - 1)) == 0) {
            // sign flip on ONE register bit j whose X is pending (predicate q): the amplitudes with bit j = 1 sit in the
            // registers with bit j = 1 where q is false and in those with bit j = 0 where q is true
            const int q = st.pend[__builtin_ctz(rcm)];
            const int p1 = g.pr(), p0 = g.pr();
            if (p >= 0) {
                g.e("and.pred %%p%d, %%p%d, !%%p%d;", p1, p, q);
                g.e("and.pred %%p%d, %%p%d, %%p%d;", p0, p, q);
            } else {
                g.e("not.pred %%p%d, %%p%d;", p1, q);
                g.e("mov.pred %%p%d, %%p%d;", p0, q);
            }
            for (int e = 0; e < NE; ++e) {
                const int pe = (e & rcm) ? p1 : p0;
                g.e("@%%p%d xor.b64 %%fd%d, %%fd%d, 0x8000000000000000;", pe, st.a[e].re, st.a[e].re);
                g.e("@%%p%d xor.b64 %%fd%d, %%fd%d, 0x8000000000000000;", pe, st.a[e].im, st.a[e].im);
            }
            return QFB_OK;
        }
        flush_mask(g, st, rcm);
        for (int e = 0; e < NE; ++e) {
            if ((e & rcm) != rcm) continue;
            if (p >= 0) {
                // sign bit of the high word, predicated: one integer-pipe instruction per component
                g.e("@%%p%d xor.b64 %%fd%d, %%fd%d, 0x8000000000000000;", p, st.a[e].re, st.a[e].re);
                g.e("@%%p%d xor.b64 %%fd%d, %%fd%d, 0x8000000000000000;", p, st.a[e].im, st.a[e].im);
            } else {
                const int re = g.fd(), im = g.fd();
                g.e("neg.f64 %%fd%d, %%fd%d;", re, st.a[e].re);        // folds into the consumer's operand
                g.e("neg.f64 %%fd%d, %%fd%d;", im, st.a[e].im);
                st.a[e].re = re;
                st.a[e].im = im;
            }
        }
        return QFB_OK;
    }
    if (hd == QFB_H_CPH_TABLE) {
        const bool whole = op.h.flag != 0;
        flush_mask(g, st, whole ? NE - 1 : rcm);
        for (int e = 0; e < NE; ++e) {
            if (!whole && (e & rcm) == 0) continue;
            const double fr = payload_f64(op, 2 * e), fi = payload_f64(op, 2 * e + 1);
            const int c = g.cst(fr), s = g.cst(fi), ns = g.neg_of(s, fi);
            cmul_fresh(g, st.a[e], c, s, ns);
        }
        return QFB_OK;
    }
    if (hd >= QFB_H_G2 && hd < QFB_H_G2 + 10) {
        const int j0 = J0[hd - QFB_H_G2], j1 = J1[hd - QFB_H_G2];
        flush_mask(g, st, rcm | (1 << j0) | (1 << j1));
        uint32_t nz;
        memcpy(&nz, op.payload + 256, 4);
        const int p = thread_predicate(g, icm, st.tfull);
        int c[32], n[32];
        for (int i = 0; i < 16; ++i) {
            c[2 * i] = c[2 * i + 1] = n[2 * i + 1] = -1;
            if (!((nz >> i) & 1)) continue;
            c[2 * i] = g.cst(payload_f64(op, 2 * i));
            c[2 * i + 1] = g.cst(payload_f64(op, 2 * i + 1));
            n[2 * i + 1] = g.neg_of(c[2 * i + 1], payload_f64(op, 2 * i + 1));
        }
        const int skip = g.label();
        if (p >= 0) g.e("@!%%p%d bra L%d;", p, skip);
        int others[RMAX], no = 0;
        for (int b = 0; b < R; ++b)
            if (b != j0 && b != j1) others[no++] = b;
        for (int q = 0; q < (1 << no); ++q) {
            int eb = 0;
            for (int i = 0; i < no; ++i) eb |= ((q >> i) & 1) << others[i];
            if ((eb & rcm) != rcm) continue;
            const int ids[4] = {eb, eb | (1 << j1), eb | (1 << j0), eb | (1 << j0) | (1 << j1)};
            int ore[4], oim[4];
            for (int r = 0; r < 4; ++r) {
                ore[r] = g.fd();
                oim[r] = g.fd();
                bool first = true;
                for (int cc = 0; cc < 4; ++cc) {
                    const int i = 4 * r + cc;
                    if (!((nz >> i) & 1)) continue;
                    const Amp &in = st.a[ids[cc]];
                    if (first) {
                        g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", ore[r], c[2 * i], in.re);
                        g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", oim[r], c[2 * i], in.im);
                        first = false;
                    } else {
                        g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", ore[r], c[2 * i], in.re, ore[r]);
                        g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", oim[r], c[2 * i], in.im, oim[r]);
                    }
                    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", ore[r], n[2 * i + 1], in.im, ore[r]);
                    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", oim[r], c[2 * i + 1], in.re, oim[r]);
                }
                if (first) {
                    g.e("mov.f64 %%fd%d, 0d0000000000000000;", ore[r]);
                    g.e("mov.f64 %%fd%d, 0d0000000000000000;", oim[r]);
                }
            }
            for (int r = 0; r < 4; ++r) {
                if (p >= 0) {
                    g.e("mov.f64 %%fd%d, %%fd%d;", st.a[ids[r]].re, ore[r]);
                    g.e("mov.f64 %%fd%d, %%fd%d;", st.a[ids[r]].im, oim[r]);
                } else {
                    st.a[ids[r]].re = ore[r];
                    st.a[ids[r]].im = oim[r];
                }
            }
        }
        if (p >= 0) g.e("L%d:", skip);
        return QFB_OK;
    }
    if (hd >= QFB_H_G2X && hd < QFB_H_G2X + 10) {
        const int j0 = J0[hd - QFB_H_G2X], j1 = J1[hd - QFB_H_G2X];
        flush_mask(g, st, rcm | (1 << j0) | (1 << j1));
        const int p = thread_predicate(g, icm, st.tfull);
        int c[8];
        for (int i = 0; i < 8; ++i) c[i] = g.cst(payload_f64(op, i));
        const int skip = g.label();
        if (p >= 0) g.e("@!%%p%d bra L%d;", p, skip);
        int others[RMAX], no = 0;
        for (int b = 0; b < R; ++b)
            if (b != j0 && b != j1) others[no++] = b;
        for (int q = 0; q < (1 << no); ++q) {
            int eb = 0;
            for (int i = 0; i < no; ++i) eb |= ((q >> i) & 1) << others[i];
            if ((eb & rcm) != rcm) continue;
            const int ids[4] = {eb, eb | (1 << j1), eb | (1 << j0), eb | (1 << j0) | (1 << j1)};
            const int pq[2][3] = {{ids[0], ids[3], 0}, {ids[1], ids[2], 4}};
            for (int w = 0; w < 2; ++w) {
                Amp &ap = st.a[pq[w][0]], &aq = st.a[pq[w][1]];
                const int base = pq[w][2];
                int *comp_p[2] = {&ap.re, &ap.im}, *comp_q[2] = {&aq.re, &aq.im};
                for (int k = 0; k < 2; ++k) {
                    const int t0 = g.fd(), t1 = g.fd(), np_ = g.fd(), nq_ = g.fd();
                    g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", t0, c[base + 2], *comp_p[k]);
                    g.e("mul.f64 %%fd%d, %%fd%d, %%fd%d;", t1, *comp_p[k], c[base]);
                    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", np_, c[base + 1], *comp_q[k], t1);
                    g.e("fma.rn.f64 %%fd%d, %%fd%d, %%fd%d, %%fd%d;", nq_, *comp_q[k], c[base + 3], t0);
                    if (p >= 0) {
                        g.e("mov.f64 %%fd%d, %%fd%d;", *comp_p[k], np_);
                        g.e("mov.f64 %%fd%d, %%fd%d;", *comp_q[k], nq_);
                    } else {
                        *comp_p[k] = np_;
                        *comp_q[k] = nq_;
                    }
                }
            }
        }
        if (p >= 0) g.e("L%d:", skip);
        return QFB_OK;
    }
    err = "jit: unknown handler " + std::to_string(hd);
    return QFB_ERR_UNSUPPORTED;
}

}  // namespace

// QFB_JIT_ASYNC (default 1): the tile of the NEXT iteration is copied global -> shared with cp.async (LDGSTS, no
// registers involved) into the exchange buffer as soon as the current tile has left it for the last time, i.e. the
// copy overlaps the last round's arithmetic and the stores; every thread copies exactly the amplitudes it will read
// back itself (its round-0 slots), so the hand-over needs no barrier, only cp.async.wait_group. 0 = LDG straight
// into registers at the top of the iteration (plus an L2 prefetch of the next tile), as the interpreter does.
static int jit_async_mode() {
    const char *v = getenv("QFB_JIT_ASYNC");
    return (v && *v) ? atoi(v) : 3;
}
// QFB_JIT_ASYNC=3 (default): as 1, but after the last exchange every thread copies the amplitudes of the next tile that
// belong to the slots it has just read ITSELF (its assignment of the last round), so the copy needs no barrier in front
// of it and is issued right behind the read-back; the hand-over barrier moves to the top of the next tile, behind
// cp.async.wait_group. With mode 1 the copy sat behind a barrier, and ptxas moved the whole register arithmetic of the
// last round in FRONT of that barrier (barriers order memory operations only): the copy was issued just before the
// stores and 19 % of the warp time went into waiting for it (profiles/r2_sweep_jit_v6_summary.txt).
static bool jit_own_slot_copy() { return jit_async_mode() == 3; }
static bool jit_async() { return jit_async_mode() != 0; }
// QFB_JIT_ASYNC=2: the copy lands in a SECOND buffer of the CTA (shared memory doubles), so it can start as soon as the
// current tile has been read out of that buffer -- a whole tile period ahead instead of one round. Worth it when the
// tile is small enough to keep three CTAs per SM (tile 2^11: 64 KiB per CTA).
static bool jit_landing_buffer() { return jit_async_mode() == 2; }

// ---- one sweep -> PTX --------------------------------------------------------------------------------------

// QFB_JIT_GROUPS: tiles a CTA works on side by side (one group of 2^(M-R) threads per tile). With G > 1 the groups
// walk the same straight-line code between the same CTA-wide barriers, so an instruction line fetched for one warp
// serves the warps of the other groups on its SM sub-partition (the code of a sweep is far larger than the
// instruction caches and every warp executes each line once per tile).
static int jit_groups(int M) {
    const char *v = getenv("QFB_JIT_GROUPS");
    int g = (v && *v) ? atoi(v) : 1;
    if (g < 1) g = 1;
    while (g > 1 && (((size_t)16 << M) * g > 200 * 1024 || (g << (M - R)) > 1024)) --g;
    return g;
}

int jit_generate(const uint8_t *rec, int nbits, int M, int reg_bits, JitSource &out, std::string &err,
                 uint64_t fix_mask) {
    R = reg_bits;
    NE = 1 << reg_bits;
    qfb_sweep_header sh;
    memcpy(&sh, rec, sizeof(sh));
    const int nholes = nbits - M, nthr = M - R, T = 1 << nthr;
    if (nthr < 3 || M > QFB_PLAN_MAX_TILE_BITS) {
        err = "jit: tile too small";
        return QFB_ERR_UNSUPPORTED;
    }
    const bool store_perm = (sh.flags & QFB_SWEEP_FLAG_STORE_PERM) != 0;
    const bool store_sync = (sh.flags & QFB_SWEEP_FLAG_STORE_SYNC) != 0;
    std::vector<RoundInfo> rounds;
    const uint8_t *rp = rec + sizeof(qfb_sweep_header);
    for (uint32_t r = 0; r < sh.nrounds + (store_perm ? 1u : 0u); ++r) {
        const qfb_round_header *rh = reinterpret_cast<const qfb_round_header *>(rp);
        rounds.push_back(RoundInfo{rh, rp + sizeof(qfb_round_header)});
        rp += rh->bytes;
    }
    const int nrounds = (int)sh.nrounds;
    const qfb_round_header *store_rh = rounds.back().rh;      // the store record when permuting, else the last round
    const uint8_t *store_ipos = store_perm ? sh.spos : sh.gpos;

    Gen g;
    g.M = M;
    g.nbits = nbits;
    g.T = T;
    const int G = jit_groups(M);
    // ---- prologue (tile independent) ----
    const int tid = g.r32(), cta = g.r32(), ncta = g.r32(), smem = g.r32(), grp = g.r32();
    g.e("mov.u32 %%r%d, %%tid.x;", tid);
    if (G > 1) {
        g.e("shr.u32 %%r%d, %%r%d, %d;", grp, tid, nthr);
        g.e("and.b32 %%r%d, %%r%d, %d;", tid, tid, T - 1);
    } else {
        g.e("mov.u32 %%r%d, 0;", grp);       // one group: the loop control is provably CTA-uniform
    }
    g.e("mov.u32 %%r%d, %%ctaid.x;", cta);
    g.e("mad.lo.u32 %%r%d, %%r%d, %d, %%r%d;", cta, cta, G, grp);          // first tile of this group
    g.e("mov.u32 %%r%d, %%nctaid.x;", ncta);
    g.e("mul.lo.u32 %%r%d, %%r%d, %d;", ncta, ncta, G);
    g.e("mov.u32 %%r%d, qfb_smem;", smem);
    g.e("mad.lo.u32 %%r%d, %%r%d, %d, %%r%d;", smem, grp, 16 << M, smem);
    const bool landing = jit_landing_buffer() && jit_async() && ((size_t)32 << M) * G <= 200 * 1024;
    int land = smem;                       // where the asynchronous copy lands: the exchange buffer, or a buffer of its own
    if (landing) {
        land = g.r32();
        g.e("add.u32 %%r%d, %%r%d, %d;", land, smem, (16 << M) * G);
    }
    const int state = g.rd(), hi = g.rd();
    g.e("ld.param.u64 %%rd%d, [p_state];", state);
    g.e("cvta.to.global.u64 %%rd%d, %%rd%d;", state, state);
    g.e("ld.param.u64 %%rd%d, [p_hi];", hi);
    const int zmask = g.rd();
    g.e("ld.param.u64 %%rd%d, [p_zero];", zmask);
    // index bits that are the same for every tile of this launch (kernel parameter): 0 for a launch over the whole
    // state; a launch over a SLICE of the state takes the index bits of `fix_mask` (non-tile bits) from here and
    // enumerates the others (jit_build_variant: sharded states pipeline a qubit remap slice by slice)
    const int fixv = g.rd();
    g.e("ld.param.u64 %%rd%d, [p_fix];", fixv);
    // thread-bit images per round: index-bit image (64-bit) and exchange offset (32-bit)
    // QFB_JIT_REMAT=1 (experiments): the images are recomputed from the thread id where they are used (a dozen integer
    // instructions each) instead of being kept across the tile loop. Measured: no gain once the coefficient loads
    // stay inside the loop (Gen::cst), which is what had caused the spills.
    const char *remat_env = getenv("QFB_JIT_REMAT");
    const bool remat = remat_env && *remat_env && atoi(remat_env) != 0;
    std::vector<int> tg_pro(nrounds, -1), stb_pro(nrounds, -1);
    auto make_tg = [&](int r) {
        int pos[16];
        for (int t = 0; t < nthr; ++t) pos[t] = sh.gpos[rounds[r].rh->thrpos[t]];
        return deposit64(g, tid, pos, nthr);
    };
    auto make_stb = [&](int r) { return thread_stb(g, tid, rounds[r].rh->thrpos, nthr); };
    auto make_tg_store = [&]() {
        int pos[16];
        for (int t = 0; t < nthr; ++t) pos[t] = store_ipos[store_rh->thrpos[t]];
        return deposit64(g, tid, pos, nthr);
    };
    int tg_store_pro = -1;
    if (!remat) {
        for (int r = 0; r < nrounds; ++r) {
            tg_pro[r] = make_tg(r);
            stb_pro[r] = make_stb(r);
        }
        tg_store_pro = make_tg_store();
    }
    auto tg = [&](int r) { return remat ? make_tg(r) : tg_pro[r]; };
    auto stb = [&](int r) { return remat ? make_stb(r) : stb_pro[r]; };
    auto tg_store_fn = [&]() { return remat ? make_tg_store() : tg_store_pro; };
    // L2 prefetch of the next tile: the 8 lanes that share the thread's 128-byte lines split its 2^R lines
    // (lane k takes the register indices whose top three bits are k); koff = byte offset of lane k's first line
    const qfb_round_header *r0 = rounds[0].rh;
    const bool lane_lines = sh.gpos[r0->thrpos[0]] == 0 && sh.gpos[r0->thrpos[1]] == 1 && sh.gpos[r0->thrpos[2]] == 2;
    int64_t step0[RMAX];
    for (int i = 0; i < R; ++i) step0[i] = (int64_t)16 << sh.gpos[r0->regpos[i]];
    int koff = -1;
    if (lane_lines) {
        const int k = g.r32();
        g.e("and.b32 %%r%d, %%r%d, 7;", k, tid);
        koff = g.rd();
        const int t = g.rd();
        g.e("mul.wide.u32 %%rd%d, %%r%d, 16;", t, k);
        g.e("neg.s64 %%rd%d, %%rd%d;", koff, t);
        for (int i = 0; i < 3; ++i) {
            const int b = g.r32(), p = g.pr(), sel = g.rd();
            g.e("and.b32 %%r%d, %%r%d, %d;", b, k, 1 << i);
            g.e("setp.ne.u32 %%p%d, %%r%d, 0;", p, b);
            g.e("selp.b64 %%rd%d, %lld, 0, %%p%d;", sel, (long long)step0[R - 3 + i], p);
            g.e("add.s64 %%rd%d, %%rd%d, %%rd%d;", koff, koff, sel);
        }
    }
    int hole[QFB_PLAN_MAX_HOLES];
    int nfree = 0;
    uint64_t hole_mask = 0;
    for (int i = 0; i < nholes; ++i) {
        hole_mask |= 1ull << sh.hole[i];
        if (!((fix_mask >> sh.hole[i]) & 1ull)) hole[nfree++] = sh.hole[i];
    }
    if (fix_mask & ~hole_mask) {
        err = "jit: a slice can only fix index bits outside the sweep's tile";
        return QFB_ERR_UNSUPPORTED;
    }
    // ---- tile loop: tile = ctaid, ctaid + nctaid, ... ----
    const int tile = g.rd(), stride = g.rd(), gb = g.rd();
    g.e("cvt.u64.u32 %%rd%d, %%r%d;", tile, cta);
    g.e("cvt.u64.u32 %%rd%d, %%r%d;", stride, ncta);
    const unsigned long long ntiles = 1ull << nfree;
    // With G groups the CTA leaves the loop as a whole (CTA-wide barriers): the loop runs while the CTA's FIRST group
    // has a tile; a group past the end works on the last tile again and keeps its stores to itself (`active`).
    const int grp64 = g.rd(), lead = g.rd(), active = g.pr(), tclamp = g.rd();
    g.e("cvt.u64.u32 %%rd%d, %%r%d;", grp64, grp);
    {
        const int p = g.pr();
        g.e("sub.u64 %%rd%d, %%rd%d, %%rd%d;", lead, tile, grp64);
        g.e("setp.ge.u64 %%p%d, %%rd%d, %llu;", p, lead, ntiles);
        g.e("@%%p%d bra L_EXIT;", p);
    }
    g.e("setp.lt.u64 %%p%d, %%rd%d, %llu;", active, tile, ntiles);
    g.e("min.u64 %%rd%d, %%rd%d, %llu;", tclamp, tile, ntiles - 1);
    {
        const int first = deposit64_from64(g, tclamp, hole, nfree);
        g.e("or.b64 %%rd%d, %%rd%d, %%rd%d;", gb, first, fixv);
    }
    {
        // QFB_JIT_STAGGER="K:D" (experiments): CTA b sleeps (b % K) * D nanoseconds before its first tile, so that the
        // CTAs of the grid do not walk through their load / compute / store phases in step
        const char *v = getenv("QFB_JIT_STAGGER");
        int K = 0, D = 0;
        if (v && sscanf(v, "%d:%d", &K, &D) == 2 && K > 1 && D > 0) {
            const int c = g.r32(), d = g.r32();
            g.e("mov.u32 %%r%d, %%ctaid.x;", c);
            g.e("rem.u32 %%r%d, %%r%d, %d;", d, c, K);
            g.e("mul.lo.u32 %%r%d, %%r%d, %d;", d, d, D);
            g.e("nanosleep.u32 %%r%d;", d);
        }
    }
    if (jit_async()) {
        // the first tile's copy (every later one is started by the iteration before it)
        const int idx = g.rd(), base = g.rd();
        g.e("or.b64 %%rd%d, %%rd%d, %%rd%d;", idx, gb, tg(0));
        g.e("shl.b64 %%rd%d, %%rd%d, 4;", idx, idx);
        g.e("add.u64 %%rd%d, %%rd%d, %%rd%d;", base, state, idx);
        const XchgAddr x = exchange_addresses(g, land, stb(0), r0->regpos);
        AddrSet as(base);
        for (int e = 0; e < NE; ++e) {
            int64_t off = 0;
            for (int i = 0; i < R; ++i)
                if ((e >> i) & 1) off += step0[i];
            const std::string src = as.operand(g, off);
            g.e("cp.async.cg.shared.global [%%r%d+%u], %s, 16;", x.base[x.low[e]], x.high[e], src.c_str());
        }
        g.e("cp.async.commit_group;");
    }
    const int iter32 = g.r32();
    g.e("mov.u32 %%r%d, 0;", iter32);
    g.e("L_TILE:");
    // next tile (for the prefetch and for the next iteration)
    const int tile_next = g.rd(), has_next = g.pr(), tnclamp = g.rd();
    g.e("add.u64 %%rd%d, %%rd%d, %%rd%d;", tile_next, tile, stride);
    g.e("sub.u64 %%rd%d, %%rd%d, %%rd%d;", lead, tile_next, grp64);
    g.e("setp.lt.u64 %%p%d, %%rd%d, %llu;", has_next, lead, ntiles);
    g.e("min.u64 %%rd%d, %%rd%d, %llu;", tnclamp, tile_next, ntiles - 1);
    const int gb_next = g.rd();
    {
        const int dep = deposit64_from64(g, tnclamp, hole, nfree);
        g.e("or.b64 %%rd%d, %%rd%d, %%rd%d;", gb_next, dep, fixv);
    }
    const int higb = g.rd();
    g.e("or.b64 %%rd%d, %%rd%d, %%rd%d;", higb, hi, gb);
    {
        const char *v = getenv("QFB_JIT_COEF_PIN");
        const int mode = (v && *v) ? atoi(v) : 1;
        if (mode == 2) {
            // the same with a CTA-uniform 32-bit iteration counter of its own (not the 64-bit tile index, which lives in
            // vector registers because thread-dependent addresses are derived from it): ptxas can prove the address
            // uniform and reads the coefficients on the uniform datapath (LDCU into uniform registers, no vector
            // registers and no LDC latency)
            const int z32 = g.r32(), z = g.rd(), cb = g.rd(), zm32 = g.r32();
            g.e("cvt.u32.u64 %%r%d, %%rd%d;", zm32, zmask);
            g.e("and.b32 %%r%d, %%r%d, %%r%d;", z32, iter32, zm32);
            g.e("add.u32 %%r%d, %%r%d, 1;", iter32, iter32);
            g.e("cvt.u64.u32 %%rd%d, %%r%d;", z, z32);
            g.e("mov.u64 %%rd%d, qfb_coef;", cb);
            g.e("add.u64 %%rd%d, %%rd%d, %%rd%d;", cb, cb, z);
            g.coef_base = cb;
        } else if (mode != 0) {
            const int z = g.rd(), cb = g.rd();
            g.e("and.b64 %%rd%d, %%rd%d, %%rd%d;", z, tile, zmask);       // zmask = 0 at run time
            g.e("mov.u64 %%rd%d, qfb_coef;", cb);
            g.e("add.u64 %%rd%d, %%rd%d, %%rd%d;", cb, cb, z);
            g.coef_base = cb;
        }
    }

    RoundState st;
    st.scalar_live = false;
    st.phr = st.phi = -1;
    const bool async = jit_async();
    int64_t off0[NEMAX];
    for (int e = 0; e < NE; ++e) {
        off0[e] = 0;
        for (int i = 0; i < R; ++i)
            if ((e >> i) & 1) off0[e] += step0[i];
    }
    // global base address of the thread's first amplitude of a tile (round-0 assignment)
    auto tile_base = [&](int gbreg) {
        const int idx = g.rd(), base = g.rd();
        g.e("or.b64 %%rd%d, %%rd%d, %%rd%d;", idx, gbreg, tg(0));
        g.e("shl.b64 %%rd%d, %%rd%d, 4;", idx, idx);
        g.e("add.u64 %%rd%d, %%rd%d, %%rd%d;", base, state, idx);
        return base;
    };
    // asynchronous copy of a tile into the thread's own round-0 slots of the exchange buffer
    auto async_copy = [&](int basereg) {
        const XchgAddr x = exchange_addresses(g, land, stb(0), r0->regpos);
        AddrSet as(basereg);
        for (int e = 0; e < NE; ++e) {
            const std::string src = as.operand(g, off0[e]);
            g.e("cp.async.cg.shared.global [%%r%d+%u], %s, 16;", x.base[x.low[e]], x.high[e], src.c_str());
        }
        g.e("cp.async.commit_group;");
    };
    // the same into the thread's own slots of round rr's assignment (the amplitudes of the tile at `gbreg` that have
    // this thread's bits of round rr): the slots the thread has just read, no barrier needed in front
    auto async_copy_round = [&](int gbreg, int rr) {
        const qfb_round_header *rh = rounds[rr].rh;
        const int idx = g.rd(), base = g.rd();
        g.e("or.b64 %%rd%d, %%rd%d, %%rd%d;", idx, gbreg, tg(rr));
        g.e("shl.b64 %%rd%d, %%rd%d, 4;", idx, idx);
        g.e("add.u64 %%rd%d, %%rd%d, %%rd%d;", base, state, idx);
        const XchgAddr x = exchange_addresses(g, land, stb(rr), rh->regpos);
        AddrSet as(base);
        for (int e = 0; e < NE; ++e) {
            int64_t off = 0;
            for (int i = 0; i < R; ++i)
                if ((e >> i) & 1) off += (int64_t)16 << sh.gpos[rh->regpos[i]];
            const std::string src = as.operand(g, off);
            g.e("cp.async.cg.shared.global [%%r%d+%u], %s, 16;", x.base[x.low[e]], x.high[e], src.c_str());
        }
        g.e("cp.async.commit_group;");
    };
    const bool own_slot = jit_own_slot_copy() && async && !landing && nrounds > 1;
    // L2 prefetch of the tile whose first amplitude (of this thread) is at `pbase`: the 8 lanes that share the thread's
    // 128-byte lines split its 2^R lines between them
    auto l2_prefetch = [&](int pbase) {
        if (lane_lines) {
            const int pb = g.rd();
            g.e("add.s64 %%rd%d, %%rd%d, %%rd%d;", pb, pbase, koff);
            for (int j = 0; j < (1 << (R - 3)); ++j) {
                int64_t off = 0;
                for (int i = 0; i < R - 3; ++i)
                    if ((j >> i) & 1) off += step0[i];
                const int q = g.rd();
                g.e("add.s64 %%rd%d, %%rd%d, %lld;", q, pb, (long long)off);
                g.e("prefetch.global.L2 [%%rd%d];", q);
                g.e("prefetch.global.L2 [%%rd%d+64];", q);
            }
        } else {
            const int m = g.r32(), p = g.pr(), skip2 = g.label();
            g.e("and.b32 %%r%d, %%r%d, 3;", m, tid);
            g.e("setp.ne.u32 %%p%d, %%r%d, 0;", p, m);
            g.e("@%%p%d bra L%d;", p, skip2);
            for (int e = 0; e < NE; ++e) {
                const int q = g.rd();
                g.e("add.s64 %%rd%d, %%rd%d, %lld;", q, pbase, (long long)off0[e]);
                g.e("prefetch.global.L2 [%%rd%d];", q);
            }
            g.e("L%d:", skip2);
        }
    };
    // QFB_JIT_L2PF=1 (experiments; measured 1 % SLOWER, default off): with the asynchronous copy the next tile is also
    // prefetched into L2 at the TOP of the iteration, most of a tile period before its copy is issued (the copy starts
    // under the last round only, because the exchange buffer is busy until then)
    const char *l2pf_env = getenv("QFB_JIT_L2PF");
    const bool l2pf = l2pf_env && *l2pf_env && atoi(l2pf_env) != 0;
    if (async) {
        if (l2pf && nrounds > 1 && !landing) {
            const int skip = g.label();
            g.e("@!%%p%d bra L%d;", has_next, skip);
            l2_prefetch(tile_base(gb_next));
            g.e("L%d:", skip);
        }
        // ---- round 0: the tile was copied into shared memory during the previous iteration ----
        g.e("cp.async.wait_group 0;");
        if (own_slot) g.e("bar.sync 0;");       // the tile was copied by the threads that owned the slots in the last round
        const XchgAddr x = exchange_addresses(g, land, stb(0), r0->regpos);
        for (int e = 0; e < NE; ++e) {
            st.a[e].re = g.fd();
            st.a[e].im = g.fd();
            g.e("ld.shared.v2.f64 {%%fd%d, %%fd%d}, [%%r%d+%u];", st.a[e].re, st.a[e].im, x.base[x.low[e]], x.high[e]);
        }
        if (nrounds == 1 || landing) {
            // no exchange in this sweep: the buffer is the thread's own, the next copy can start at once
            const int skip = g.label();
            g.e("@!%%p%d bra L%d;", has_next, skip);
            async_copy(tile_base(gb_next));
            g.e("L%d:", skip);
        }
    } else {
    // ---- round 0: coalesced loads straight into registers ----
    const int base0 = tile_base(gb);
    AddrSet as0(base0);
    for (int e = 0; e < NE; ++e) {
        st.a[e].re = g.fd();
        st.a[e].im = g.fd();
        const std::string src = as0.operand(g, off0[e]);
        g.e("ld.global.cs.v2.f64 {%%fd%d, %%fd%d}, %s;", st.a[e].re, st.a[e].im, src.c_str());
    }
    {
        // prefetch: delta = 16 * (gb_next - gb)
        const int skip = g.label();
        g.e("@!%%p%d bra L%d;", has_next, skip);
        const int delta = g.rd(), pbase = g.rd();
        g.e("sub.s64 %%rd%d, %%rd%d, %%rd%d;", delta, gb_next, gb);
        g.e("shl.b64 %%rd%d, %%rd%d, 4;", delta, delta);
        g.e("add.s64 %%rd%d, %%rd%d, %%rd%d;", pbase, base0, delta);
        l2_prefetch(pbase);
        g.e("L%d:", skip);
    }
    }
    // ---- rounds ----
    const char *barb_env = getenv("QFB_JIT_BARB");
    const bool keep_barriers = barb_env && *barb_env && atoi(barb_env) != 0;
    for (int r = 0; r < nrounds; ++r) {
        const qfb_round_header *rh = rounds[r].rh;
        st.tfull = g.rd();
        g.e("or.b64 %%rd%d, %%rd%d, %%rd%d;", st.tfull, higb, tg(r));
        st.scalar_live = false;
        for (int j = 0; j < RMAX; ++j) st.pend[j] = -1;
        const uint8_t *op = rounds[r].ops;
        for (;;) {
            OpView v;
            memcpy(&v.h, op, sizeof(v.h));
            v.payload = op + sizeof(qfb_op_header);
            if (v.h.handler == QFB_H_END) break;
            const int rc = emit_op(g, st, v, err);
            if (rc != QFB_OK) return rc;
            op += v.h.bytes;
        }
        if (st.scalar_live) {
            const int ns = g.fd();
            g.e("neg.f64 %%fd%d, %%fd%d;", ns, st.phi);
            for (int e = 0; e < NE; ++e) cmul_fresh(g, st.a[e], st.phr, st.phi, ns);
        }
        if (r + 1 == nrounds) break;
        // exchange: this round's assignment out, the next round's in
        {
            if (pending_mask(st) == 0) {
                const XchgAddr x = exchange_addresses(g, smem, stb(r), rh->regpos);
                for (int e = 0; e < NE; ++e)
                    g.e("st.shared.v2.f64 [%%r%d+%u], {%%fd%d, %%fd%d};", x.base[x.low[e]], x.high[e], st.a[e].re,
                        st.a[e].im);
            } else {
                const XchgStore x = exchange_store_addresses(g, smem, stb(r), rh->regpos, st.pend);
                for (int e = 0; e < NE; ++e)
                    g.e("st.shared.v2.f64 [%%r%d+%u], {%%fd%d, %%fd%d};", x.reg[e], x.imm[e], st.a[e].re, st.a[e].im);
            }
        }
        g.e("bar.sync 0;");
        {
            const XchgAddr x = exchange_addresses(g, smem, stb(r + 1), rounds[r + 1].rh->regpos);
            for (int e = 0; e < NE; ++e) {
                st.a[e].re = g.fd();
                st.a[e].im = g.fd();
                g.e("ld.shared.v2.f64 {%%fd%d, %%fd%d}, [%%r%d+%u];", st.a[e].re, st.a[e].im, x.base[x.low[e]], x.high[e]);
            }
        }
        // The barrier after the read-back protects the buffer against the NEXT writer. Between two exchanges of a tile
        // that writer is the thread itself (it writes the slots of its own assignment of round r + 1, which nobody else
        // read), so the barrier is only needed after the LAST exchange: there the asynchronous copy of the next tile (or,
        // without it, the next tile's first exchange) writes round-0 slots that other threads may still be reading.
        // QFB_JIT_BARB=1 keeps every barrier (experiments).
        if (own_slot && r + 2 == nrounds) {
            const int skip = g.label();
            g.e("@!%%p%d bra L%d;", has_next, skip);
            async_copy_round(gb_next, r + 1);
            g.e("L%d:", skip);
            // The copy has no consumer inside the iteration, so ptxas gives it the lowest priority and lets it sink
            // behind the whole last round (measured: issued just before the stores, 19 % of the warp time waiting for
            // it at the top of the next tile). A shared-memory load of one of the slots the copy writes cannot move
            // in front of it; its value (masked to zero at run time, made warp uniform by a vote so that the
            // coefficient loads stay on the uniform datapath) goes into the address of the last round's coefficients,
            // which puts the copy on the critical path of the round's arithmetic.
            if (g.coef_base >= 0) {
                const XchgAddr x = exchange_addresses(g, land, stb(r + 1), rounds[r + 1].rh->regpos);
                const int z = g.r32(), zm = g.r32(), pz = g.pr(), b = g.r32(), b64 = g.rd(), cb2 = g.rd();
                g.e("ld.shared.u32 %%r%d, [%%r%d+%u];", z, x.base[x.low[0]], x.high[0]);
                g.e("cvt.u32.u64 %%r%d, %%rd%d;", zm, zmask);
                g.e("and.b32 %%r%d, %%r%d, %%r%d;", z, z, zm);
                g.e("setp.ne.u32 %%p%d, %%r%d, 0;", pz, z);
                g.e("vote.sync.ballot.b32 %%r%d, %%p%d, 0xffffffff;", b, pz);
                g.e("and.b32 %%r%d, %%r%d, %%r%d;", b, b, zm);
                g.e("cvt.u64.u32 %%rd%d, %%r%d;", b64, b);
                g.e("add.u64 %%rd%d, %%rd%d, %%rd%d;", cb2, g.coef_base, b64);
                g.coef_base = cb2;
            }
            continue;
        }
        if (r + 2 == nrounds || keep_barriers) g.e("bar.sync 0;");
        if (async && !landing && r + 2 == nrounds) {
            // the tile has left the exchange buffer for the last time: the next tile's copy runs under the last round
            const int skip = g.label();
            g.e("@!%%p%d bra L%d;", has_next, skip);
            async_copy(tile_base(gb_next));
            g.e("L%d:", skip);
        }
    }
    // ---- store: amplitude i goes to address i ^ store_xor; tile bit j is stored at spos[j] ----
    if (store_sync) g.e("bar.sync 0;");
    {
        const int idx = g.rd(), sbase = g.rd();
        g.e("or.b64 %%rd%d, %%rd%d, %%rd%d;", idx, gb, tg_store_fn());
        uint64_t regmask = 0;
        for (int i = 0; i < R; ++i) regmask |= 1ull << store_ipos[store_rh->regpos[i]];
        const uint64_t fixed_xor = sh.store_xor & ~regmask;
        if (fixed_xor) g.e("xor.b64 %%rd%d, %%rd%d, %llu;", idx, idx, (unsigned long long)fixed_xor);
        g.e("shl.b64 %%rd%d, %%rd%d, 4;", idx, idx);
        g.e("add.u64 %%rd%d, %%rd%d, %%rd%d;", sbase, state, idx);
        // pending conditional flips of the last round go into the store address: bit j of the destination is
        // e_j ^ (static flip) ^ (predicate), so every value v of the pending bits gets a root address of its own
        const int jm = pending_mask(st);
        std::map<int, AddrSet> roots;
        for (int v = 0; v < NE; ++v) {
            if (v & ~jm) continue;
            int root = sbase;
            for (int j = 0; j < R; ++j) {
                if (!((jm >> j) & 1)) continue;
                const int bit = store_ipos[store_rh->regpos[j]];
                const int a = ((v >> j) & 1) ^ (int)((sh.store_xor >> bit) & 1ull);
                const int sel = g.rd(), nr = g.rd();
                const long long step = (long long)16 << bit;
                g.e("selp.b64 %%rd%d, %lld, %lld, %%p%d;", sel, a ? 0ll : step, a ? step : 0ll, st.pend[j]);
                g.e("add.u64 %%rd%d, %%rd%d, %%rd%d;", nr, root, sel);
                root = nr;
            }
            roots.emplace(v, AddrSet(root));
        }
        for (int e = 0; e < NE; ++e) {
            int64_t off = 0;
            for (int i = 0; i < R; ++i) {
                if ((jm >> i) & 1) continue;
                const int bit = store_ipos[store_rh->regpos[i]];
                const int flipped = (int)((sh.store_xor >> bit) & 1ull);
                if (((e >> i) & 1) ^ flipped) off += (int64_t)16 << bit;
            }
            const std::string dst = roots.at(e & jm).operand(g, off);
            if (G > 1)
                g.e("@%%p%d st.global.cs.v2.f64 %s, {%%fd%d, %%fd%d};", active, dst.c_str(), st.a[e].re, st.a[e].im);
            else
                g.e("st.global.cs.v2.f64 %s, {%%fd%d, %%fd%d};", dst.c_str(), st.a[e].re, st.a[e].im);
        }
    }
    // ---- next tile ----
    g.e("mov.u64 %%rd%d, %%rd%d;", tile, tile_next);
    g.e("mov.u64 %%rd%d, %%rd%d;", gb, gb_next);
    g.e("setp.lt.u64 %%p%d, %%rd%d, %llu;", active, tile, ntiles);
    g.e("@%%p%d bra L_TILE;", has_next);
    g.e("L_EXIT:");
    g.e("ret;");

    // resident CTAs the register file allows: 2^R amplitudes = 4 * 2^R registers + ~40 per thread
    // (R = 4: 96 registers = 5 CTAs of 128 threads; ptxas then keeps the coefficients in uniform registers (LDCU) and
    // spills ~10 words per thread; measured 137.3 against 141.3 ms per step with 128 registers = 4 CTAs)
    const int regs_per_thread = (R == 5) ? 168 : (R == 4) ? 96 : 64;
    const size_t smem_per_cta = ((size_t)16 << M) * G * (landing ? 2 : 1);
    const int by_regs = (65536 / (regs_per_thread * T)) / G, by_smem = (int)((227 * 1024) / (smem_per_cta + 1024));
    int minb = std::max(1, std::min(8, std::min(by_regs, by_smem)));
    if (const char *v = getenv("QFB_JIT_MINB")) {        // experiments: resident CTAs per SM the register budget is cut for
        if (*v && atoi(v) > 0) minb = std::min(atoi(v), by_smem);
    }
    const size_t coef_bytes = std::max<size_t>(16, (g.coef.size() * 8 + 15) / 16 * 16);
    char head[1024];
    snprintf(head, sizeof(head),
             ".version 8.7\n.target sm_100a\n.address_size 64\n"
             ".const .align 16 .b8 qfb_coef[%zu];\n"
             ".extern .shared .align 128 .b8 qfb_smem[];\n"
             ".visible .entry qfb_sweep(.param .u64 p_state, .param .u64 p_hi, .param .u64 p_zero, .param .u64 p_fix)\n"
             ".maxntid %d, 1, 1\n.minnctapersm %d\n{\n"
             ".reg .f64 %%fd<%d>;\n.reg .b64 %%rd<%d>;\n.reg .b32 %%r<%d>;\n.reg .pred %%p<%d>;\n",
             coef_bytes, T * G, minb, g.nfd + 1, g.nrd + 1, g.nr + 1, g.np + 1);
    out.ptx = std::string(head) + g.body + "}\n";
    out.coef = g.coef;
    out.coef_bytes = coef_bytes;
    out.threads = T * G;
    out.smem_bytes = smem_per_cta;
    out.nholes = nfree;
    out.groups = G;
    return QFB_OK;
}

// ---- PTX -> cubin (no GPU needed) ---------------------------------------------------------------------------

int jit_compile(const std::string &ptx, std::vector<char> &cubin, std::string &log) {
    nvPTXCompilerHandle c = nullptr;
    if (nvPTXCompilerCreate(&c, ptx.size(), ptx.c_str()) != NVPTXCOMPILE_SUCCESS) {
        log = "nvPTXCompilerCreate failed";
        return QFB_ERR_CUDA;
    }
    const char *opts[] = {"--gpu-name=sm_100a", "-O3", "--verbose"};
    const nvPTXCompileResult res = nvPTXCompilerCompile(c, 3, opts);
    size_t n = 0;
    if (res != NVPTXCOMPILE_SUCCESS) {
        nvPTXCompilerGetErrorLogSize(c, &n);
        std::vector<char> buf(n + 1, 0);
        if (n) nvPTXCompilerGetErrorLog(c, buf.data());
        log = std::string("ptx compilation failed: ") + buf.data();
        nvPTXCompilerDestroy(&c);
        return QFB_ERR_CUDA;
    }
    nvPTXCompilerGetInfoLogSize(c, &n);
    if (n) {
        std::vector<char> buf(n + 1, 0);
        nvPTXCompilerGetInfoLog(c, buf.data());
        log = buf.data();
    }
    nvPTXCompilerGetCompiledProgramSize(c, &n);
    cubin.resize(n);
    nvPTXCompilerGetCompiledProgram(c, cubin.data());
    nvPTXCompilerDestroy(&c);
    return QFB_OK;
}

// ---- compiled-image cache: one image per sweep STRUCTURE (the PTX text holds no coefficient value) ----------

namespace {
std::mutex g_cache_mutex;
std::map<std::string, std::shared_ptr<std::vector<char>>> g_cache;
uint64_t g_cache_hits = 0, g_cache_misses = 0;
}  // namespace

static int compile_cached(const std::string &ptx, std::shared_ptr<std::vector<char>> &image, std::string &log) {
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        auto it = g_cache.find(ptx);
        if (it != g_cache.end()) {
            image = it->second;
            ++g_cache_hits;
            return QFB_OK;
        }
    }
    auto cubin = std::make_shared<std::vector<char>>();
    const int rc = jit_compile(ptx, *cubin, log);
    if (rc != QFB_OK) return rc;
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    ++g_cache_misses;
    image = g_cache.emplace(ptx, cubin).first->second;
    return QFB_OK;
}

void jit_cache_stats(uint64_t *hits, uint64_t *misses) {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    if (hits) *hits = g_cache_hits;
    if (misses) *misses = g_cache_misses;
}

// ---- driver API (resolved at run time: the library must load on a machine without libcuda) -------------------

namespace {
struct Driver {
    CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*ModuleGetGlobal)(CUdeviceptr *, size_t *, CUmodule, const char *) = nullptr;
    CUresult (*MemcpyHtoD)(CUdeviceptr, const void *, size_t) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int *, CUfunction, int, size_t) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                             void **, void **) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
    bool ok = false;
};

const Driver &driver() {
    static Driver d = [] {
        Driver x;
        void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return x;
#define QFB_SYM(field, name) *(void **)(&x.field) = dlsym(h, name)
        QFB_SYM(ModuleLoadData, "cuModuleLoadData");
        QFB_SYM(ModuleUnload, "cuModuleUnload");
        QFB_SYM(ModuleGetFunction, "cuModuleGetFunction");
        QFB_SYM(ModuleGetGlobal, "cuModuleGetGlobal_v2");
        QFB_SYM(MemcpyHtoD, "cuMemcpyHtoD_v2");
        QFB_SYM(FuncSetAttribute, "cuFuncSetAttribute");
        QFB_SYM(OccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor");
        QFB_SYM(LaunchKernel, "cuLaunchKernel");
        QFB_SYM(GetErrorString, "cuGetErrorString");
#undef QFB_SYM
        x.ok = x.ModuleLoadData && x.ModuleUnload && x.ModuleGetFunction && x.ModuleGetGlobal && x.MemcpyHtoD &&
               x.FuncSetAttribute && x.OccupancyMaxActiveBlocksPerMultiprocessor && x.LaunchKernel;
        return x;
    }();
    return d;
}

const char *cu_error(CUresult r) {
    const char *s = nullptr;
    if (driver().GetErrorString) driver().GetErrorString(r, &s);
    return s ? s : "unknown driver error";
}
}  // namespace

#define QFB_CU(call)                                                                          \
    do {                                                                                      \
        CUresult r__ = (call);                                                                \
        if (r__ != CUDA_SUCCESS) {                                                            \
            set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cu_error(r__));    \
            return QFB_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

struct JitSweep {
    CUmodule module = nullptr;
    CUfunction func = nullptr;
    int threads = 0, grid = 0, resident = 0;
    uint64_t nctas = 0;
    size_t smem = 0;
};

void jit_destroy(JitSweep *s) {
    if (!s) return;
    if (s->module && driver().ok) driver().ModuleUnload(s->module);
    delete s;
}

// load a compiled sweep into the current context, write its coefficients, size its grid
static int load_sweep(const JitSource &src, const std::vector<char> &image, JitSweep **out) {
    const Driver &d = driver();
    JitSweep *s = new JitSweep();
    *out = s;
    QFB_CU(d.ModuleLoadData(&s->module, image.data()));
    QFB_CU(d.ModuleGetFunction(&s->func, s->module, "qfb_sweep"));
    CUdeviceptr cptr = 0;
    size_t cbytes = 0;
    QFB_CU(d.ModuleGetGlobal(&cptr, &cbytes, s->module, "qfb_coef"));
    if (cbytes < src.coef.size() * 8) {
        set_error("jit: coefficient bank too small");
        return QFB_ERR_CUDA;
    }
    if (!src.coef.empty()) QFB_CU(d.MemcpyHtoD(cptr, src.coef.data(), src.coef.size() * 8));
    s->threads = src.threads;
    s->smem = src.smem_bytes;
    QFB_CU(d.FuncSetAttribute(s->func, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)s->smem));
    int resident = 0;
    QFB_CU(d.OccupancyMaxActiveBlocksPerMultiprocessor(&resident, s->func, s->threads, s->smem));
    resident = std::max(1, resident);
    const uint64_t ntiles = 1ull << src.nholes;
    const uint64_t nctas = (ntiles + src.groups - 1) / src.groups;
    s->grid = (int)std::min<uint64_t>(nctas, (uint64_t)sm_count_cached() * resident);
    s->resident = resident;
    s->nctas = nctas;
    return QFB_OK;
}

// Generate + compile every sweep of a validated plan (parallel over host threads), then load the images into the
// current context and write the coefficients. `offsets` = byte offset of every sweep record in the plan.
int jit_build_plan(const uint8_t *plan, const std::vector<size_t> &offsets, int nbits, int M, int reg_bits,
                   std::vector<JitSweep *> &out) {
    const Driver &d = driver();
    if (!d.ok) {
        set_error("jit: the CUDA driver library (libcuda.so.1) is not available");
        return QFB_ERR_CUDA;
    }
    const size_t n = offsets.size();
    std::vector<JitSource> src(n);
    std::vector<std::shared_ptr<std::vector<char>>> image(n);
    std::vector<std::string> logs(n);
    std::vector<int> rcs(n, QFB_OK);
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t nthreads = std::min<size_t>(n, std::min<unsigned>(hw, 16));
    std::vector<std::thread> pool;
    for (size_t w = 0; w < nthreads; ++w) {
        pool.emplace_back([&, w] {
            for (size_t i = w; i < n; i += nthreads) {
                rcs[i] = jit_generate(plan + offsets[i], nbits, M, reg_bits, src[i], logs[i]);
                if (rcs[i] == QFB_OK) rcs[i] = compile_cached(src[i].ptx, image[i], logs[i]);
            }
        });
    }
    for (auto &t : pool) t.join();
    for (size_t i = 0; i < n; ++i) {
        if (rcs[i] != QFB_OK) {
            set_error("jit: sweep %zu: %s", i, logs[i].c_str());
            return rcs[i];
        }
    }
    for (size_t i = 0; i < n; ++i) {
        JitSweep *s = nullptr;
        const int rc = load_sweep(src[i], *image[i], &s);
        if (s) out.push_back(s);
        if (rc != QFB_OK) return rc;
    }
    return QFB_OK;
}

// One sweep restricted to the slice of the state whose index bits `fix_mask` (non-tile bits of the sweep) are given at
// launch time (jit_launch's fix_value): same code, the tile loop enumerates the remaining non-tile bits only.
int jit_build_variant(const uint8_t *rec, int nbits, int M, int reg_bits, uint64_t fix_mask, JitSweep **out) {
    *out = nullptr;
    if (!driver().ok) {
        set_error("jit: the CUDA driver library (libcuda.so.1) is not available");
        return QFB_ERR_CUDA;
    }
    JitSource src;
    std::string log;
    int rc = jit_generate(rec, nbits, M, reg_bits, src, log, fix_mask);
    std::shared_ptr<std::vector<char>> image;
    if (rc == QFB_OK) rc = compile_cached(src.ptx, image, log);
    if (rc != QFB_OK) {
        set_error("jit: slice variant: %s", log.c_str());
        return rc;
    }
    rc = load_sweep(src, *image, out);
    if (rc != QFB_OK && *out) {
        jit_destroy(*out);
        *out = nullptr;
    }
    return rc;
}

int jit_launch(JitSweep *s, void *state, uint64_t hi_shifted, cudaStream_t st, uint64_t fix_value, int ctas_per_sm) {
    uint64_t zero = 0;      // see Gen::cst
    void *args[4] = {&state, &hi_shifted, &zero, &fix_value};
    // the CTAs are persistent (a tile loop each), so a kernel on another stream only finds room on an SM if this launch
    // leaves it: ctas_per_sm > 0 bounds the resident CTAs per SM (negative: that many fewer than fit)
    int grid = s->grid;
    if (ctas_per_sm != 0) {
        const int per_sm = ctas_per_sm > 0 ? std::min(ctas_per_sm, s->resident) : std::max(1, s->resident + ctas_per_sm);
        grid = (int)std::min<uint64_t>(s->nctas, (uint64_t)sm_count_cached() * per_sm);
    }
    QFB_CU(driver().LaunchKernel(s->func, grid, 1, 1, s->threads, 1, 1, (unsigned)s->smem, (CUstream)st, args, nullptr));
    count_launch();
    return QFB_OK;
}

}  // namespace qfb
