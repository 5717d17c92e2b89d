// qfb_sweep.cu -- tiled multi-gate executor: the performance path behind Circuit.run / Circuit.evolve
// (reference loop: quantumflow/circuits.py:87-109, one np.einsum sweep per gate).
//
// One launch = one sweep = one read + one write of the state (algorithmic traffic 32 B per amplitude,
// 32 * 2^n bytes per launch), no matter how many gates the planner packed into it. See qfb_plan.h for the
// plan layout. Per tile of 2^M amplitudes:
//
//   round 0 : coalesced LDG.128 straight into registers (lanes span the low index bits -> every warp
//             request covers whole 128-byte lines), apply the round's ops on the 2^R register amplitudes
//   round r : STS.128 -> barrier -> LDS.128 with the next register/thread bit assignment -> barrier -> ops
//   last    : coalesced STG.128 from registers
//
// Shared-memory layout is XOR swizzled at 16-byte granularity: phys = idx ^ fold3(idx >> 3) where fold3 XORs
// the upper index bits into 3 bits (bit p >= 3 lands on bit (p-3) % 3). The planner assigns lane bits 0..2 of
// every round to tile bits of three different classes, which makes every quarter-warp LDS/STS.128 hit 8
// distinct 16-byte bank groups (conflict free) for any choice of register bits.
//
// Roofline: HBM bound by design (planner caps the FP64 work per sweep); FP64 FMA pipe and shared-memory
// bandwidth are the secondary limits ("sweep kernel").
#include <stdlib.h>
#include <algorithm>
#include <map>
#include <vector>
#include "qfb_common.cuh"
#include "qfb_jit.h"
#include "qfb_plan.h"
#include "qfb_oploop.inc"

namespace qfb {

constexpr int R = QFB_PLAN_REG_BITS;
constexpr int NE = 1 << R;  // amplitudes per thread

// x = x * s + c with the result in x's own register. The amplitudes are pinned to registers for the whole tile
// (the op interpreter is PTX with tied operands, qfb_oploop.inc); the few C++ updates go through the same form so
// that the compiler has no reason to rename them around the interpreter.
__device__ __forceinline__ void ip_fma_scale(double &x, double s, double c) {
    asm("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(s), "d"(c));
}
// a *= (fr + i fi)
__device__ __forceinline__ void cmul_inplace(c128 &a, double fr, double fi) {
    const double t0 = -fi * a.im, t1 = fi * a.re;
    ip_fma_scale(a.re, fr, t0);
    ip_fma_scale(a.im, fr, t1);
}

template <int M>
struct SweepCfg {
    static constexpr int T = 1 << (M - R);                 // threads per CTA
    // 2^R amplitudes = 4 * 2^R registers per thread: 170 registers at 3 CTAs of 128 threads (M = 12)
    static constexpr int MINB = (M >= 13) ? 1 : ((M == 12) ? 3 : ((M == 11) ? 6 : 8));
    static constexpr int TILE_BYTES = 16 << M;
};

// HAS_G2 = false drops the dense 2-bit handlers: circuits made of 1-bit, controlled-1-bit and diagonal gates
// (every workload of BASELINE.json except the density channels) run the smaller kernel, which is kinder to the
// instruction cache.
// q[e] for all 2^R register indices from a base and one step per register bit: 2^R - 1 operations instead of
// recombining the steps for every e
template <typename T, typename S, typename F>
__device__ __forceinline__ void spread(T (&q)[NE], T base, const S (&step)[R], F combine) {
    q[0] = base;
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int e = 0; e < (1 << i); ++e) q[e | (1 << i)] = combine(q[e], step[i]);
}

constexpr int HLUT_BITS = 6;                                  // tile-id bits per look-up

template <int M, bool HAS_G2>
__global__ void __launch_bounds__(SweepCfg<M>::T, SweepCfg<M>::MINB)
sweep_kernel(c128 *__restrict__ state, const uint8_t *__restrict__ rec_g, uint32_t rec_bytes, int nholes,
             uint64_t hi_shifted, int contiguous, int pf_mask) {
    constexpr int T = SweepCfg<M>::T;
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *tile = smem;
    uint64_t *hlut = reinterpret_cast<uint64_t *>(smem + SweepCfg<M>::TILE_BYTES);
    const int nlut = (nholes + HLUT_BITS - 1) / HLUT_BITS;
    uint8_t *rec = smem + SweepCfg<M>::TILE_BYTES + (nlut << (HLUT_BITS + 3));
    const int tid = threadIdx.x;
    {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(rec_g);
        uint4 *d4 = reinterpret_cast<uint4 *>(rec);
        for (uint32_t i = tid; i < rec_bytes / 16; i += T) d4[i] = s4[i];
    }
    __syncthreads();
    const qfb_sweep_header *sh = reinterpret_cast<const qfb_sweep_header *>(rec);
    const int nrounds = (int)sh->nrounds;
    const uint64_t ntiles = 1ull << nholes;
    const uint64_t store_xor = sh->store_xor;
    const bool store_sync = (sh->flags & QFB_SWEEP_FLAG_STORE_SYNC) != 0;
    const bool store_perm = (sh->flags & QFB_SWEEP_FLAG_STORE_PERM) != 0;
    // tile id -> index bits (deposit through hole[]): HLUT_BITS tile-id bits per table look-up
    for (int i = tid; i < nlut << HLUT_BITS; i += T) {
        const int k = i >> HLUT_BITS;
        uint64_t v = 0;
        for (int b = 0; b < HLUT_BITS && k * HLUT_BITS + b < nholes; ++b)
            v |= (uint64_t)((i >> b) & 1) << sh->hole[k * HLUT_BITS + b];
        hlut[i] = v;
    }
    __syncthreads();
    auto tile_base = [&](uint64_t tile_id) {
        uint64_t gb = 0;
        for (int k = 0; k < nlut; ++k) gb |= hlut[(k << HLUT_BITS) | ((tile_id >> (k * HLUT_BITS)) & ((1 << HLUT_BITS) - 1))];
        return gb;
    };
    // do the lanes of the load round walk index bits 0, 1, 2 (one 128-byte line per 8 lanes)?
    bool lane_lines = false;
    if (R >= 3 && M - R >= 3) {
        const qfb_round_header *r0 = reinterpret_cast<const qfb_round_header *>(rec + sizeof(qfb_sweep_header));
        lane_lines = sh->gpos[r0->thrpos[0]] == 0 && sh->gpos[r0->thrpos[1]] == 1 && sh->gpos[r0->thrpos[2]] == 2;
    }
    const auto add64 = [](const char *p, int64_t s) { return p + s; };
    const auto xor32 = [](uint32_t x, uint32_t s) { return x ^ s; };

    // Order in which a CTA walks the tiles (QFB_TILE_ORDER, default 0): 0 = strided over the CTAs (the CTAs
    // that run together touch neighbouring 128-byte lines), 1 = a contiguous range per CTA, 2 = strided pairs of
    // neighbouring tiles (a tile and the prefetch of the next one fall into the same DRAM page).
    const uint64_t per_cta = (ntiles + gridDim.x - 1) / gridDim.x;
    auto tile_of = [&](uint64_t i) -> uint64_t {
        if (contiguous == 1) return i < per_cta ? blockIdx.x * per_cta + i : ntiles;
        if (contiguous == 2) return ((i >> 1) * gridDim.x + blockIdx.x) * 2 + (i & 1);
        return blockIdx.x + i * gridDim.x;
    };
    for (uint64_t it = 0;; ++it) {
        const uint64_t tile_id = tile_of(it);
        if (tile_id >= ntiles) break;
        const uint64_t tile_next = tile_of(it + 1);
        const uint64_t gb = tile_base(tile_id);
        const uint8_t *rp = rec + sizeof(qfb_sweep_header);
        const qfb_round_header *rh = reinterpret_cast<const qfb_round_header *>(rp);
        // thread id -> exchange-buffer offset and index-bit image of the thread bits: two table look-ups
        uint32_t stb;
        uint64_t tg;
        {
            const uint4 l0 = *reinterpret_cast<const uint4 *>(&rh->lut_lo[tid & 15]);
            const uint4 l1 = *reinterpret_cast<const uint4 *>(&rh->lut_hi[(tid >> 4) & 31]);
            stb = l0.x ^ l1.x;
            tg = (((uint64_t)l0.w << 32) | l0.z) | (((uint64_t)l1.w << 32) | l1.z);
        }
        c128 a[NE];
        {   // round 0: coalesced LDG.128 straight into registers
            int64_t step[R];
#pragma unroll
            for (int i = 0; i < R; ++i) step[i] = rh->rgb[i];
            const char *p[NE];
            spread(p, reinterpret_cast<const char *>(state + (gb | tg)), step, add64);
#pragma unroll
            for (int e = 0; e < NE; ++e) a[e] = ldg_stream(reinterpret_cast<const c128 *>(p[e]));
            if (tile_next < ntiles) {
                // Warm L2 with the next tile's lines, one request per 64 bytes (the L2 fetch granularity: one
                // request per 128-byte line left half of the sectors to miss, ncu 49 % hit rate).
                const int64_t delta = (int64_t)(tile_base(tile_next) - gb) * 16;
                if (lane_lines) {
                    // The 8 lanes that share the thread's 128-byte lines split the 2^R lines among themselves:
                    // lane k takes the register indices whose top three bits are k (2 requests per line).
                    constexpr int LOWB = R - 3;
                    const int k = tid & 7;
                    const char *base = p[0] + delta - (k << 4);
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        if ((k >> i) & 1) base += step[LOWB + i];
#pragma unroll
                    for (int j = 0; j < (1 << LOWB); ++j) {
                        const char *q = base;
#pragma unroll
                        for (int i = 0; i < LOWB; ++i)
                            if ((j >> i) & 1) q += step[i];
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 64));
                    }
                } else if ((tid & pf_mask) == 0) {
#pragma unroll
                    for (int e = 0; e < NE; ++e) asm volatile("prefetch.global.L2 [%0];" ::"l"(p[e] + delta));
                }
            }
        }

        int round = 0;
        for (;;) {
            const uint64_t tfull = hi_shifted | gb | tg;
            double phr = 1.0, phi = 0.0;  // running per-thread scalar phase of this round
            // ---- op interpreter (PTX, qfb_oploop.inc): one brx.idx per op, amplitudes pinned to registers ----
            uint32_t op = (uint32_t)__cvta_generic_to_shared(rp + sizeof(qfb_round_header));
#define QFB_AMP(e) "+d"(a[e].re), "+d"(a[e].im)
#define QFB_OPLOOP_OPERANDS                                                                                   \
    QFB_AMP(0), QFB_AMP(1), QFB_AMP(2), QFB_AMP(3), QFB_AMP(4), QFB_AMP(5), QFB_AMP(6), QFB_AMP(7), QFB_AMP(8),    \
        QFB_AMP(9), QFB_AMP(10), QFB_AMP(11), QFB_AMP(12), QFB_AMP(13), QFB_AMP(14), QFB_AMP(15), QFB_AMP(16),    \
        QFB_AMP(17), QFB_AMP(18), QFB_AMP(19), QFB_AMP(20), QFB_AMP(21), QFB_AMP(22), QFB_AMP(23), QFB_AMP(24),   \
        QFB_AMP(25), QFB_AMP(26), QFB_AMP(27), QFB_AMP(28), QFB_AMP(29), QFB_AMP(30), QFB_AMP(31), "+d"(phr),     \
        "+d"(phi), "+r"(op)                                                                                   \
        : "l"(tfull)
            if constexpr (HAS_G2) {
                asm volatile(QFB_OPLOOP_PTX_G2 : QFB_OPLOOP_OPERANDS);
            } else {
                asm volatile(QFB_OPLOOP_PTX : QFB_OPLOOP_OPERANDS);
            }
#undef QFB_AMP
#undef QFB_OPLOOP_OPERANDS
            if (rh->has_scalar) {
#pragma unroll
                for (int e = 0; e < NE; ++e) cmul_inplace(a[e], phr, phi);
            }
            if (++round == nrounds) break;

            // ---- exchange: this round's assignment out, the next round's in (swizzled, conflict free) ----
            {
                uint32_t ps[R], q[NE];
#pragma unroll
                for (int i = 0; i < R; ++i) ps[i] = rh->ps_b[i];
                spread(q, stb, ps, xor32);
#pragma unroll
                for (int e = 0; e < NE; ++e) *reinterpret_cast<c128 *>(tile + q[e]) = a[e];
            }
            __syncthreads();
            rp += rh->bytes;
            rh = reinterpret_cast<const qfb_round_header *>(rp);
            {
                const uint4 l0 = *reinterpret_cast<const uint4 *>(&rh->lut_lo[tid & 15]);
                const uint4 l1 = *reinterpret_cast<const uint4 *>(&rh->lut_hi[(tid >> 4) & 31]);
                stb = l0.x ^ l1.x;
                tg = (((uint64_t)l0.w << 32) | l0.z) | (((uint64_t)l1.w << 32) | l1.z);
                uint32_t ps[R], q[NE];
#pragma unroll
                for (int i = 0; i < R; ++i) ps[i] = rh->ps_b[i];
                spread(q, stb, ps, xor32);
#pragma unroll
                for (int e = 0; e < NE; ++e) a[e] = *reinterpret_cast<const c128 *>(tile + q[e]);
            }
            __syncthreads();  // everyone has read before anyone overwrites the tile again
        }

        // ---- last round: coalesced STG.128; amplitude i goes to address i ^ store_xor (pending X flips): the
        // base takes the XOR, the per-register-bit steps of a flipped bit are negative (rst[], set by the planner)
        if (store_sync) __syncthreads();   // single-round sweep: all loads of the tile precede the permuted stores
        if (store_perm) {
            // in-place permutation of the tile bits (qubit remap of a sharded state): a header-only record after
            // the last round holds the images of the thread and register bits under the STORE bit positions
            rh = reinterpret_cast<const qfb_round_header *>(rp + rh->bytes);
            const uint4 l0 = *reinterpret_cast<const uint4 *>(&rh->lut_lo[tid & 15]);
            const uint4 l1 = *reinterpret_cast<const uint4 *>(&rh->lut_hi[(tid >> 4) & 31]);
            tg = (((uint64_t)l0.w << 32) | l0.z) | (((uint64_t)l1.w << 32) | l1.z);
        }
        {
            int64_t step[R];
#pragma unroll
            for (int i = 0; i < R; ++i) step[i] = rh->rst[i];
            const char *p[NE];
            spread(p, reinterpret_cast<const char *>(state + ((gb | tg) ^ store_xor)), step, add64);
#pragma unroll
            for (int e = 0; e < NE; ++e) stg_stream(reinterpret_cast<c128 *>(const_cast<char *>(p[e])), a[e]);
        }
    }
}

// ---- host side ----
struct SweepInfo {
    size_t offset;  // byte offset of the sweep record inside the device copy
    uint32_t bytes;
    bool has_g2;    // selects the kernel variant with the dense 2-bit handlers
};

struct PlanHandle {
    uint32_t magic;
    int nbits, tile_bits, reg_bits, device;
    std::vector<SweepInfo> sweeps;
    std::vector<JitSweep *> jit;   // sweep-specialised kernels (qfb_jit.cu); empty = the interpreter runs the plan
    void *dev;
    size_t dev_bytes;
    // slice launches (qfb_plan_launch_part): host copy of the plan and the kernels built for (sweep, fixed bits)
    std::vector<uint8_t> host;
    std::map<std::pair<int, uint64_t>, JitSweep *> variants;
};

// QFB_JIT: 0 = interpreter only, 1 = sweep-specialised kernels for every plan the generator supports, unset = for
// states of at least QFB_JIT_MIN_BITS index bits (default 24: below that a sweep is so short that compiling it --
// a few hundred milliseconds per distinct sweep structure -- never pays off)
static bool jit_wanted(int nbits, int tile_bits, int reg_bits) {
    if (reg_bits != R) return true;     // only the sweep-specialised kernels run plans with 4 register bits
    if (tile_bits - R < 3) return false;
    const char *v = getenv("QFB_JIT");
    if (v && *v) return atoi(v) != 0;
    const char *m = getenv("QFB_JIT_MIN_BITS");
    return nbits >= (m && *m ? atoi(m) : 24);
}

// register-bit pairs (j0 > j1) in handler order
static const int J0[10] = {1, 2, 2, 3, 3, 3, 4, 4, 4, 4}, J1[10] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3};
constexpr int NPAIRS = R * (R - 1) / 2;
static_assert(R == 5 && QFB_H_CPH_NEG2 + NPAIRS == QFB_H_CPH_REGM && QFB_H_G2 + NPAIRS == QFB_H_G2X && QFB_H_G2X + NPAIRS == QFB_H_CPH_TABLE,
              "handler ids in qfb_plan.h assume R = 5");

static uint32_t swz_host(uint32_t idx) {
    const uint32_t x = idx >> 3;
    return idx ^ ((x ^ (x >> 3) ^ (x >> 6) ^ (x >> 9)) & 7u);
}

static int validate_plan(const uint8_t *p, size_t nbytes, std::vector<SweepInfo> &sweeps, int &nbits, int &M,
                         int *reg_bits_out = nullptr) {
    QFB_CHECK_ARG(p && nbytes >= sizeof(qfb_plan_header), "plan: too small");
    qfb_plan_header h;
    memcpy(&h, p, sizeof(h));
    QFB_CHECK_ARG(h.magic == QFB_PLAN_MAGIC && h.version == QFB_PLAN_VERSION, "plan: bad magic/version");
    QFB_CHECK_ARG(h.total_bytes == nbytes, "plan: size mismatch (%llu vs %llu)", (unsigned long long)h.total_bytes,
                  (unsigned long long)nbytes);
    // 5 register bits: interpreter (sweep_kernel) and sweep-specialised kernels; 3 or 4: sweep-specialised kernels only
    QFB_CHECK_ARG(h.reg_bits >= 3u && h.reg_bits <= 5u, "plan: reg_bits=%u unsupported", h.reg_bits);
    const int RB = (int)h.reg_bits, NEB = 1 << RB;   // handler ids keep their stride of R = 5 register bits
    if (reg_bits_out) *reg_bits_out = RB;
    QFB_CHECK_ARG(h.tile_bits >= QFB_PLAN_MIN_TILE_BITS && h.tile_bits <= QFB_PLAN_MAX_TILE_BITS &&
                      h.tile_bits <= h.nbits,
                  "plan: tile_bits=%u unsupported (nbits=%u)", h.tile_bits, h.nbits);
    QFB_CHECK_ARG(h.nbits - h.tile_bits <= QFB_PLAN_MAX_HOLES, "plan: nbits=%u too large", h.nbits);
    nbits = (int)h.nbits;
    M = (int)h.tile_bits;
    const uint32_t nholes = h.nbits - h.tile_bits;
    size_t off = sizeof(h);
    for (uint32_t s = 0; s < h.nsweeps; ++s) {
        QFB_CHECK_ARG(off + sizeof(qfb_sweep_header) <= nbytes, "plan: truncated sweep %u", s);
        qfb_sweep_header sh;
        memcpy(&sh, p + off, sizeof(sh));
        QFB_CHECK_ARG(sh.bytes % 16 == 0 && sh.bytes >= sizeof(sh) && off + sh.bytes <= nbytes &&
                          sh.bytes <= QFB_PLAN_MAX_SWEEP_BYTES,
                      "plan: sweep %u has bad size %u", s, sh.bytes);
        QFB_CHECK_ARG(sh.nrounds >= 1, "plan: sweep %u has no rounds", s);
        // tile bits and holes must partition [0, nbits)
        uint64_t seen = 0, tilemask = 0;
        for (int j = 0; j < M; ++j) {
            QFB_CHECK_ARG(sh.gpos[j] < h.nbits && !((seen >> sh.gpos[j]) & 1ull), "plan: sweep %u bad gpos", s);
            seen |= 1ull << sh.gpos[j];
        }
        tilemask = seen;
        for (uint32_t i = 0; i < nholes; ++i) {
            QFB_CHECK_ARG(sh.hole[i] < h.nbits && !((seen >> sh.hole[i]) & 1ull), "plan: sweep %u bad hole", s);
            seen |= 1ull << sh.hole[i];
        }
        QFB_CHECK_ARG((sh.store_xor & ~tilemask) == 0, "plan: sweep %u store_xor leaves the tile", s);
        bool perm = false;
        {
            uint64_t sseen = 0;
            for (int j = 0; j < M; ++j) {
                QFB_CHECK_ARG(sh.spos[j] < h.nbits && ((tilemask >> sh.spos[j]) & 1ull) && !((sseen >> sh.spos[j]) & 1ull),
                              "plan: sweep %u spos is not a permutation of the tile bits", s);
                sseen |= 1ull << sh.spos[j];
                perm = perm || sh.spos[j] != sh.gpos[j];
            }
        }
        QFB_CHECK_ARG(((sh.flags & QFB_SWEEP_FLAG_STORE_PERM) != 0) == perm, "plan: sweep %u bad store-perm flag", s);
        const bool want_sync = (sh.store_xor != 0 || perm) && sh.nrounds == 1;
        QFB_CHECK_ARG(((sh.flags & QFB_SWEEP_FLAG_STORE_SYNC) != 0) == want_sync, "plan: sweep %u bad store-sync flag", s);
        size_t roff = off + sizeof(sh);
        const size_t send = off + sh.bytes;
        bool sweep_g2 = false;
        qfb_round_header last_rh;
        memset(&last_rh, 0, sizeof(last_rh));
        // rounds, then (with a store permutation) the header-only store record, checked against spos[]
        for (uint32_t r = 0; r < sh.nrounds + (perm ? 1u : 0u); ++r) {
            QFB_CHECK_ARG(roff + sizeof(qfb_round_header) <= send, "plan: sweep %u truncated round %u", s, r);
            qfb_round_header rh;
            memcpy(&rh, p + roff, sizeof(rh));
            const bool store_rec = r == sh.nrounds;
            const uint8_t *ipos = store_rec ? sh.spos : sh.gpos;   // index-bit images used by this record
            if (store_rec) {
                QFB_CHECK_ARG(rh.nops == 0 && memcmp(rh.regpos, last_rh.regpos, RB) == 0 &&
                                  memcmp(rh.thrpos, last_rh.thrpos, 12) == 0,
                              "plan: sweep %u store record does not match the last round", s);
            }
            sweep_g2 = sweep_g2 || rh.has_g2;
            uint32_t tseen = 0;
            for (int i = 0; i < RB; ++i) {
                QFB_CHECK_ARG(rh.regpos[i] < M && !((tseen >> rh.regpos[i]) & 1u), "plan: bad regpos");
                tseen |= 1u << rh.regpos[i];
                const int64_t gbytes = (int64_t)16 << sh.gpos[rh.regpos[i]];
                const int64_t sbytes = (int64_t)16 << ipos[rh.regpos[i]];
                const bool flipped = (sh.store_xor >> ipos[rh.regpos[i]]) & 1ull;
                QFB_CHECK_ARG(rh.ps_b[i] == (swz_host(1u << rh.regpos[i]) << 4) && rh.rgb[i] == gbytes &&
                                  rh.rst[i] == (flipped ? -sbytes : sbytes),
                              "plan: sweep %u round %u bad register-bit images", s, r);
            }
            for (int t = 0; t < M - RB; ++t) {
                QFB_CHECK_ARG(rh.thrpos[t] < M && !((tseen >> rh.thrpos[t]) & 1u), "plan: bad thrpos");
                tseen |= 1u << rh.thrpos[t];
            }
            // thread LUTs must reproduce the deposit of the thread bits through thrpos / gpos
            for (int t = 0; t < (1 << (M - RB)); ++t) {
                uint32_t tb = 0;
                uint64_t tg = 0;
                for (int b = 0; b < M - RB; ++b) {
                    if ((t >> b) & 1) {
                        tb |= 1u << rh.thrpos[b];
                        tg |= 1ull << ipos[rh.thrpos[b]];
                    }
                }
                const qfb_thread_lut &lo = rh.lut_lo[t & 15], &hi = rh.lut_hi[(t >> 4) & 31];
                QFB_CHECK_ARG((lo.tb | hi.tb) == tb && (lo.tg | hi.tg) == tg && (lo.stb ^ hi.stb) == (swz_host(tb) << 4),
                              "plan: sweep %u round %u bad thread LUT", s, r);
            }
            size_t ooff = roff + sizeof(rh);
            const size_t rend = roff + rh.bytes;
            bool ended = false;
            for (uint32_t o = 0; o <= rh.nops; ++o) {
                QFB_CHECK_ARG(ooff + sizeof(qfb_op_header) <= rend, "plan: truncated op");
                qfb_op_header oh;
                memcpy(&oh, p + ooff, sizeof(oh));
                const uint32_t bytes = oh.bytes;
                QFB_CHECK_ARG(bytes % 16 == 0 && bytes >= sizeof(oh) && ooff + bytes <= rend, "plan: op bad size");
                const int hd = (int)oh.handler;
                const int rcm = oh.reg_cmask;
                // register bits an op names must exist in this plan
                if (hd < QFB_H_CPH_SCALAR) QFB_CHECK_ARG(hd % R < RB, "plan: op on a register bit the plan does not have");
                if (o == rh.nops) {
                    QFB_CHECK_ARG(hd == QFB_H_END && bytes == 16, "plan: round does not end with an END record");
                    ended = true;
                } else if (hd >= QFB_H_G1_GENERAL && hd < QFB_H_G1_SUMDIFF) {
                    QFB_CHECK_ARG(bytes == 16 + 64 && rcm == 0 && oh.idx_cmask == 0, "plan: bad G1 op");
                } else if (hd >= QFB_H_G1_SUMDIFF && hd < QFB_H_G1C_GENERAL) {
                    QFB_CHECK_ARG(bytes == 32 && rcm == 0 && oh.idx_cmask == 0, "plan: bad pivoted G1 op");
                } else if (hd >= QFB_H_G1C_GENERAL && hd < QFB_H_G1C_SWAPX + R) {
                    const int j = (hd - QFB_H_G1C_GENERAL) % R;
                    const uint32_t want = hd < QFB_H_G1C_SWAPX ? 16 + 64 : 32;   // controlled X: the constant 1.0
                    if (hd >= QFB_H_G1C_SWAPX) {
                        double one;
                        memcpy(&one, p + ooff + 16, 8);
                        QFB_CHECK_ARG(bytes == 32 && one == 1.0, "plan: controlled X needs the payload 1.0");
                    }
                    QFB_CHECK_ARG(bytes == want && !((rcm >> j) & 1) && rcm < NEB, "plan: bad controlled G1 op");
                } else if (hd == QFB_H_CPH_SCALAR) {
                    QFB_CHECK_ARG(bytes == 32 && rcm == 0 && rh.has_scalar == 1, "plan: bad scalar CPH op");
                } else if (hd >= QFB_H_CPH_REG1 && hd < QFB_H_CPH_NEG2) {
                    QFB_CHECK_ARG(bytes == 32 && rcm == (1 << ((hd - QFB_H_CPH_REG1) % R)), "plan: bad 1-bit CPH op");
                    if (hd >= QFB_H_CPH_RSC1 && hd < QFB_H_CPH_NEG1) {
                        double im;
                        memcpy(&im, p + ooff + 24, 8);
                        QFB_CHECK_ARG(im == 0.0, "plan: real-scale CPH op with a complex factor");
                    }
                } else if (hd >= QFB_H_CPH_NEG2 && hd < QFB_H_CPH_REGM) {
                    const int pi = hd - QFB_H_CPH_NEG2;
                    QFB_CHECK_ARG(bytes == 32 && rcm == ((1 << J0[pi]) | (1 << J1[pi])), "plan: bad 2-bit CPH op");
                } else if (hd == QFB_H_CPH_REGM || hd == QFB_H_CPH_NEGM) {
                    QFB_CHECK_ARG(bytes == 32 && rcm > 0 && rcm < NEB, "plan: bad CPH op");
                } else if (hd == QFB_H_CPH_TABLE) {
                    QFB_CHECK_ARG(bytes == 16 + 16 * NE && oh.idx_cmask == 0 && rcm < NEB && oh.flag <= 1 &&
                                      (rcm != 0 || oh.flag == 1),
                                  "plan: bad diagonal table op");
                } else if (hd >= QFB_H_G2 && hd < QFB_H_G2 + NPAIRS) {
                    const int j0 = J0[hd - QFB_H_G2], j1 = J1[hd - QFB_H_G2];
                    QFB_CHECK_ARG(rh.has_g2 == 1 && bytes == 16 + 272 && !((rcm >> j0) & 1) && !((rcm >> j1) & 1) &&
                                      rcm < NEB,
                                  "plan: bad G2 op");
                } else if (hd >= QFB_H_G2X && hd < QFB_H_G2X + NPAIRS) {
                    const int j0 = J0[hd - QFB_H_G2X], j1 = J1[hd - QFB_H_G2X];
                    QFB_CHECK_ARG(rh.has_g2 == 1 && bytes == 16 + 64 && !((rcm >> j0) & 1) && !((rcm >> j1) & 1) &&
                                      rcm < NEB,
                                  "plan: bad X-shaped G2 op");
                } else {
                    QFB_CHECK_ARG(false, "plan: unknown handler %d", hd);
                }
                ooff += bytes;
            }
            QFB_CHECK_ARG(ended, "plan: missing END record");
            QFB_CHECK_ARG(ooff == rend, "plan: round size mismatch");
            roff += rh.bytes;
            last_rh = rh;
        }
        QFB_CHECK_ARG(roff == send, "plan: sweep size mismatch");
        QFB_CHECK_ARG(((sh.flags & QFB_SWEEP_FLAG_G2) != 0) == sweep_g2, "plan: sweep %u bad G2 flag", s);
        sweeps.push_back(SweepInfo{off, sh.bytes, sweep_g2});
        off += sh.bytes;
    }
    QFB_CHECK_ARG(off == nbytes, "plan: trailing bytes");
    return QFB_OK;
}

template <int M, bool G2>
static int launch_sweep(c128 *state, const uint8_t *rec_dev, uint32_t rec_bytes, int nbits, uint64_t index_hi,
                        cudaStream_t st) {
    constexpr int T = SweepCfg<M>::T;
    // +32: the op loop reads one header + 16 payload bytes past the END record of a round
    const int nholes = nbits - M;
    const size_t hlut_bytes = (size_t)((nholes + HLUT_BITS - 1) / HLUT_BITS) << (HLUT_BITS + 3);
    // QFB_SMEM_PAD (bytes): occupancy experiments only -- extra dynamic shared memory lowers the CTAs per SM
    static const size_t smem_pad = [] {
        const char *v = getenv("QFB_SMEM_PAD");
        return v ? (size_t)atol(v) : (size_t)0;
    }();
    const size_t smem = (size_t)SweepCfg<M>::TILE_BYTES + hlut_bytes + rec_bytes + 32 + smem_pad;
    static thread_local size_t configured[64] = {0};
    int dev = 0;
    QFB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 0;
    if (smem > configured[dev]) {
        // grow in 8 KiB steps so the attribute is set a handful of times per process
        const size_t want = std::min<size_t>(((smem + 8191) / 8192) * 8192, 227 * 1024);
        QFB_CHECK_ARG(smem <= want, "sweep: %zu bytes of shared memory exceed the 227 KiB limit", smem);
        QFB_CUDA(cudaFuncSetAttribute(sweep_kernel<M, G2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
        configured[dev] = want;
    }
    int resident = 0;  // CTAs per SM for this launch's shared-memory footprint (host-side arithmetic)
    QFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, sweep_kernel<M, G2>, T, smem));
    resident = std::max(1, resident);
    const uint64_t ntiles = 1ull << nholes;
    const uint64_t cap = (uint64_t)sm_count_cached() * resident;
    const int grid = (int)std::min<uint64_t>(ntiles, cap);
    const uint64_t hi_shifted = (nbits >= 64) ? 0ull : (index_hi << nbits);
    static const int tile_order = [] {
        const char *v = getenv("QFB_TILE_ORDER");
        return v ? atoi(v) : 0;
    }();
    static const int pf_mask = [] {     // QFB_PF_MASK: lanes whose (tid & mask) == 0 issue the L2 prefetches
        const char *v = getenv("QFB_PF_MASK");
        return v ? atoi(v) : 3;
    }();
    sweep_kernel<M, G2><<<grid, T, smem, st>>>(state, rec_dev, rec_bytes, nholes, hi_shifted, tile_order, pf_mask);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

#define QFB_SWEEP_CASE(MM)                                                                          \
    case MM:                                                                                        \
        return g2 ? launch_sweep<MM, true>(state, rec_dev, rec_bytes, nbits, index_hi, st)          \
                  : launch_sweep<MM, false>(state, rec_dev, rec_bytes, nbits, index_hi, st);

static int launch_sweep_dispatch(int M, bool g2, c128 *state, const uint8_t *rec_dev, uint32_t rec_bytes,
                                 int nbits, uint64_t index_hi, cudaStream_t st) {
    switch (M) {
        QFB_SWEEP_CASE(6)
        QFB_SWEEP_CASE(7)
        QFB_SWEEP_CASE(8)
        QFB_SWEEP_CASE(9)
        QFB_SWEEP_CASE(10)
        QFB_SWEEP_CASE(11)
        QFB_SWEEP_CASE(12)
        QFB_SWEEP_CASE(13)
        default: break;
    }
    set_error("sweep: tile_bits=%d unsupported", M);
    return QFB_ERR_UNSUPPORTED;
}

constexpr uint32_t HANDLE_MAGIC = 0x48424651u;

}  // namespace qfb

using namespace qfb;

extern "C" {

int qfb_plan_validate(const void *plan_host, size_t plan_bytes) {
    std::vector<SweepInfo> sweeps;
    int nbits = 0, M = 0;
    return validate_plan((const uint8_t *)plan_host, plan_bytes, sweeps, nbits, M);
}

int qfb_plan_upload(const void *plan_host, size_t plan_bytes, void **handle_out, void *stream) {
    QFB_CHECK_ARG(handle_out, "qfb_plan_upload: null handle_out");
    *handle_out = nullptr;
    PlanHandle *h = new PlanHandle();
    h->magic = HANDLE_MAGIC;
    h->dev = nullptr;
    int rc = validate_plan((const uint8_t *)plan_host, plan_bytes, h->sweeps, h->nbits, h->tile_bits, &h->reg_bits);
    if (rc != QFB_OK) {
        delete h;
        return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaGetDevice(&h->device);
    if (e == cudaSuccess) e = cudaMalloc(&h->dev, plan_bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->dev, plan_host, plan_bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // pageable source
    if (e != cudaSuccess) {
        set_error("qfb_plan_upload: %s", cudaGetErrorString(e));
        if (h->dev) cudaFree(h->dev);
        delete h;
        return QFB_ERR_CUDA;
    }
    h->dev_bytes = plan_bytes;
    if (jit_wanted(h->nbits, h->tile_bits, h->reg_bits)) {
        std::vector<size_t> offsets;
        for (const SweepInfo &s : h->sweeps) offsets.push_back(s.offset);
        rc = jit_build_plan((const uint8_t *)plan_host, offsets, h->nbits, h->tile_bits, h->reg_bits, h->jit);
        if (rc != QFB_OK) {
            for (JitSweep *j : h->jit) jit_destroy(j);
            cudaFree(h->dev);
            delete h;
            return rc;
        }
    }
    h->host.assign((const uint8_t *)plan_host, (const uint8_t *)plan_host + plan_bytes);
    *handle_out = h;
    return QFB_OK;
}

int qfb_plan_sweep_info(void *handle, int sweep, int *nsweeps, uint64_t *nontile_mask, int *specialised) {
    PlanHandle *h = (PlanHandle *)handle;
    QFB_CHECK_ARG(h && h->magic == HANDLE_MAGIC, "qfb_plan_sweep_info: bad handle");
    if (nsweeps) *nsweeps = (int)h->sweeps.size();
    if (specialised) *specialised = h->jit.empty() ? 0 : 1;
    if (nontile_mask) {
        QFB_CHECK_ARG(sweep >= 0 && (size_t)sweep < h->sweeps.size(), "qfb_plan_sweep_info: no sweep %d", sweep);
        qfb_sweep_header sh;
        memcpy(&sh, h->host.data() + h->sweeps[sweep].offset, sizeof(sh));
        uint64_t m = 0;
        for (int i = 0; i < h->nbits - h->tile_bits; ++i) m |= 1ull << sh.hole[i];
        *nontile_mask = m;
    }
    return QFB_OK;
}

int qfb_plan_launch_part(void *handle, void *state, int nbits, uint64_t index_hi, int first_sweep, int nsweeps,
                         uint64_t fix_mask, uint64_t fix_value, int ctas_per_sm, void *stream) {
    PlanHandle *h = (PlanHandle *)handle;
    QFB_CHECK_ARG(h && h->magic == HANDLE_MAGIC, "qfb_plan_launch_part: bad handle");
    QFB_CHECK_ARG(state, "qfb_plan_launch_part: null state");
    QFB_CHECK_ARG(nbits == h->nbits, "qfb_plan_launch_part: plan built for %d bits, state has %d", h->nbits, nbits);
    QFB_CHECK_ARG(first_sweep >= 0 && nsweeps >= 0 && (size_t)first_sweep + nsweeps <= h->sweeps.size(),
                  "qfb_plan_launch_part: sweeps [%d, %d) out of range", first_sweep, first_sweep + nsweeps);
    QFB_CHECK_ARG((fix_value & ~fix_mask) == 0 && (nbits >= 64 || (fix_mask >> nbits) == 0),
                  "qfb_plan_launch_part: fixed bits outside the mask / the state");
    const uint64_t hi_shifted = (nbits >= 64) ? 0ull : (index_hi << nbits);
    for (int i = first_sweep; i < first_sweep + nsweeps; ++i) {
        const SweepInfo &s = h->sweeps[i];
        if (fix_mask == 0) {
            int rc = h->jit.empty() ? launch_sweep_dispatch(h->tile_bits, s.has_g2, (c128 *)state,
                                                            (const uint8_t *)h->dev + s.offset, s.bytes, nbits, index_hi,
                                                            (cudaStream_t)stream)
                                    : jit_launch(h->jit[i], state, hi_shifted, (cudaStream_t)stream);
            if (rc != QFB_OK) return rc;
            continue;
        }
        if (h->jit.empty()) {
            set_error("qfb_plan_launch_part: slices need the sweep-specialised kernels (QFB_JIT, QFB_JIT_MIN_BITS)");
            return QFB_ERR_UNSUPPORTED;
        }
        const auto key = std::make_pair(i, fix_mask);
        auto it = h->variants.find(key);
        if (it == h->variants.end()) {
            JitSweep *v = nullptr;
            int rc = jit_build_variant(h->host.data() + s.offset, h->nbits, h->tile_bits, h->reg_bits, fix_mask, &v);
            if (rc != QFB_OK) return rc;
            it = h->variants.emplace(key, v).first;
        }
        int rc = jit_launch(it->second, state, hi_shifted, (cudaStream_t)stream, fix_value, ctas_per_sm);
        if (rc != QFB_OK) return rc;
    }
    return QFB_OK;
}

int qfb_plan_launch(void *handle, void *state, int nbits, uint64_t index_hi, void *stream) {
    PlanHandle *h = (PlanHandle *)handle;
    QFB_CHECK_ARG(h && h->magic == HANDLE_MAGIC, "qfb_plan_launch: bad handle");
    QFB_CHECK_ARG(state, "qfb_plan_launch: null state");
    QFB_CHECK_ARG(nbits == h->nbits, "qfb_plan_launch: plan built for %d bits, state has %d", h->nbits, nbits);
    if (!h->jit.empty()) {
        const uint64_t hi_shifted = (nbits >= 64) ? 0ull : (index_hi << nbits);
        for (JitSweep *j : h->jit) {
            int rc = jit_launch(j, state, hi_shifted, (cudaStream_t)stream);
            if (rc != QFB_OK) return rc;
        }
        return QFB_OK;
    }
    for (const SweepInfo &s : h->sweeps) {
        int rc = launch_sweep_dispatch(h->tile_bits, s.has_g2, (c128 *)state, (const uint8_t *)h->dev + s.offset, s.bytes,
                                       nbits, index_hi, (cudaStream_t)stream);
        if (rc != QFB_OK) return rc;
    }
    return QFB_OK;
}

int qfb_plan_destroy(void *handle) {
    PlanHandle *h = (PlanHandle *)handle;
    QFB_CHECK_ARG(h && h->magic == HANDLE_MAGIC, "qfb_plan_destroy: bad handle");
    h->magic = 0;
    for (JitSweep *j : h->jit) jit_destroy(j);
    for (auto &kv : h->variants) jit_destroy(kv.second);
    if (h->dev) cudaFree(h->dev);
    delete h;
    return QFB_OK;
}

int qfb_jit_ptx(const void *plan_host, size_t plan_bytes, int sweep, char *buf, size_t cap, size_t *needed,
                size_t *ncoef) {
    std::vector<SweepInfo> sweeps;
    int nbits = 0, M = 0, RB = 0;
    int rc = validate_plan((const uint8_t *)plan_host, plan_bytes, sweeps, nbits, M, &RB);
    if (rc != QFB_OK) return rc;
    QFB_CHECK_ARG(sweep >= 0 && (size_t)sweep < sweeps.size(), "qfb_jit_ptx: no sweep %d", sweep);
    JitSource src;
    std::string err;
    rc = jit_generate((const uint8_t *)plan_host + sweeps[sweep].offset, nbits, M, RB, src, err);
    if (rc != QFB_OK) {
        set_error("%s", err.c_str());
        return rc;
    }
    if (needed) *needed = src.ptx.size() + 1;
    if (ncoef) *ncoef = src.coef.size();
    if (buf && cap) {
        const size_t n = std::min(cap - 1, src.ptx.size());
        memcpy(buf, src.ptx.data(), n);
        buf[n] = 0;
    }
    return QFB_OK;
}

int qfb_jit_source(const void *plan_host, size_t plan_bytes, int sweep, uint64_t fix_mask, char *ptx, size_t ptx_cap,
                   size_t *ptx_needed, double *coef, size_t coef_cap, size_t *ncoef, int *threads,
                   size_t *smem_bytes, int *groups) {
    std::vector<SweepInfo> sweeps;
    int nbits = 0, M = 0, RB = 0;
    int rc = validate_plan((const uint8_t *)plan_host, plan_bytes, sweeps, nbits, M, &RB);
    if (rc != QFB_OK) return rc;
    QFB_CHECK_ARG(sweep >= 0 && (size_t)sweep < sweeps.size(), "qfb_jit_source: no sweep %d", sweep);
    JitSource src;
    std::string err;
    rc = jit_generate((const uint8_t *)plan_host + sweeps[sweep].offset, nbits, M, RB, src, err, fix_mask);
    if (rc != QFB_OK) {
        set_error("%s", err.c_str());
        return rc;
    }
    if (ptx_needed) *ptx_needed = src.ptx.size() + 1;
    if (ncoef) *ncoef = src.coef.size();
    if (threads) *threads = src.threads;
    if (smem_bytes) *smem_bytes = src.smem_bytes;
    if (groups) *groups = src.groups;
    if (ptx && ptx_cap) {
        const size_t n = std::min(ptx_cap - 1, src.ptx.size());
        memcpy(ptx, src.ptx.data(), n);
        ptx[n] = 0;
    }
    if (coef && coef_cap) memcpy(coef, src.coef.data(), sizeof(double) * std::min(coef_cap, src.coef.size()));
    return QFB_OK;
}

int qfb_jit_check(const void *plan_host, size_t plan_bytes, char *log, size_t cap) {
    std::vector<SweepInfo> sweeps;
    int nbits = 0, M = 0, RB = 0;
    int rc = validate_plan((const uint8_t *)plan_host, plan_bytes, sweeps, nbits, M, &RB);
    if (rc != QFB_OK) return rc;
    std::string all;
    for (size_t i = 0; i < sweeps.size(); ++i) {
        JitSource src;
        std::string err, info;
        rc = jit_generate((const uint8_t *)plan_host + sweeps[i].offset, nbits, M, RB, src, err);
        if (rc != QFB_OK) {
            set_error("sweep %zu: %s", i, err.c_str());
            return rc;
        }
        std::vector<char> cubin;
        rc = jit_compile(src.ptx, cubin, info);
        if (rc != QFB_OK) {
            set_error("sweep %zu: %s", i, info.c_str());
            return rc;
        }
        all += "sweep " + std::to_string(i) + ": " + std::to_string(src.ptx.size()) + " bytes of PTX, " +
               std::to_string(src.coef.size()) + " coefficients, image " + std::to_string(cubin.size()) + " bytes\n" + info + "\n";
    }
    if (log && cap) {
        const size_t n = std::min(cap - 1, all.size());
        memcpy(log, all.data(), n);
        log[n] = 0;
    }
    return QFB_OK;
}

int qfb_run_plan(void *state, int nbits, uint64_t index_hi, const void *plan_host, size_t plan_bytes,
                 void *stream) {
    void *h = nullptr;
    int rc = qfb_plan_upload(plan_host, plan_bytes, &h, stream);
    if (rc != QFB_OK) return rc;
    rc = qfb_plan_launch(h, state, nbits, index_hi, stream);
    // the device copy must outlive the queued kernels
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    qfb_plan_destroy(h);
    if (rc == QFB_OK && e != cudaSuccess) {
        set_error("qfb_run_plan: %s", cudaGetErrorString(e));
        return QFB_ERR_CUDA;
    }
    return rc;
}

}  // extern "C"
