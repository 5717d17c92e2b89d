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
#include <vector>
#include "qfb_common.cuh"
#include "qfb_plan.h"

namespace qfb {

constexpr int R = QFB_PLAN_REG_BITS;
constexpr int NE = 1 << R;  // amplitudes per thread

__device__ __forceinline__ uint32_t swz(uint32_t idx) {
    const uint32_t x = idx >> 3;
    return idx ^ ((x ^ (x >> 3) ^ (x >> 6) ^ (x >> 9)) & 7u);
}

template <typename T>
__device__ __forceinline__ T pick(const T (&s)[R], int e) {
    T r = 0;
#pragma unroll
    for (int i = 0; i < R; ++i)
        if ((e >> i) & 1) r |= s[i];
    return r;
}

// XOR-combine (swizzled offsets overlap in their low bits, so OR would be wrong)
__device__ __forceinline__ uint32_t pick_xor(const uint32_t (&s)[R], int e) {
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < R; ++i)
        if ((e >> i) & 1) r ^= s[i];
    return r;
}

// ---------------------------------------------------------------------------------------------------------
// Two-address FP64 primitives. Every update of a register amplitude goes through one of these: the tied
// "+d" operand keeps each amplitude component in ONE virtual register for the whole op loop. Written as plain
// C++ (new SSA value per update) the compiler renames the 2^R amplitudes inside handlers and then re-copies all
// of them at every merge point of the interpreter loop: 58% of all executed instructions were MOVs in the
// first profile (profiles/r1_sweep_v1_summary.txt).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ip_mul(double &x, double s) {            // x = x * s
    asm("mul.f64 %0, %0, %1;" : "+d"(x) : "d"(s));
}
__device__ __forceinline__ void ip_fma_acc(double &acc, double a, double b) {   // acc = a * b + acc
    asm("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc) : "d"(a), "d"(b));
}
__device__ __forceinline__ void ip_fma_scale(double &x, double s, double c) {   // x = x * s + c
    asm("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(s), "d"(c));
}
__device__ __forceinline__ void ip_set(double &x, double v) {             // x = v (keeps x's register)
    asm("mov.f64 %0, %1;" : "+d"(x) : "d"(v));
}
__device__ __forceinline__ void ip_neg(double &x) {                       // x = -x (sign-bit flip, ALU pipe)
    asm("xor.b64 %0, %0, 0x8000000000000000;" : "+d"(x));
}
__device__ __forceinline__ void ip_swap(double &x, double &y) {
    asm("{\n\t.reg .f64 t;\n\tmov.f64 t, %0;\n\tmov.f64 %0, %1;\n\tmov.f64 %1, t;\n\t}" : "+d"(x), "+d"(y));
}
// a *= (fr + i fi)
__device__ __forceinline__ void cmul_inplace(c128 &a, double fr, double fi) {
    const double t0 = -fi * a.im, t1 = fi * a.re;
    ip_fma_scale(a.re, fr, t0);
    ip_fma_scale(a.im, fr, t1);
}
// (x, y) <- (m00 x + m01 y, m10 x + m11 y), all complex; 16 FP64 ops, cross terms first
__device__ __forceinline__ void pair_general(c128 &x, c128 &y, const double *__restrict__ m) {
    const double m00r = m[0], m00i = m[1], m01r = m[2], m01i = m[3], m10r = m[4], m10i = m[5], m11r = m[6],
                 m11i = m[7];
    const double pr = fma(m01r, y.re, -m01i * y.im), pi = fma(m01r, y.im, m01i * y.re);   // m01 y
    const double qr = fma(m10r, x.re, -m10i * x.im), qi = fma(m10r, x.im, m10i * x.re);   // m10 x
    const double rx = fma(m00i, x.re, pi);   // imaginary part of m00 x + m01 y, minus m00r x.im
    const double ry = fma(m11i, y.re, qi);
    ip_fma_scale(x.re, m00r, pr);
    ip_fma_acc(x.re, -m00i, x.im);
    ip_fma_scale(x.im, m00r, rx);
    ip_fma_scale(y.re, m11r, qr);
    ip_fma_acc(y.re, -m11i, y.im);
    ip_fma_scale(y.im, m11r, ry);
}

// ---- 1-bit operator on register bit J; RC: honour the register control mask ----
// p0, p1: the first two payload doubles, prefetched together with the op header
template <int J, int KIND, bool RC>
__device__ __forceinline__ void g1_apply(c128 (&a)[NE], const double *__restrict__ m, double p0, double p1,
                                         uint32_t rc) {
#define QFB_PAIR_LOOP                                                                      \
    _Pragma("unroll") for (int p = 0; p < NE / 2; ++p) {                                   \
        const int e0 = ((p >> J) << (J + 1)) | (p & ((1 << J) - 1)), e1 = e0 | (1 << J);   \
        if (RC && (e0 & rc) != rc) continue;
    // controlled operators only come as SWAPX or GENERAL (pivoting needs an unconditional, uniform scalar)
    if constexpr (KIND == QFB_G1_SUMDIFF) {
        // x' = x + r0 y, y' = x + r1 y with r = +-1. y' is formed from x' (y' = x' + (r1 - r0) y) so that no
        // temporary copy is needed: 4 FP64 per pair, nothing else. Exact zeros under destructive interference
        // are kept: x = -r0 y gives x' = 0 exactly, x = -r1 y gives x' = (r0 - r1) y exactly and y' = 0.
        const double r0 = p0, d = p1 - p0;
        QFB_PAIR_LOOP
            c128 &x = a[e0], &y = a[e1];
            ip_fma_acc(x.re, r0, y.re);
            ip_fma_acc(x.im, r0, y.im);
            ip_fma_scale(y.re, d, x.re);
            ip_fma_scale(y.im, d, x.im);
        }
    } else if constexpr (KIND == QFB_G1_ROT_R) {
        const double r = p0, s = p1;    // x' = x + r y, y' = y + s x
        QFB_PAIR_LOOP
            c128 &x = a[e0], &y = a[e1];
            const double xr = x.re, xi = x.im;
            ip_fma_acc(x.re, r, y.re);
            ip_fma_acc(x.im, r, y.im);
            ip_fma_acc(y.re, s, xr);
            ip_fma_acc(y.im, s, xi);
        }
    } else if constexpr (KIND == QFB_G1_ROT_I) {
        const double ca = p0, cb = p1;  // x' = x + i ca y, y' = y + i cb x
        QFB_PAIR_LOOP
            c128 &x = a[e0], &y = a[e1];
            const double xr = x.re, xi = x.im;
            ip_fma_acc(x.re, -ca, y.im);
            ip_fma_acc(x.im, ca, y.re);
            ip_fma_acc(y.re, -cb, xi);
            ip_fma_acc(y.im, cb, xr);
        }
    } else if constexpr (KIND == QFB_G1_REAL) {
        const double m00 = m[0], m01 = m[2], m10 = m[4], m11 = m[6];
        QFB_PAIR_LOOP
            c128 &x = a[e0], &y = a[e1];
            const double pr = m01 * y.re, pi = m01 * y.im, qr = m10 * x.re, qi = m10 * x.im;
            ip_fma_scale(x.re, m00, pr);
            ip_fma_scale(x.im, m00, pi);
            ip_fma_scale(y.re, m11, qr);
            ip_fma_scale(y.im, m11, qi);
        }
    } else if constexpr (KIND == QFB_G1_RXLIKE) {
        const double d0 = m[0], o01 = m[3], o10 = m[5], d1 = m[6];
        QFB_PAIR_LOOP
            c128 &x = a[e0], &y = a[e1];
            // (d0) x + (i o01) y ; (i o10) x + (d1) y
            const double pr = -o01 * y.im, pi = o01 * y.re, qr = -o10 * x.im, qi = o10 * x.re;
            ip_fma_scale(x.re, d0, pr);
            ip_fma_scale(x.im, d0, pi);
            ip_fma_scale(y.re, d1, qr);
            ip_fma_scale(y.im, d1, qi);
        }
    } else if constexpr (KIND == QFB_G1_SWAPX) {
        QFB_PAIR_LOOP
            ip_swap(a[e0].re, a[e1].re);
            ip_swap(a[e0].im, a[e1].im);
        }
    } else if constexpr (KIND == QFB_G1_ANTIDIAG) {
        const double ar = m[2], ai = m[3], br = m[4], bi = m[5];
        QFB_PAIR_LOOP
            c128 &x = a[e0], &y = a[e1];
            const double qr = fma(br, x.re, -bi * x.im), qi = fma(br, x.im, bi * x.re);   // b x
            const double pr = fma(ar, y.re, -ai * y.im), pi = fma(ar, y.im, ai * y.re);   // a y
            ip_set(x.re, pr);
            ip_set(x.im, pi);
            ip_set(y.re, qr);
            ip_set(y.im, qi);
        }
    } else {
        QFB_PAIR_LOOP
            pair_general(a[e0], a[e1], m);
        }
    }
#undef QFB_PAIR_LOOP
}

// ---- 2-bit operator on register bits J0 > J1 (operator index = bit(J0) << 1 | bit(J1)) ----
template <int J0, int J1>
__device__ __forceinline__ void g2_apply(c128 (&a)[NE], const double *__restrict__ m, uint32_t nz, uint32_t rc) {
    static_assert(J0 > J1, "planner normalises j0 > j1");
    // the two register bits that enumerate the 4 independent groups
    constexpr int O0 = (J1 != 0) ? 0 : ((J0 != 1) ? 1 : 2);
    constexpr int O1 = 6 - J0 - J1 - O0;  // bits sum to 0+1+2+3
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int eb = ((g & 1) << O0) | ((g >> 1) << O1);
        if ((eb & rc) != rc) continue;
        const int id[4] = {eb, eb | (1 << J1), eb | (1 << J0), eb | (1 << J0) | (1 << J1)};
        double outr[4], outi[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            double accr = 0.0, acci = 0.0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if ((nz >> (4 * r + c)) & 1u) {
                    const double mr = m[2 * (4 * r + c)], mi = m[2 * (4 * r + c) + 1];
                    accr = fma(mr, a[id[c]].re, accr);
                    accr = fma(-mi, a[id[c]].im, accr);
                    acci = fma(mr, a[id[c]].im, acci);
                    acci = fma(mi, a[id[c]].re, acci);
                }
            }
            outr[r] = accr;
            outi[r] = acci;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            ip_set(a[id[r]].re, outr[r]);
            ip_set(a[id[r]].im, outi[r]);
        }
    }
}

// ---- controlled phase on the register elements whose index contains MASK ----
template <int MASK>
__device__ __forceinline__ void cph_apply(c128 (&a)[NE], double fr, double fi, bool neg) {
    if (neg) {
#pragma unroll
        for (int e = 0; e < NE; ++e)
            if ((e & MASK) == MASK) {
                ip_neg(a[e].re);
                ip_neg(a[e].im);
            }
    } else {
#pragma unroll
        for (int e = 0; e < NE; ++e)
            if ((e & MASK) == MASK) cmul_inplace(a[e], fr, fi);
    }
}

__device__ __forceinline__ void cph_dispatch(c128 (&a)[NE], uint32_t rc, double fr, double fi, bool neg) {
    uint32_t opaque0;   // see the op interpreter: forces a jump table instead of a compare tree
    asm volatile("mov.u32 %0, 0;" : "=r"(opaque0));
    switch (rc + opaque0) {
        case 1: cph_apply<1>(a, fr, fi, neg); break;
        case 2: cph_apply<2>(a, fr, fi, neg); break;
        case 3: cph_apply<3>(a, fr, fi, neg); break;
        case 4: cph_apply<4>(a, fr, fi, neg); break;
        case 5: cph_apply<5>(a, fr, fi, neg); break;
        case 6: cph_apply<6>(a, fr, fi, neg); break;
        case 7: cph_apply<7>(a, fr, fi, neg); break;
        case 8: cph_apply<8>(a, fr, fi, neg); break;
        case 9: cph_apply<9>(a, fr, fi, neg); break;
        case 10: cph_apply<10>(a, fr, fi, neg); break;
        case 11: cph_apply<11>(a, fr, fi, neg); break;
        case 12: cph_apply<12>(a, fr, fi, neg); break;
        case 13: cph_apply<13>(a, fr, fi, neg); break;
        case 14: cph_apply<14>(a, fr, fi, neg); break;
        default: cph_apply<15>(a, fr, fi, neg); break;
    }
}

template <int M>
struct SweepCfg {
    static constexpr int T = 1 << (M - R);                 // threads per CTA
    static constexpr int MINB = (M >= 13) ? 1 : ((M == 12) ? 2 : ((M == 11) ? 4 : 8));
    static constexpr int TILE_BYTES = 16 << M;
};

// HAS_G2 = false drops the dense 2-bit handlers (40% of the code): circuits made of 1-bit, controlled-1-bit and
// diagonal gates (every workload of BASELINE.json) run the smaller kernel, which is kinder to the instruction cache.
template <int M, bool HAS_G2>
__global__ void __launch_bounds__(SweepCfg<M>::T, SweepCfg<M>::MINB)
sweep_kernel(c128 *__restrict__ state, const uint8_t *__restrict__ rec_g, uint32_t rec_bytes, int nholes,
             uint64_t hi_shifted, int async_load) {
    constexpr int T = SweepCfg<M>::T;
    extern __shared__ __align__(16) uint8_t smem[];
    c128 *tile = reinterpret_cast<c128 *>(smem);
    uint8_t *rec = smem + SweepCfg<M>::TILE_BYTES;
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

    // ---- asynchronous tile loader: the NEXT tile streams into the exchange buffer (cp.async, 16 B per request,
    // L1 bypass) while the last round of the current tile computes and stores. Thread t moves tile-local elements
    // t + c*T, c = 0..NE-1: consecutive threads read consecutive 16-byte words, i.e. whole 128-byte lines.
    uint64_t lin_tg = 0;   // index-bit image of the tile-local bits carried by the thread id
#pragma unroll
    for (int b = 0; b < M - R; ++b) lin_tg |= (uint64_t)((tid >> b) & 1) << sh->gpos[b];
    uint64_t lin_cg[R];    // index-bit image of tile-local bits M-R .. M-1 (the chunk number c)
#pragma unroll
    for (int i = 0; i < R; ++i) lin_cg[i] = 1ull << sh->gpos[M - R + i];
    const uint32_t lin_sw = swz((uint32_t)tid);
    const uint32_t tile_smem = (uint32_t)__cvta_generic_to_shared(tile);
    auto tile_base = [&](uint64_t tile_id) {
        uint64_t gb = 0;
        for (int i = 0; i < nholes; ++i) gb |= ((tile_id >> i) & 1ull) << sh->hole[i];
        return gb;
    };
    auto prefetch_tile = [&](uint64_t gb) {
        const c128 *src = state + (gb | lin_tg);
#pragma unroll
        for (int c = 0; c < NE; ++c) {
            const uint32_t dst = tile_smem + ((lin_sw ^ swz((uint32_t)c << (M - R))) << 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + pick(lin_cg, c)) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // Two loaders (launch parameter): async_load = 1 streams the next tile into the exchange buffer (above);
    // async_load = 0 loads round 0 with LDG straight into registers and only warms L2 with the next tile's lines
    // (no extra shared-memory pass, but round 0 must keep the low bits on the lanes).
    auto prefetch_l2 = [&](uint64_t gb) {
        const c128 *src = state + (gb | lin_tg);
        if ((tid & 7) == 0) {   // one request per 128-byte line
#pragma unroll
            for (int c = 0; c < NE; ++c)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(src + pick(lin_cg, c)));
        }
    };
    if (async_load && (uint64_t)blockIdx.x < ntiles) prefetch_tile(tile_base(blockIdx.x));

    for (uint64_t tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
        const uint64_t gb = tile_base(tile_id);
        if (async_load) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();   // the whole tile has landed in shared memory
        }

        c128 a[NE];
        const uint8_t *rp = rec + sizeof(qfb_sweep_header);
        for (int round = 0; round < nrounds; ++round) {
            const qfb_round_header *rh = reinterpret_cast<const qfb_round_header *>(rp);
            // tile-local index of this thread (register bits zero) and its global image: two table look-ups
            const uint4 l0 = *reinterpret_cast<const uint4 *>(&rh->lut_lo[tid & 15]);
            const uint4 l1 = *reinterpret_cast<const uint4 *>(&rh->lut_hi[(tid >> 4) & 31]);
            const uint32_t tb = l0.x | l1.x;
            const uint64_t tg = (((uint64_t)l0.w << 32) | l0.z) | (((uint64_t)l1.w << 32) | l1.z);
            uint32_t ps[R];  // swizzled image of each register bit
#pragma unroll
            for (int i = 0; i < R; ++i) ps[i] = swz(1u << rh->regpos[i]);
            const uint32_t ptb = swz(tb);  // swz is linear over XOR and tb, register offsets are disjoint

            if (round == 0 && !async_load) {
                uint64_t sg[R];
#pragma unroll
                for (int i = 0; i < R; ++i) sg[i] = 1ull << sh->gpos[rh->regpos[i]];
                const c128 *src = state + (gb | tg);
#pragma unroll
                for (int e = 0; e < NE; ++e) a[e] = ldg_stream(src + pick(sg, e));
            } else {
#pragma unroll
                for (int e = 0; e < NE; ++e) a[e] = tile[ptb ^ pick_xor(ps, e)];
                __syncthreads();  // everyone has read before anyone overwrites the tile again
            }
            if (round + 1 == nrounds && tile_id + gridDim.x < ntiles) {
                if (async_load) prefetch_tile(tile_base(tile_id + gridDim.x));   // buffer is free until the next tile
                else prefetch_l2(tile_base(tile_id + gridDim.x));
            }

            const uint64_t tfull = hi_shifted | gb | tg;
            double phr = 1.0, phi = 0.0;  // running per-thread scalar phase of this round
            // ---- op interpreter: one table jump per op; the next header is in flight while a handler runs ----
            const uint8_t *op = rp + sizeof(qfb_round_header);
            uint4 hw = *reinterpret_cast<const uint4 *>(op);
            double2 pw = *reinterpret_cast<const double2 *>(op + sizeof(qfb_op_header));
            for (;;) {
                // `opaque0` is 0, but the compiler cannot know and must treat the handler id as per-thread data:
                // for warp-uniform values ptxas only builds compare trees (~15 cycles per level of serial
                // branches, profiles/r1_microbench_v5.jsonl: a CZ sign flip cost as much as a Hadamard), for
                // per-thread values it emits jump tables (BRX).
                uint32_t opaque0;
                asm volatile("mov.u32 %0, 0;" : "=r"(opaque0));
                const uint32_t handler = (hw.x & 0xffu) + opaque0;
                if (handler == QFB_H_END) break;
                const uint32_t rc = (hw.x >> 8) & 0xffu;
                const uint64_t cm = ((uint64_t)hw.w << 32) | hw.z;
                const double *m = reinterpret_cast<const double *>(op + sizeof(qfb_op_header));
                op += ((hw.x >> 16) & 0xffu) << 4;
                const uint4 hwn = *reinterpret_cast<const uint4 *>(op);   // every round ends with an END record
                const double2 pwn = *reinterpret_cast<const double2 *>(op + sizeof(qfb_op_header));
#define QFB_G1_CASES(BASE, KIND)                                                               \
    case BASE + 0: g1_apply<0, KIND, false>(a, m, pw.x, pw.y, 0u); break;                      \
    case BASE + 1: g1_apply<1, KIND, false>(a, m, pw.x, pw.y, 0u); break;                      \
    case BASE + 2: g1_apply<2, KIND, false>(a, m, pw.x, pw.y, 0u); break;                      \
    case BASE + 3: g1_apply<3, KIND, false>(a, m, pw.x, pw.y, 0u); break;
#define QFB_G1C_CASES(BASE, KIND)                                                              \
    case BASE + 0: if ((tfull & cm) == cm) g1_apply<0, KIND, true>(a, m, pw.x, pw.y, rc); break;           \
    case BASE + 1: if ((tfull & cm) == cm) g1_apply<1, KIND, true>(a, m, pw.x, pw.y, rc); break;           \
    case BASE + 2: if ((tfull & cm) == cm) g1_apply<2, KIND, true>(a, m, pw.x, pw.y, rc); break;           \
    case BASE + 3: if ((tfull & cm) == cm) g1_apply<3, KIND, true>(a, m, pw.x, pw.y, rc); break;
#define QFB_G2_CASE(IDX, J0, J1)                                                               \
    case QFB_H_G2 + IDX:                                                                       \
        if (HAS_G2 && (tfull & cm) == cm)                                                      \
            g2_apply<J0, J1>(a, m, *reinterpret_cast<const uint32_t *>(m + 32), rc);           \
        break;
                switch (handler) {
                    QFB_G1_CASES(QFB_H_G1_GENERAL, QFB_G1_GENERAL)
                    QFB_G1_CASES(QFB_H_G1_SWAPX, QFB_G1_SWAPX)
                    QFB_G1_CASES(QFB_H_G1_SUMDIFF, QFB_G1_SUMDIFF)
                    QFB_G1_CASES(QFB_H_G1_ROT_R, QFB_G1_ROT_R)
                    QFB_G1_CASES(QFB_H_G1_ROT_I, QFB_G1_ROT_I)
                    QFB_G1C_CASES(QFB_H_G1C_GENERAL, QFB_G1_GENERAL)
                    QFB_G1C_CASES(QFB_H_G1C_SWAPX, QFB_G1_SWAPX)
                    case QFB_H_CPH_SCALAR:
                        if ((tfull & cm) == cm) {
                            const double fr = pw.x, fi = pw.y;
                            const double t0 = fi * phi, t1 = fi * phr;
                            phr = fma(fr, phr, -t0);
                            phi = fma(fr, phi, t1);
                        }
                        break;
                    case QFB_H_CPH_REG:
                        if ((tfull & cm) == cm) cph_dispatch(a, rc, pw.x, pw.y, false);
                        break;
                    case QFB_H_CPH_NEG:
                        if ((tfull & cm) == cm) cph_dispatch(a, rc, 0.0, 0.0, true);
                        break;
                    QFB_G2_CASE(0, 1, 0)
                    QFB_G2_CASE(1, 2, 0)
                    QFB_G2_CASE(2, 2, 1)
                    QFB_G2_CASE(3, 3, 0)
                    QFB_G2_CASE(4, 3, 1)
                    QFB_G2_CASE(5, 3, 2)
                    default: break;
                }
#undef QFB_G1_CASES
#undef QFB_G1C_CASES
#undef QFB_G2_CASE
                // hand the prefetched header over at the very end (volatile: keeps the copy, and with it the
                // wait for the load, below the handler instead of right behind the LDS)
                asm volatile("mov.b32 %0, %4;\n\tmov.b32 %1, %5;\n\tmov.b32 %2, %6;\n\tmov.b32 %3, %7;"
                             : "=r"(hw.x), "=r"(hw.y), "=r"(hw.z), "=r"(hw.w)
                             : "r"(hwn.x), "r"(hwn.y), "r"(hwn.z), "r"(hwn.w));
                asm volatile("mov.f64 %0, %2;\n\tmov.f64 %1, %3;" : "=d"(pw.x), "=d"(pw.y) : "d"(pwn.x), "d"(pwn.y));
            }
            if (rh->has_scalar) {
#pragma unroll
                for (int e = 0; e < NE; ++e) cmul_inplace(a[e], phr, phi);
            }

            if (round + 1 < nrounds) {
#pragma unroll
                for (int e = 0; e < NE; ++e) tile[ptb ^ pick_xor(ps, e)] = a[e];
                __syncthreads();
            } else {
                uint64_t sg[R];
#pragma unroll
                for (int i = 0; i < R; ++i) sg[i] = 1ull << sh->gpos[rh->regpos[i]];
                c128 *dst = state + (gb | tg);
#pragma unroll
                for (int e = 0; e < NE; ++e) stg_stream(dst + pick(sg, e), a[e]);
            }
            rp += rh->bytes;
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
    int nbits, tile_bits, device;
    std::vector<SweepInfo> sweeps;
    void *dev;
    size_t dev_bytes;
};

static int validate_plan(const uint8_t *p, size_t nbytes, std::vector<SweepInfo> &sweeps, int &nbits, int &M) {
    QFB_CHECK_ARG(p && nbytes >= sizeof(qfb_plan_header), "plan: too small");
    qfb_plan_header h;
    memcpy(&h, p, sizeof(h));
    QFB_CHECK_ARG(h.magic == QFB_PLAN_MAGIC && h.version == QFB_PLAN_VERSION, "plan: bad magic/version");
    QFB_CHECK_ARG(h.total_bytes == nbytes, "plan: size mismatch (%llu vs %llu)", (unsigned long long)h.total_bytes,
                  (unsigned long long)nbytes);
    QFB_CHECK_ARG(h.reg_bits == (uint32_t)R, "plan: reg_bits=%u unsupported", h.reg_bits);
    QFB_CHECK_ARG(h.tile_bits >= QFB_PLAN_MIN_TILE_BITS && h.tile_bits <= QFB_PLAN_MAX_TILE_BITS &&
                      h.tile_bits <= h.nbits,
                  "plan: tile_bits=%u unsupported (nbits=%u)", h.tile_bits, h.nbits);
    QFB_CHECK_ARG(h.nbits - h.tile_bits <= QFB_PLAN_MAX_HOLES, "plan: nbits=%u too large", h.nbits);
    nbits = (int)h.nbits;
    M = (int)h.tile_bits;
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
        uint64_t seen = 0;
        for (int j = 0; j < M; ++j) {
            QFB_CHECK_ARG(sh.gpos[j] < h.nbits && !((seen >> sh.gpos[j]) & 1ull), "plan: sweep %u bad gpos", s);
            seen |= 1ull << sh.gpos[j];
        }
        for (uint32_t i = 0; i < h.nbits - h.tile_bits; ++i) {
            QFB_CHECK_ARG(sh.hole[i] < h.nbits && !((seen >> sh.hole[i]) & 1ull), "plan: sweep %u bad hole", s);
            seen |= 1ull << sh.hole[i];
        }
        size_t roff = off + sizeof(sh);
        const size_t send = off + sh.bytes;
        bool sweep_g2 = false;
        for (uint32_t r = 0; r < sh.nrounds; ++r) {
            QFB_CHECK_ARG(roff + sizeof(qfb_round_header) <= send, "plan: sweep %u truncated round %u", s, r);
            qfb_round_header rh;
            memcpy(&rh, p + roff, sizeof(rh));
            QFB_CHECK_ARG(rh.bytes % 16 == 0 && rh.bytes >= sizeof(rh) && roff + rh.bytes <= send,
                          "plan: sweep %u round %u bad size", s, r);
            sweep_g2 = sweep_g2 || rh.has_g2;
            // thread LUTs must reproduce the deposit of the thread bits through thrpos / gpos
            for (int t = 0; t < (1 << (M - R)); ++t) {
                uint32_t tb = 0;
                uint64_t tg = 0;
                for (int b = 0; b < M - R; ++b) {
                    if ((t >> b) & 1) {
                        QFB_CHECK_ARG(rh.thrpos[b] < M, "plan: bad thrpos");
                        tb |= 1u << rh.thrpos[b];
                        tg |= 1ull << sh.gpos[rh.thrpos[b]];
                    }
                }
                const qfb_thread_lut &lo = rh.lut_lo[t & 15], &hi = rh.lut_hi[(t >> 4) & 31];
                QFB_CHECK_ARG((lo.tb | hi.tb) == tb && (lo.tg | hi.tg) == tg, "plan: sweep %u round %u bad thread LUT",
                              s, r);
            }
            uint32_t tseen = 0;
            for (int i = 0; i < R; ++i) {
                QFB_CHECK_ARG(rh.regpos[i] < M && !((tseen >> rh.regpos[i]) & 1u), "plan: bad regpos");
                tseen |= 1u << rh.regpos[i];
            }
            for (int t = 0; t < M - R; ++t) {
                QFB_CHECK_ARG(rh.thrpos[t] < M && !((tseen >> rh.thrpos[t]) & 1u), "plan: bad thrpos");
                tseen |= 1u << rh.thrpos[t];
            }
            size_t ooff = roff + sizeof(rh);
            const size_t rend = roff + rh.bytes;
            bool ended = false;
            for (uint32_t o = 0; o <= rh.nops; ++o) {
                QFB_CHECK_ARG(ooff + sizeof(qfb_op_header) <= rend, "plan: truncated op");
                qfb_op_header oh;
                memcpy(&oh, p + ooff, sizeof(oh));
                const uint32_t bytes = (uint32_t)oh.size16 * 16u;
                QFB_CHECK_ARG(bytes >= sizeof(oh) && ooff + bytes <= rend, "plan: op bad size");
                const int h = oh.handler;
                if (o == rh.nops) {
                    QFB_CHECK_ARG(h == QFB_H_END && bytes == 16, "plan: round does not end with an END record");
                    ended = true;
                } else if (h >= QFB_H_G1_GENERAL && h < QFB_H_G1C_GENERAL) {
                    QFB_CHECK_ARG(bytes == 16 + 64 && oh.reg_cmask == 0 && oh.idx_cmask == 0, "plan: bad G1 op");
                } else if (h >= QFB_H_G1C_GENERAL && h < QFB_H_G1C_SWAPX + 4) {
                    const int j = (h - QFB_H_G1C_GENERAL) & 3;
                    QFB_CHECK_ARG(bytes == 16 + 64 && !((oh.reg_cmask >> j) & 1) && oh.reg_cmask < NE,
                                  "plan: bad controlled G1 op");
                } else if (h == QFB_H_CPH_SCALAR) {
                    QFB_CHECK_ARG(bytes == 32 && oh.reg_cmask == 0 && rh.has_scalar == 1, "plan: bad scalar CPH op");
                } else if (h == QFB_H_CPH_REG || h == QFB_H_CPH_NEG) {
                    QFB_CHECK_ARG(bytes == 32 && oh.reg_cmask > 0 && oh.reg_cmask < NE, "plan: bad CPH op");
                } else if (h >= QFB_H_G2 && h < QFB_H_G2 + 6) {
                    static const int J0[6] = {1, 2, 2, 3, 3, 3}, J1[6] = {0, 0, 1, 0, 1, 2};
                    const int j0 = J0[h - QFB_H_G2], j1 = J1[h - QFB_H_G2];
                    QFB_CHECK_ARG(rh.has_g2 == 1 && bytes == 16 + 272 && !((oh.reg_cmask >> j0) & 1) &&
                                      !((oh.reg_cmask >> j1) & 1) && oh.reg_cmask < NE,
                                  "plan: bad G2 op");
                } else {
                    QFB_CHECK_ARG(false, "plan: unknown handler %d", h);
                }
                ooff += bytes;
            }
            QFB_CHECK_ARG(ended, "plan: missing END record");
            QFB_CHECK_ARG(ooff == rend, "plan: round size mismatch");
            roff += rh.bytes;
        }
        QFB_CHECK_ARG(roff == send, "plan: sweep size mismatch");
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
    // +32: the op loop prefetches one header + 16 payload bytes past the END record of a round
    const size_t smem = (size_t)SweepCfg<M>::TILE_BYTES + rec_bytes + 32;
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
    const int nholes = nbits - M;
    const uint64_t ntiles = 1ull << nholes;
    const uint64_t cap = (uint64_t)sm_count_cached() * resident;
    const int grid = (int)std::min<uint64_t>(ntiles, cap);
    const uint64_t hi_shifted = (nbits >= 64) ? 0ull : (index_hi << nbits);
    static const int async_load = [] {
        const char *v = getenv("QFB_LOADER");
        return (v && strcmp(v, "async") == 0) ? 1 : 0;
    }();
    sweep_kernel<M, G2><<<grid, T, smem, st>>>(state, rec_dev, rec_bytes, nholes, hi_shifted, async_load);
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
        QFB_SWEEP_CASE(5)
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
    int rc = validate_plan((const uint8_t *)plan_host, plan_bytes, h->sweeps, h->nbits, h->tile_bits);
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
    *handle_out = h;
    return QFB_OK;
}

int qfb_plan_launch(void *handle, void *state, int nbits, uint64_t index_hi, void *stream) {
    PlanHandle *h = (PlanHandle *)handle;
    QFB_CHECK_ARG(h && h->magic == HANDLE_MAGIC, "qfb_plan_launch: bad handle");
    QFB_CHECK_ARG(state, "qfb_plan_launch: null state");
    QFB_CHECK_ARG(nbits == h->nbits, "qfb_plan_launch: plan built for %d bits, state has %d", h->nbits, nbits);
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
    if (h->dev) cudaFree(h->dev);
    delete h;
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
