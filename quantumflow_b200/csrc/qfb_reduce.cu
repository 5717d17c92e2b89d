// qfb_reduce.cu -- read-out reductions and the few elementwise sweeps the State/Density API needs.
//
// Reductions are two-pass and deterministic: pass 1 = per-thread serial accumulation over a grid-stride range
// -> warp shuffle tree -> per-block partial in a stream-ordered workspace; pass 2 = one block folds the partials
// in a fixed order. Algorithmic traffic is one read of the operand(s): 16 B (32 B for vdot) per amplitude.
// References: State.norm quantumflow/qubits.py:180-182, bk.inner numpybk.py:125-128, State.probabilities
// states.py:113-119, State.expectation states.py:131-147, Measure.run stdops.py:53-65.
#include <algorithm>
#include <vector>
#include "qfb_common.cuh"

namespace qfb {

constexpr int RT = 256;  // threads per reduction block

static int reduce_blocks(uint64_t n) {
    const uint64_t need = (n + RT * 4 - 1) / (RT * 4);
    const uint64_t cap = (uint64_t)sm_count_cached() * 8;
    return (int)std::max<uint64_t>(1, std::min<uint64_t>(need, cap));
}

// MODE 0: sum |a|^2                -> 1 output
// MODE 1: sum conj(a)*b           -> 2 outputs
// MODE 2: sum diag[i]*|a[i]|^2    -> 1 output
// MODE 3: marginal on bit         -> 2 outputs (bit=0, bit=1)
// MODE 4: trace of [dim,dim]      -> 2 outputs (n = dim, stride dim+1)
template <int MODE>
__global__ void __launch_bounds__(RT) reduce_pass1(const c128 *__restrict__ a, const c128 *__restrict__ b,
                                                   const double *__restrict__ diag, uint64_t n, int bit,
                                                   double *__restrict__ partial) {
    __shared__ double scratch[RT / 32];
    double s0 = 0.0, s1 = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (MODE == 0) {
            const c128 x = ldg128(a + i);
            s0 = fma(x.re, x.re, s0);
            s0 = fma(x.im, x.im, s0);
        } else if (MODE == 1) {
            const c128 x = ldg128(a + i), y = ldg128(b + i);
            s0 = fma(x.re, y.re, s0);
            s0 = fma(x.im, y.im, s0);
            s1 = fma(x.re, y.im, s1);
            s1 = fma(-x.im, y.re, s1);
        } else if (MODE == 2) {
            const c128 x = ldg128(a + i);
            s0 = fma(diag[i], fma(x.re, x.re, x.im * x.im), s0);
        } else if (MODE == 3) {
            const c128 x = ldg128(a + i);
            const double p = fma(x.re, x.re, x.im * x.im);
            if ((i >> bit) & 1ull) s1 += p;
            else s0 += p;
        } else {
            const c128 x = ldg128(a + i * (n + 1));
            s0 += x.re;
            s1 += x.im;
        }
    }
    const double r0 = block_sum<RT>(s0, scratch);
    double r1 = 0.0;
    if (MODE == 1 || MODE == 3 || MODE == 4) r1 = block_sum<RT>(s1, scratch);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = r0;
        partial[2 * blockIdx.x + 1] = r1;
    }
}

__global__ void __launch_bounds__(RT) reduce_pass2(const double *__restrict__ partial, int nblocks, int nout,
                                                   double *__restrict__ out) {
    __shared__ double scratch[RT / 32];
    double s0 = 0.0, s1 = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += RT) {
        s0 += partial[2 * i];
        s1 += partial[2 * i + 1];
    }
    const double r0 = block_sum<RT>(s0, scratch);
    const double r1 = block_sum<RT>(s1, scratch);
    if (threadIdx.x == 0) {
        out[0] = r0;
        if (nout > 1) out[1] = r1;
    }
}

template <int MODE>
static int run_reduce(const void *a, const void *b, const double *diag, uint64_t n, int bit, int nout,
                      double *out_dev, cudaStream_t st) {
    const int blocks = reduce_blocks(n);
    double *partial = nullptr;
    QFB_CUDA(cudaMallocAsync((void **)&partial, sizeof(double) * 2 * blocks, st));
    reduce_pass1<MODE><<<blocks, RT, 0, st>>>((const c128 *)a, (const c128 *)b, diag, n, bit, partial);
    QFB_LAUNCH_CHECK();
    reduce_pass2<<<1, RT, 0, st>>>(partial, blocks, nout, out_dev);
    QFB_LAUNCH_CHECK();
    QFB_CUDA(cudaFreeAsync(partial, st));
    return QFB_OK;
}

// ---- elementwise ----
// OP 0: dst = s*src   OP 1: dst = src * rsqrt(*dev)   OP 2: dst = src / complex(*dev)   OP 3: dst = conj(src)
template <int OP>
__global__ void __launch_bounds__(256) scale_kernel(c128 *__restrict__ dst, const c128 *__restrict__ src,
                                                    uint64_t n, double sre, double sim,
                                                    const double *__restrict__ dev) {
    c128 s = cmake(sre, sim);
    if (OP == 1) s = cmake(1.0 / sqrt(dev[0]), 0.0);
    if (OP == 2) {
        const double dr = dev[0], di = dev[1], den = dr * dr + di * di;
        s = cmake(dr / den, -di / den);
    }
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const c128 x = ldg128(src + i);
        if (OP == 3) stg128(dst + i, cmake(x.re, -x.im));
        else stg128(dst + i, cmul(s, x));
    }
}

__global__ void __launch_bounds__(256) axpby_kernel(c128 *__restrict__ dst, const c128 *__restrict__ a, c128 alpha,
                                                    const c128 *__restrict__ b, c128 beta, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        c128 r = cmul(alpha, ldg128(a + i));
        if (b) cfma(r, beta, ldg128(b + i));
        stg128(dst + i, r);
    }
}

__global__ void __launch_bounds__(256) probs_kernel(const c128 *__restrict__ a, uint64_t n,
                                                    double *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const c128 x = ldg128(a + i);
        // same operation order as numpy: |z| then squared would lose bits; the reference computes
        // absolute(z)**2 -> hypot then square. We keep re^2+im^2 (differs by <= 1 ulp, inside 1e-10).
        out[i] = fma(x.re, x.re, x.im * x.im);
    }
}

__global__ void __launch_bounds__(256) collapse_kernel(c128 *__restrict__ dst, const c128 *__restrict__ src,
                                                       uint64_t n, int bit, int value, double scale) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if ((int)((i >> bit) & 1ull) == value) {
            const c128 x = ldg128(src + i);
            stg128(dst + i, cmake(x.re * scale, x.im * scale));
        } else {
            stg128(dst + i, cmake(0.0, 0.0));
        }
    }
}

__global__ void __launch_bounds__(256) outer_kernel(c128 *__restrict__ dst, const c128 *__restrict__ a,
                                                    uint64_t na, const c128 *__restrict__ b, uint64_t nb,
                                                    int lognb, int conj_b) {
    const uint64_t n = na * nb;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t ia, ib;
        if (lognb >= 0) {
            ia = i >> lognb;
            ib = i & (nb - 1);
        } else {
            ia = i / nb;
            ib = i - ia * nb;
        }
        c128 y = ldg128(b + ib);
        if (conj_b) y.im = -y.im;
        stg128(dst + i, cmul(ldg128(a + ia), y));
    }
}

__global__ void __launch_bounds__(256) density_diag_kernel(const c128 *__restrict__ rho, uint64_t dim,
                                                           c128 *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride)
        stg128(out + i, ldg128(rho + i * (dim + 1)));
}

// ---- partial trace (reference: quantumflow/qubits.py:201-227, an np.einsum with repeated subscripts) ----
// out[j] = sum over s in {0,1}^ntr of in[deposit(j, keep_pos) | spread(s, tmask)]: keep_pos[b] is the input index
// bit of output bit b, tmask[t] the OR of the index bits of traced qubit t in all rank blocks of the tensor
// (ket and bra bit of a density), so only the "all copies equal" elements are summed.
struct PTraceParams {
    int keep_pos[62];
    uint64_t tmask[31];
};

__device__ __forceinline__ uint64_t ptrace_deposit(uint64_t j, int nkeep, const PTraceParams &pp) {
    uint64_t base = 0;
    for (int b = 0; b < nkeep; ++b) base |= ((j >> b) & 1ull) << pp.keep_pos[b];
    return base;
}

// many outputs: one thread per output element walks the traced combinations in Gray-code order (one XOR per
// term; the order is fixed, so the sum is deterministic)
__global__ void __launch_bounds__(256) ptrace_thread_kernel(const c128 *__restrict__ in, c128 *__restrict__ out,
                                                            uint64_t nout, int nkeep, int ntr,
                                                            const __grid_constant__ PTraceParams pp) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t terms = 1ull << ntr;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nout; j += stride) {
        uint64_t idx = ptrace_deposit(j, nkeep, pp);
        c128 acc = ldg128(in + idx);
        for (uint64_t s = 1; s < terms; ++s) {
            idx ^= pp.tmask[__ffsll((long long)s) - 1];
            const c128 v = ldg128(in + idx);
            acc.re += v.re;
            acc.im += v.im;
        }
        stg128(out + j, acc);
    }
}

// few outputs: one CTA per output element, the threads stride over the traced combinations, fixed-order block sum
__global__ void __launch_bounds__(256) ptrace_block_kernel(const c128 *__restrict__ in, c128 *__restrict__ out,
                                                           uint64_t nout, int nkeep, int ntr,
                                                           const __grid_constant__ PTraceParams pp) {
    __shared__ double scratch[8];
    const uint64_t terms = 1ull << ntr;
    for (uint64_t j = blockIdx.x; j < nout; j += gridDim.x) {
        const uint64_t base = ptrace_deposit(j, nkeep, pp);
        double re = 0.0, im = 0.0;
        for (uint64_t s = threadIdx.x; s < terms; s += 256) {
            uint64_t idx = base;
            for (int t = 0; t < ntr; ++t)
                if ((s >> t) & 1ull) idx |= pp.tmask[t];
            const c128 v = ldg128(in + idx);
            re += v.re;
            im += v.im;
        }
        re = block_sum<256>(re, scratch);
        im = block_sum<256>(im, scratch);
        if (threadIdx.x == 0) stg128(out + j, cmake(re, im));
    }
}

struct PermParams {
    c128 *dst;
    const c128 *src;
    uint64_t n;
    int nbits;
    int conj;
    int perm[64];  // dst bit j <- src bit perm[j]
};

// dst-coalesced gather; source index assembled bit by bit (only used for State.permute / QubitVector.H and
// the remap pack step, never inside a circuit sweep)
__global__ void __launch_bounds__(256) permute_kernel(const __grid_constant__ PermParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        uint64_t s = 0;
        for (int j = 0; j < p.nbits; ++j) s |= ((i >> j) & 1ull) << p.perm[j];
        c128 x = ldg128(p.src + s);
        if (p.conj) x.im = -x.im;
        stg128(p.dst + i, x);
    }
}

// ---- sampling: hierarchical CDF search ----
constexpr int SB = 1024;  // amplitudes per chunk
__global__ void __launch_bounds__(256) chunk_sum_kernel(const double *__restrict__ probs, uint64_t n,
                                                        double *__restrict__ sums) {
    __shared__ double scratch[256 / 32];
    const uint64_t c = blockIdx.x;
    const uint64_t lo = c * SB;
    double s = 0.0;
    for (int t = threadIdx.x; t < SB; t += 256) {
        const uint64_t i = lo + t;
        if (i < n) s += probs[i];
    }
    const double r = block_sum<256>(s, scratch);
    if (threadIdx.x == 0) sums[c] = r;
}

static int grid1d(uint64_t n) {
    const uint64_t need = (n + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count_cached() * 8;
    return (int)std::max<uint64_t>(1, std::min<uint64_t>(need, cap));
}

// ---- gate gradient ----
struct GradParams {
    const c128 *g;
    const c128 *psi;
    uint64_t ngroups;
    int k;
    int sorted[3];
    uint64_t off[8];
    double *partial;  // [blocks][2*4^k]
};

template <int K>
__global__ void __launch_bounds__(RT) gate_grad_pass1(const __grid_constant__ GradParams p) {
    constexpr int D = 1 << K;
    __shared__ double scratch[RT / 32];
    double acc[2 * D * D];
#pragma unroll
    for (int i = 0; i < 2 * D * D; ++i) acc[i] = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < p.ngroups; g += stride) {
        uint64_t base = g;
#pragma unroll
        for (int j = 0; j < K; ++j) base = insert_zero(base, p.sorted[j]);
        c128 gv[D], pv[D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
            gv[c] = ldg128(p.g + (base | p.off[c]));
            pv[c] = ldg128(p.psi + (base | p.off[c]));
        }
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
            for (int c = 0; c < D; ++c) {
                // g[r] * conj(psi[c])
                acc[2 * (r * D + c)] += gv[r].re * pv[c].re + gv[r].im * pv[c].im;
                acc[2 * (r * D + c) + 1] += gv[r].im * pv[c].re - gv[r].re * pv[c].im;
            }
    }
#pragma unroll
    for (int i = 0; i < 2 * D * D; ++i) {
        const double r = block_sum<RT>(acc[i], scratch);
        if (threadIdx.x == 0) p.partial[(size_t)blockIdx.x * 2 * D * D + i] = r;
    }
}

__global__ void __launch_bounds__(RT) gate_grad_pass2(const double *__restrict__ partial, int nblocks, int nvals,
                                                      double *__restrict__ out) {
    __shared__ double scratch[RT / 32];
    for (int v = 0; v < nvals; ++v) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nblocks; i += RT) s += partial[(size_t)i * nvals + v];
        const double r = block_sum<RT>(s, scratch);
        if (threadIdx.x == 0) out[v] = r;
    }
}

template <int K>
static int launch_grad(GradParams &p, int blocks, cudaStream_t st) {
    gate_grad_pass1<K><<<blocks, RT, 0, st>>>(p);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

}  // namespace qfb

using namespace qfb;

extern "C" {

int qfb_norm2(const void *a, uint64_t n, double *out_dev, void *stream) {
    QFB_CHECK_ARG(a && out_dev, "qfb_norm2: null pointer");
    return run_reduce<0>(a, nullptr, nullptr, n, 0, 1, out_dev, (cudaStream_t)stream);
}

int qfb_vdot(const void *a, const void *b, uint64_t n, double *out_dev2, void *stream) {
    QFB_CHECK_ARG(a && b && out_dev2, "qfb_vdot: null pointer");
    return run_reduce<1>(a, b, nullptr, n, 0, 2, out_dev2, (cudaStream_t)stream);
}

int qfb_expect_diag(const void *a, const double *diag_dev, uint64_t n, double *out_dev, void *stream) {
    QFB_CHECK_ARG(a && diag_dev && out_dev, "qfb_expect_diag: null pointer");
    return run_reduce<2>(a, nullptr, diag_dev, n, 0, 1, out_dev, (cudaStream_t)stream);
}

int qfb_marginal(const void *a, int nbits, int bit, double *out_dev2, void *stream) {
    QFB_CHECK_ARG(a && out_dev2, "qfb_marginal: null pointer");
    QFB_CHECK_ARG(bit >= 0 && bit < nbits && nbits <= 62, "qfb_marginal: bit %d out of range (nbits=%d)", bit,
                  nbits);
    return run_reduce<3>(a, nullptr, nullptr, 1ull << nbits, bit, 2, out_dev2, (cudaStream_t)stream);
}

int qfb_density_trace(const void *rho, int nq, double *out_dev2, void *stream) {
    QFB_CHECK_ARG(rho && out_dev2, "qfb_density_trace: null pointer");
    QFB_CHECK_ARG(nq >= 0 && nq <= 31, "qfb_density_trace: nq=%d out of range", nq);
    return run_reduce<4>(rho, nullptr, nullptr, 1ull << nq, 0, 2, out_dev2, (cudaStream_t)stream);
}

int qfb_density_diag(const void *rho, int nq, void *out_dev_c128, void *stream) {
    QFB_CHECK_ARG(rho && out_dev_c128, "qfb_density_diag: null pointer");
    QFB_CHECK_ARG(nq >= 0 && nq <= 31, "qfb_density_diag: nq=%d out of range", nq);
    const uint64_t dim = 1ull << nq;
    density_diag_kernel<<<grid1d(dim), 256, 0, (cudaStream_t)stream>>>((const c128 *)rho, dim,
                                                                       (c128 *)out_dev_c128);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

int qfb_partial_trace(void *dst, const void *src, int nbits, int nkeep, const int *keep_pos, int ntr,
                      const uint64_t *trace_masks, void *stream) {
    QFB_CHECK_ARG(dst && src && dst != src, "qfb_partial_trace: null or aliased pointer");
    QFB_CHECK_ARG(nbits >= 1 && nbits <= 62 && nkeep >= 0 && ntr >= 0 && ntr <= 31 && nkeep + ntr <= nbits,
                  "qfb_partial_trace: nbits=%d nkeep=%d ntr=%d out of range", nbits, nkeep, ntr);
    QFB_CHECK_ARG((nkeep == 0 || keep_pos) && (ntr == 0 || trace_masks), "qfb_partial_trace: null bit list");
    PTraceParams pp;
    memset(&pp, 0, sizeof(pp));
    uint64_t seen = 0;
    for (int b = 0; b < nkeep; ++b) {
        QFB_CHECK_ARG(keep_pos[b] >= 0 && keep_pos[b] < nbits && !((seen >> keep_pos[b]) & 1ull),
                      "qfb_partial_trace: bad kept bit %d", keep_pos[b]);
        seen |= 1ull << keep_pos[b];
        pp.keep_pos[b] = keep_pos[b];
    }
    for (int t = 0; t < ntr; ++t) {
        QFB_CHECK_ARG(trace_masks[t] != 0 && (trace_masks[t] >> nbits) == 0 && !(trace_masks[t] & seen),
                      "qfb_partial_trace: bad trace mask %d", t);
        seen |= trace_masks[t];
        pp.tmask[t] = trace_masks[t];
    }
    QFB_CHECK_ARG(seen == ((nbits == 64) ? ~0ull : ((1ull << nbits) - 1ull)),
                  "qfb_partial_trace: kept and traced bits do not cover the %d index bits", nbits);
    const uint64_t nout = 1ull << nkeep;
    if (nout >= 4096) {
        ptrace_thread_kernel<<<grid1d(nout), 256, 0, (cudaStream_t)stream>>>((const c128 *)src, (c128 *)dst, nout,
                                                                             nkeep, ntr, pp);
    } else {
        ptrace_block_kernel<<<(int)nout, 256, 0, (cudaStream_t)stream>>>((const c128 *)src, (c128 *)dst, nout, nkeep,
                                                                         ntr, pp);
    }
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

int qfb_probs(const void *a, uint64_t n, double *out_dev, void *stream) {
    QFB_CHECK_ARG(a && out_dev, "qfb_probs: null pointer");
    probs_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>((const c128 *)a, n, out_dev);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

int qfb_collapse(void *dst, const void *src, int nbits, int bit, int value, double scale, void *stream) {
    QFB_CHECK_ARG(dst && src, "qfb_collapse: null pointer");
    QFB_CHECK_ARG(bit >= 0 && bit < nbits && nbits <= 62, "qfb_collapse: bit %d out of range", bit);
    const uint64_t n = 1ull << nbits;
    collapse_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>((c128 *)dst, (const c128 *)src, n, bit,
                                                                 value ? 1 : 0, scale);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

int qfb_scale(void *dst, const void *src, uint64_t n, double scale_re, double scale_im, void *stream) {
    QFB_CHECK_ARG(dst && src, "qfb_scale: null pointer");
    scale_kernel<0><<<grid1d(n), 256, 0, (cudaStream_t)stream>>>((c128 *)dst, (const c128 *)src, n, scale_re,
                                                                  scale_im, nullptr);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

int qfb_scale_rsqrt_dev(void *dst, const void *src, uint64_t n, const double *norm2_dev, void *stream) {
    QFB_CHECK_ARG(dst && src && norm2_dev, "qfb_scale_rsqrt_dev: null pointer");
    scale_kernel<1><<<grid1d(n), 256, 0, (cudaStream_t)stream>>>((c128 *)dst, (const c128 *)src, n, 0.0, 0.0,
                                                                  norm2_dev);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

int qfb_scale_cdiv_dev(void *dst, const void *src, uint64_t n, const double *cdiv_dev2, void *stream) {
    QFB_CHECK_ARG(dst && src && cdiv_dev2, "qfb_scale_cdiv_dev: null pointer");
    scale_kernel<2><<<grid1d(n), 256, 0, (cudaStream_t)stream>>>((c128 *)dst, (const c128 *)src, n, 0.0, 0.0,
                                                                  cdiv_dev2);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

int qfb_conj(void *dst, const void *src, uint64_t n, void *stream) {
    QFB_CHECK_ARG(dst && src, "qfb_conj: null pointer");
    scale_kernel<3><<<grid1d(n), 256, 0, (cudaStream_t)stream>>>((c128 *)dst, (const c128 *)src, n, 0.0, 0.0,
                                                                  nullptr);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

int qfb_axpby(void *dst, const void *a, double alpha_re, double alpha_im, const void *b, double beta_re,
              double beta_im, uint64_t n, void *stream) {
    QFB_CHECK_ARG(dst && a, "qfb_axpby: null pointer");
    axpby_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>((c128 *)dst, (const c128 *)a,
                                                               c128{alpha_re, alpha_im}, (const c128 *)b,
                                                               c128{beta_re, beta_im}, n);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

int qfb_outer(void *dst, const void *a, uint64_t na, const void *b, uint64_t nb, int conj_b, void *stream) {
    QFB_CHECK_ARG(dst && a && b && na > 0 && nb > 0, "qfb_outer: bad argument");
    int lognb = -1;
    if ((nb & (nb - 1)) == 0) {
        lognb = 0;
        while ((1ull << lognb) < nb) ++lognb;
    }
    outer_kernel<<<grid1d(na * nb), 256, 0, (cudaStream_t)stream>>>((c128 *)dst, (const c128 *)a, na,
                                                                     (const c128 *)b, nb, lognb, conj_b);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

int qfb_permute_bits(void *dst, const void *src, int nbits, const int *perm, int conj, void *stream) {
    QFB_CHECK_ARG(dst && src && (perm || nbits == 0), "qfb_permute_bits: null pointer");
    QFB_CHECK_ARG(dst != src, "qfb_permute_bits: must be out of place");
    QFB_CHECK_ARG(nbits >= 0 && nbits <= 62, "qfb_permute_bits: nbits=%d out of range", nbits);
    PermParams p;
    p.dst = (c128 *)dst;
    p.src = (const c128 *)src;
    p.n = 1ull << nbits;
    p.nbits = nbits;
    p.conj = conj;
    uint64_t seen = 0;
    for (int j = 0; j < nbits; ++j) {
        QFB_CHECK_ARG(perm[j] >= 0 && perm[j] < nbits && !((seen >> perm[j]) & 1ull),
                      "qfb_permute_bits: perm is not a permutation");
        seen |= 1ull << perm[j];
        p.perm[j] = perm[j];
    }
    permute_kernel<<<grid1d(p.n), 256, 0, (cudaStream_t)stream>>>(p);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

int qfb_sample_search(const double *probs_dev, uint64_t n, const double *u_host, int nu, uint64_t *out_idx_host,
                      void *stream) {
    QFB_CHECK_ARG(probs_dev && u_host && out_idx_host && nu >= 0 && n > 0, "qfb_sample_search: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t nchunks = (n + SB - 1) / SB;
    QFB_CHECK_ARG(nchunks <= 0x7fffffffull, "qfb_sample_search: too many chunks");
    double *sums_dev = nullptr;
    QFB_CUDA(cudaMallocAsync((void **)&sums_dev, sizeof(double) * nchunks, st));
    chunk_sum_kernel<<<(unsigned)nchunks, 256, 0, st>>>(probs_dev, n, sums_dev);
    QFB_LAUNCH_CHECK();
    std::vector<double> sums(nchunks);
    QFB_CUDA(cudaMemcpyAsync(sums.data(), sums_dev, sizeof(double) * nchunks, cudaMemcpyDeviceToHost, st));
    QFB_CUDA(cudaStreamSynchronize(st));
    QFB_CUDA(cudaFreeAsync(sums_dev, st));
    // host: sequential cumulative sum over chunk totals, then one chunk fetched per draw
    std::vector<double> cum(nchunks);
    double run = 0.0;
    for (uint64_t c = 0; c < nchunks; ++c) {
        run += sums[c];
        cum[c] = run;
    }
    const double total = run;
    std::vector<double> chunk(SB);
    // draws sorted by chunk: every chunk of probabilities is fetched once, whatever the number of draws in it
    std::vector<std::pair<uint64_t, int>> order(nu);
    for (int j = 0; j < nu; ++j) {
        const double target = u_host[j] * total;
        uint64_t c = std::upper_bound(cum.begin(), cum.end(), target) - cum.begin();
        if (c >= nchunks) c = nchunks - 1;
        order[j] = std::make_pair(c, j);
    }
    std::sort(order.begin(), order.end());
    uint64_t loaded = ~0ull;
    for (int k = 0; k < nu; ++k) {
        const uint64_t c = order[k].first;
        const int j = order[k].second;
        const double target = u_host[j] * total;
        const uint64_t lo = c * SB;
        const uint64_t len = std::min<uint64_t>(SB, n - lo);
        if (c != loaded) {
            QFB_CUDA(cudaMemcpyAsync(chunk.data(), probs_dev + lo, sizeof(double) * len, cudaMemcpyDeviceToHost, st));
            QFB_CUDA(cudaStreamSynchronize(st));
            loaded = c;
        }
        double acc = (c == 0) ? 0.0 : cum[c - 1];
        // The chunk totals come from a tree reduction, this walk is sequential: their roundings differ, so the walk
        // may end without exceeding the target. The fallback is the LAST entry with a non-zero probability (never
        // a zero-probability basis state such as the padded tail of a chunk).
        uint64_t pick = lo, last_nonzero = lo;
        bool found = false, any = false;
        for (uint64_t t = 0; t < len; ++t) {
            if (chunk[t] > 0.0) {
                last_nonzero = lo + t;
                any = true;
            }
            acc += chunk[t];
            if (acc > target) {
                pick = lo + t;
                found = true;
                break;
            }
        }
        if (!found) pick = any ? last_nonzero : lo + len - 1;
        out_idx_host[j] = pick;
    }
    return QFB_OK;
}

int qfb_gate_grad(const void *g, const void *psi, int nbits, int k, const int *bits, void *out_dev,
                  void *stream) {
    QFB_CHECK_ARG(g && psi && bits && out_dev, "qfb_gate_grad: null pointer");
    QFB_CHECK_ARG(k >= 1 && k <= 3 && k <= nbits && nbits <= 62, "qfb_gate_grad: k=%d unsupported", k);
    cudaStream_t st = (cudaStream_t)stream;
    GradParams p;
    p.g = (const c128 *)g;
    p.psi = (const c128 *)psi;
    p.ngroups = 1ull << (nbits - k);
    p.k = k;
    const int dim = 1 << k;
    for (int c = 0; c < dim; ++c) {
        uint64_t o = 0;
        for (int j = 0; j < k; ++j)
            if ((c >> (k - 1 - j)) & 1) o |= 1ull << bits[j];
        p.off[c] = o;
    }
    for (int j = 0; j < k; ++j) {
        QFB_CHECK_ARG(bits[j] >= 0 && bits[j] < nbits, "qfb_gate_grad: bit %d out of range", bits[j]);
        p.sorted[j] = bits[j];
    }
    std::sort(p.sorted, p.sorted + k);
    const int nvals = 2 * dim * dim;
    const uint64_t need = (p.ngroups + RT - 1) / RT;
    const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>(need, (uint64_t)sm_count_cached() * 4));
    QFB_CUDA(cudaMallocAsync((void **)&p.partial, sizeof(double) * (size_t)blocks * nvals, st));
    int rc = QFB_OK;
    if (k == 1) rc = launch_grad<1>(p, blocks, st);
    else if (k == 2) rc = launch_grad<2>(p, blocks, st);
    else rc = launch_grad<3>(p, blocks, st);
    if (rc != QFB_OK) return rc;
    gate_grad_pass2<<<1, RT, 0, st>>>(p.partial, blocks, nvals, (double *)out_dev);
    QFB_LAUNCH_CHECK();
    QFB_CUDA(cudaFreeAsync(p.partial, st));
    return QFB_OK;
}

}  // extern "C"
