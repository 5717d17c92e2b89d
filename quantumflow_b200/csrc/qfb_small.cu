// qfb_small.cu -- whole circuits on small states in ONE launch (the QAOA gradient step of
// examples/qaoa_maxcut.py:37-87 is a 6-qubit circuit of ~100 gates; gate by gate it is ~250 launches and one autograd
// node per gate, i.e. host bound).
//
// A state of up to 13 qubits (2^13 x 16 B = 128 KiB) lives in the shared memory of one CTA for the whole circuit:
//   small_forward   psi <- U_G ... U_1 psi                      (one barrier per gate)
//   small_adjoint   the reverse sweep of the adjoint method for UNITARY gates: starting from the final state and the
//                   incoming gradient lambda = dL/dpsi_final, for g = G .. 1
//                       psi    <- U_g^H psi                      (the state BEFORE gate g: nothing was saved)
//                       grad_g  = sum_groups lambda[r] conj(psi[c])          (same convention as qfb_gate_grad;
//                                                                               laid out like the matrix array)
//                       lambda <- U_g^H lambda
//                   in one pass over the groups per gate; both vectors stay in shared memory (<= 12 qubits).
// The gate list (k <= 2 qubits per gate) and the matrices are device arrays, so a parametrised circuit re-uses the
// descriptor and uploads only the matrices. blockIdx.x = batch item (independent parameter sets / start states).
// Gradient sums are deterministic: fixed group order per thread, warp shuffle tree, warps summed in index order.
#include "qfb_common.cuh"

namespace qfb {

constexpr int SMALL_THREADS = 128;
constexpr int SMALL_WARPS = SMALL_THREADS / 32;

struct SmallGate {
    int k;          // 1 or 2
    int bit0;       // index bit of gate qubit 0 (the MSB of the matrix row index)
    int bit1;       // index bit of gate qubit 1 (k = 2)
    int mat_off;    // offset of the row-major 2^k x 2^k matrix in the matrix array, in complex elements
};

__device__ __forceinline__ c128 cconj(c128 a) { return cmake(a.re, -a.im); }

// element (r, c) of the operator that is applied: the matrix, or its conjugate transpose
template <bool ADJ>
__device__ __forceinline__ c128 mat_at(const c128 *m, int dim, int r, int c) {
    return ADJ ? cconj(m[c * dim + r]) : m[r * dim + c];
}

template <bool ADJ>
__device__ __forceinline__ void apply_gate(c128 *psi, int nbits, const SmallGate &g, const c128 *m) {
    if (g.k == 1) {
        const c128 u00 = mat_at<ADJ>(m, 2, 0, 0), u01 = mat_at<ADJ>(m, 2, 0, 1);
        const c128 u10 = mat_at<ADJ>(m, 2, 1, 0), u11 = mat_at<ADJ>(m, 2, 1, 1);
        const uint32_t groups = 1u << (nbits - 1);
        for (uint32_t t = threadIdx.x; t < groups; t += SMALL_THREADS) {
            const uint32_t i0 = (uint32_t)insert_zero(t, g.bit0), i1 = i0 | (1u << g.bit0);
            const c128 x = psi[i0], y = psi[i1];
            c128 a = cmul(u00, x), b = cmul(u10, x);
            cfma(a, u01, y);
            cfma(b, u11, y);
            psi[i0] = a;
            psi[i1] = b;
        }
    } else {
        c128 u[16];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) u[4 * r + c] = mat_at<ADJ>(m, 4, r, c);
        const int lo = min(g.bit0, g.bit1), hi = max(g.bit0, g.bit1);
        const uint32_t groups = 1u << (nbits - 2);
        for (uint32_t t = threadIdx.x; t < groups; t += SMALL_THREADS) {
            const uint32_t base = (uint32_t)insert_zero(insert_zero(t, lo), hi);
            uint32_t idx[4];
            c128 v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                idx[r] = base | (((r >> 1) & 1u) << g.bit0) | ((r & 1u) << g.bit1);
                v[r] = psi[idx[r]];
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                c128 acc = cmul(u[4 * r], v[0]);
                cfma(acc, u[4 * r + 1], v[1]);
                cfma(acc, u[4 * r + 2], v[2]);
                cfma(acc, u[4 * r + 3], v[3]);
                psi[idx[r]] = acc;
            }
        }
    }
}

__global__ void __launch_bounds__(SMALL_THREADS) small_forward_kernel(c128 *out, const c128 *in, int nbits, int ngates,
                                                                      const SmallGate *gates, const c128 *mats,
                                                                      size_t mats_stride) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    c128 *psi = reinterpret_cast<c128 *>(small_smem);
    const size_t n = (size_t)1 << nbits;
    const c128 *src = in + blockIdx.x * n;
    const c128 *m = mats + blockIdx.x * mats_stride;
    for (size_t i = threadIdx.x; i < n; i += SMALL_THREADS) psi[i] = src[i];
    __syncthreads();
    for (int gi = 0; gi < ngates; ++gi) {
        const SmallGate g = gates[gi];
        apply_gate<false>(psi, nbits, g, m + g.mat_off);
        __syncthreads();
    }
    c128 *dst = out + blockIdx.x * n;
    for (size_t i = threadIdx.x; i < n; i += SMALL_THREADS) dst[i] = psi[i];
}

// warp-level sums of `cnt` doubles held per thread -> partial[warp][0..cnt) (lane 0 writes)
template <int CNT>
__device__ __forceinline__ void warp_partials(const double *acc, double *partial) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < CNT; ++i) {
        const double s = warp_sum(acc[i]);
        if (lane == 0) partial[warp * 32 + i] = s;
    }
}

__global__ void __launch_bounds__(SMALL_THREADS) small_adjoint_kernel(const c128 *psi_final, const c128 *grad_out, int nbits,
                                                                      int ngates, const SmallGate *gates, const c128 *mats,
                                                                      size_t mats_stride, c128 *grad_mats, c128 *grad_in,
                                                                      double *scratch) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    const size_t n = (size_t)1 << nbits;
    c128 *psi = reinterpret_cast<c128 *>(small_smem);
    c128 *lam = psi + n;
    const c128 *m = mats + blockIdx.x * mats_stride;
    double *partial = scratch + (size_t)blockIdx.x * ngates * SMALL_WARPS * 32;
    for (size_t i = threadIdx.x; i < n; i += SMALL_THREADS) {
        psi[i] = psi_final[blockIdx.x * n + i];
        lam[i] = grad_out[blockIdx.x * n + i];
    }
    __syncthreads();
    for (int gi = ngates - 1; gi >= 0; --gi) {
        const SmallGate g = gates[gi];
        const c128 *u = m + g.mat_off;
        double acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.0;
        if (g.k == 1) {
            const c128 a00 = cconj(u[0]), a01 = cconj(u[2]), a10 = cconj(u[1]), a11 = cconj(u[3]);     // U^H
            const uint32_t groups = 1u << (nbits - 1);
            for (uint32_t t = threadIdx.x; t < groups; t += SMALL_THREADS) {
                const uint32_t i0 = (uint32_t)insert_zero(t, g.bit0), i1 = i0 | (1u << g.bit0);
                const c128 x = psi[i0], y = psi[i1], lx = lam[i0], ly = lam[i1];
                c128 p0 = cmul(a00, x), p1 = cmul(a10, x);
                cfma(p0, a01, y);
                cfma(p1, a11, y);
                // grad[r][c] += lambda_out[r] * conj(psi_in[c])
                const c128 l[2] = {lx, ly}, q[2] = {cconj(p0), cconj(p1)};
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const c128 z = cmul(l[r], q[c]);
                        acc[2 * (2 * r + c)] += z.re;
                        acc[2 * (2 * r + c) + 1] += z.im;
                    }
                c128 m0 = cmul(a00, lx), m1 = cmul(a10, lx);
                cfma(m0, a01, ly);
                cfma(m1, a11, ly);
                psi[i0] = p0;
                psi[i1] = p1;
                lam[i0] = m0;
                lam[i1] = m1;
            }
            warp_partials<8>(acc, partial + (size_t)gi * SMALL_WARPS * 32);
        } else {
            c128 a[16];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) a[4 * r + c] = cconj(u[4 * c + r]);
            const int lo = min(g.bit0, g.bit1), hi = max(g.bit0, g.bit1);
            const uint32_t groups = 1u << (nbits - 2);
            for (uint32_t t = threadIdx.x; t < groups; t += SMALL_THREADS) {
                const uint32_t base = (uint32_t)insert_zero(insert_zero(t, lo), hi);
                uint32_t idx[4];
                c128 v[4], l[4], p[4], q[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    idx[r] = base | (((r >> 1) & 1u) << g.bit0) | ((r & 1u) << g.bit1);
                    v[r] = psi[idx[r]];
                    l[r] = lam[idx[r]];
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    p[r] = cmul(a[4 * r], v[0]);
                    cfma(p[r], a[4 * r + 1], v[1]);
                    cfma(p[r], a[4 * r + 2], v[2]);
                    cfma(p[r], a[4 * r + 3], v[3]);
                    q[r] = cmul(a[4 * r], l[0]);
                    cfma(q[r], a[4 * r + 1], l[1]);
                    cfma(q[r], a[4 * r + 2], l[2]);
                    cfma(q[r], a[4 * r + 3], l[3]);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const c128 z = cmul(l[r], cconj(p[c]));
                        acc[2 * (4 * r + c)] += z.re;
                        acc[2 * (4 * r + c) + 1] += z.im;
                    }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    psi[idx[r]] = p[r];
                    lam[idx[r]] = q[r];
                }
            }
            warp_partials<32>(acc, partial + (size_t)gi * SMALL_WARPS * 32);
        }
        __syncthreads();
    }
    // gradient of the input state, and the gate gradients: warps summed in index order
    for (size_t i = threadIdx.x; i < n; i += SMALL_THREADS) grad_in[blockIdx.x * n + i] = lam[i];
    for (int e = threadIdx.x; e < ngates * 16; e += SMALL_THREADS) {
        const int gi = e / 16, j = e % 16;
        const int dim2 = gates[gi].k == 1 ? 4 : 16;
        c128 s = cmake(0.0, 0.0);
        if (j < dim2) {
            const double *p = partial + (size_t)gi * SMALL_WARPS * 32;
            for (int w = 0; w < SMALL_WARPS; ++w) {
                s.re += p[w * 32 + 2 * j];
                s.im += p[w * 32 + 2 * j + 1];
            }
        }
        if (j < dim2) grad_mats[(size_t)blockIdx.x * mats_stride + gates[gi].mat_off + j] = s;
    }
}

static int check_small(int nbits, int batch, int ngates, int max_bits, const char *who) {
    QFB_CHECK_ARG(nbits >= 1 && nbits <= max_bits, "%s: %d qubits (1..%d: the state lives in the shared memory of one CTA)",
                  who, nbits, max_bits);
    QFB_CHECK_ARG(batch >= 1 && batch <= 65535 && ngates >= 0, "%s: batch=%d ngates=%d", who, batch, ngates);
    return QFB_OK;
}

}  // namespace qfb

using namespace qfb;

extern "C" {

int qfb_small_circuit_run(void *out, const void *in, int nbits, int batch, int ngates, const void *gates_dev,
                          const void *mats_dev, size_t mats_stride, void *stream) {
    int rc = check_small(nbits, batch, ngates, 13, "qfb_small_circuit_run");
    if (rc != QFB_OK) return rc;
    QFB_CHECK_ARG(out && in && (ngates == 0 || (gates_dev && mats_dev)), "qfb_small_circuit_run: null pointer");
    const size_t smem = (size_t)16 << nbits;
    QFB_CUDA(cudaFuncSetAttribute(small_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    small_forward_kernel<<<batch, SMALL_THREADS, smem, (cudaStream_t)stream>>>(
        (c128 *)out, (const c128 *)in, nbits, ngates, (const SmallGate *)gates_dev, (const c128 *)mats_dev, mats_stride);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

int qfb_small_circuit_adjoint(const void *psi_final, const void *grad_out, int nbits, int batch, int ngates,
                              const void *gates_dev, const void *mats_dev, size_t mats_stride, void *grad_mats_dev,
                              void *grad_in_dev, void *scratch_dev, void *stream) {
    int rc = check_small(nbits, batch, ngates, 12, "qfb_small_circuit_adjoint");
    if (rc != QFB_OK) return rc;
    QFB_CHECK_ARG(psi_final && grad_out && grad_in_dev && (ngates == 0 || (gates_dev && mats_dev && grad_mats_dev && scratch_dev)),
                  "qfb_small_circuit_adjoint: null pointer");
    const size_t smem = (size_t)32 << nbits;
    QFB_CUDA(cudaFuncSetAttribute(small_adjoint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    small_adjoint_kernel<<<batch, SMALL_THREADS, smem, (cudaStream_t)stream>>>(
        (const c128 *)psi_final, (const c128 *)grad_out, nbits, ngates, (const SmallGate *)gates_dev, (const c128 *)mats_dev,
        mats_stride, (c128 *)grad_mats_dev, (c128 *)grad_in_dev, (double *)scratch_dev);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

size_t qfb_small_circuit_scratch_doubles(int batch, int ngates) {
    return (size_t)batch * (size_t)ngates * SMALL_WARPS * 32;
}

}  // extern "C"
