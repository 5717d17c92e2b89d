// qfb_batch.cu -- batched stochastic trajectories: B = 2^b pure states of n qubits in ONE buffer,
// trajectory index in the top b index bits. Reference: Kraus.run / UnitaryMixture.run, one state at a time
// (quantumflow/channels.py:70-77, 119-125): every Kraus branch K_k psi is computed, its norm gives the branch
// probability, numpy draws the branch, the state is renormalised -- (number of operators + 2) sweeps and a host
// round trip per state. Here a 1-qubit Kraus channel costs the whole batch two passes:
//
//   qfb_batch_rho1     per trajectory the 1-qubit reduced density of the target bit (p0, p1, Re rho01, Im rho01):
//                      w_k |K_k psi|^2 = w_k tr(K_k rho K_k^dagger) follows on the host for every branch at once
//                      (one read of the batch, 16 B per amplitude; deterministic two-pass reduction);
//   qfb_batch_apply1   psi_t <- M_t psi_t with a 2x2 operator PER TRAJECTORY, picked from a device table by the
//                      trajectory bits of the index (the branch the host drew, already divided by the branch
//                      norm): one read + one write of the batch, in place.
//
// Unitary gates need nothing new: the batch is an (n+b)-bit state whose top b bits no gate touches, so the planner
// and the sweep kernels run every trajectory at once. Mixtures of unitaries (Depolarizing, Dephasing) skip the
// first pass (their probabilities are the weights).
#include <algorithm>
#include "qfb_common.cuh"

namespace qfb {

constexpr int RHO_THREADS = 256;
constexpr int RHO_PAIRS_PER_THREAD = 8;

// partial[(traj * nchunks + chunk) * 4 + {0,1,2,3}] = sum over the chunk's pairs of |x|^2, |y|^2, Re(x conj y), Im(x conj y)
__global__ void __launch_bounds__(RHO_THREADS) batch_rho1_pass1(const c128 *__restrict__ state, int nstate, int bit,
                                                                uint32_t nchunks, double *__restrict__ partial) {
    __shared__ double scratch[RHO_THREADS / 32];
    const uint64_t traj = blockIdx.y;
    const uint64_t npairs = 1ull << (nstate - 1);
    const c128 *psi = state + (traj << nstate);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const uint64_t per_chunk = (npairs + nchunks - 1) / nchunks;
    const uint64_t lo = (uint64_t)blockIdx.x * per_chunk, hi = min(npairs, lo + per_chunk);
    for (uint64_t g = lo + threadIdx.x; g < hi; g += RHO_THREADS) {
        const uint64_t i0 = insert_zero(g, bit);
        const c128 x = ldg_stream(psi + i0), y = ldg_stream(psi + (i0 | (1ull << bit)));
        acc[0] += x.re * x.re + x.im * x.im;
        acc[1] += y.re * y.re + y.im * y.im;
        acc[2] += x.re * y.re + x.im * y.im;      // Re(x conj(y)) = rho01
        acc[3] += x.im * y.re - x.re * y.im;      // Im(x conj(y))
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double total = block_sum<RHO_THREADS>(acc[c], scratch);
        if (threadIdx.x == 0) partial[((traj * nchunks) + blockIdx.x) * 4 + c] = total;
    }
}

// fixed-order sum of the chunks of every trajectory: out[traj * 4 + c]
__global__ void batch_rho1_pass2(const double *__restrict__ partial, uint32_t nchunks, uint32_t ntraj,
                                 double *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntraj * 4u) return;
    const uint32_t traj = i >> 2, c = i & 3;
    double s = 0.0;
    for (uint32_t k = 0; k < nchunks; ++k) s += partial[((uint64_t)traj * nchunks + k) * 4 + c];
    out[i] = s;
}

// psi_t <- M_t psi_t; mats[t] = row-major 2x2 complex (8 doubles)
__global__ void __launch_bounds__(256) batch_apply1_kernel(c128 *__restrict__ state, int nstate, int ntotal, int bit,
                                                           const double *__restrict__ mats) {
    const uint64_t npairs = 1ull << (ntotal - 1);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < npairs; g += stride) {
        const uint64_t i0 = insert_zero(g, bit);
        const uint64_t traj = i0 >> nstate;
        const double2 *mt = reinterpret_cast<const double2 *>(mats + traj * 8);
        const double2 a00 = __ldg(mt), a01 = __ldg(mt + 1), a10 = __ldg(mt + 2), a11 = __ldg(mt + 3);
        const c128 x = ldg_stream(state + i0), y = ldg_stream(state + (i0 | (1ull << bit)));
        c128 nx = cmake(0.0, 0.0), ny = cmake(0.0, 0.0);
        cfma(nx, cmake(a00.x, a00.y), x);
        cfma(nx, cmake(a01.x, a01.y), y);
        cfma(ny, cmake(a10.x, a10.y), x);
        cfma(ny, cmake(a11.x, a11.y), y);
        stg_stream(state + i0, nx);
        stg_stream(state + (i0 | (1ull << bit)), ny);
    }
}

// chunks per trajectory of the first reduction pass: enough CTAs to fill the GPU, at least RHO_PAIRS_PER_THREAD pairs
// per thread
static uint32_t rho1_chunks(int nstate, int nbatch_bits) {
    const uint64_t ntraj = 1ull << nbatch_bits, npairs = 1ull << (nstate - 1);
    const uint64_t want = std::max<uint64_t>(1, (uint64_t)sm_count_cached() * 8 / ntraj);
    const uint64_t cap = std::max<uint64_t>(1, npairs / ((uint64_t)RHO_THREADS * RHO_PAIRS_PER_THREAD));
    return (uint32_t)std::min<uint64_t>(std::min(want, cap), 4096);
}

}  // namespace qfb

using namespace qfb;

extern "C" {

int qfb_batch_rho1(const void *state, int nstate, int nbatch_bits, int bit, double *out_dev, void *workspace_dev,
                   size_t workspace_bytes, void *stream) {
    QFB_CHECK_ARG(state && out_dev && workspace_dev, "qfb_batch_rho1: null pointer");
    QFB_CHECK_ARG(nstate >= 1 && nstate <= 40 && nbatch_bits >= 0 && nbatch_bits <= 16 && bit >= 0 && bit < nstate,
                  "qfb_batch_rho1: bad sizes (nstate=%d batch bits=%d bit=%d)", nstate, nbatch_bits, bit);
    const uint32_t ntraj = 1u << nbatch_bits;
    const uint32_t nchunks = rho1_chunks(nstate, nbatch_bits);
    QFB_CHECK_ARG(workspace_bytes >= (size_t)ntraj * nchunks * 4 * sizeof(double),
                  "qfb_batch_rho1: workspace of %zu bytes needed", (size_t)ntraj * nchunks * 4 * sizeof(double));
    cudaStream_t st = (cudaStream_t)stream;
    batch_rho1_pass1<<<dim3(nchunks, ntraj), RHO_THREADS, 0, st>>>((const c128 *)state, nstate, bit, nchunks,
                                                                 (double *)workspace_dev);
    QFB_LAUNCH_CHECK();
    batch_rho1_pass2<<<(ntraj * 4 + 127) / 128, 128, 0, st>>>((const double *)workspace_dev, nchunks, ntraj, out_dev);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

size_t qfb_batch_rho1_workspace(int nstate, int nbatch_bits) {
    if (nstate < 1 || nbatch_bits < 0 || nbatch_bits > 16) return 0;
    return ((size_t)1 << nbatch_bits) * rho1_chunks(nstate, nbatch_bits) * 4 * sizeof(double);
}

int qfb_batch_apply1(void *state, int nstate, int nbatch_bits, int bit, const double *mats_dev, void *stream) {
    QFB_CHECK_ARG(state && mats_dev, "qfb_batch_apply1: null pointer");
    QFB_CHECK_ARG(nstate >= 1 && nbatch_bits >= 0 && nstate + nbatch_bits <= 40 && bit >= 0 && bit < nstate,
                  "qfb_batch_apply1: bad sizes (nstate=%d batch bits=%d bit=%d)", nstate, nbatch_bits, bit);
    QFB_CHECK_ARG(((uintptr_t)mats_dev % 16) == 0, "qfb_batch_apply1: operator table must be 16-byte aligned");
    const int ntotal = nstate + nbatch_bits;
    const uint64_t npairs = 1ull << (ntotal - 1);
    const uint64_t blocks = std::min<uint64_t>((npairs + 255) / 256, (uint64_t)sm_count_cached() * 16);
    batch_apply1_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((c128 *)state, nstate, ntotal, bit, mats_dev);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

}  // extern "C"
