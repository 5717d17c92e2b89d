// qfb_apply.cu -- one-gate-per-sweep kernels: the direct restatement of bk.tensormul
// (reference: quantumflow/backend/numpybk.py:159-214) on a flat complex128 vector in HBM.
//
//   dense_small<K,G>  K = 1..4 target bits, matrix in the kernel parameter block (constant bank, compile-time
//                     offsets after unrolling), one closed group of 2^K amplitudes per thread -> in-place safe.
//                     Algorithmic traffic: 32 B per amplitude (16 B read + 16 B write) = 32 * 2^n per launch.
//   dense_generic     any K <= QFB_MAX_DENSE_K, matrix in global memory, one output amplitude per thread,
//                     out-of-place (in-place callers go through a temporary).
//   diag              out[i] = d[sel(i)] * in[i]
//
// The multi-gate tiled executor (qfb_sweep.cu) is the performance path for circuits; these kernels are what a
// single Gate.run / Channel.evolve call costs and they serve every shape the executor does not handle.
#include <algorithm>
#include <vector>
#include "qfb_common.cuh"

namespace qfb {

template <int K>
struct DenseSmallParams {
    c128 *dst;
    const c128 *src;
    uint64_t ngroups;
    uint64_t ctrl_mask;            // control bits (local index); all must be 1
    uint64_t off[1 << K];          // off[c]: address offset of matrix index c (gate qubit 0 = MSB of c)
    int sorted[K];                 // target bit positions, ascending
    double mat[2 << (2 * K)];      // row-major, interleaved re/im
};

template <int K, int G>
__global__ void __launch_bounds__(256) dense_small_kernel(const __grid_constant__ DenseSmallParams<K> p) {
    constexpr int D = 1 << K;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const bool inplace = (p.dst == p.src);
    for (uint64_t g0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g0 < p.ngroups; g0 += stride * G) {
        uint64_t base[G];
        bool live[G], act[G];
        c128 a[G][D];
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const uint64_t g = g0 + (uint64_t)u * stride;
            live[u] = g < p.ngroups;
            uint64_t b = g;
#pragma unroll
            for (int j = 0; j < K; ++j) b = insert_zero(b, p.sorted[j]);
            base[u] = b;
            act[u] = live[u] && ((b & p.ctrl_mask) == p.ctrl_mask);
            // inactive (control = 0) groups are only read when they must be copied to a distinct dst
            if (live[u] && (act[u] || !inplace)) {
#pragma unroll
                for (int c = 0; c < D; ++c) a[u][c] = ldg128(p.src + (b | p.off[c]));
            }
        }
#pragma unroll
        for (int u = 0; u < G; ++u) {
            if (!live[u]) continue;
            if (act[u]) {
#pragma unroll
                for (int r = 0; r < D; ++r) {
                    c128 acc = cmake(0.0, 0.0);
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        const c128 m = cmake(p.mat[2 * (r * D + c)], p.mat[2 * (r * D + c) + 1]);
                        cfma(acc, m, a[u][c]);
                    }
                    stg128(p.dst + (base[u] | p.off[r]), acc);
                }
            } else if (!inplace) {
#pragma unroll
                for (int c = 0; c < D; ++c) stg128(p.dst + (base[u] | p.off[c]), a[u][c]);
            }
        }
    }
}

struct DenseGenericParams {
    c128 *dst;
    const c128 *src;
    uint64_t n;
    uint64_t target_mask;
    uint64_t ctrl_mask;
    const c128 *mat;       // device, row-major 2^k x 2^k
    const uint64_t *off;   // device, 2^k offsets
    int k;
    int bits[QFB_MAX_DENSE_K];
};

__global__ void __launch_bounds__(256) dense_generic_kernel(const __grid_constant__ DenseGenericParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const int dim = 1 << p.k;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        if ((i & p.ctrl_mask) != p.ctrl_mask) {
            stg128(p.dst + i, ldg128(p.src + i));
            continue;
        }
        int r = 0;
        for (int j = 0; j < p.k; ++j) r = (r << 1) | (int)((i >> p.bits[j]) & 1ull);
        const uint64_t base = i & ~p.target_mask;
        const c128 *row = p.mat + (size_t)r * dim;
        c128 acc = cmake(0.0, 0.0);
        for (int c = 0; c < dim; ++c) cfma(acc, row[c], ldg128(p.src + (base | p.off[c])));
        stg128(p.dst + i, acc);
    }
}

constexpr int DIAG_PARAM_K = 6;
struct DiagParams {
    c128 *dst;
    const c128 *src;
    uint64_t n;
    const c128 *table_dev;  // used when k > DIAG_PARAM_K
    int k;
    int bits[QFB_MAX_DIAG_K];
    double table[2 << DIAG_PARAM_K];
};

__global__ void __launch_bounds__(256) diag_kernel(const __grid_constant__ DiagParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        int s = 0;
        for (int j = 0; j < p.k; ++j) s = (s << 1) | (int)((i >> p.bits[j]) & 1ull);
        c128 d;
        if (p.k <= DIAG_PARAM_K) d = cmake(p.table[2 * s], p.table[2 * s + 1]);
        else d = p.table_dev[s];
        stg128(p.dst + i, cmul(d, ldg128(p.src + i)));
    }
}

static int grid_for(uint64_t work_items, int threads, int per_sm) {
    const uint64_t need = (work_items + threads - 1) / threads;
    const uint64_t cap = (uint64_t)sm_count_cached() * per_sm;
    return (int)std::max<uint64_t>(1, std::min<uint64_t>(need, cap));
}

template <int K, int G>
static int launch_dense_small(c128 *dst, const c128 *src, int nbits, const double *mat, const int *bits,
                              uint64_t ctrl_mask, cudaStream_t st) {
    DenseSmallParams<K> q;
    q.dst = dst;
    q.src = src;
    q.ngroups = 1ull << (nbits - K);
    q.ctrl_mask = ctrl_mask;
    constexpr int D = 1 << K;
    for (int c = 0; c < D; ++c) {
        uint64_t o = 0;
        for (int j = 0; j < K; ++j)
            if ((c >> (K - 1 - j)) & 1) o |= 1ull << bits[j];
        q.off[c] = o;
    }
    for (int j = 0; j < K; ++j) q.sorted[j] = bits[j];
    std::sort(q.sorted, q.sorted + K);
    memcpy(q.mat, mat, sizeof(double) * 2 * D * D);
    const uint64_t items = (q.ngroups + G - 1) / G;
    dense_small_kernel<K, G><<<grid_for(items, 256, 8), 256, 0, st>>>(q);
    QFB_LAUNCH_CHECK();
    return QFB_OK;
}

}  // namespace qfb

using namespace qfb;

extern "C" int qfb_apply_dense(void *dst, const void *src, int nbits, const double *mat_host, int k,
                               const int *bits, int nctrl, const int *ctrl_bits, uint64_t index_hi,
                               void *stream) {
    QFB_CHECK_ARG(dst && src && mat_host && bits, "qfb_apply_dense: null pointer");
    QFB_CHECK_ARG(nbits >= 0 && nbits <= 62, "qfb_apply_dense: nbits=%d out of range", nbits);
    QFB_CHECK_ARG(k >= 0 && k <= QFB_MAX_DENSE_K && k <= nbits, "qfb_apply_dense: k=%d (nbits=%d) unsupported",
                  k, nbits);
    QFB_CHECK_ARG(nctrl >= 0 && nctrl <= QFB_MAX_CTRL, "qfb_apply_dense: nctrl=%d out of range", nctrl);
    uint64_t tmask = 0, cmask = 0;
    for (int j = 0; j < k; ++j) {
        QFB_CHECK_ARG(bits[j] >= 0 && bits[j] < nbits,
                      "qfb_apply_dense: target bit %d not local (nbits=%d); remap first", bits[j], nbits);
        QFB_CHECK_ARG(!((tmask >> bits[j]) & 1ull), "qfb_apply_dense: duplicate target bit %d", bits[j]);
        tmask |= 1ull << bits[j];
    }
    bool hi_ok = true;
    for (int j = 0; j < nctrl; ++j) {
        const int b = ctrl_bits[j];
        QFB_CHECK_ARG(b >= 0 && b < 64 + nbits, "qfb_apply_dense: control bit %d out of range", b);
        if (b >= nbits) {
            if (!((index_hi >> (b - nbits)) & 1ull)) hi_ok = false;
        } else {
            QFB_CHECK_ARG(!((tmask >> b) & 1ull), "qfb_apply_dense: control bit %d is also a target", b);
            cmask |= 1ull << b;
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t n = 1ull << nbits;
    if (!hi_ok) {  // a control held in the rank bits is 0 on this shard: identity
        if (dst != src) QFB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(c128), cudaMemcpyDeviceToDevice, st));
        return QFB_OK;
    }
    c128 *d = (c128 *)dst;
    const c128 *s = (const c128 *)src;
    switch (k) {
        case 0: {
            return qfb_scale(dst, src, n, mat_host[0], mat_host[1], stream);
        }
        case 1: return launch_dense_small<1, 4>(d, s, nbits, mat_host, bits, cmask, st);
        case 2: return launch_dense_small<2, 2>(d, s, nbits, mat_host, bits, cmask, st);
        case 3: return launch_dense_small<3, 1>(d, s, nbits, mat_host, bits, cmask, st);
        case 4: return launch_dense_small<4, 1>(d, s, nbits, mat_host, bits, cmask, st);
        default: break;
    }
    // generic path: matrix and offsets staged in device memory (stream-ordered allocation)
    const int dim = 1 << k;
    std::vector<uint64_t> off(dim);
    for (int c = 0; c < dim; ++c) {
        uint64_t o = 0;
        for (int j = 0; j < k; ++j)
            if ((c >> (k - 1 - j)) & 1) o |= 1ull << bits[j];
        off[c] = o;
    }
    void *mat_dev = nullptr, *off_dev = nullptr, *tmp = nullptr;
    const size_t mat_bytes = sizeof(c128) * (size_t)dim * dim;
    QFB_CUDA(cudaMallocAsync(&mat_dev, mat_bytes, st));
    QFB_CUDA(cudaMallocAsync(&off_dev, sizeof(uint64_t) * dim, st));
    QFB_CUDA(cudaMemcpyAsync(mat_dev, mat_host, mat_bytes, cudaMemcpyHostToDevice, st));
    QFB_CUDA(cudaMemcpyAsync(off_dev, off.data(), sizeof(uint64_t) * dim, cudaMemcpyHostToDevice, st));
    // pageable host sources: make sure the staging copies are finished before the vectors go away
    QFB_CUDA(cudaStreamSynchronize(st));
    DenseGenericParams p;
    p.src = s;
    p.n = n;
    p.target_mask = tmask;
    p.ctrl_mask = cmask;
    p.mat = (const c128 *)mat_dev;
    p.off = (const uint64_t *)off_dev;
    p.k = k;
    for (int j = 0; j < k; ++j) p.bits[j] = bits[j];
    if (dst == src) {
        QFB_CUDA(cudaMallocAsync(&tmp, n * sizeof(c128), st));
        p.dst = (c128 *)tmp;
    } else {
        p.dst = d;
    }
    dense_generic_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(p);
    QFB_LAUNCH_CHECK();
    if (tmp) {
        QFB_CUDA(cudaMemcpyAsync(dst, tmp, n * sizeof(c128), cudaMemcpyDeviceToDevice, st));
        QFB_CUDA(cudaFreeAsync(tmp, st));
    }
    QFB_CUDA(cudaFreeAsync(mat_dev, st));
    QFB_CUDA(cudaFreeAsync(off_dev, st));
    return QFB_OK;
}

extern "C" int qfb_apply_diag(void *dst, const void *src, int nbits, const double *diag_host, int k,
                              const int *bits, uint64_t index_hi, void *stream) {
    QFB_CHECK_ARG(dst && src && diag_host && (bits || k == 0), "qfb_apply_diag: null pointer");
    QFB_CHECK_ARG(nbits >= 0 && nbits <= 62, "qfb_apply_diag: nbits=%d out of range", nbits);
    QFB_CHECK_ARG(k >= 0 && k <= QFB_MAX_DIAG_K, "qfb_apply_diag: k=%d unsupported", k);
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t n = 1ull << nbits;
    // fold bits held by index_hi (rank bits of a sharded state) into a smaller table over the local bits
    int lbits[QFB_MAX_DIAG_K];
    int kl = 0;
    uint64_t seen = 0;
    for (int j = 0; j < k; ++j) {
        QFB_CHECK_ARG(bits[j] >= 0 && bits[j] < nbits + 64, "qfb_apply_diag: bit %d out of range", bits[j]);
        if (bits[j] < nbits) {
            QFB_CHECK_ARG(!((seen >> bits[j]) & 1ull), "qfb_apply_diag: duplicate bit %d", bits[j]);
            seen |= 1ull << bits[j];
            lbits[kl++] = bits[j];
        }
    }
    const int diml = 1 << kl;
    std::vector<double> table(2 * (size_t)diml);
    for (int sl = 0; sl < diml; ++sl) {
        int s = 0, jl = 0;
        for (int j = 0; j < k; ++j) {
            int bit;
            if (bits[j] < nbits) {
                bit = (sl >> (kl - 1 - jl)) & 1;
                ++jl;
            } else {
                bit = (int)((index_hi >> (bits[j] - nbits)) & 1ull);
            }
            s = (s << 1) | bit;
        }
        table[2 * sl] = diag_host[2 * s];
        table[2 * sl + 1] = diag_host[2 * s + 1];
    }
    if (kl == 0) return qfb_scale(dst, src, n, table[0], table[1], stream);
    DiagParams p;
    p.dst = (c128 *)dst;
    p.src = (const c128 *)src;
    p.n = n;
    p.k = kl;
    p.table_dev = nullptr;
    for (int j = 0; j < kl; ++j) p.bits[j] = lbits[j];
    void *tdev = nullptr;
    if (kl <= DIAG_PARAM_K) {
        memcpy(p.table, table.data(), sizeof(double) * 2 * diml);
    } else {
        QFB_CUDA(cudaMallocAsync(&tdev, sizeof(c128) * diml, st));
        QFB_CUDA(cudaMemcpyAsync(tdev, table.data(), sizeof(c128) * diml, cudaMemcpyHostToDevice, st));
        QFB_CUDA(cudaStreamSynchronize(st));
        p.table_dev = (const c128 *)tdev;
    }
    diag_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(p);
    QFB_LAUNCH_CHECK();
    if (tdev) QFB_CUDA(cudaFreeAsync(tdev, st));
    return QFB_OK;
}
