/*
 * CPU ORACLE (C) -- TEST INFRASTRUCTURE ONLY. Flat-vector restatement of the reference's tensormul
 * (quantumflow/backend/numpybk.py:159-214) for sizes where np.einsum is too slow (N = 22..30): for every
 * assignment of the non-target bits gather 2^k amplitudes (gate qubit 0 = MSB of the matrix index), multiply by
 * the row-major matrix, scatter. OpenMP over groups; used by tests and by
 * bench.py's cpu_baseline / --impl reference legs, never by the product.
 * Parity: pinned against oracle/qf_oracle.py::tensormul (the einsum restatement) in tests/test_oracle.py,
 * which is itself pinned against reference-generated fixtures.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define QFO_MAX_K 6

static inline uint64_t insert_zero(uint64_t x, int b) {
    uint64_t lo = x & ((1ull << b) - 1ull);
    return ((x >> b) << (b + 1)) | lo;
}

int qfo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void qfo_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* state: 2^nbits complex128 (interleaved), updated in place. mat: row-major 2^k x 2^k complex128. */
int qfo_apply_dense(double *state, int nbits, const double *mat, int k, const int *bits) {
    if (k < 1 || k > QFO_MAX_K || k > nbits) return 1;
    const int dim = 1 << k;
    int sorted[QFO_MAX_K];
    uint64_t off[1 << QFO_MAX_K];
    for (int j = 0; j < k; ++j) sorted[j] = bits[j];
    for (int i = 0; i < k; ++i)
        for (int j = i + 1; j < k; ++j)
            if (sorted[j] < sorted[i]) { int t = sorted[i]; sorted[i] = sorted[j]; sorted[j] = t; }
    for (int c = 0; c < dim; ++c) {
        uint64_t o = 0;
        for (int j = 0; j < k; ++j)
            if ((c >> (k - 1 - j)) & 1) o |= 1ull << bits[j];
        off[c] = o;
    }
    const int64_t ngroups = (int64_t)1 << (nbits - k);
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < ngroups; ++g) {
        uint64_t base = (uint64_t)g;
        for (int j = 0; j < k; ++j) base = insert_zero(base, sorted[j]);
        double in[2 << QFO_MAX_K];
        for (int c = 0; c < dim; ++c) {
            in[2 * c] = state[2 * (base | off[c])];
            in[2 * c + 1] = state[2 * (base | off[c]) + 1];
        }
        for (int r = 0; r < dim; ++r) {
            double re = 0.0, im = 0.0;
            const double *row = mat + 2 * (size_t)r * dim;
            for (int c = 0; c < dim; ++c) {
                re += row[2 * c] * in[2 * c] - row[2 * c + 1] * in[2 * c + 1];
                im += row[2 * c] * in[2 * c + 1] + row[2 * c + 1] * in[2 * c];
            }
            state[2 * (base | off[r])] = re;
            state[2 * (base | off[r]) + 1] = im;
        }
    }
    return 0;
}

double qfo_norm2(const double *state, int nbits) {
    const int64_t n = (int64_t)1 << nbits;
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < n; ++i) s += state[2 * i] * state[2 * i] + state[2 * i + 1] * state[2 * i + 1];
    return s;
}
