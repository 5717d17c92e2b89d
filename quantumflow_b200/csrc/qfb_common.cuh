// qfb_common.cuh -- shared helpers for libqfb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/qfb200.h"

namespace qfb {

// complex128 as two doubles; 16-byte aligned so every access is one LDG/STG.128
struct __align__(16) c128 {
    double re, im;
};

__device__ __forceinline__ c128 cmake(double r, double i) { c128 z; z.re = r; z.im = i; return z; }
__device__ __forceinline__ c128 cmul(c128 a, c128 b) {
    return cmake(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
// acc += a*b
__device__ __forceinline__ void cfma(c128 &acc, c128 a, c128 b) {
    acc.re = fma(a.re, b.re, acc.re);
    acc.re = fma(-a.im, b.im, acc.re);
    acc.im = fma(a.re, b.im, acc.im);
    acc.im = fma(a.im, b.re, acc.im);
}

__device__ __forceinline__ c128 ldg128(const c128 *p) {
    double2 v = *reinterpret_cast<const double2 *>(p);
    return cmake(v.x, v.y);
}
__device__ __forceinline__ void stg128(c128 *p, c128 v) {
    *reinterpret_cast<double2 *>(p) = make_double2(v.re, v.im);
}
// streaming variants: the state is touched once per sweep, keep it out of L1
__device__ __forceinline__ c128 ldg_stream(const c128 *p) {
    double2 v = __ldcs(reinterpret_cast<const double2 *>(p));
    return cmake(v.x, v.y);
}
__device__ __forceinline__ void stg_stream(c128 *p, c128 v) {
    __stcs(reinterpret_cast<double2 *>(p), make_double2(v.re, v.im));
}

// insert a zero bit at position b (bits >= b move up by one)
__device__ __host__ __forceinline__ uint64_t insert_zero(uint64_t x, int b) {
    uint64_t lo = x & ((1ull << b) - 1ull);
    return ((x >> b) << (b + 1)) | lo;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum (deterministic for a fixed block size); result valid in thread 0
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double *scratch /* THREADS/32 doubles */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        r = (lane < THREADS / 32) ? scratch[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// ---- host side error plumbing ----
void set_error(const char *fmt, ...);
void count_launch(uint64_t n = 1);
int sm_count_cached();

#define QFB_CHECK_ARG(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            qfb::set_error(__VA_ARGS__);    \
            return QFB_ERR_ARG;             \
        }                                   \
    } while (0)

#define QFB_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            qfb::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,              \
                           cudaGetErrorString(e__));                                         \
            return QFB_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

#define QFB_LAUNCH_CHECK()                                                                   \
    do {                                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess) {                                                            \
            qfb::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,          \
                           cudaGetErrorString(e__));                                         \
            return QFB_ERR_CUDA;                                                             \
        }                                                                                    \
        qfb::count_launch();                                                                 \
    } while (0)

}  // namespace qfb
