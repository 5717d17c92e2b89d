// fp64_peaks.cu -- measured FP64 ceilings of the box (SURVEY 8d / VERDICT r1 item 4): DFMA (vector FP64 pipe),
// DMMA (mma.sync.m8n8k4.f64, the FP64 tensor path of sm_100a), both at once (are the two pipes additive?), and
// the two ways of moving a 64-bit register pair (DMUL by 1.0 on the FP64 pipe, SEL/MOV on the integer pipe).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/fp64_peaks tools/fp64_peaks.cu
// Prints one JSON line per case.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;

// mode bit 0: DFMA warps, bit 1: DMMA warps; mode 3: even warps DFMA, odd warps DMMA
__global__ void __launch_bounds__(1024) fp64_kernel(double *out, int mode, double a, double b) {
    const int warp = threadIdx.x >> 5;
    const bool do_fma = (mode == 1) || (mode == 3 && (warp & 1) == 0);
    const bool do_mma = (mode == 2) || (mode == 3 && (warp & 1) == 1);
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    if (do_fma) {
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(a), "d"(b));
        }
    }
    if (do_mma) {
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(acc[2 * i]), "+d"(acc[2 * i + 1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

// 64-bit register moves: mode 0 = DMUL by 1.0 (FP64 pipe), mode 1 = predicated swap through SEL (integer pipe)
__global__ void __launch_bounds__(1024) move_kernel(double *out, int mode, double one, int flag) {
    double x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; y[i] = threadIdx.x - i; }
    const bool p = ((threadIdx.x ^ flag) & 1) != 0;
    for (int it = 0; it < ITERS; ++it) {
        if (mode == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double t;
                asm volatile("mul.f64 %0, %1, %2;" : "=d"(t) : "d"(x[i]), "d"(one));
                asm volatile("mul.f64 %0, %1, %2;" : "=d"(x[i]) : "d"(y[i]), "d"(one));
                asm volatile("mul.f64 %0, %1, %2;" : "=d"(y[i]) : "d"(t), "d"(one));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double t;
                asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.f64 %0, %1, %2, q; }" : "=d"(t) : "d"(y[i]), "d"(x[i]), "r"((int)p));
                asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.f64 %0, %1, %2, q; }" : "=d"(y[i]) : "d"(x[i]), "d"(y[i]), "r"((int)p));
                x[i] = t;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] - y[i];
    if (s == 12345.678) out[0] = s;
}

int main() {
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double *out;
    CHECK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    const char *names[4] = {"", "dfma", "dmma_m8n8k4", "dfma_even_warps+dmma_odd_warps"};
    for (int threads : {256, 512, 1024}) {
        for (int mode = 1; mode <= 3; ++mode) {
            const int blocks = sms * (2048 / threads);
            fp64_kernel<<<blocks, threads>>>(out, mode, 1.0000001, 1e-9);
            CHECK(cudaDeviceSynchronize());
            CHECK(cudaEventRecord(e0));
            fp64_kernel<<<blocks, threads>>>(out, mode, 1.0000001, 1e-9);
            CHECK(cudaEventRecord(e1));
            CHECK(cudaDeviceSynchronize());
            float ms;
            CHECK(cudaEventElapsedTime(&ms, e0, e1));
            const double warps = (double)blocks * threads / 32;
            const double fma_warps = mode == 1 ? warps : mode == 3 ? warps / 2 : 0;
            const double mma_warps = mode == 2 ? warps : mode == 3 ? warps / 2 : 0;
            const double fma_flop = fma_warps * ITERS * 16 * 32 * 2.0;
            const double mma_flop = mma_warps * ITERS * 8 * 512.0;
            printf("{\"case\": \"%s\", \"threads_per_cta\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"dfma_tflops\": %.2f, "
                   "\"dmma_tflops\": %.2f, \"total_tflops\": %.2f, \"dfma_warp_instr_per_clk_per_sm_at_1965MHz\": %.3f, "
                   "\"dmma_instr_per_clk_per_sm_at_1965MHz\": %.3f}\n",
                   names[mode], threads, 2048 / 32, ms, fma_flop / ms / 1e9, mma_flop / ms / 1e9,
                   (fma_flop + mma_flop) / ms / 1e9, fma_warps * ITERS * 16 / (ms * 1e-3) / sms / 1.965e9,
                   mma_warps * ITERS * 8 / (ms * 1e-3) / sms / 1.965e9);
        }
    }
    for (int mode = 0; mode < 2; ++mode) {
        const int threads = 512, blocks = sms * 4;
        move_kernel<<<blocks, threads>>>(out, mode, 1.0, 0);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaEventRecord(e0));
        move_kernel<<<blocks, threads>>>(out, mode, 1.0, 0);
        CHECK(cudaEventRecord(e1));
        CHECK(cudaDeviceSynchronize());
        float ms;
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double swaps = (double)blocks * threads * ITERS * 8;   // 64-bit pair swaps (per lane)
        printf("{\"case\": \"swap64_%s\", \"ms\": %.3f, \"lane_swaps_per_clk_per_sm_at_1965MHz\": %.2f}\n",
               mode == 0 ? "dmul_by_one" : "selp", ms, swaps / (ms * 1e-3) / sms / 1.965e9);
    }
    return 0;
}
