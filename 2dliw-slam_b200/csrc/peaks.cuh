// peaks.cuh — measured fp64 roofline denominators for the two kernels of the LM iteration that are NOT HBM-bound
// (factor_pair_kernel, window_kernel): the achievable DFMA rate of the vector pipe and the DMMA (mma.sync m8n8k4 f64)
// rate of the tensor pipe, on this device, at the clocks it runs at right now.  bench.py calls
// lvio2d_measure_fp64_peak once per run and quotes the fp64 fractions of those kernels against these numbers
// ("the build must add one").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lv {

// 16 independent FMA chains per thread: enough ILP to cover the fp64 latency at any occupancy
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = (double)(threadIdx.x + k) * 1e-3;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = fma(v[k], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += v[k];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the chains alive
}

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 8 independent accumulator tiles per warp
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
    double c[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) c[k] = 0.0;
    const double av = a + threadIdx.x * 1e-9, bv = b;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) dmma_m8n8k4(c[2 * k], c[2 * k + 1], av, bv);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += c[k];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace lv
