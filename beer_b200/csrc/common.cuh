// Shared device helpers for the beer_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>

#define BEER_OK 0
#define BEER_ERR_ARG (-1)       // bad argument combination
#define BEER_ERR_UNSUPPORTED (-2)  // shape outside what the kernels were built for
#define BEER_ERR_ALLOC (-3)

#define BEER_CUDA_TRY(expr)                      \
    do {                                         \
        cudaError_t _e = (expr);                 \
        if (_e != cudaSuccess) return (int)_e;   \
    } while (0)

#define BEER_LAUNCH_CHECK()                          \
    do {                                             \
        cudaError_t _e = cudaGetLastError();         \
        if (_e != cudaSuccess) return (int)_e;       \
    } while (0)

namespace beer {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kNegInf = -INFINITY;
constexpr int kNumSMs = 148;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Warp-wide max in one instruction (CREDUX.MAX.F32, sm_100a).
__device__ __forceinline__ float warp_max(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// digamma for x > 0 in double: upward recurrence to x >= 8, then the
// asymptotic series (abs error < 1e-14 there).
__host__ __device__ inline double digamma_d(double x) {
    double r = 0.0;
    while (x < 8.0) {
        r -= 1.0 / x;
        x += 1.0;
    }
    double f = 1.0 / (x * x);
    double t = f * (-1.0 / 12.0 +
                    f * (1.0 / 120.0 +
                         f * (-1.0 / 252.0 +
                              f * (1.0 / 240.0 + f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
    return r + log(x) - 0.5 / x + t;
}

}  // namespace beer
