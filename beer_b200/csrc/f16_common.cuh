// fp16-split tcgen05 helpers shared by the kernels that feed the tensor core with 3-pass fp16 hi / lo operands
// (mix16.cu: emission + statistics of mixtures; emission_bwd.cu: gradient of the expected llh w.r.t. the frames).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace beer {
namespace mix16 {

using namespace tcu;

constexpr int TILE = 64;            // frames per image tile
constexpr int HI_EXP = 12;          // scaled statistics: per-dimension maximum in [2^12, 2^13)
constexpr int W_EXP = 13;           // scaled weights: per-Gaussian maximum in [2^13, 2^14)

__host__ __device__ inline int kp_of(int D) { return (2 * D + 15) / 16 * 16; }

// offsets in halfs inside one image half-tile
__device__ __forceinline__ int off1(int f, int k, int KP) { return (f >> 3) * (KP * 8) + (k >> 3) * 64 + (f & 7) * 8 + (k & 7); }
__device__ __forceinline__ int off2(int k, int f) { return (k >> 3) * (TILE * 8) + (f >> 3) * 64 + (k & 7) * 8 + (f & 7); }

__device__ __forceinline__ uint32_t pack_h2(float lo_elem, float hi_elem) {
    __half2 h = __floats2half2_rn(lo_elem, hi_elem);     // .x (low 16 bits) = first argument
    return *reinterpret_cast<uint32_t*>(&h);
}
// fp32 rounded to 11 significant bits (what fp16 keeps of a normal number): two integer instructions
__device__ __forceinline__ float h_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// A operand in tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// balanced reductions (a serial chain of 7 dependent max / add operations per pdf was what the epilogue waited on)
template <int N>
__device__ __forceinline__ float tree_max(const float* v) {
    if constexpr (N == 1) return v[0];
    else return fmaxf(tree_max<N / 2>(v), tree_max<N - N / 2>(v + N / 2));
}
template <int N>
__device__ __forceinline__ float tree_sum(const float* v) {
    if constexpr (N == 1) return v[0];
    else return tree_sum<N / 2>(v) + tree_sum<N - N / 2>(v + N / 2);
}

// ring position + pass parity without divisions
struct Ring {
    int pos, n;
    uint32_t phase;
    __device__ __forceinline__ Ring(int n_) : pos(0), n(n_), phase(0) {}
    __device__ __forceinline__ void next() {
        if (++pos == n) {
            pos = 0;
            phase ^= 1;
        }
    }
};

}  // namespace mix16
}  // namespace beer
