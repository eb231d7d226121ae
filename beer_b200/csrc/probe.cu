// Roofline probes: the denominators bench.py reports fractions against are MEASURED on the GPU the bench runs on.
//
//   beer_probe_mma   : back-to-back tcgen05.mma (M = 128, N = 256, one k-step each) on resident shared-memory
//                      operands, one CTA per SM -> the dispatch-limited peak of the tensor pipe for the MMA kind the
//                      statistics / emission kernels use (kind::tf32 or kind::f16); their 3-pass split runs at a
//                      third of it.
//   beer_probe_fill  : write-only streams (float4 stores, or 1-D bulk copies shared -> global) -> the DRAM WRITE
//                      ceiling an llh-producing kernel (KA) can reach.
//   beer_probe_read  : read-only stream (float4 loads folded into one value per thread).
//
// Timing is done by the caller with CUDA events on the stream the probe is launched on.
#include <cuda.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/beer_b200.h"

namespace beer {
namespace probe {

using namespace tcu;

template <int KIND>   // 0 = tf32 (K = 8 per MMA), 1 = f16 (K = 16 per MMA)
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(int n_mma) {
    constexpr int M = 128, N = 256;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    // one k-step = two 16-byte core-matrix columns: A 128 x 32 B, B 256 x 32 B (contents irrelevant: zeros)
    for (int i = threadIdx.x; i < (M + N) * 32 / 16; i += blockDim.x)
        reinterpret_cast<float4*>(smem_raw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t fmt = KIND == 0 ? 2u : 0u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t da = make_desc(smem_u32(smem_raw), 128, 256);
        const uint64_t db = make_desc(smem_u32(smem_raw) + M * 32, 128, 256);
        for (int i = 0; i < n_mma; ++i) {
            const uint32_t d = tmem + (uint32_t)((i >> 2) & 1) * N;     // four accumulating k-steps per tile
            if (KIND == 0) {
                umma_tf32(d, da, db, idesc, (i & 3) != 0);
            } else {
                asm volatile(
                    "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d),
                    "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)((i & 3) != 0))
                    : "memory");
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// tcgen05.mma kind::f16 rate as a function of the tile width N, the number of independent accumulators the k-steps
// rotate over (n_acc = 1: every MMA accumulates onto the previous one), the number of distinct shared-memory operand
// buffers (n_buf = 1: the same 12 KB over and over) and the A operand's home (a_tmem != 0: tensor memory).
template <bool ELECT>
__global__ void __launch_bounds__(128, 1) mma_shape_kernel(int n_mma, int N, int n_acc, int n_buf, int a_tmem) {
    constexpr int M = 128;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int buf_bytes = (M + N) * 32;
    for (int i = threadIdx.x; i < n_buf * buf_bytes / 16; i += blockDim.x)
        reinterpret_cast<float4*>(smem_raw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x < 32 && (ELECT ? elect_one() : threadIdx.x == 0)) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a_col = 496;       // 8 columns of A in tensor memory (contents irrelevant)
        int acc = 0, buf = 0;
        for (int i = 0; i < n_mma; ++i) {
            const uint32_t base = smem_u32(smem_raw) + (uint32_t)(buf * buf_bytes);
            const uint64_t da = make_desc(base, 128, 256), db = make_desc(base + M * 32, 128, 256);
            const uint32_t d = tmem + (uint32_t)(acc * N);
            if (a_tmem) {
                asm volatile(
                    "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d),
                    "r"(tmem + a_col), "l"(db), "r"(idesc), "r"(1u)
                    : "memory");
            } else {
                asm volatile(
                    "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d),
                    "l"(da), "l"(db), "r"(idesc), "r"(1u)
                    : "memory");
            }
            if (++acc == n_acc) acc = 0;
            if (++buf == n_buf) buf = 0;
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

__global__ void __launch_bounds__(256) fill_st_kernel(float4* __restrict__ dst, int64_t n4) {
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = v;
}

constexpr int BULK_BYTES = 32 * 1024;
__global__ void __launch_bounds__(128) fill_bulk_kernel(float* __restrict__ dst, int64_t n_chunks) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    for (int i = threadIdx.x; i < BULK_BYTES / 16; i += blockDim.x)
        reinterpret_cast<float4*>(smem_raw)[i] = make_float4(1.f, 2.f, 3.f, 4.f);
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        int pending = 0;
        for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            bulk_s2g(reinterpret_cast<uint8_t*>(dst) + c * BULK_BYTES, smem_raw, BULK_BYTES);
            bulk_commit();
            if (++pending == 8) {                 // at most 8 bulk stores (256 KB) in flight per SM
                asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
                pending = 4;
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

__global__ void __launch_bounds__(256) read_kernel(const float4* __restrict__ src, int64_t n4, float* __restrict__ sink) {
    float acc = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(src + i);
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456f) *sink = acc;      // never true: keeps the loads alive
}

// Global -> shared copy engine throughput per SM: every CTA (one per SM) streams `copies` chunks of `chunk` bytes from
// a buffer that fits in L2 through a ring of `stages` shared-memory slots; a copy is re-issued as soon as it landed.
//   MODE 0: cp.async.bulk (1-D bulk copy, UBLKCP)     MODE 1: cp.async.bulk.tensor.2d (tensor map, UTMALDG)
template <int MODE>
__global__ void __launch_bounds__(256, 1) tma_read_kernel(const uint8_t* __restrict__ src, int64_t src_bytes, int chunk,
                                                          int stages, int copies, int issuers, int shared_walk,
                                                          const __grid_constant__ CUtensorMap map) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t full[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int w = threadIdx.x >> 5;
    if (w < issuers && elect_one()) {       // (under `lane == 0` each copy cost its issuing thread ~500 cycles)
        // issuer w owns the slots s = w, w + issuers, ...
        const int64_t n_chunks = src_bytes / chunk;
        // issuers > 0: every SM walks its own part of the buffer; the caller passes stages < 0 ... (see `shared_walk`)
        int64_t pos = shared_walk ? ((int64_t)w * 1009) % n_chunks : ((int64_t)blockIdx.x * 37 + (int64_t)w * 1009) % n_chunks;
        const int my_stages = (stages - w + issuers - 1) / issuers, my_copies = copies / issuers;
        for (int i = 0; i < my_copies + my_stages; ++i) {
            const int s = w + (i % my_stages) * issuers;
            if (i >= my_stages) mbar_wait(&full[s], ((i / my_stages) - 1) & 1);
            if (i < my_copies) {
                mbar_arrive_expect_tx(&full[s], (uint32_t)chunk);
                if (MODE == 0) {
                    bulk_g2s(smem_raw + (size_t)s * chunk, src + pos * chunk, (uint32_t)chunk, &full[s]);
                } else {
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                            smem_u32(smem_raw + (size_t)s * chunk)),
                        "l"(&map), "r"(0), "r"((int)(pos * (chunk / 256))), "r"(smem_u32(&full[s]))
                        : "memory");
                }
                pos = (pos + 1) % n_chunks;
            }
        }
    }
}

}  // namespace probe
}  // namespace beer

using namespace beer;

extern "C" {

int beer_probe_mma(int kind, int n_mma, double* flops_out, void* stream) {
    if ((kind != 0 && kind != 1) || n_mma <= 0) return BEER_ERR_ARG;
    const size_t smem = (128 + 256) * 32 + 1024;
    if (kind == 0)
        probe::mma_peak_kernel<0><<<kNumSMs, 128, smem, (cudaStream_t)stream>>>(n_mma);
    else
        probe::mma_peak_kernel<1><<<kNumSMs, 128, smem, (cudaStream_t)stream>>>(n_mma);
    BEER_LAUNCH_CHECK();
    if (flops_out) *flops_out = (double)kNumSMs * (double)n_mma * 2.0 * 128.0 * 256.0 * (kind == 0 ? 8.0 : 16.0);
    return BEER_OK;
}

int beer_probe_mma_shape(int n_mma, int N, int n_acc, int n_buf, int a_tmem, int elect, double* flops_out, void* stream) {
    if (n_mma <= 0 || N < 16 || N > 256 || N % 16 != 0 || n_acc < 1 || n_acc * N > 480 || n_buf < 1 ||
        (size_t)n_buf * (128 + N) * 32 > 200 * 1024)
        return BEER_ERR_ARG;
    const size_t smem = (size_t)n_buf * (128 + N) * 32 + 1024;
    if (elect) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(probe::mma_shape_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        probe::mma_shape_kernel<true><<<kNumSMs, 128, smem, (cudaStream_t)stream>>>(n_mma, N, n_acc, n_buf, a_tmem);
    } else {
        BEER_CUDA_TRY(cudaFuncSetAttribute(probe::mma_shape_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        probe::mma_shape_kernel<false><<<kNumSMs, 128, smem, (cudaStream_t)stream>>>(n_mma, N, n_acc, n_buf, a_tmem);
    }
    BEER_LAUNCH_CHECK();
    if (flops_out) *flops_out = (double)kNumSMs * (double)n_mma * 2.0 * 128.0 * (double)N * 16.0;
    return BEER_OK;
}

int beer_probe_fill(float* dst, int64_t bytes, int mode, void* stream) {
    if (!dst || bytes <= 0 || (((uintptr_t)dst) & 127) != 0) return BEER_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) {
        probe::fill_st_kernel<<<kNumSMs * 8, 256, 0, st>>>(reinterpret_cast<float4*>(dst), bytes / 16);
    } else if (mode == 1) {
        static bool attr = false;
        if (!attr) {
            BEER_CUDA_TRY(cudaFuncSetAttribute(probe::fill_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               probe::BULK_BYTES));
            attr = true;
        }
        probe::fill_bulk_kernel<<<kNumSMs * 4, 128, probe::BULK_BYTES, st>>>(dst, bytes / probe::BULK_BYTES);
    } else {
        return BEER_ERR_ARG;
    }
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_probe_read(const float* src, int64_t bytes, float* sink, void* stream) {
    if (!src || !sink || bytes <= 0 || (((uintptr_t)src) & 15) != 0) return BEER_ERR_ARG;
    probe::read_kernel<<<kNumSMs * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(src), bytes / 16, sink);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

typedef CUresult (*ProbeEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int beer_probe_tma(const float* src, int64_t src_bytes, int mode, int chunk_bytes, int stages, int copies_per_sm,
                   int issuers, int shared_walk, void* stream) {
    if (issuers < 1 || issuers > 8 || issuers > stages) return BEER_ERR_ARG;
    if (!src || (mode != 0 && mode != 1) || chunk_bytes < 256 || chunk_bytes % 256 != 0 || stages < 1 || stages > 16 ||
        (size_t)chunk_bytes * stages > 200 * 1024 || src_bytes < chunk_bytes || chunk_bytes / 256 > 256)
        return BEER_ERR_ARG;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (mode == 1) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        BEER_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (fn == nullptr || q != cudaDriverEntryPointSuccess) return BEER_ERR_UNSUPPORTED;
        const cuuint64_t gdim[2] = {64, (cuuint64_t)(src_bytes / 256)};
        const cuuint64_t gstride[1] = {256};
        const cuuint32_t box[2] = {64, (cuuint32_t)(chunk_bytes / 256)};
        const cuuint32_t estr[2] = {1, 1};
        if (reinterpret_cast<ProbeEncodeTiled>(fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(src), gdim,
                                                   gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return BEER_ERR_UNSUPPORTED;
    }
    const size_t smem = (size_t)chunk_bytes * stages + 1024;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(probe::tma_read_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        probe::tma_read_kernel<0><<<kNumSMs, 256, smem, st>>>((const uint8_t*)src, src_bytes, chunk_bytes, stages, copies_per_sm, issuers, shared_walk, map);
    } else {
        BEER_CUDA_TRY(cudaFuncSetAttribute(probe::tma_read_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        probe::tma_read_kernel<1><<<kNumSMs, 256, smem, st>>>((const uint8_t*)src, src_bytes, chunk_bytes, stages, copies_per_sm, issuers, shared_walk, map);
    }
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

}  // extern "C"
