// Mixture path (C Gaussians per pdf, M = Kp * C large: BASELINE configs[2]) on tcgen05 kind::f16 with a 3-pass
// fp16 split (hi.hi + lo.hi + hi.lo: 22-bit operands, the accuracy of the 3xTF32 kernels at twice the MMA rate),
// built so that the per-Gaussian llhs [N, M] NEVER reach HBM:
//
//   feature images (once per resident chunk)   X -> [x | -x^2/2] scaled per dimension by a power of two, split into
//       fp16 hi / lo, stored as ready-made UMMA operand tiles of 64 frames: img1 = frame-major (K = statistics),
//       img2 = statistic-major (K = frames).  TMA bulk copies land them in shared memory as they are: nobody
//       transposes, splits or converts inside the hot kernels.
//   KA16 (emission16_kernel)   S = img1 . W'^T per (128 frames x NB Gaussians), epilogue: z = S k1 + k2 (log2 domain),
//       log-sum-exp over the C components -> llh2 [N, Kp] (log2 units, offset form).  Nothing per Gaussian is written.
//   KCF (mixstats16_kernel)    the flash-attention structure: per (128 Gaussians x 64 frames)
//         S^T = W' . img1^T           W' RESIDENT IN TENSOR MEMORY (A operand), accumulator lanes = Gaussians
//         w   = 2^(S k1 + k2 + lg2(post) - llh2)   responsibilities x pdf posteriors, per thread = one Gaussian
//         A2  = fp16 hi / lo of w written back IN PLACE over S^T with tcgen05.st
//         acc += A2 . img2^T          second MMA with A from tensor memory, K = frames
//       (mixtureset.py:100-112, normalset.py:121-123).  Costs 2 Q M more flop per frame than reading stored
//       responsibilities, removes 2 x 4 M bytes per frame of HBM traffic (cfg3: 80 of the 121 GB per step).
//
// Both kernels compute S with the SAME operands (same images, same packed weights), the same k-step and pass order,
// so the z of KCF is the z KA16 normalised: the responsibilities sum to one without a second normalisation.
//
// Reference semantics: beer/dists/normalgamma.py:55-59, beer/models/mixtureset.py:85-112, normalset.py:121-123.
#include <cuda.h>
#include <cuda_fp16.h>
#include <type_traits>
#include "common.cuh"
#include "tc_common.cuh"
#include "f16_common.cuh"
#include "../../include/beer_b200.h"

namespace beer {
namespace mix16 {

constexpr int TRACE_SLOTS = 8, TRACE_ITEMS = 256;     // per traced item (chunk / tile): 8 time stamps of CTA 0
static unsigned long long* g_trace = nullptr;        // device buffer [TRACE_ITEMS][TRACE_SLOTS] or null (beer_mix16_set_trace)

__device__ __forceinline__ void trace(unsigned long long* buf, uint32_t item, int slot) {
    if (buf != nullptr && blockIdx.x == 0 && item < TRACE_ITEMS) buf[item * TRACE_SLOTS + slot] = clock64();
}

// ---------------------------------------------------------------------------------------------
// feature images
// ---------------------------------------------------------------------------------------------

// absmax[d] = max_t |x_td| (bits of a non-negative float compare like integers)
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ X, int64_t N, int D, uint32_t* __restrict__ absmax) {
    __shared__ uint32_t s_max[256];
    for (int i = threadIdx.x; i < D; i += blockDim.x) s_max[i] = 0u;
    __syncthreads();
    const int D4 = D >> 2, rpb = 256 / D4;
    const int c4 = threadIdx.x % D4, rl = threadIdx.x / D4;
    float m[4] = {0.f, 0.f, 0.f, 0.f};
    if (rl < rpb) {
        for (int64_t r = (int64_t)blockIdx.x * rpb + rl; r < N; r += (int64_t)gridDim.x * rpb) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(X + (size_t)r * D) + c4);
            m[0] = fmaxf(m[0], fabsf(v.x)); m[1] = fmaxf(m[1], fabsf(v.y));
            m[2] = fmaxf(m[2], fabsf(v.z)); m[3] = fmaxf(m[3], fabsf(v.w));
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) atomicMax(&s_max[4 * c4 + e], __float_as_uint(m[e]));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) atomicMax(&absmax[i], s_max[i]);
}

// alpha[k] = power of two that brings the largest |statistic k| into [2^HI_EXP, 2^(HI_EXP+1)); statistics = [x | -x^2/2]
__global__ void scales_kernel(const uint32_t* __restrict__ absmax, int D, float* __restrict__ alpha) {
    const int k = threadIdx.x;
    if (k >= 2 * D) return;
    const float mx = __uint_as_float(absmax[k < D ? k : k - D]);
    const float v = (k < D) ? mx : 0.5f * mx * mx;
    float a = 1.f;
    if (v > 0.f && isfinite(v)) {
        int e = HI_EXP - ilogbf(v);
        e = max(-100, min(100, e));
        a = ldexpf(1.f, e);
    }
    alpha[k] = a;
}

// One CTA per 64-frame tile: X -> img1 (frame-major) and img2 (statistic-major), fp16 hi | lo halves.
__global__ void __launch_bounds__(256) feat_image_kernel(const float* __restrict__ X, int64_t N, int D,
                                                         const float* __restrict__ alpha, __half* __restrict__ img1,
                                                         __half* __restrict__ img2) {
    extern __shared__ float xs[];      // [TILE][D + 1]
    __shared__ float s_alpha[256];
    const int KP = kp_of(D), D4 = D >> 2, ldx = D + 1;
    const int64_t tile = blockIdx.x, t0 = tile * TILE;
    for (int i = threadIdx.x; i < KP; i += blockDim.x) s_alpha[i] = (i < 2 * D) ? alpha[i] : 0.f;
    for (int e = threadIdx.x; e < TILE * D4; e += blockDim.x) {
        const int r = e / D4, c4 = e - r * D4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t0 + r < N) v = __ldg(reinterpret_cast<const float4*>(X + (size_t)(t0 + r) * D) + c4);
        float* dst = xs + r * ldx + 4 * c4;
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    __syncthreads();
    auto stat = [&](int f, int k) -> float {
        if (k >= 2 * D) return 0.f;
        const float x = xs[f * ldx + (k < D ? k : k - D)];
        return ((k < D) ? x : -0.5f * x * x) * s_alpha[k];
    };
    __half* t1 = img1 + (size_t)tile * (2 * TILE * KP);
    __half* t2 = img2 + (size_t)tile * (2 * TILE * KP);
    const int KG = KP >> 3;
    for (int item = threadIdx.x; item < TILE * KG; item += blockDim.x) {
        const int f = item / KG, k8 = item - f * KG;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float a = stat(f, 8 * k8 + 2 * e), b = stat(f, 8 * k8 + 2 * e + 1);
            const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
            hi[e] = (uint32_t)__half_as_ushort(ha) | ((uint32_t)__half_as_ushort(hb) << 16);
            lo[e] = pack_h2(a - __half2float(ha), b - __half2float(hb));
        }
        const int o = off1(f, 8 * k8, KP);
        *reinterpret_cast<uint4*>(t1 + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(t1 + TILE * KP + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    for (int item = threadIdx.x; item < KP * (TILE / 8); item += blockDim.x) {
        const int k = item >> 3, f8 = item & 7;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float a = stat(8 * f8 + 2 * e, k), b = stat(8 * f8 + 2 * e + 1, k);
            const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
            hi[e] = (uint32_t)__half_as_ushort(ha) | ((uint32_t)__half_as_ushort(hb) << 16);
            lo[e] = pack_h2(a - __half2float(ha), b - __half2float(hb));
        }
        const int o = off2(k, 8 * f8);
        *reinterpret_cast<uint4*>(t2 + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(t2 + TILE * KP + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// frame_ref[t] = sum_d (-x_td^2 / 2) ref[d] + ref[D]: the per-frame constant of the offset form (emission_prepare)
__global__ void __launch_bounds__(256) frame_ref_kernel(const float* __restrict__ X, int64_t N, int D,
                                                        const float* __restrict__ ref, float* __restrict__ out) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= N) return;
    const float4* row = reinterpret_cast<const float4*>(X + (size_t)t * D);
    float acc = 0.f;
    for (int c = 0; c < (D >> 2); ++c) {
        const float4 v = __ldg(row + c);
        acc = fmaf(-0.5f * v.x * v.x, __ldg(ref + 4 * c), acc);
        acc = fmaf(-0.5f * v.y * v.y, __ldg(ref + 4 * c + 1), acc);
        acc = fmaf(-0.5f * v.z * v.z, __ldg(ref + 4 * c + 2), acc);
        acc = fmaf(-0.5f * v.w * v.w, __ldg(ref + 4 * c + 3), acc);
    }
    out[t] = acc + __ldg(ref + D);
}

// ---------------------------------------------------------------------------------------------
// weight pack: one warp per Gaussian (rows M..Mp-1 are padding: zero weights, z = -inf)
//   wimg   [n_chunks][hi | lo][NB x KP] halfs, core-matrix layout    (B operand of KA16)
//   wtm    [Mp][KP/2 hi words | KP/2 lo words]                       (A operand of KCF, copied to tensor memory)
//   k12    [Mp] = (log2(e) / beta_j, bias_j log2(e))                  (z = S k1 + k2, log2 domain)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pack16_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                     const float* __restrict__ alpha, int M, int Mp, int D, int NB,
                                                     __half* __restrict__ wimg, uint32_t* __restrict__ wtm,
                                                     float2* __restrict__ k12) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= Mp) return;
    const int KP = kp_of(D), NPAIR = KP >> 1;
    float v[3][2];
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int p = lane + 32 * i;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int k = 2 * p + e;
            float w = 0.f;
            if (j < M && p < NPAIR && k < 2 * D) w = W[(size_t)j * 2 * D + k] / alpha[k];     // exact: alpha = 2^e
            v[i][e] = w;
            mx = fmaxf(mx, fabsf(w));
        }
    }
    mx = warp_max(mx);
    int be = 0;
    if (mx > 0.f && isfinite(mx)) be = max(-100, min(100, W_EXP - ilogbf(mx)));
    const float beta = ldexpf(1.f, be);
    const int chunk = j / NB, n = j - chunk * NB;
    __half* img = wimg + (size_t)chunk * (2 * NB * KP);
    uint32_t* tm = wtm + (size_t)j * KP;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int p = lane + 32 * i;
        if (p >= NPAIR) continue;
        const float a = v[i][0] * beta, b = v[i][1] * beta;
        const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
        const uint32_t hi = (uint32_t)__half_as_ushort(ha) | ((uint32_t)__half_as_ushort(hb) << 16);
        const uint32_t lo = pack_h2(a - __half2float(ha), b - __half2float(hb));
        const int o = off1(n, 2 * p, KP);
        *reinterpret_cast<uint32_t*>(img + o) = hi;
        *reinterpret_cast<uint32_t*>(img + NB * KP + o) = lo;
        tm[p] = hi;
        tm[NPAIR + p] = lo;
    }
    if (lane == 0) k12[j] = (j < M) ? make_float2(kLog2e * ldexpf(1.f, -be), bias[j] * kLog2e) : make_float2(0.f, kNegInf);
}

// ---------------------------------------------------------------------------------------------
// KA16: llh2 [N, Kp] = log2 sum_c 2^(z_tkc)
// ---------------------------------------------------------------------------------------------
constexpr int KA_NCQ = 3;               // epilogue warps per (lane quarter, group): each takes 1 / KA_NCQ of the chunk's units
constexpr int KA_WORKERS = 4 * 2 * KA_NCQ * 32, KA_MMA_WARP = KA_WORKERS / 32, KA_LOAD_WARP = KA_MMA_WARP + 1;
constexpr int KA_THREADS = KA_WORKERS + 64;
constexpr int KA_STAGES = 3;            // weight-chunk ring
constexpr int K12_RING = 8;             // the loader runs up to KA_STAGES + 2 chunks ahead of the epilogue that reads (k1, k2)

struct KaArgs {
    const __half* img1;
    int64_t N;
    const __half* wimg;
    const float2* k12;       // [Mp] (k1, k2)
    int Kp, NB, n_chunks;
    float* llh2;
    int64_t ld;
    int staged;              // llh tiles leave through shared memory + TMA tensor stores (per lane quarter)
    unsigned long long* trace;
};

struct KaBarriers {
    uint64_t a_full[2], a_empty[2];
    uint64_t b_full[KA_STAGES], b_empty[KA_STAGES];
    uint64_t t_full[2], t_empty[2];
    uint32_t tmem_base;
    uint32_t pad[3];
};

// The statistics tile of 128 frames is the A operand and lives in TENSOR MEMORY (lane = frame, 40 columns of fp16 pairs
// for hi and 40 for lo): with both operands in shared memory a 128 x 160 x 16 MMA took ~190 cycles against an 80-cycle
// floor (9 KB of operand fetch per MMA); from tensor memory only the 5 KB weight slice is fetched.
template <int KP, int C>
__global__ void __launch_bounds__(KA_THREADS, 1) emission16_kernel(KaArgs a, const __grid_constant__ CUtensorMap omap) {
    constexpr int FR = 128, KSTEPS = KP / 16;
    constexpr uint32_t LBO = 128, SBO = KP * 16;
    constexpr int UNIT = C < 8 ? 8 : C;                 // columns per epilogue unit (whole pdfs, one tcgen05.ld)
    constexpr int PPU = UNIT / C;                       // pdfs per unit
    constexpr uint32_t COL_A0 = 160, COL_A1 = 416;      // accumulators at columns 0 and 256 (<= 160 wide), A tiles behind them
    static_assert(KP <= 96, "tensor memory");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __half* Bs = reinterpret_cast<__half*>(smem_raw);                        // [KA_STAGES][hi | lo] of NB x KP
    const int b_stage = 2 * a.NB * KP;
    const int n_stages = a.n_chunks > 1 ? KA_STAGES : 1, n_k12 = a.n_chunks > 1 ? K12_RING : 1;
    float2* s_k12 = reinterpret_cast<float2*>(Bs + n_stages * b_stage);      // [K12_RING][NB] (k1, k2)
    KaBarriers* bars = reinterpret_cast<KaBarriers*>(s_k12 + n_k12 * a.NB);
    // staged output: [4 lane quarters][2 epilogue groups][2 buffers][32 frames][NB / C pdfs] (128-byte aligned)
    float* s_out = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(bars + 1) + 127) & ~uintptr_t(127));
    const int PC = a.NB / C;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = (a.N + FR - 1) / FR, n_tiles64 = (a.N + TILE - 1) / TILE;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->a_full[i], 128);
            mbar_init(&bars->a_empty[i], 1);
            mbar_init(&bars->t_full[i], 1);
            mbar_init(&bars->t_empty[i], KA_WORKERS / 2);       // one epilogue group per accumulator buffer
        }
        for (int i = 0; i < KA_STAGES; ++i) {
            mbar_init(&bars->b_full[i], 1);
            mbar_init(&bars->b_empty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == KA_MMA_WARP) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == KA_LOAD_WARP) {
        // weight chunks + their (k1, k2) through a KA_STAGES-deep ring
        if (elect_one()) {
            uint32_t it = 0;
            int st = 0;
            uint32_t ph = 0;
            const uint32_t b_bytes = (uint32_t)b_stage * 2u, k_bytes = (uint32_t)a.NB * 8u;
            if (a.n_chunks == 1) {          // the whole weight image stays resident: loaded once
                mbar_arrive_expect_tx(&bars->b_full[0], b_bytes + k_bytes);
                bulk_g2s(Bs, a.wimg, b_bytes, &bars->b_full[0]);
                bulk_g2s(s_k12, a.k12, k_bytes, &bars->b_full[0]);
            }
            for (int64_t tile = blockIdx.x; a.n_chunks > 1 && tile < n_tiles; tile += gridDim.x) {
                for (int c = 0; c < a.n_chunks; ++c, ++it) {
                    mbar_wait_relaxed(&bars->b_empty[st], ph ^ 1, 200);
                    trace(a.trace, it, 0);
                    mbar_arrive_expect_tx(&bars->b_full[st], b_bytes + k_bytes);
                    bulk_g2s(Bs + (size_t)st * b_stage, a.wimg + (size_t)c * b_stage, b_bytes, &bars->b_full[st]);
                    bulk_g2s(s_k12 + (it & (K12_RING - 1)) * a.NB, a.k12 + (size_t)c * a.NB, k_bytes, &bars->b_full[st]);
                    if (++st == KA_STAGES) {
                        st = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == KA_MMA_WARP) {
        if (elect_one()) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(a.NB >> 3) << 17) | ((uint32_t)(FR >> 4) << 24);   // f16 x f16 -> f32
            uint32_t it = 0, tile_it = 0;
            int st = 0;
            uint32_t ph = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
                const int ab = tile_it & 1;
                mbar_wait(&bars->a_full[ab], (tile_it >> 1) & 1);
                const uint32_t a_hi = tmem_base + (ab ? COL_A1 : COL_A0), a_lo = a_hi + KP / 2;
                for (int c = 0; c < a.n_chunks; ++c, ++it) {
                    const int buf = it & 1;
                    if (a.n_chunks > 1 || it == 0) mbar_wait(&bars->b_full[st], ph);      // (a resident image lands once)
                    trace(a.trace, it, 1);
                    mbar_wait_relaxed(&bars->t_empty[buf], ((it >> 1) & 1) ^ 1, 32);
                    trace(a.trace, it, 2);
                    tc_fence_after();
                    const uint32_t b_hi = smem_u32(Bs + (size_t)st * b_stage), b_lo = b_hi + (uint32_t)a.NB * KP * 2u;
                    const uint32_t d_tmem = tmem_base + (uint32_t)buf * 256u;
                    // descriptors of k-step s = the first one + 16 s (256 bytes >> 4): additions on the uniform datapath
                    const uint64_t dbh = make_desc(b_hi, LBO, SBO), dbl = make_desc(b_lo, LBO, SBO);
#pragma unroll
                    for (int s = 0; s < KSTEPS; ++s) {
                        umma_f16_ts(d_tmem, a_hi + 8u * s, dbh + 16u * s, idesc, s > 0);
                        umma_f16_ts(d_tmem, a_hi + 8u * s, dbl + 16u * s, idesc, 1);      // statistics hi x weights lo
                        umma_f16_ts(d_tmem, a_lo + 8u * s, dbh + 16u * s, idesc, 1);      // statistics lo x weights hi
                    }
                    umma_commit(&bars->t_full[buf]);
                    trace(a.trace, it, 3);
                    if (a.n_chunks > 1) {
                        umma_commit(&bars->b_empty[st]);
                        if (++st == KA_STAGES) {
                            st = 0;
                            ph ^= 1;
                        }
                    }
                }
                umma_commit(&bars->a_empty[ab]);
            }
        }
    } else {
        // epilogue: TMEM lane = frame row 32 (warp % 4) + lane.  Two groups of eight warps take alternate chunks
        // (= alternate accumulator buffers): while one group sits in its barrier / tensor-memory / store latencies the
        // other one keeps the special-function unit busy (one exp2 per frame and Gaussian is the floor of this kernel;
        // all sixteen warps on one chunk left it idle 40 % of the time).  The two warps a group has per lane quarter
        // take half of the chunk's units (whole pdfs) each.
        const int r = (warp & 3) * 32 + lane, part = warp >> 2;
        const int group = part & 1, cq = part >> 1;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const int n_units = a.NB / UNIT;
        const int u0 = cq * n_units / KA_NCQ, u1 = (cq + 1) * n_units / KA_NCQ;
        const int qq = warp & 3;
        // the statistics row of this thread's frame: image tile -> registers -> tensor memory (first warp of each quarter)
        auto load_a = [&](int64_t tile, uint32_t tile_it) {
            const int ab = tile_it & 1;
            mbar_wait(&bars->a_empty[ab], ((tile_it >> 1) & 1) ^ 1);
            tc_fence_after();
            const int64_t t64 = 2 * tile + (r >> 6);
            const int f = r & 63;
            const uint4* src = reinterpret_cast<const uint4*>(a.img1 + (size_t)t64 * (2 * TILE * KP) + (f >> 3) * (KP * 8) + (f & 7) * 8);
            const bool have = t64 < n_tiles64;
            const uint32_t taddr = tmem_base + lane_addr + (ab ? COL_A1 : COL_A0);
#pragma unroll
            for (int h = 0; h < 2; ++h) {                 // hi image, lo image
                uint32_t w[KP / 2];
#pragma unroll
                for (int k8 = 0; k8 < KP / 8; ++k8) {     // 16-byte row pieces, 128 bytes apart
                    const uint4 v = have ? __ldg(src + h * (TILE * KP / 8) + k8 * 8) : make_uint4(0u, 0u, 0u, 0u);
                    w[4 * k8] = v.x; w[4 * k8 + 1] = v.y; w[4 * k8 + 2] = v.z; w[4 * k8 + 3] = v.w;
                }
#pragma unroll
                for (int c = 0; c < KP / 2; c += 8) {
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(
                                     taddr + (uint32_t)(h * (KP / 2) + c)),
                                 "r"(w[c]), "r"(w[c + 1]), "r"(w[c + 2]), "r"(w[c + 3]), "r"(w[c + 4]), "r"(w[c + 5]),
                                 "r"(w[c + 6]), "r"(w[c + 7])
                                 : "memory");
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars->a_full[ab]);
        };
        uint32_t it = 0, tile_it = 0;
        if (part == 0 && (int64_t)blockIdx.x < n_tiles) load_a(blockIdx.x, 0);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
            // the next tile's statistics go to the other A buffer while this tile's chunks run
            if (part == 0 && tile + gridDim.x < n_tiles) load_a(tile + gridDim.x, tile_it + 1);
            const int64_t t = tile * FR + r;
            const bool valid = t < a.N;
            float* orow = a.llh2 + (size_t)(valid ? t : 0) * a.ld;
            for (int c = 0; c < a.n_chunks; ++c, ++it) {
                if ((int)(it & 1) != group) continue;
                const int buf = group;
                mbar_wait(&bars->t_full[buf], (it >> 1) & 1);
                if (lane == 0 && (warp == 0 || warp == 15)) trace(a.trace, it, warp == 0 ? 4 : 6);
                tc_fence_after();
                const float2* kk = s_k12 + (a.n_chunks > 1 ? (it & (K12_RING - 1)) * a.NB : 0);
                const uint32_t taddr = tmem_base + lane_addr + (uint32_t)buf * 256u;
                // staging buffers [lane quarter][group][2]: this frame's row of the chunk
                float* stile = s_out + (size_t)((qq * 2 + group) * 2 + ((it >> 1) & 1)) * 32 * PC;
                float* sbuf = stile + (size_t)lane * PC;
                // G units at a time: all their TMEM loads in flight together, independent max / exp / sum chains
                auto process = [&](int u, auto gtag) {
                    constexpr int G = decltype(gtag)::value;
                    const int p = u * UNIT;
                    float v[G * UNIT];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        if constexpr (UNIT == 8) tmem_ld8_nowait(taddr + (uint32_t)(p + g * UNIT), v + g * UNIT);
                        else tmem_ld16_nowait(taddr + (uint32_t)(p + g * UNIT), v + g * UNIT);
                    }
                    tmem_ld_wait();
                    if (u + G == u1) {
                        // the warp's last columns are in registers: the accumulator goes back to the MMA thread now, the
                        // math of this batch runs under the next chunk's MMAs (an accumulator belongs to one group: held
                        // to the end of the epilogue, the MMAs of chunk i + 2 started only then and the group waited
                        // ~1100 cycles per chunk for them)
                        tc_fence_before();
                        mbar_arrive(&bars->t_empty[buf]);
                    }
#pragma unroll
                    for (int i = 0; i < G * UNIT; i += 2) {
                        const float4 k = *reinterpret_cast<const float4*>(kk + p + i);      // (k1, k2) of two columns
                        v[i] = fmaf(v[i], k.x, k.y);
                        v[i + 1] = fmaf(v[i + 1], k.z, k.w);
                    }
                    float o[G * PPU];
#pragma unroll
                    for (int k = 0; k < G * PPU; ++k) {
                        if constexpr (C == 1) {       // single-Gaussian pdfs: z is the llh
                            o[k] = v[k];
                            continue;
                        }
                        const float m = tree_max<C>(v + k * C);
                        const float ms = (m == kNegInf) ? 0.f : m;
#pragma unroll
                        for (int j = 0; j < C; ++j) v[k * C + j] = ex2(v[k * C + j] - ms);
                        o[k] = ms + lg2(tree_sum<C>(v + k * C));
                    }
                    if (a.staged) {
#pragma unroll
                        for (int k = 0; k < G * PPU; ++k) sbuf[p / C + k] = o[k];
                    } else if (valid) {
                        const int k0 = (c * a.NB + p) / C;
#pragma unroll
                        for (int k = 0; k < G * PPU; ++k)
                            if (k0 + k < a.Kp) orow[k0 + k] = o[k];
                    }
                };
                // batches of GMAX units, the remainder first: the LAST batch is a full one (half of the warp's columns at
                // the cfg3 shape), so the accumulator is released halfway through the epilogue
                constexpr int GMAX = UNIT == 8 ? (KA_NCQ >= 3 ? 3 : 5) : 2;
                int u = u0;
                int rem = (u1 - u0) % GMAX;
                if (rem >= 4) { process(u, std::integral_constant<int, 4>()); u += 4; rem -= 4; }
                if (rem >= 2) { process(u, std::integral_constant<int, 2>()); u += 2; rem -= 2; }
                if (rem >= 1) { process(u, std::integral_constant<int, 1>()); u += 1; }
                for (; u + GMAX <= u1; u += GMAX) process(u, std::integral_constant<int, GMAX>());
                if (u1 == u0) {          // (no units for this warp: it still owes its arrival)
                    tc_fence_before();
                    mbar_arrive(&bars->t_empty[buf]);
                }
                if (a.staged) {
                    // one tensor store per lane quarter and chunk: [32 frames x NB / C pdfs], rows past N and columns
                    // past Kp clipped by the map.  Two staging buffers per (quarter, group): the issuing lane waits for
                    // its previous store (the other buffer, issued a whole chunk ago) BEFORE the barrier, so behind
                    // the barrier both warps may write the other buffer.
                    fence_proxy_async();
                    if (cq == 0 && elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + qq + 4 * group), "n"(32 * KA_NCQ) : "memory");
                    if (cq == 0 && elect_one()) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&omap),
                                     "r"(smem_u32(stile)), "r"(c * PC), "r"((int)(tile * FR) + 32 * qq)
                                     : "memory");
                        bulk_commit();
                    }
                }
                if (lane == 0 && (warp == 0 || warp == 15)) trace(a.trace, it, warp == 0 ? 5 : 7);
            }
        }
        if (a.staged && cq == 0 && elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == KA_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static size_t ka_smem(int KP, int NB, int C, int n_chunks, bool staged) {
    const int stages = n_chunks > 1 ? KA_STAGES : 1, k12 = n_chunks > 1 ? K12_RING : 1;
    return (size_t)stages * 2 * NB * KP * 2 + (size_t)k12 * NB * 8 + sizeof(KaBarriers) + 128 +
           (staged ? (size_t)4 * 4 * 32 * (NB / C) * 4 : 0) + 1024;
}

static int encode_rows(CUtensorMap* map, const float* base, int64_t N, int Kp, int64_t ld, int box_cols, int box_rows);

template <int KP, int C>
static int launch_ka(const KaArgs& a0, cudaStream_t st) {
    KaArgs a = a0;
    // staged tensor stores when the staging buffers fit next to the weight ring (wide single-Gaussian chunks do not)
    const bool fits = ka_smem(KP, a.NB, C, a.n_chunks, true) <= 227 * 1024;
    const size_t smem = ka_smem(KP, a.NB, C, a.n_chunks, fits);
    if (smem > 227 * 1024) return BEER_ERR_UNSUPPORTED;
    static bool attr = false;
    if (!attr) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(emission16_kernel<KP, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    // staged tensor stores need 16-byte rows in shared memory and in HBM
    CUtensorMap omap;
    memset(&omap, 0, sizeof(omap));
    const int PC = a.NB / C;
    a.staged = (fits && PC % 4 == 0 && a.ld % 4 == 0 && ((uintptr_t)a.llh2 & 15) == 0 && a.N < ((int64_t)1 << 31)) ? 1 : 0;
    if (a.staged && encode_rows(&omap, a.llh2, a.N, a.Kp, a.ld, PC, 32) != BEER_OK) a.staged = 0;
    const int64_t n_tiles = (a.N + 127) / 128;
    const int grid = (int)std::min<int64_t>(n_tiles, kNumSMs);
    emission16_kernel<KP, C><<<grid, KA_THREADS, smem, st>>>(a, omap);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

// ---------------------------------------------------------------------------------------------
// KCF: acc [M, 2D+2] += sum_t w_tj T(x_t),  w = pdf posterior x responsibility, responsibilities recomputed on chip
// ---------------------------------------------------------------------------------------------
constexpr int GM = 128;                  // Gaussians per CTA (UMMA M, TMEM lanes)
constexpr int EPI = 512;                 // 16 epilogue warps: 4 per TMEM lane quarter, each a quarter of the 64 frames
constexpr int KC_MMA_WARP = EPI / 32, KC_LOAD_WARP = KC_MMA_WARP + 1;
constexpr int KC_LOADERS = 3;            // one issuing thread per copy stream (20 warps in all: 96 registers per thread)
constexpr int KC_THREADS = EPI + 32 + 32 * KC_LOADERS;
constexpr int RING_MAX = 8;              // upper bound of either shared-memory ring
constexpr int NSB = 4;                   // S^T / A2 buffers in tensor memory (64 columns each; the first MMA fills a PAIR: N = 128)
constexpr int DR = 4;                    // tiles per drain of the statistics accumulator (48 truncating accumulations)

struct KcArgs {
    const __half* img1;
    const __half* img2;
    int64_t N;
    const uint32_t* wtm;
    const float2* k12;
    const float* alpha;
    int M, Kp, n_gtiles;
    int64_t frames_per_cta;  // multiple of TILE
    float wexp;              // w is carried as w 2^wexp (top of the fp16 range)
    int na, nb;              // depth of the two shared-memory rings
    double* acc;
    int D;
    unsigned long long* trace;
    const uint8_t* block_active;   // [ceil(N / TILE), ld_active] or NULL: tiles of 64 frames in which this Gaussian tile's pdfs
    int64_t ld_active;             // carry weight at all (beer_hmm_forward_backward_blocks); the others are skipped
    int act_cap;                   // tiles per CTA (capacity of the list of active tiles in shared memory)
};

struct KcBarriers {
    uint64_t a_full[RING_MAX], a_empty[RING_MAX];     // ring A: img1 tiles (first MMA), freed as soon as S^T is computed
    uint64_t b_full[RING_MAX], b_empty[RING_MAX];     // ring B: img2 tile + llh / posterior blocks (epilogue, second MMA)
    uint64_t s_full[NSB], a2_full[NSB];
    uint64_t a2_empty[NSB];                           // single-Gaussian mode: the second MMA has read the A2 buffer
    uint64_t d2_full[2], d2_empty[2];
    uint32_t tmem_base;
    int n_act;                                        // active tiles of this CTA
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// map_l2 / map_lp: [N, Kp] fp32 arrays (log2 pdf llhs of KA16, log2 pdf posteriors of the forward-backward), boxes of
// [TILE frames x GM / C pdfs]; rows past N and columns past Kp arrive as zeros.
// REL: map_lp holds d = log2(scale posterior) - llh2 in ONE array (beer_hmm_forward_backward_ex, lpost_relative): one
// block per stage, one load and one addition per (frame, Gaussian) instead of two and three.
template <int KP, int C, bool REL>
__global__ void __launch_bounds__(KC_THREADS, 1)
mixstats16_kernel(KcArgs a, const __grid_constant__ CUtensorMap map_l2, const __grid_constant__ CUtensorMap map_lp) {
    constexpr int KS1 = KP / 16;                       // k-steps of S^T = W' . img1^T
    constexpr int KS2 = TILE / 16;                     // k-steps of acc += A2 . img2^T
    constexpr int IMG_HALF = TILE * KP;                // halfs of one image half-tile
    constexpr bool SINGLE = C == 1;                    // single-Gaussian pdfs: w = the pdf posterior itself, no first MMA
    constexpr int NK = GM / C;                         // pdfs of a Gaussian tile
    constexpr int RAW_FLOATS = TILE * NK;              // one [frame][pdf] block
    constexpr int STAGE_A = 2 * IMG_HALF * 2;          // img1 hi | lo
    constexpr int NRAW = (SINGLE || REL) ? 1 : 2;
    constexpr int STAGE_B = 2 * IMG_HALF * 2 + NRAW * RAW_FLOATS * 4;    // img2 hi | lo | llh2 | lpost (SINGLE: posteriors, REL: d)
    constexpr uint32_t COL_W = 0, COL_S = KP, COL_D2 = KP + NSB * TILE;     // tensor-memory columns
    static_assert(COL_D2 + 2 * KP <= 512, "tensor memory");
    static_assert(STAGE_A % 128 == 0 && STAGE_B % 128 == 0, "stage alignment");
    constexpr int NCH = KP / 4;                        // 4-column chunks of the accumulator
    constexpr int MYCH = (NCH + 3) / 4;                // per thread (four warps share a lane quarter)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* ring_a = smem_raw;
    uint8_t* ring_b = smem_raw + (size_t)a.na * STAGE_A;
    KcBarriers* bars = reinterpret_cast<KcBarriers*>(ring_b + (size_t)a.nb * STAGE_B);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gtile = blockIdx.x % a.n_gtiles;
    const int g0 = gtile * GM;
    const int64_t f_begin = (int64_t)(blockIdx.x / a.n_gtiles) * a.frames_per_cta;
    const int64_t f_end = min(a.N, f_begin + a.frames_per_cta);
    const int n_all = (f_end > f_begin) ? (int)((f_end - f_begin + TILE - 1) / TILE) : 0;
    const int64_t tile0 = f_begin / TILE;
    // the tiles this CTA works on: all of its range, or (activity map) those in which the pdfs of its Gaussian tile have
    // posterior mass -- elsewhere the weights are zero in both fp16 halves and the tile adds exactly nothing
    const bool sparse = !SINGLE && a.block_active != nullptr;
    uint16_t* act = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(bars + 1));
    auto tix = [&](int k) { return sparse ? (int)act[k] : k; };

    if (tid == 0) {
        for (int i = 0; i < RING_MAX; ++i) {
            mbar_init(&bars->a_full[i], 1);
            mbar_init(&bars->a_empty[i], 1);
            mbar_init(&bars->b_full[i], 2);
            mbar_init(&bars->b_empty[i], 1);
        }
        for (int i = 0; i < NSB; ++i) {
            mbar_init(&bars->s_full[i], 1);
            mbar_init(&bars->a2_full[i], EPI / 2);
            mbar_init(&bars->a2_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->d2_full[i], 1);
            mbar_init(&bars->d2_empty[i], EPI);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == KC_MMA_WARP) tmem_alloc(&bars->tmem_base, 512);
    if (warp == 4) {
        int count = n_all;
        if (sparse) {
            count = 0;
            for (int base = 0; base < n_all; base += 32) {
                const int t = base + lane;
                const bool f = t < n_all && a.block_active[(size_t)(tile0 + t) * a.ld_active + gtile] != 0;
                const uint32_t mask = __ballot_sync(0xffffffffu, f);
                if (f) act[count + __popc(mask & ((1u << lane) - 1u))] = (uint16_t)t;
                count += __popc(mask);
            }
        }
        if (lane == 0) bars->n_act = count;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const int n_tiles = bars->n_act;

    // the packed weights of this Gaussian tile -> tensor memory (A operand of the first MMA), once
    if (!SINGLE && warp < 4) {
        const uint32_t* row = a.wtm + (size_t)(g0 + warp * 32 + lane) * KP;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + COL_W;
#pragma unroll 1
        for (int c = 0; c < KP; c += 16) {
            uint32_t r[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + c) + q);
                r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
            }
            tmem_st16(taddr + (uint32_t)c, r);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp >= KC_LOAD_WARP) {
        // ------------------------------ loaders (TMA) ------------------------------
        // three warps, one copy stream each: img1 tiles -> ring A; img2 tiles | llh + posterior blocks -> ring B
        if (elect_one()) {
            const int which = warp - KC_LOAD_WARP;
            const uint32_t bytes = STAGE_A;                   // hi + lo of one image
            const int k0 = g0 / C;
            if (which == 0) {
                // img1 tiles in PAIRS: [hi(t) | hi(t + 1) | lo(t) | lo(t + 1)], so that one descriptor spans the 128
                // frames of the pair (a tile's hi / lo halves are adjacent in HBM: four copies per pair)
                Ring rp(a.na / 2);
                const uint32_t half_bytes = IMG_HALF * 2u;
                for (int i = 0; !SINGLE && i < n_tiles; i += 2, rp.next()) {
                    const int nt = min(2, n_tiles - i);
                    mbar_wait_relaxed(&bars->a_empty[rp.pos], rp.phase ^ 1, 200);
                    trace(a.trace, i, 0);
                    mbar_arrive_expect_tx(&bars->a_full[rp.pos], (uint32_t)nt * bytes);
                    uint8_t* dst = ring_a + (size_t)rp.pos * (2 * STAGE_A);
                    for (int t = 0; t < nt; ++t) {
                        const __half* src = a.img1 + (size_t)(tile0 + tix(i + t)) * (2 * IMG_HALF);
                        bulk_g2s(dst + (size_t)t * half_bytes, src, half_bytes, &bars->a_full[rp.pos]);
                        bulk_g2s(dst + (size_t)(2 + t) * half_bytes, src + IMG_HALF, half_bytes, &bars->a_full[rp.pos]);
                    }
                }
            }
            Ring r(a.nb);
            for (int i = 0; i < (which == 0 ? 0 : n_tiles); ++i, r.next()) {
                const int t0 = (int)(f_begin + (int64_t)tix(i) * TILE);
                mbar_wait_relaxed(&bars->b_empty[r.pos], r.phase ^ 1, 200);
                uint8_t* dst = ring_b + (size_t)r.pos * STAGE_B;
                if (which == 1) {
                    mbar_arrive_expect_tx(&bars->b_full[r.pos], bytes);
                    bulk_g2s(dst, a.img2 + (size_t)(tile0 + tix(i)) * (2 * IMG_HALF), bytes, &bars->b_full[r.pos]);
                } else if (SINGLE || REL) {
                    mbar_arrive_expect_tx(&bars->b_full[r.pos], RAW_FLOATS * 4u);
                    tma_load_2d(dst + bytes, &map_lp, k0, t0, &bars->b_full[r.pos]);       // [64 frames x 128 posteriors]
                } else {
                    mbar_arrive_expect_tx(&bars->b_full[r.pos], 2u * RAW_FLOATS * 4u);
                    tma_load_2d(dst + bytes, &map_l2, k0, t0, &bars->b_full[r.pos]);
                    tma_load_2d(dst + bytes + RAW_FLOATS * 4, &map_lp, k0, t0, &bars->b_full[r.pos]);
                }
            }
        }
    } else if (warp == KC_MMA_WARP) {
        // ------------------------------ MMA issuer ---------------------------------
        if (n_tiles > 0 && elect_one()) {
            const uint32_t idesc1 = (1u << 4) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);
            const uint32_t w_hi = tmem_base + COL_W, w_lo = w_hi + KP / 2;
            // S^T of TWO tiles per MMA (N = 128 frames): a 128 x 64 x 16 MMA with A in tensor memory takes ~51 cycles
            // against 33 of tensor work, the 27 MMAs of a tile were the floor of this kernel (1400 cycles per tile)
            const uint32_t idesc1p = (1u << 4) | ((uint32_t)((2 * TILE) >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);
            Ring ra(a.na / 2), rb(a.nb);
            int b2 = 0;                              // S^T buffer of the next second MMA
            uint32_t ph2 = 0;
            auto issue_g1 = [&](int i) {             // tiles i, i + 1 (i even) -> buffers (i % 4), (i % 4) + 1
                mbar_wait(&bars->a_full[ra.pos], ra.phase);
                trace(a.trace, i, 1);
                tc_fence_after();
                const uint32_t base = smem_u32(ring_a + (size_t)ra.pos * (2 * STAGE_A));
                const uint64_t dbh = make_desc(base, 128, KP * 16), dbl = make_desc(base + 2u * IMG_HALF * 2u, 128, KP * 16);
                const int b1 = i & 2;
                const uint32_t d = tmem_base + COL_S + (uint32_t)b1 * TILE;
                const uint32_t idesc = (i + 1 < n_tiles) ? idesc1p : idesc1;
#pragma unroll
                for (int ks = 0; ks < KS1; ++ks) {
                    umma_f16_ts(d, w_hi + 8u * ks, dbh + 16u * ks, idesc, ks > 0);
                    umma_f16_ts(d, w_lo + 8u * ks, dbh + 16u * ks, idesc, 1);        // weights lo x statistics hi
                    umma_f16_ts(d, w_hi + 8u * ks, dbl + 16u * ks, idesc, 1);        // weights hi x statistics lo
                }
                umma_commit(&bars->s_full[b1]);
                umma_commit(&bars->a_empty[ra.pos]);          // img1 tiles free: their loader runs ahead of the epilogue
                trace(a.trace, i, 2);
                ra.next();
            };
            auto issue_g2 = [&](int i) {
                const int grp = i / DR, dbuf = grp & 1;
                const bool first = (i % DR) == 0, last = (i % DR) == DR - 1 || i == n_tiles - 1;
                mbar_wait_relaxed(&bars->a2_full[b2], ph2, 32);
                trace(a.trace, i, 3);
                if (first) mbar_wait(&bars->d2_empty[dbuf], ((grp >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t base = smem_u32(ring_b + (size_t)rb.pos * STAGE_B);
                const uint64_t dbh = make_desc(base, 128, 1024), dbl = make_desc(base + IMG_HALF * 2u, 128, 1024);
                const uint32_t d = tmem_base + COL_D2 + (uint32_t)dbuf * KP;
                const uint32_t a2 = tmem_base + COL_S + (uint32_t)b2 * TILE;
#pragma unroll
                for (int ks = 0; ks < KS2; ++ks) {
                    umma_f16_ts(d, a2 + 16u * ks, dbh + 16u * ks, idesc2, !(first && ks == 0));
                    umma_f16_ts(d, a2 + 16u * ks + 8u, dbh + 16u * ks, idesc2, 1);   // w lo x statistics hi
                    umma_f16_ts(d, a2 + 16u * ks, dbl + 16u * ks, idesc2, 1);        // w hi x statistics lo
                }
                umma_commit(&bars->b_empty[rb.pos]);    // the epilogue finished with the stage before a2_full completed
                if constexpr (SINGLE) umma_commit(&bars->a2_empty[b2]);     // (mixtures: implied by the next s_full of the buffer)
                if (last) umma_commit(&bars->d2_full[dbuf]);
                trace(a.trace, i, 4);
                rb.next();
                if (++b2 == NSB) {
                    b2 = 0;
                    ph2 ^= 1;
                }
            };
            if constexpr (SINGLE) {
                for (int i = 0; i < n_tiles; ++i) issue_g2(i);
            } else {
                issue_g1(0);
                for (int i = 0; i < n_tiles; i += 2) {
                    if (i + 2 < n_tiles) issue_g1(i + 2);     // (overwrites the buffers of the pair before this one:
                    issue_g2(i);                              //  their second MMAs were issued an iteration ago)
                    if (i + 1 < n_tiles) issue_g2(i + 1);
                }
            }
        }
    } else {
        // ------------------------------ epilogue warps -------------------------------
        // Two groups of eight warps take alternate tiles: while one group sits in its barrier / tensor-memory
        // latencies the other one computes (all sixteen warps on one tile spent a third of the tile time waiting).
        const int q = warp & 3, part = warp >> 2;            // TMEM lane quarter; quarter of the accumulator columns (drain)
        const int group = part & 1, half = part >> 1;        // tile parity; half of the tile's frames
        const int g = q * 32 + lane;                         // Gaussian (local) = TMEM lane
        const int pl = g / C;                                // its pdf (local)
        float k1 = 0.f, k2 = 0.f;
        if constexpr (!SINGLE) {
            const float2 k12 = __ldg(a.k12 + g0 + g);
            k1 = k12.x;
            k2 = k12.y;
        }
        const float wscale = exp2f(a.wexp);
        const float k2w = k2 + a.wexp;                       // REL: w 2^wexp = 2^(S k1 + (k2 + wexp) + d)
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        float sums[MYCH][4], comp[MYCH][4];
#pragma unroll
        for (int m = 0; m < MYCH; ++m)
#pragma unroll
            for (int i = 0; i < 4; ++i) sums[m][i] = comp[m][i] = 0.f;
        float wsum = 0.f, wcomp = 0.f;                       // sum_t w 2^wexp of this thread's frames (Kahan over tiles)
        int next_drain = 0;
        auto drain = [&](int grp) {
            const int dbuf = grp & 1;
            mbar_wait(&bars->d2_full[dbuf], (grp >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int m = 0; m < MYCH; ++m) {
                const int ch = part + 4 * m;
                if (ch < NCH) {
                    float v[4];
                    tmem_ld4(tmem_base + lane_addr + COL_D2 + (uint32_t)(dbuf * KP + ch * 4), v);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float y = v[i] - comp[m][i];
                        const float t = sums[m][i] + y;
                        comp[m][i] = (t - sums[m][i]) - y;
                        sums[m][i] = t;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&bars->d2_empty[dbuf]);
        };

        Ring rb(a.nb);
        if (group) rb.next();
        int b = group;
        uint32_t phs = 0;
        const int late = group ? 1 : 2;          // a finished drain group is collected one / two tiles into the next one
        for (int i = group; i < n_tiles; i += 2) {
            mbar_wait(&bars->b_full[rb.pos], rb.phase);          // the llh / posterior blocks of the tile (TMA)
            if (tid == 0) trace(a.trace, i, 5);
            if constexpr (SINGLE) mbar_wait(&bars->a2_empty[b], phs ^ 1);     // the MMA of the tile three back has read the buffer
            else mbar_wait(&bars->s_full[b & ~1], phs);          // (one completion per pair of tiles)
            if (tid == 0) trace(a.trace, i, 6);
            tc_fence_after();
            const float* raw = reinterpret_cast<const float*>(ring_b + (size_t)rb.pos * STAGE_B + STAGE_A);
            const int n_left = (int)min((int64_t)TILE, f_end - (f_begin + (int64_t)tix(i) * TILE)) - half * 32;
            float ts[4] = {0.f, 0.f, 0.f, 0.f};
            // one block of 16 frames: w 2^wexp -> fp16 hi (the top 11 significant bits, by truncation: exact in fp16)
            // and lo = w - hi (exact in fp32, rounded to fp16), written back in place of S^T.  MASKED (the last tile of
            // the batch only): frames past the end carry no weight.
            auto block16 = [&](int sub, auto masked_tag) {
                constexpr bool MASKED = decltype(masked_tag)::value;
                const int f0 = half * 32 + sub * 16;             // first frame (in the tile) of this block of 16
                const uint32_t taddr = tmem_base + lane_addr + COL_S + (uint32_t)(b * TILE + f0);
                float v[16];
                if constexpr (SINGLE) {
                    // w = the pdf posterior (rows past N and columns past Kp arrive as zeros)
                    const float* rp = raw + f0 * NK + pl;
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = rp[e * NK] * wscale;
                } else if constexpr (REL) {
                    tmem_ld16(taddr, v);
                    const float* rd = raw + f0 * NK + pl;
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = ex2(fmaf(v[e], k1, k2w) + rd[e * NK]);
                } else {
                    tmem_ld16(taddr, v);
                    const float* rl2 = raw + f0 * NK + pl;
                    const float* rlp = rl2 + RAW_FLOATS;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        // z = S k1 + k2 is bit for bit the value KA16 normalised: z - llh2 = log2 responsibility
                        const float z = fmaf(v[e], k1, k2);
                        v[e] = ex2((z - rl2[e * NK]) + (rlp[e * NK] + a.wexp));
                    }
                }
                if constexpr (MASKED) {
                    const int nvalid = n_left - sub * 16;
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (e >= nvalid) v[e] = 0.f;
                }
                uint32_t out[16];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    float wh[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        ts[e] += v[4 * h + e];
                        wh[e] = __uint_as_float(__float_as_uint(v[4 * h + e]) & 0xffffe000u);
                    }
                    out[2 * h] = pack_h2(wh[0], wh[1]);
                    out[2 * h + 1] = pack_h2(wh[2], wh[3]);
                    out[8 + 2 * h] = pack_h2(v[4 * h] - wh[0], v[4 * h + 1] - wh[1]);
                    out[8 + 2 * h + 1] = pack_h2(v[4 * h + 2] - wh[2], v[4 * h + 3] - wh[3]);
                }
                tmem_st16(taddr, out);      // in place: [hi of 16 frames (8 columns) | lo (8 columns)]
            };
            if (n_left >= 32) {
                block16(0, std::false_type());
                block16(1, std::false_type());
            } else {
                block16(0, std::true_type());
                block16(1, std::true_type());
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars->a2_full[b]);
            if (tid == 0) trace(a.trace, i, 7);
            {
                const float y = ((ts[0] + ts[1]) + (ts[2] + ts[3])) - wcomp;
                const float t = wsum + y;
                wcomp = (t - wsum) - y;
                wsum = t;
            }
            rb.next();
            rb.next();
            b += 2;
            if (b >= NSB) {
                b -= NSB;
                phs ^= 1;
            }
            if (i % DR == late && i / DR - 1 >= next_drain) drain(next_drain++);
        }
        if (n_tiles > 0) {
            const int last_g = (n_tiles - 1) / DR;
            while (next_drain <= last_g) drain(next_drain++);
            if (g0 + g < a.M) {
                const int D = a.D, Q = 2 * D + 2;
                double* row = a.acc + (size_t)(g0 + g) * Q;
                const double unscale = exp2(-(double)a.wexp);
#pragma unroll
                for (int m = 0; m < MYCH; ++m) {
                    const int ch = part + 4 * m;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = ch * 4 + i;
                        if (ch < NCH && c < 2 * D) {
                            const double val = ((double)sums[m][i] - (double)comp[m][i]) * unscale / (double)__ldg(a.alpha + c);
                            if (val != 0.0) atomicAdd(row + c, val);
                        }
                    }
                }
                const double cnt = ((double)wsum - (double)wcomp) * unscale;
                if (cnt != 0.0) {
                    atomicAdd(row + 2 * D, -0.5 * cnt);
                    atomicAdd(row + 2 * D + 1, 0.5 * cnt);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == KC_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static size_t kc_smem(int KP, int C, bool rel, int na, int nb, int act_cap = 0) {
    return (size_t)na * (2 * TILE * KP * 2) + (size_t)nb * (2 * TILE * KP * 2 + ((C == 1 || rel) ? 1 : 2) * TILE * (GM / C) * 4) +
           sizeof(KcBarriers) + (size_t)act_cap * 2 + 16 + 1024;
}

// log2 of pdf posteriors (the forward-backward kernels that cannot write them themselves)
__global__ void __launch_bounds__(256) log2_post_kernel(const float* __restrict__ post, int64_t N, int Kp, int64_t ld,
                                                        float* __restrict__ out, int64_t ld_out) {
    const int64_t total = N * Kp;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = i / Kp;
        const int k = (int)(i - t * Kp);
        out[t * ld_out + k] = lg2(post[t * ld + k]);
    }
}

// GMM without an HMM on top (Mixture.expected_log_likelihood, mixture.py:70-93): the M components are treated as Kp
// pseudo-pdfs of C each; KA16 has written their log2-sum-exp2, this kernel finishes the softmax over the frame:
// L_t = log2 sum_k 2^llh2[t, k] (the expected llh of the frame, = LSE over all components), lpost[t, k] = llh2[t, k] - L_t
// (so that KCF's w = 2^(z - llh2 + lpost) = 2^(z - L_t) is the component responsibility), and the per-utterance sums.
// One warp per frame.
__global__ void __launch_bounds__(256) gmm_post_kernel(const float* __restrict__ llh2, int64_t N, int Kp, int64_t ld,
                                                       const float* __restrict__ frame_ref, const int64_t* __restrict__ utt_off,
                                                       int n_utts, float scale, float* __restrict__ lpost, int64_t ld_post,
                                                       float* __restrict__ frame_llh, double* __restrict__ utt_exp_llh) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float lscale = lg2(scale);
    constexpr int RUN = 32;           // consecutive frames per warp: one atomic per run and utterance, not per frame
    for (int64_t r0 = warp0 * RUN; r0 < N; r0 += nwarps * RUN) {
        int cur_utt = -1;
        int64_t cur_end = -1;
        double acc = 0.0;
        for (int64_t t = r0; t < min(N, r0 + RUN); ++t) {
            const float* row = llh2 + (size_t)t * ld;
            float m = kNegInf;
            for (int k = lane; k < Kp; k += 32) m = fmaxf(m, row[k]);
            m = warp_max(m);
            const float ms = (m == kNegInf) ? 0.f : m;
            float sm = 0.f;
            for (int k = lane; k < Kp; k += 32) sm += ex2(row[k] - ms);
            sm = warp_sum(sm);
            const float L = ms + lg2(sm);
            for (int k = lane; k < Kp; k += 32) lpost[(size_t)t * ld_post + k] = row[k] - L + lscale;
            const float f = scale * (L * kLn2 + (frame_ref != nullptr ? frame_ref[t] : 0.f));
            if (lane == 0 && frame_llh != nullptr) frame_llh[t] = f;
            if (utt_exp_llh != nullptr) {
                if (t >= cur_end) {       // (warp-uniform) the run enters another utterance
                    if (lane == 0 && cur_utt >= 0) atomicAdd(utt_exp_llh + cur_utt, acc);
                    int lo = 0, hi = n_utts;          // last utterance with utt_off[u] <= t
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (utt_off[mid] <= t) lo = mid; else hi = mid;
                    }
                    cur_utt = lo;
                    cur_end = utt_off[lo + 1];
                    acc = 0.0;
                }
                acc += (double)f;
            }
        }
        if (lane == 0 && cur_utt >= 0) atomicAdd(utt_exp_llh + cur_utt, acc);
    }
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_rows(CUtensorMap* map, const float* base, int64_t N, int Kp, int64_t ld, int box_cols, int box_rows) {
    static EncodeTiled encode = nullptr;
    if (encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        BEER_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (fn == nullptr || q != cudaDriverEntryPointSuccess) return BEER_ERR_UNSUPPORTED;
        encode = reinterpret_cast<EncodeTiled>(fn);
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)N};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return BEER_ERR_UNSUPPORTED;
    return BEER_OK;
}

template <int KP, int C, bool REL = false>
static int launch_kc(const KcArgs& a0, const float* llh2, int64_t ld_llh, const float* lpost, int64_t ld_lpost,
                     int64_t ranges, cudaStream_t st) {
    if constexpr (C != 1 && !REL) {
        if (llh2 == nullptr) return launch_kc<KP, C, true>(a0, lpost, ld_lpost, lpost, ld_lpost, ranges, st);
    }
    KcArgs a = a0;
    // ring A (img1 tiles): 4 deep; ring B (img2 tile + llh / posterior blocks): whatever else fits
    a.na = C == 1 ? 0 : 4;
    a.nb = RING_MAX;
    const int cap = a.block_active != nullptr ? a.act_cap : 0;
    while (a.nb > 2 && kc_smem(KP, C, REL, a.na, a.nb, cap) > 227 * 1024) --a.nb;
    while (C != 1 && a.na > 2 && kc_smem(KP, C, REL, a.na, a.nb, cap) > 227 * 1024) a.na -= 2;     // (pairs of tiles)
    const size_t smem = kc_smem(KP, C, REL, a.na, a.nb, cap);
    if (smem > 227 * 1024) return BEER_ERR_UNSUPPORTED;
    CUtensorMap m1, m2;
    int rc = encode_rows(&m1, llh2, a.N, a.Kp, ld_llh, GM / C, TILE);
    if (rc != BEER_OK) return rc;
    rc = encode_rows(&m2, lpost, a.N, a.Kp, ld_lpost, GM / C, TILE);
    if (rc != BEER_OK) return rc;
    static bool attr = false;
    if (!attr) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(mixstats16_kernel<KP, C, REL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    mixstats16_kernel<KP, C, REL><<<(int)(ranges * a.n_gtiles), KC_THREADS, smem, st>>>(a, m1, m2);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

static int nb_of(int M, int C) {
    // Gaussians per KA16 chunk: the largest divisor of M that is a multiple of 16 and of C, at most 256 (no padded
    // columns in the last chunk); 128 with padding when M has none
    if (M <= 160 && C <= 16) return (M + 15) / 16 * 16;   // one chunk, no wider than needed
    int best = 0;
    for (int nb = 16; nb <= 160; nb += 16)      // two weight stages, two statistics tiles and the staged output fit in shared memory
        if (M % nb == 0 && nb % C == 0) best = nb;
    if (best >= 64) return best;
    return 128;
}

}  // namespace mix16
}  // namespace beer

using namespace beer;

extern "C" {

// Debug: device buffer of 256 x 8 time stamps (clock64 of CTA 0) filled by the next emission / statistics launches:
// emission, per chunk: 0 weight copy issued, 1 weights landed, 2 accumulator free, 3 MMAs issued, 4/5 first epilogue warp
// starts / ends, 6/7 last epilogue warp; statistics, per tile: 0 image copy issued, 1 stage landed (MMA thread), 2 first
// MMAs issued, 3 responsibilities ready, 4 second MMAs issued, 5 stage landed (epilogue), 6 S ready, 7 epilogue done.
void beer_mix16_set_trace(void* dev_buf) { mix16::g_trace = (unsigned long long*)dev_buf; }

int beer_mix16_supported(int M, int D, int C) {
    if (M <= 0 || C <= 0 || M % C != 0) return 0;
    if (!(D == 20 || D == 40)) return 0;
    if (!(C == 1 || C == 4 || C == 8 || C == 16)) return 0;      // C | 128, llh / posterior block of a tile <= 8 KB
    if ((M / C) % 4 != 0) return 0;                      // 16-byte rows of posteriors / llhs
    return 1;
}

int beer_mix16_geometry(int M, int D, int C, int64_t N, int64_t* sizes) {
    if (!beer_mix16_supported(M, D, C) || N < 0 || !sizes) return BEER_ERR_UNSUPPORTED;
    const int KP = mix16::kp_of(D), NB = mix16::nb_of(M, C);
    const int n_chunks = (M + NB - 1) / NB;
    int Mp = n_chunks * NB;
    Mp = (Mp + mix16::GM - 1) / mix16::GM * mix16::GM;
    const int64_t n_tiles = (N + mix16::TILE - 1) / mix16::TILE;
    sizes[0] = n_tiles * 2 * mix16::TILE * KP;                 // halfs of img1 (and of img2)
    sizes[1] = (int64_t)((Mp + NB - 1) / NB) * 2 * NB * KP;    // halfs of wimg
    sizes[2] = (int64_t)Mp * KP;                               // words of wtm
    sizes[3] = 2 * (int64_t)Mp;                                // floats of k12 = (k1, k2) per Gaussian
    sizes[4] = NB;
    sizes[5] = KP;
    return BEER_OK;
}

int beer_mix16_feature_images(const float* X, int64_t N, int D, float* alpha, uint32_t* absmax_scratch, void* img1,
                              void* img2, void* stream) {
    if (!X || !alpha || !absmax_scratch || !img1 || !img2 || N < 0 || !(D == 20 || D == 40)) return BEER_ERR_ARG;
    if (((uintptr_t)X & 15) != 0 || ((uintptr_t)img1 & 127) != 0 || ((uintptr_t)img2 & 127) != 0) return BEER_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    BEER_CUDA_TRY(cudaMemsetAsync(absmax_scratch, 0, (size_t)D * 4, st));
    if (N == 0) return BEER_OK;
    int blocks = (int)std::min<int64_t>((N + 24) / 25, kNumSMs * 8);
    mix16::absmax_kernel<<<blocks, 256, 0, st>>>(X, N, D, absmax_scratch);
    BEER_LAUNCH_CHECK();
    mix16::scales_kernel<<<1, 128, 0, st>>>(absmax_scratch, D, alpha);
    BEER_LAUNCH_CHECK();
    const int64_t n_tiles = (N + mix16::TILE - 1) / mix16::TILE;
    mix16::feat_image_kernel<<<(unsigned)n_tiles, 256, (size_t)mix16::TILE * (D + 1) * 4, st>>>(
        X, N, D, alpha, (__half*)img1, (__half*)img2);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_mix16_pack(const float* W, const float* bias, const float* alpha, int M, int D, int C, void* wimg,
                    uint32_t* wtm, float* k12, void* stream) {
    if (!W || !bias || !alpha || !wimg || !wtm || !k12) return BEER_ERR_ARG;
    if (!beer_mix16_supported(M, D, C)) return BEER_ERR_UNSUPPORTED;
    int64_t sz[6];
    beer_mix16_geometry(M, D, C, 0, sz);
    const int Mp = (int)sz[3] / 2, NB = (int)sz[4];
    mix16::pack16_kernel<<<(Mp * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(W, bias, alpha, M, Mp, D, NB,
                                                                                 (__half*)wimg, wtm, (float2*)k12);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_mix16_frame_ref(const float* X, int64_t N, int D, const float* ref, float* frame_ref, void* stream) {
    if (!X || !ref || !frame_ref || N < 0 || D % 4 != 0 || ((uintptr_t)X & 15) != 0) return BEER_ERR_ARG;
    if (N == 0) return BEER_OK;
    mix16::frame_ref_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, N, D, ref, frame_ref);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_mix16_emission(const void* img1, int64_t N, int D, const void* wimg, const float* k12, int M, int C,
                        float* llh2, int64_t ld, void* stream) {
    if (!img1 || !wimg || !k12 || !llh2 || N < 0) return BEER_ERR_ARG;
    if (!beer_mix16_supported(M, D, C)) return BEER_ERR_UNSUPPORTED;
    if (ld < M / C) return BEER_ERR_ARG;
    if (N == 0) return BEER_OK;
    int64_t sz[6];
    beer_mix16_geometry(M, D, C, N, sz);
    mix16::KaArgs a;
    a.img1 = (const __half*)img1; a.N = N; a.wimg = (const __half*)wimg; a.k12 = (const float2*)k12;
    a.Kp = M / C; a.NB = (int)sz[4];
    a.n_chunks = (M + a.NB - 1) / a.NB;
    a.llh2 = llh2; a.ld = ld; a.staged = 0; a.trace = mix16::g_trace;
    const int KP = (int)sz[5];
    cudaStream_t st = (cudaStream_t)stream;
#define BEER_KA_CASE(kp, c) \
    if (KP == kp && C == c) return mix16::launch_ka<kp, c>(a, st);
    BEER_KA_CASE(80, 1) BEER_KA_CASE(80, 4) BEER_KA_CASE(80, 8) BEER_KA_CASE(80, 16)
    BEER_KA_CASE(48, 1) BEER_KA_CASE(48, 4) BEER_KA_CASE(48, 8) BEER_KA_CASE(48, 16)
#undef BEER_KA_CASE
    return BEER_ERR_UNSUPPORTED;
}

int beer_mix16_log2_posteriors(const float* pdf_post, int64_t N, int Kp, int64_t ld_post, float* pdf_lpost,
                               int64_t ld_lpost, void* stream) {
    if (!pdf_post || !pdf_lpost || N < 0 || Kp <= 0 || ld_post < Kp || ld_lpost < Kp) return BEER_ERR_ARG;
    if (N == 0) return BEER_OK;
    const int blocks = (int)std::min<int64_t>((N * Kp + 255) / 256, kNumSMs * 16);
    mix16::log2_post_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pdf_post, N, Kp, ld_post, pdf_lpost, ld_lpost);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_mix16_gmm_posteriors(const float* llh2, int64_t N, int Kp, int64_t ld_llh, const float* frame_ref,
                              const int64_t* utt_off, int n_utts, float scale, float* pdf_lpost, int64_t ld_lpost,
                              float* frame_exp_llh, double* utt_exp_llh, void* stream) {
    if (!llh2 || !pdf_lpost || N < 0 || Kp <= 0 || ld_llh < Kp || ld_lpost < Kp) return BEER_ERR_ARG;
    if (utt_exp_llh != nullptr && (!utt_off || n_utts <= 0)) return BEER_ERR_ARG;
    if (N == 0) return BEER_OK;
    const int blocks = (int)std::min<int64_t>((N + 8 * 32 - 1) / (8 * 32), kNumSMs * 8);
    mix16::gmm_post_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(llh2, N, Kp, ld_llh, frame_ref, utt_off, n_utts, scale,
                                                                     pdf_lpost, ld_lpost, frame_exp_llh, utt_exp_llh);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_mix16_weight_exponent(float scale) {
    // the posteriors carry `scale`: keep w 2^wexp <= 2^14 (fp16 overflows at 2^16)
    int e = 14;
    if (scale > 1.f) e -= (int)ceilf(log2f(scale));
    return e;
}

int beer_mix16_accumulate(const void* img1, const void* img2, int64_t N, int D, const uint32_t* wtm, const float* k12,
                          const float* alpha, int M, int C, const float* pdf_lpost, int64_t ld_lpost,
                          const float* llh2, int64_t ld_llh, float scale, double* acc_normal, void* stream) {
    return beer_mix16_accumulate_blocks(img1, img2, N, D, wtm, k12, alpha, M, C, pdf_lpost, ld_lpost, llh2, ld_llh, scale,
                                        nullptr, 0, acc_normal, stream);
}

int beer_mix16_accumulate_blocks(const void* img1, const void* img2, int64_t N, int D, const uint32_t* wtm,
                                 const float* k12, const float* alpha, int M, int C, const float* pdf_lpost,
                                 int64_t ld_lpost, const float* llh2, int64_t ld_llh, float scale,
                                 const uint8_t* block_active, int64_t ld_active, double* acc_normal, void* stream) {
    if (!img2 || !alpha || !pdf_lpost || !acc_normal || N < 0) return BEER_ERR_ARG;
    if (C != 1 && (!img1 || !wtm || !k12)) return BEER_ERR_ARG;
    if (!beer_mix16_supported(M, D, C)) return BEER_ERR_UNSUPPORTED;
    const bool one_array = C == 1 || llh2 == nullptr;    // C = 1: the posteriors themselves; else the relative form
    if (one_array) ld_llh = ld_lpost;
    if (ld_lpost < M / C || ld_llh < M / C || ld_lpost % 4 != 0 || ld_llh % 4 != 0) return BEER_ERR_ARG;
    if (((uintptr_t)pdf_lpost & 15) != 0 || ((uintptr_t)llh2 & 15) != 0 || N >= (int64_t)1 << 31) return BEER_ERR_ARG;
    if (C == 1) llh2 = pdf_lpost;
    if (N == 0) return BEER_OK;
    mix16::KcArgs a;
    a.img1 = (const __half*)img1; a.img2 = (const __half*)img2; a.N = N; a.wtm = wtm; a.k12 = (const float2*)k12; a.alpha = alpha;
    a.M = M; a.Kp = M / C; a.D = D; a.acc = acc_normal; a.na = a.nb = 0; a.trace = mix16::g_trace;
    a.n_gtiles = (M + mix16::GM - 1) / mix16::GM;
    a.wexp = (float)beer_mix16_weight_exponent(scale);
    a.block_active = (C != 1) ? block_active : nullptr;
    a.ld_active = ld_active;
    if (a.block_active != nullptr && ld_active < a.n_gtiles) return BEER_ERR_ARG;
    // CTAs = Gaussian tiles x frame ranges, consecutive CTAs = the Gaussian tiles of ONE range (they stream the same
    // image tiles: L2 reuse); the largest grid of whole tile sets within 3 waves
    int64_t ranges = std::max<int64_t>(1, (3 * kNumSMs) / a.n_gtiles);
    const int64_t max_ranges = (N + mix16::TILE - 1) / mix16::TILE;
    if (ranges > max_ranges) ranges = max_ranges;
    int64_t fpc = (N + ranges - 1) / ranges;
    fpc = (fpc + mix16::TILE - 1) / mix16::TILE * mix16::TILE;
    ranges = (N + fpc - 1) / fpc;
    a.frames_per_cta = fpc;
    a.act_cap = (int)(fpc / mix16::TILE);
    if (a.act_cap > 60000) a.block_active = nullptr;         // (16-bit tile indices: such ranges run dense)
    cudaStream_t st = (cudaStream_t)stream;
    const int KP = mix16::kp_of(D);
#define BEER_KC_CASE(kp, c) \
    if (KP == kp && C == c) return mix16::launch_kc<kp, c>(a, llh2, ld_llh, pdf_lpost, ld_lpost, ranges, st);
    BEER_KC_CASE(80, 1) BEER_KC_CASE(80, 4) BEER_KC_CASE(80, 8) BEER_KC_CASE(80, 16)
    BEER_KC_CASE(48, 1) BEER_KC_CASE(48, 4) BEER_KC_CASE(48, 8) BEER_KC_CASE(48, 16)
#undef BEER_KC_CASE
    return BEER_ERR_UNSUPPORTED;
}

}  // extern "C"
