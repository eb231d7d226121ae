// KA backward: gradient of the expected log-likelihood w.r.t. the frames with the posteriors held fixed -- what an
// encoder in front of the model receives (HMM-VAE: beer/models/vae.py:63-89 calls prior.expected_log_likelihood on
// reparameterised samples; hmm.py:79-87 / mixtureset.py:85-98 detach the posteriors).  With
//     w_tj = scale * gamma_t,pdf(j) * r_tj          (r = responsibilities inside the pdf, 1 for single-Gaussian pdfs)
// the derivative of sum_t go_t sum_j w_tj llh_j(x_t), llh_j(x) = E[T(theta_j)] . T(x), T(x) = [x, -x^2/2, ...] is
//     grad_t = go_t * ( sum_j w_tj E[lambda_j mu_j]  -  x_t o sum_j w_tj E[lambda_j] )
// i.e. one [N x M] x [M x 2D] product ((gamma (x) r) @ E[T(theta)]) and an elementwise finish.
//
// tcgen05 kind::f16, 3-pass fp16 split (the precision scheme of mix16.cu), A OPERAND IN TENSOR MEMORY:
//   per CTA a tile of 128 frames (TMEM lanes); per chunk of 64 Gaussians the 8 worker warps compute w in registers
//   (thread = frame, 32 Gaussians), split it into fp16 hi / lo and store it with tcgen05.st as the A operand; the image
//   of E[T(theta)] (columns scaled by powers of two, hi | lo, K-major core-matrix layout, built once by
//   beer_emission_bwd_pack) streams through a shared-memory ring by TMA bulk copies; the accumulator [128 x 2D] is
//   drained into registers every DRG chunks (tensor-core accumulation truncates) and finished through a
//   shared-memory transpose so that the gradient rows are stored coalesced.  w [N, M] never exists in memory.
#include <algorithm>
#include "common.cuh"
#include "tc_common.cuh"
#include "f16_common.cuh"
#include "../../include/beer_b200.h"

namespace beer {
namespace bwd {

using namespace mix16;

constexpr int GC = TILE;                 // Gaussians per chunk (K of one MMA group: 4 k-steps of 16); image layout = off2
constexpr int FM = 128;                  // frames per CTA tile (UMMA M, TMEM lanes)
constexpr int NAB = 3;                   // A buffers in tensor memory
constexpr int WORKERS = 256;             // 8 warps: TMEM lane quarter = warp & 3, half of the chunk's Gaussians = warp >> 2
constexpr int MMA_WARP = WORKERS / 32, LOAD_WARP = MMA_WARP + 1;
constexpr int THREADS = WORKERS + 64;
constexpr int STAGES = 4;                // ring of image chunks
constexpr int DRG = 8;                   // chunks per drain of the accumulator (96 truncating accumulations)

__host__ __device__ inline int kpb_of(int D) { return (2 * D + 15) / 16 * 16; }

// colmax[n] = max_j |ets[j][n]| (bits of a non-negative float compare like integers); caller zeroes colmax
__global__ void __launch_bounds__(128) colmax_kernel(const float* __restrict__ ets, int M, int64_t ld, int D2,
                                                     uint32_t* __restrict__ colmax) {
    const int n = threadIdx.x;
    if (n >= D2) return;
    float m = 0.f;
    for (int j = blockIdx.x; j < M; j += gridDim.x) m = fmaxf(m, fabsf(__ldg(ets + (size_t)j * ld + n)));
    if (isfinite(m)) atomicMax(colmax + n, __float_as_uint(m));
}

// image [n_chunks][hi | lo][KPB x GC] halfs: row n = column n of E[T(theta)] times 2^e_n (largest entry in
// [2^HI_EXP, 2^(HI_EXP+1))), K = the Gaussians of the chunk; inv_scale[n] = 2^-e_n.  One block per chunk.
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ ets, int M, int64_t ld, int D, int KPB,
                                                   const uint32_t* __restrict__ colmax, __half* __restrict__ img,
                                                   float* __restrict__ inv_scale) {
    __shared__ float s_scale[256];
    for (int n = threadIdx.x; n < KPB; n += blockDim.x) {
        float sc = 0.f;
        if (n < 2 * D) {
            const float mx = __uint_as_float(colmax[n]);
            int e = 0;
            if (mx > 0.f) e = max(-100, min(100, HI_EXP - ilogbf(mx)));
            sc = ldexpf(1.f, e);
            if (blockIdx.x == 0) inv_scale[n] = ldexpf(1.f, -e);
        }
        s_scale[n] = sc;
    }
    __syncthreads();
    const int j0 = blockIdx.x * GC;
    __half* hi_img = img + (size_t)blockIdx.x * (2 * KPB * GC);
    __half* lo_img = hi_img + KPB * GC;
    for (int item = threadIdx.x; item < KPB * (GC / 8); item += blockDim.x) {
        const int j8 = item / KPB, n = item - j8 * KPB;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = j0 + 8 * j8 + 2 * e + h;
                v[h] = (j < M && n < 2 * D) ? __ldg(ets + (size_t)j * ld + n) * s_scale[n] : 0.f;
            }
            const __half ha = __float2half_rn(v[0]), hb = __float2half_rn(v[1]);
            hi[e] = (uint32_t)__half_as_ushort(ha) | ((uint32_t)__half_as_ushort(hb) << 16);
            lo[e] = pack_h2(v[0] - __half2float(ha), v[1] - __half2float(hb));
        }
        const int o = off2(n, 8 * j8);
        *reinterpret_cast<uint4*>(hi_img + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(lo_img + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

struct BwArgs {
    const float* X;
    int64_t N;
    int D;
    const float* post;      // [N, ld_post] scale * pdf posteriors
    int64_t ld_post;
    const float* comp;      // [N, ld_comp] per-Gaussian llhs, or null: single-Gaussian pdfs
    int64_t ld_comp;
    const float* pdf_llh;   // [N, ld_pdf] (with comp; same offset form as comp)
    int64_t ld_pdf;
    const int* pdf_of;      // [M] pdf of every Gaussian (with comp)
    const float* grad_out;  // [N] upstream gradient of the per-frame values, or null = ones
    const __half* img;
    const float* inv_scale;
    int M, n_chunks;
    float wexp;             // w is carried as w 2^wexp
    float* grad;            // [N, D]
    int64_t n_tiles;
};

struct BwBarriers {
    uint64_t b_full[STAGES], b_empty[STAGES];
    uint64_t a_full[NAB], a_empty[NAB];
    uint64_t d_full[2], d_empty[2];
    uint32_t tmem_base;
    uint32_t pad[3];
};

template <int KPB>
__global__ void __launch_bounds__(THREADS, 1) emission_bwd_kernel(BwArgs a) {
    constexpr int HALF_BYTES = KPB * GC * 2;           // one half (hi or lo) of an image chunk
    constexpr int STAGE = 2 * HALF_BYTES;
    constexpr int LDS = KPB + 1;                       // staging row (floats), odd: conflict-free column reads
    constexpr uint32_t COL_A = 0, COL_D = NAB * GC;
    constexpr int NCH = KPB / 4, MYCH = NCH / 2;       // 4-column chunks of the accumulator; per thread
    static_assert(COL_D + 2 * KPB <= 512 && NCH % 2 == 0, "tensor memory");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* ring = smem_raw;
    float* stage_s = reinterpret_cast<float*>(smem_raw + (size_t)STAGES * STAGE);
    BwBarriers* bars = reinterpret_cast<BwBarriers*>(stage_s + FM * LDS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&bars->b_full[i], 1);
            mbar_init(&bars->b_empty[i], 1);
        }
        for (int i = 0; i < NAB; ++i) {
            mbar_init(&bars->a_full[i], WORKERS);
            mbar_init(&bars->a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->d_full[i], 1);
            mbar_init(&bars->d_empty[i], WORKERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const int64_t my_tiles = (a.n_tiles > blockIdx.x) ? (a.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int n_groups = (a.n_chunks + DRG - 1) / DRG;

    if (warp == LOAD_WARP) {
        if (elect_one()) {
            Ring r(STAGES);
            for (int64_t it = 0; it < my_tiles; ++it)
                for (int c = 0; c < a.n_chunks; ++c, r.next()) {
                    mbar_wait_relaxed(&bars->b_empty[r.pos], r.phase ^ 1, 100);
                    mbar_arrive_expect_tx(&bars->b_full[r.pos], STAGE);
                    bulk_g2s(ring + (size_t)r.pos * STAGE, a.img + (size_t)c * (STAGE / 2), STAGE, &bars->b_full[r.pos]);
                }
        }
    } else if (warp == MMA_WARP) {
        if (elect_one()) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(KPB >> 3) << 17) | ((uint32_t)(FM >> 4) << 24);
            Ring rb(STAGES), ra(NAB);
            uint32_t G = 0;                             // drain groups issued so far
            for (int64_t it = 0; it < my_tiles; ++it)
                for (int c = 0; c < a.n_chunks; ++c) {
                    const bool first = (c % DRG) == 0, last = (c % DRG) == DRG - 1 || c == a.n_chunks - 1;
                    const uint32_t dbuf = G & 1;
                    mbar_wait(&bars->b_full[rb.pos], rb.phase);
                    mbar_wait(&bars->a_full[ra.pos], ra.phase);
                    if (first) mbar_wait(&bars->d_empty[dbuf], ((G >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t base = smem_u32(ring + (size_t)rb.pos * STAGE);
                    const uint64_t dbh = make_desc(base, 128, GC * 16), dbl = make_desc(base + HALF_BYTES, 128, GC * 16);
                    const uint32_t d = tmem_base + COL_D + dbuf * KPB;
                    const uint32_t a2 = tmem_base + COL_A + (uint32_t)ra.pos * GC;
#pragma unroll
                    for (int ks = 0; ks < GC / 16; ++ks) {
                        umma_f16_ts(d, a2 + 16u * ks, dbh + 16u * ks, idesc, !(first && ks == 0));
                        umma_f16_ts(d, a2 + 16u * ks + 8u, dbh + 16u * ks, idesc, 1);     // w lo x E[T] hi
                        umma_f16_ts(d, a2 + 16u * ks, dbl + 16u * ks, idesc, 1);          // w hi x E[T] lo
                    }
                    umma_commit(&bars->b_empty[rb.pos]);
                    umma_commit(&bars->a_empty[ra.pos]);
                    if (last) {
                        umma_commit(&bars->d_full[dbuf]);
                        ++G;
                    }
                    rb.next();
                    ra.next();
                }
        }
    } else {
        const int q = warp & 3, half = warp >> 2;
        const int fl = q * 32 + lane;                   // frame inside the tile = TMEM lane
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const float wscale = exp2f(a.wexp);
        Ring ra(NAB);
        uint32_t G = 0;                                 // drain groups collected so far
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t t0 = ((int64_t)blockIdx.x + it * gridDim.x) * FM;
            const int64_t t = t0 + fl;
            const bool valid = t < a.N;
            const float* prow = a.post + (size_t)(valid ? t : 0) * a.ld_post;
            const float* crow = a.comp ? a.comp + (size_t)(valid ? t : 0) * a.ld_comp : nullptr;
            const float* lrow = a.comp ? a.pdf_llh + (size_t)(valid ? t : 0) * a.ld_pdf : nullptr;
            float sums[MYCH][4];
#pragma unroll
            for (int m = 0; m < MYCH; ++m)
#pragma unroll
                for (int i = 0; i < 4; ++i) sums[m][i] = 0.f;
            auto drain = [&]() {
                const uint32_t dbuf = G & 1;
                mbar_wait(&bars->d_full[dbuf], (G >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int m = 0; m < MYCH; ++m) {
                    float v[4];
                    tmem_ld4(tmem_base + lane_addr + COL_D + dbuf * KPB + (uint32_t)((half + 2 * m) * 4), v);
#pragma unroll
                    for (int i = 0; i < 4; ++i) sums[m][i] += v[i];
                }
                tc_fence_before();
                mbar_arrive(&bars->d_empty[dbuf]);
                ++G;
            };
            for (int c = 0; c < a.n_chunks; ++c, ra.next()) {
                mbar_wait(&bars->a_empty[ra.pos], ra.phase ^ 1);     // the MMAs of the chunk NAB back have read the buffer
                tc_fence_after();
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    const int jb = c * GC + half * 32 + sub * 16;
                    float v[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = 0.f;
                    if (valid) {
                        if (crow == nullptr) {
                            if (jb + 16 <= a.M && ((a.ld_post | jb) & 3) == 0 && ((uintptr_t)a.post & 15) == 0) {
#pragma unroll
                                for (int e4 = 0; e4 < 4; ++e4) {
                                    const float4 p = __ldg(reinterpret_cast<const float4*>(prow + jb) + e4);
                                    v[4 * e4] = p.x; v[4 * e4 + 1] = p.y; v[4 * e4 + 2] = p.z; v[4 * e4 + 3] = p.w;
                                }
                            } else {
#pragma unroll
                                for (int e = 0; e < 16; ++e)
                                    if (jb + e < a.M) v[e] = __ldg(prow + jb + e);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                const int j = jb + e;
                                if (j < a.M) {
                                    const int k = __ldg(a.pdf_of + j);
                                    const float p = __ldg(prow + k);
                                    // responsibilities from the stored llhs (mixtureset.py:100-104): exp(comp - pdf)
                                    v[e] = (p > 0.f) ? p * __expf(__ldg(crow + j) - __ldg(lrow + k)) : 0.f;
                                }
                            }
                        }
                    }
                    uint32_t out[16];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        float w[4], wh[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            w[e] = v[4 * h + e] * wscale;
                            wh[e] = h_rn(w[e]);
                        }
                        out[2 * h] = pack_h2(wh[0], wh[1]);
                        out[2 * h + 1] = pack_h2(wh[2], wh[3]);
                        out[8 + 2 * h] = pack_h2(w[0] - wh[0], w[1] - wh[1]);
                        out[8 + 2 * h + 1] = pack_h2(w[2] - wh[2], w[3] - wh[3]);
                    }
                    tmem_st16(tmem_base + lane_addr + COL_A + (uint32_t)(ra.pos * GC + half * 32 + sub * 16), out);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bars->a_full[ra.pos]);
                // the group that ended with the previous chunk is collected one chunk late (its MMAs are in flight now)
                if (c > 0 && (c % DRG) == 0) drain();
            }
            drain();                                    // the last group of the tile
            (void)n_groups;
            // raw sums -> shared memory, then the gradient rows coalesced
            asm volatile("bar.sync 1, 256;" ::: "memory");      // the previous tile's readers are done with the staging
#pragma unroll
            for (int m = 0; m < MYCH; ++m)
#pragma unroll
                for (int i = 0; i < 4; ++i) stage_s[fl * LDS + (half + 2 * m) * 4 + i] = sums[m][i];
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int D = a.D;
            const float unscale = exp2f(-a.wexp);
            const int64_t rows = min((int64_t)FM, a.N - t0);
            for (int idx = tid; idx < (int)rows * D; idx += WORKERS) {
                const int r = idx / D, d = idx - r * D;
                const float s1 = stage_s[r * LDS + d] * __ldg(a.inv_scale + d);
                const float s2 = stage_s[r * LDS + D + d] * __ldg(a.inv_scale + D + d);
                const float go = a.grad_out ? __ldg(a.grad_out + t0 + r) : 1.f;
                const float x = __ldg(a.X + (size_t)(t0 + r) * D + d);
                a.grad[(size_t)(t0 + r) * D + d] = go * unscale * (s1 - x * s2);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int KPB>
static int launch(const BwArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)STAGES * (2 * KPB * GC * 2) + (size_t)FM * (KPB + 1) * 4 + sizeof(BwBarriers) + 1024;
    auto kern = emission_bwd_kernel<KPB>;
    static bool configured = false;
    if (!configured) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int grid = (int)std::min<int64_t>(a.n_tiles, kNumSMs);
    kern<<<grid, THREADS, smem, st>>>(a);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

}  // namespace bwd
}  // namespace beer

using namespace beer;

extern "C" {

int beer_emission_bwd_supported(int M, int D) {
    const int kpb = bwd::kpb_of(D);
    return (M >= 1 && D >= 1 && kpb <= 128) ? 1 : 0;
}

int64_t beer_emission_bwd_image_bytes(int M, int D) {
    const int64_t n_chunks = (M + bwd::GC - 1) / bwd::GC;
    return n_chunks * 2 * bwd::kpb_of(D) * bwd::GC * 2;
}

int beer_emission_bwd_pack(const float* exp_stats, int M, int D, int64_t ld, void* image, float* inv_scale,
                           uint32_t* colmax_scratch, void* stream) {
    if (!exp_stats || !image || !inv_scale || !colmax_scratch || ld < 2 * D) return BEER_ERR_ARG;
    if (!beer_emission_bwd_supported(M, D)) return BEER_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    BEER_CUDA_TRY(cudaMemsetAsync(colmax_scratch, 0, sizeof(uint32_t) * 2 * D, st));
    bwd::colmax_kernel<<<std::min(M, 4 * kNumSMs), 128, 0, st>>>(exp_stats, M, ld, 2 * D, colmax_scratch);
    BEER_LAUNCH_CHECK();
    const int n_chunks = (M + bwd::GC - 1) / bwd::GC;
    bwd::pack_kernel<<<n_chunks, 256, 0, st>>>(exp_stats, M, ld, D, bwd::kpb_of(D), colmax_scratch, (__half*)image, inv_scale);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_emission_llh_bwd(const float* X, int64_t N, int D, const void* image, const float* inv_scale, int M,
                          const float* pdf_post, int64_t ld_post, const float* comp_llh, int64_t ld_comp,
                          const float* pdf_llh, int64_t ld_pdf, const int* pdf_of, const float* grad_out, float scale,
                          float* grad_X, void* stream) {
    if (!X || !image || !inv_scale || !pdf_post || !grad_X || N < 0) return BEER_ERR_ARG;
    if (comp_llh != nullptr && (!pdf_llh || !pdf_of || ld_comp < M)) return BEER_ERR_ARG;
    if (comp_llh == nullptr && ld_post < M) return BEER_ERR_ARG;
    if (!beer_emission_bwd_supported(M, D)) return BEER_ERR_UNSUPPORTED;
    if (N == 0) return BEER_OK;
    bwd::BwArgs a;
    a.X = X; a.N = N; a.D = D; a.post = pdf_post; a.ld_post = ld_post; a.comp = comp_llh; a.ld_comp = ld_comp;
    a.pdf_llh = pdf_llh; a.ld_pdf = ld_pdf; a.pdf_of = pdf_of; a.grad_out = grad_out; a.img = (const __half*)image;
    a.inv_scale = inv_scale; a.M = M; a.n_chunks = (M + bwd::GC - 1) / bwd::GC; a.grad = grad_X;
    a.n_tiles = (N + bwd::FM - 1) / bwd::FM;
    int e = 14;                       // the posteriors carry `scale`: keep w 2^wexp <= 2^14 (fp16 overflows at 2^16)
    if (scale > 1.f) e -= (int)ceilf(log2f(scale));
    a.wexp = (float)e;
    cudaStream_t st = (cudaStream_t)stream;
    switch (bwd::kpb_of(D)) {
        case 16: return bwd::launch<16>(a, st);
        case 32: return bwd::launch<32>(a, st);
        case 48: return bwd::launch<48>(a, st);
        case 64: return bwd::launch<64>(a, st);
        case 80: return bwd::launch<80>(a, st);
        case 96: return bwd::launch<96>(a, st);
        case 112: return bwd::launch<112>(a, st);
        case 128: return bwd::launch<128>(a, st);
    }
    return BEER_ERR_UNSUPPORTED;
}

}  // extern "C"
