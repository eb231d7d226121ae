// KA on the 5th-generation tensor cores (tcgen05 + TMEM), 3xTF32 split for fp32 accuracy.
//
//   D[frame, gaussian] = [x, -x^2/2]_hi . W_hi + [x, -x^2/2]_lo . W_hi + [x, -x^2/2]_hi . W_lo
//
// One CTA = 128 frames (UMMA M) x chunks of <= 128 Gaussians (UMMA N), K = 2D features.
// Operands live in shared memory in the canonical no-swizzle K-major core-matrix layout
// ([row/8][k/4][row%8][k%4], 128-byte core matrices).  The weight image is packed once per VB
// iteration by beer_emission_tc_pack and streamed in with cp.async.bulk (TMA, 1-D); the
// statistics tile is built by the worker warps from X (hi = rn_tf32(x), lo = x - hi).
// Accumulators are fp32 in TMEM (double buffered); the epilogue reads them back with
// tcgen05.ld, adds the bias, takes the log-sum-exp over the C components of each pdf and
// stores the offset-form llh.
//
// Warp roles: warps 0-7 = workers (tile builder: two threads per frame row; epilogue: one frame
// row per thread, the two warps of a TMEM lane quarter split the columns), warp 8 = MMA issuer
// (one elected lane) + TMEM allocator, warp 9 = weight-image producer.
//
// Reference semantics: beer/dists/normalgamma.py:55-59, beer/models/mixtureset.py:85-98.
#include <cuda.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/beer_b200.h"

namespace beer {
namespace tc {

constexpr int FR = 128;        // frames per tile (UMMA M)
constexpr int NB_MAX = 128;    // Gaussians per chunk (UMMA N)
constexpr int WORKERS = 256;          // 8 worker warps: two per scheduler (one per scheduler left them latency-bound)
constexpr int MMA_WARP = WORKERS / 32, LOAD_WARP = MMA_WARP + 1;
constexpr int THREADS = WORKERS + 64;
// Streamed chunks: the bias of chunk i rides with its weight image into slot i % BIAS_RING.  Slot reuse
// (chunk i + 4) waits for the MMAs of chunk i + 2, which wait for the epilogue of chunk i (TMEM buffer).
constexpr int BIAS_RING = 4;
constexpr int CBOX = 32;       // columns of one staged per-Gaussian llh box: 128-byte rows, 128B-swizzled (TMA store)

using namespace tcu;

// offset (in floats) of element (row, k) inside a [rows x Kd] core-matrix image
__host__ __device__ __forceinline__ int img_off(int row, int k, int Kd) {
    return (row >> 3) * (Kd * 8) + (k >> 2) * 32 + (row & 7) * 4 + (k & 3);
}

struct Args {
    const float* X;
    int64_t N;
    const float* img;    // [n_chunks][2][NB * Kd] packed weights (hi image, lo image)
    const float* bias;   // [n_chunks * NB], zero padded
    const float* ref;    // [D + 1]
    int M, C, Kp, NB, n_chunks, stages;
    int logC;            // log2(C)
    int staged;          // llh tile staged in shared memory and written with one bulk store
    int cstaged;         // streamed chunks: per-Gaussian llh tile staged in shared memory, rows bulk-stored
    float* pdf_llh;
    int64_t ld;
    float* comp_llh;
    float* frame_ref;
};

struct Barriers {
    uint64_t a_ready, a_free;
    uint64_t b_full[2], b_empty[2];
    uint64_t t_full[2], t_empty[2];
    uint32_t tmem_base;
    uint32_t pad[3];     // keeps what follows (the staged llh tile) 16-byte aligned
};

template <int D4>
__global__ void __launch_bounds__(THREADS, 1) emission_tc_kernel(Args a, const __grid_constant__ CUtensorMap cmap) {
    constexpr int D = 4 * D4, Kd = 2 * D, KSTEPS = Kd / 8;
    constexpr uint32_t LBO = 128, SBO = (Kd / 4) * 128;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    float* A_hi = reinterpret_cast<float*>(smem_raw);
    float* A_lo = A_hi + FR * Kd;
    float* Bst = A_lo + FR * Kd;                           // stages x [hi image | lo image]
    const int b_stage_floats = 2 * a.NB * Kd;
    float* s_bias = Bst + (size_t)a.stages * b_stage_floats;  // [NB] resident, or a BIAS_RING-deep ring when streaming
    const int bias_floats = (a.stages == 1) ? a.NB : BIAS_RING * a.NB;
    float* s_ref = s_bias + bias_floats;                   // [D + 1]
    Barriers* bars = reinterpret_cast<Barriers*>(s_ref + ((D + 1 + 3) & ~3));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = (a.N + FR - 1) / FR;
    const uint32_t tmem_cols = (2 * a.NB <= 32) ? 32 : (2 * a.NB <= 64 ? 64 : (2 * a.NB <= 128 ? 128 : 256));

    if (tid == 0) {
        mbar_init(&bars->a_ready, WORKERS);
        mbar_init(&bars->a_free, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->b_full[i], 1);
            mbar_init(&bars->b_empty[i], 1);
            mbar_init(&bars->t_full[i], 1);
            mbar_init(&bars->t_empty[i], WORKERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) tmem_alloc(&bars->tmem_base, tmem_cols);
    if (a.stages == 1)
        for (int i = tid; i < bias_floats; i += THREADS) s_bias[i] = a.bias[i];
    for (int i = tid; i <= D; i += THREADS) s_ref[i] = a.ref[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == LOAD_WARP) {
        // ------------------------- weight-image producer -------------------------
        if (elect_one()) {
            const uint32_t bytes = (uint32_t)b_stage_floats * 4u;
            if (a.stages == 1) {
                // a single chunk stays resident for the whole kernel
                mbar_arrive_expect_tx(&bars->b_full[0], bytes);
                bulk_g2s(Bst, a.img, bytes, &bars->b_full[0]);
            } else {
                uint32_t it = 0;
                for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                    for (int c = 0; c < a.n_chunks; ++c, ++it) {
                        const int st = it & 1;
                        mbar_wait(&bars->b_empty[st], ((it >> 1) & 1) ^ 1);
                        mbar_arrive_expect_tx(&bars->b_full[st], bytes + (uint32_t)a.NB * 4u);
                        bulk_g2s(Bst + (size_t)st * b_stage_floats, a.img + (size_t)c * b_stage_floats, bytes,
                                 &bars->b_full[st]);
                        bulk_g2s(s_bias + (it & (BIAS_RING - 1)) * a.NB, a.bias + (size_t)c * a.NB, (uint32_t)a.NB * 4u,
                                 &bars->b_full[st]);
                    }
            }
        }
    } else if (warp == MMA_WARP) {
        // ------------------------------ MMA issuer -------------------------------
        if (elect_one()) {
            // instruction descriptor: D=f32, A=B=tf32, both K-major, N = NB, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.NB >> 3) << 17) |
                                   ((uint32_t)(FR >> 4) << 24);
            const uint32_t a_hi = smem_u32(A_hi), a_lo = smem_u32(A_lo);
            uint32_t it = 0, tile_it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
                mbar_wait(&bars->a_ready, tile_it & 1);
                for (int c = 0; c < a.n_chunks; ++c, ++it) {
                    const int st = (a.stages == 1) ? 0 : (it & 1);
                    const int buf = it & 1;
                    if (a.stages == 1) {
                        if (it == 0) mbar_wait(&bars->b_full[0], 0);
                    } else {
                        mbar_wait(&bars->b_full[st], (it >> 1) & 1);
                    }
                    mbar_wait(&bars->t_empty[buf], ((it >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t b_hi = smem_u32(Bst + (size_t)st * b_stage_floats);
                    const uint32_t b_lo = b_hi + (uint32_t)a.NB * Kd * 4u;
                    const uint32_t d_tmem = tmem_base + (uint32_t)buf * (uint32_t)a.NB;
#pragma unroll 1
                    for (int s = 0; s < KSTEPS; ++s) {
                        const uint32_t ko = (uint32_t)s * 256u;
                        const uint64_t dah = make_desc(a_hi + ko, LBO, SBO), dal = make_desc(a_lo + ko, LBO, SBO);
                        const uint64_t dbh = make_desc(b_hi + ko, LBO, SBO), dbl = make_desc(b_lo + ko, LBO, SBO);
                        umma_tf32(d_tmem, dah, dbh, idesc, s > 0);
                        umma_tf32(d_tmem, dal, dbh, idesc, 1);
                        umma_tf32(d_tmem, dah, dbl, idesc, 1);
                    }
                    umma_commit(&bars->t_full[buf]);
                    if (a.stages > 1) umma_commit(&bars->b_empty[st]);
                }
                umma_commit(&bars->a_free);
            }
        }
    } else {
        // ------------------------ workers: build tile, epilogue ------------------
        // build: lanes l and l ^ 8 share a frame row (interleaved 16-byte chunks of x);
        // epilogue: TMEM lane = frame row 32 (warp % 4) + lane, warps w and w + 4 split the columns
        const int rb = (tid & 7) + 8 * (tid >> 4), hb = (tid >> 3) & 1;   // 8 lanes = 8 rows: conflict-free stores
        const int r = (warp & 3) * 32 + lane, he = warp >> 2;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const int rbase = (rb >> 3) * (Kd * 8) + (rb & 7) * 4;
        constexpr int XC = (D4 + 1) / 2;     // x chunks per builder thread
        float* s_out = a.staged ? reinterpret_cast<float*>(bars + 1) : nullptr;   // [FR][Kp] llh tile

        auto load_x = [&](int64_t tile, float4 (&xv)[XC]) {
            const int64_t t = tile * FR + rb;
            const float4* xrow = reinterpret_cast<const float4*>(a.X + (size_t)(t < a.N ? t : 0) * D);
#pragma unroll
            for (int i = 0; i < XC; ++i) {
                const int c = 2 * i + hb;
                xv[i] = (c < D4) ? __ldg(xrow + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        // statistics tile [x | -x^2/2] of this thread's frame, hi / lo split, + the per-frame constant
        auto build = [&](int64_t tile, const float4 (&xv)[XC]) {
            const int64_t t = tile * FR + rb;
            const bool valid = t < a.N;
            float rt = 0.f;
#pragma unroll
            for (int i = 0; i < XC; ++i) {
                const int c = 2 * i + hb;
                if (c >= D4) continue;
                const float x[4] = {valid ? xv[i].x : 0.f, valid ? xv[i].y : 0.f, valid ? xv[i].z : 0.f,
                                    valid ? xv[i].w : 0.f};
                float h[4], l[4], qh[4], ql[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    h[e] = tf32_rn(x[e]);
                    l[e] = x[e] - h[e];
                    const float q = -0.5f * x[e] * x[e];
                    qh[e] = tf32_rn(q);
                    ql[e] = q - qh[e];
                    rt = fmaf(q, s_ref[4 * c + e], rt);
                }
                *reinterpret_cast<float4*>(A_hi + rbase + c * 32) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(A_lo + rbase + c * 32) = make_float4(l[0], l[1], l[2], l[3]);
                *reinterpret_cast<float4*>(A_hi + rbase + (D4 + c) * 32) = make_float4(qh[0], qh[1], qh[2], qh[3]);
                *reinterpret_cast<float4*>(A_lo + rbase + (D4 + c) * 32) = make_float4(ql[0], ql[1], ql[2], ql[3]);
            }
            rt += __shfl_xor_sync(0xffffffffu, rt, 8);      // the two halves of the row
            if (valid && hb == 0 && a.frame_ref != nullptr) a.frame_ref[t] = rt + s_ref[D];
        };
        // TMEM -> registers -> bias, log-sum-exp over the C components -> llh (chunk c of a tile)
        auto epilogue = [&](int64_t tile, uint32_t it, int c) {
            const int64_t t = tile * FR + r;
            const bool valid = t < a.N;
            const int buf = it & 1;
            if (a.staged) {
                // the previous tile's bulk store must have read the staging tile before it is rewritten
                if (warp == 0 && elect_one()) bulk_wait_read0();
                asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory");
            }
            mbar_wait(&bars->t_full[buf], (it >> 1) & 1);
            tc_fence_after();
            const int g0 = c * a.NB;
            const float* bias_c = s_bias + ((a.stages == 1) ? 0 : (int)(it & (BIAS_RING - 1)) * a.NB);
            const uint32_t taddr = tmem_base + lane_addr + (uint32_t)buf * (uint32_t)a.NB;
            const int nch = a.NB >> 4, ch0 = he ? (nch + 1) / 2 : 0, ch1 = he ? nch : (nch + 1) / 2;
            for (int p = ch0 * 16; p < ch1 * 16; p += 16) {
                float v[16];
                tmem_ld16(taddr + (uint32_t)p, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += bias_c[p + i];
                if (valid || a.staged) {
                    const int g = g0 + p;
                    if (valid && a.comp_llh != nullptr) {
                        float* dst = a.comp_llh + (size_t)t * a.M + g;
                        if (g + 16 <= a.M && (a.M & 3) == 0) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                reinterpret_cast<float4*>(dst)[q] =
                                    make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (g + i < a.M) dst[i] = v[i];
                        }
                    }
                    // log-sum-exp over the C components of each pdf (C divides 16)
                    const int C = a.C;
                    float o[16];
                    int no;
                    if (C == 1) {
                        no = 16;
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = v[i];
                    } else {
                        no = 16 / C;
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = 0.f;
#pragma unroll
                        for (int lg = 1; lg <= 4; ++lg) {
                            if (C == (1 << lg)) {
                                const int CC = 1 << lg;
#pragma unroll
                                for (int k = 0; k < 16 / CC; ++k) {
                                    float m = v[k * CC];
#pragma unroll
                                    for (int j = 1; j < CC; ++j) m = fmaxf(m, v[k * CC + j]);
                                    float sm = 0.f;
#pragma unroll
                                    for (int j = 0; j < CC; ++j) sm += __expf(v[k * CC + j] - m);
                                    o[k] = m + __logf(sm);
                                }
                            }
                        }
                    }
                    const int k0 = g >> a.logC;          // C is a power of two
                    // staged: the tile [FR][Kp] is contiguous in HBM, rows go to shared memory first
                    float* dst = a.staged ? s_out + (size_t)r * a.Kp + k0 : a.pdf_llh + (size_t)t * a.ld + k0;
                    if (no == 16 && k0 + 16 <= a.Kp && ((a.staged ? a.Kp : a.ld) & 3) == 0) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            reinterpret_cast<float4*>(dst)[q] =
                                make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (i < no && k0 + i < a.Kp) dst[i] = o[i];
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&bars->t_empty[buf]);
            if (a.staged) {
                // one bulk (TMA) store of the whole tile instead of 128 strided row stores
                fence_proxy_async();
                asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory");
                if (warp == 0 && elect_one()) {
                    const int64_t rows = min((int64_t)FR, a.N - tile * FR);
                    bulk_s2g(a.pdf_llh + (size_t)tile * FR * a.Kp, s_out, (uint32_t)(rows * a.Kp * 4));
                    bulk_commit();
                }
            }
        };

        // Streamed chunks of a mixture: the per-Gaussian llh tile goes through shared memory and out as TMA tensor
        // stores (a thread's own row stores are 16-byte pieces of 128 distinct lines per instruction, and one bulk
        // row copy per frame cost ~35 cycles of TMA issue each: both bounded the kernel well below HBM).  The two
        // warps of a TMEM lane quarter own one [FR x 32] box each (128-byte rows, 128B swizzle: conflict-free
        // float4 stores); the log-sum-exp runs under the store.
        auto epilogue_cs = [&](int64_t tile, uint32_t it, int c) {
            const int64_t t = tile * FR + r;
            const bool valid = t < a.N;
            const int buf = it & 1;
            // boxes: 1024-byte aligned (the swizzle is a function of the address)
            uint8_t* s_box = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bars + 1) + 1023) & ~uintptr_t(1023)) +
                             (size_t)he * (FR * CBOX * 4);
            bool issuer = false;                      // one lane of the first warp of each column half
            if ((warp & 3) == 0) issuer = elect_one();
            if (issuer) bulk_wait_read0();            // the previous chunk's store has read the box
            asm volatile("bar.sync %0, 128;" ::"r"(1 + he) : "memory");
            mbar_wait(&bars->t_full[buf], (it >> 1) & 1);
            tc_fence_after();
            const int g0 = c * a.NB;
            // the bias slot landed with the weight image the MMAs behind t_full have read (no wait on b_full here:
            // it may already be two phases ahead, which a parity wait cannot tell from "not yet")
            const float* bias_c = s_bias + (int)(it & (BIAS_RING - 1)) * a.NB;
            const uint32_t taddr = tmem_base + lane_addr + (uint32_t)buf * (uint32_t)a.NB;
            float v[2][16];
            const int p0 = he * 32;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                tmem_ld16(taddr + (uint32_t)(p0 + 16 * h), v[h]);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[h][i] += bias_c[p0 + 16 * h + i];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<float4*>(s_box + (size_t)r * 128 + (((4 * h + q) ^ (r & 7)) << 4)) =
                        make_float4(v[h][4 * q], v[h][4 * q + 1], v[h][4 * q + 2], v[h][4 * q + 3]);
            }
            tc_fence_before();
            mbar_arrive(&bars->t_empty[buf]);
            fence_proxy_async();
            asm volatile("bar.sync %0, 128;" ::"r"(1 + he) : "memory");
            if (issuer) {
                // rows past N and columns past M are clipped by the tensor map
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&cmap),
                             "r"(smem_u32(s_box)), "r"(g0 + p0), "r"((int)(tile * FR))
                             : "memory");
                bulk_commit();
            }
            if (!valid) return;
            // log-sum-exp over the C components of each pdf
            const int C = a.C, no = 32 / C;
            float o[8];
#pragma unroll
            for (int lg = 2; lg <= 4; ++lg) {
                if (C == (1 << lg)) {
                    const int CC = 1 << lg;
#pragma unroll
                    for (int k = 0; k < 32 / CC; ++k) {
                        const float* vv = &v[0][0] + k * CC;
                        float m = vv[0];
#pragma unroll
                        for (int j = 1; j < CC; ++j) m = fmaxf(m, vv[j]);
                        float sm = 0.f;
#pragma unroll
                        for (int j = 0; j < CC; ++j) sm += __expf(vv[j] - m);
                        o[k] = m + __logf(sm);
                    }
                }
            }
            const int k0 = (g0 + p0) >> a.logC;
            float* dst = a.pdf_llh + (size_t)t * a.ld + k0;
            if (no == 4 && k0 + 4 <= a.Kp && (a.ld & 3) == 0) {
                *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i < no && k0 + i < a.Kp) dst[i] = o[i];
            }
        };

        if (a.n_chunks == 1) {
            // Software pipeline: the epilogue of tile i-1 runs under the MMAs of tile i and the
            // frames of tile i+1 are already in flight.
            float4 xv[XC];
            int64_t tile = blockIdx.x, prev = -1;
            uint32_t it = 0;
            if (tile < n_tiles) load_x(tile, xv);
            for (; tile < n_tiles; tile += gridDim.x, ++it) {
                mbar_wait(&bars->a_free, (it & 1) ^ 1);   // the previous tile's MMAs have read A
                build(tile, xv);
                fence_proxy_async();                      // generic-proxy smem writes -> tensor core
                mbar_arrive(&bars->a_ready);
                if (tile + gridDim.x < n_tiles) load_x(tile + gridDim.x, xv);
                if (prev >= 0) epilogue(prev, it - 1, 0);
                prev = tile;
            }
            if (prev >= 0) epilogue(prev, it - 1, 0);
            if (a.staged && warp == 0 && elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        } else {
            uint32_t it = 0, tile_it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
                float4 xv[XC];
                load_x(tile, xv);
                mbar_wait(&bars->a_free, (tile_it & 1) ^ 1);
                build(tile, xv);
                fence_proxy_async();
                mbar_arrive(&bars->a_ready);
                if (a.cstaged)
                    for (int c = 0; c < a.n_chunks; ++c, ++it) epilogue_cs(tile, it, c);
                else
                    for (int c = 0; c < a.n_chunks; ++c, ++it) epilogue(tile, it, c);
            }
            if (a.cstaged && (warp & 3) == 0 && elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// Pack W [M, 2D] + bias [M] into per-chunk core-matrix images (hi / lo split).
__global__ void emission_tc_pack_kernel(const float* __restrict__ W, const float* __restrict__ bias, int M, int D,
                                        int NB, int n_chunks, float* __restrict__ img, float* __restrict__ bias_pad) {
    const int Kd = 2 * D;
    const int per_chunk = 2 * NB * Kd;
    const int64_t total = (int64_t)n_chunks * NB * Kd;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(e / (NB * Kd));
        int rem = (int)(e - (int64_t)c * NB * Kd);
        int n = rem / Kd, k = rem - n * Kd;
        int g = c * NB + n;
        float w = (g < M) ? W[(size_t)g * Kd + k] : 0.f;
        float hi = tf32_rn(w);
        float lo = tf32_rn(w - hi);
        float* base = img + (size_t)c * per_chunk;
        base[img_off(n, k, Kd)] = hi;
        base[NB * Kd + img_off(n, k, Kd)] = lo;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_chunks * NB; i += gridDim.x * blockDim.x)
        bias_pad[i] = (i < M) ? bias[i] : 0.f;
}

struct Geometry { int NB, n_chunks, stages; };

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool geometry(int M, int D, int C, Geometry* g) {
    if (D % 4 != 0) return false;
    int d4 = D / 4;
    if (!(d4 == 5 || d4 == 10 || d4 == 16 || d4 == 20)) return false;
    if (!(C == 1 || C == 2 || C == 4 || C == 8 || C == 16)) return false;
    if (M <= NB_MAX) {
        g->NB = (M + 15) / 16 * 16;
        g->n_chunks = 1;
        g->stages = 1;
    } else {
        g->NB = 64;
        g->n_chunks = (M + 63) / 64;
        g->stages = 2;
    }
    return true;
}

static size_t smem_bytes(int D, const Geometry& g, int staged_kp = 0) {
    int Kd = 2 * D;
    size_t f = (size_t)2 * FR * Kd + (size_t)g.stages * 2 * g.NB * Kd +
               (size_t)(g.stages == 1 ? g.NB : BIAS_RING * g.NB) + ((D + 1 + 3) & ~3) + (size_t)FR * staged_kp;
    return f * 4 + sizeof(Barriers) + 1024;
}

template <int D4>
static int launch(const Args& a, const Geometry& g, cudaStream_t st) {
    size_t smem = smem_bytes(4 * D4, g, a.staged ? a.Kp : (a.cstaged ? 2 * CBOX + 8 : 0));   // + 1 KB alignment slack
    if (smem > 227 * 1024) return BEER_ERR_UNSUPPORTED;
    static bool attr_set = false;
    if (!attr_set) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(emission_tc_kernel<D4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           227 * 1024));
        attr_set = true;
    }
    int64_t n_tiles = (a.N + FR - 1) / FR;
    int grid = (int)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
    CUtensorMap cmap;
    memset(&cmap, 0, sizeof(cmap));
    if (a.cstaged) {
        // comp_llh [N, M] fp32 row-major, boxes of [FR rows x CBOX columns], 128B swizzle in shared memory
        static EncodeTiled encode = nullptr;
        if (encode == nullptr) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            BEER_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
            if (fn == nullptr || q != cudaDriverEntryPointSuccess) return BEER_ERR_UNSUPPORTED;
            encode = reinterpret_cast<EncodeTiled>(fn);
        }
        const cuuint64_t gdim[2] = {(cuuint64_t)a.M, (cuuint64_t)a.N};
        const cuuint64_t gstride[1] = {(cuuint64_t)a.M * 4};
        const cuuint32_t box[2] = {CBOX, FR};
        const cuuint32_t estr[2] = {1, 1};
        if (encode(&cmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.comp_llh, gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return BEER_ERR_UNSUPPORTED;
    }
    emission_tc_kernel<D4><<<grid, THREADS, smem, st>>>(a, cmap);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

}  // namespace tc
}  // namespace beer

using namespace beer;

extern "C" {

int beer_emission_tc_supported(int M, int D, int C) {
    tc::Geometry g;
    if (M <= 0 || D <= 0 || C <= 0 || M % C != 0) return 0;
    if (!tc::geometry(M, D, C, &g)) return 0;
    return tc::smem_bytes(D, g) <= 227 * 1024 ? 1 : 0;
}

int64_t beer_emission_tc_image_floats(int M, int D, int C) {
    tc::Geometry g;
    if (M <= 0 || D <= 0 || C <= 0 || !tc::geometry(M, D, C, &g)) return BEER_ERR_UNSUPPORTED;
    return (int64_t)g.n_chunks * (2 * g.NB * 2 * D + g.NB);
}

int beer_emission_tc_pack(const float* W, const float* bias, int M, int D, int C, float* image, void* stream) {
    tc::Geometry g;
    if (!W || !bias || !image || M <= 0 || D <= 0 || !tc::geometry(M, D, C, &g)) return BEER_ERR_UNSUPPORTED;
    float* bias_pad = image + (size_t)g.n_chunks * 2 * g.NB * 2 * D;
    int64_t total = (int64_t)g.n_chunks * g.NB * 2 * D;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    tc::emission_tc_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, bias, M, D, g.NB, g.n_chunks, image,
                                                                          bias_pad);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_emission_llh_tc(const float* X, int64_t N, int D, const float* image, const float* ref, int M, int C,
                         float* pdf_llh, int64_t ld_pdf, float* comp_llh, float* frame_ref, void* stream) {
    tc::Geometry g;
    if (!X || !image || !ref || !pdf_llh || N < 0 || M <= 0 || C <= 0 || M % C != 0) return BEER_ERR_ARG;
    if (!tc::geometry(M, D, C, &g)) return BEER_ERR_UNSUPPORTED;
    if (ld_pdf < M / C) return BEER_ERR_ARG;
    if (((uintptr_t)X & 15) != 0 || ((uintptr_t)image & 15) != 0) return BEER_ERR_ARG;
    if (N == 0) return BEER_OK;
    tc::Args a;
    a.X = X; a.N = N; a.img = image;
    a.bias = image + (size_t)g.n_chunks * 2 * g.NB * 2 * D;
    a.ref = ref; a.M = M; a.C = C; a.Kp = M / C; a.NB = g.NB; a.n_chunks = g.n_chunks; a.stages = g.stages;
    a.pdf_llh = pdf_llh; a.ld = ld_pdf; a.comp_llh = comp_llh; a.frame_ref = frame_ref;
    a.logC = 0;
    while ((1 << a.logC) < C) ++a.logC;
    // stage + bulk-store the llh tile when its rows are contiguous in HBM and it fits next to the operands
    a.staged = (g.n_chunks == 1 && ld_pdf == a.Kp && (a.Kp & 3) == 0 && ((uintptr_t)pdf_llh & 15) == 0 &&
                tc::smem_bytes(D, g, a.Kp) <= 227 * 1024) ? 1 : 0;
    // streamed chunks of a mixture (C = 4, 8, 16): per-Gaussian llh rows leave through shared memory + bulk copies
    a.cstaged = (g.n_chunks > 1 && g.NB == 64 && comp_llh != nullptr && (C == 4 || C == 8 || C == 16) && (M & 3) == 0 &&
                 ((uintptr_t)comp_llh & 15) == 0 && tc::smem_bytes(D, g, 2 * tc::CBOX + 8) <= 227 * 1024) ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (D / 4) {
        case 5: return tc::launch<5>(a, g, st);
        case 10: return tc::launch<10>(a, g, st);
        case 16: return tc::launch<16>(a, g, st);
        case 20: return tc::launch<20>(a, g, st);
    }
    return BEER_ERR_UNSUPPORTED;
}

}  // extern "C"
