// KC on the 5th-generation tensor cores (tcgen05 + TMEM), 3xTF32 split for fp32 accuracy.
//
//   acc[j, :] += sum_t w_tj [x_t, x_t^2, 1]        w_tj = g_t,pdf(j) * r_tj
//
// A tall-skinny (M x T).(T x (2D+1)) contraction whose reduction dimension is time: both
// operands are stored frame-major in HBM and are transposed while they are staged.  One CTA owns a tile of
// 128 Gaussians (UMMA M) and a contiguous range of frames.  The fp32 accumulation in TMEM
// truncates (measured: sums come out ~1e-7 low per 24 accumulations), so a TMEM accumulator
// only ever holds ONE stage (64 frames); the stage results are drained into round-to-nearest
// register sums (double-buffered TMEM, so the drain overlaps the next stage's MMAs) and the
// CTA flushes once with fp64 atomics.
//
//   D[gauss, feat] += W_hi^T S_hi + W_lo^T S_hi + W_hi^T S_lo     (K = 8 frames per MMA)
//
// Warp roles: warps 0-7 = producers (global -> registers -> hi/lo split -> shared memory in
// the canonical no-swizzle K-major core-matrix layout, i.e. transposed on the way: tf32
// operands with the MN-major descriptor bit came back as zeros on B200; 2-stage ring, loads of
// the next stage are in flight while the current one is split) and epilogue; warp 8 = MMA
// issuer + TMEM allocator.  The count column (sum_t w) rides along as a constant-one feature.
//
// Reference semantics: beer/models/mixtureset.py:100-112, beer/models/normalset.py:121-123.
#include <cstdlib>
#include <type_traits>
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/beer_b200.h"

namespace beer {
namespace kctc {

using namespace tcu;

constexpr int GM = 128;         // Gaussians per tile (UMMA M)
constexpr int PRODUCERS = 256;
constexpr int THREADS = PRODUCERS + 32;
constexpr int MAX_STAGES = 4;

struct Args {
    const float* X;
    int64_t N;
    const float* pdf_post;
    int64_t ld_post;
    const float* pdf_llh;
    int64_t ld_pdf;
    const float* comp_llh;
    const int32_t* comp_off;
    const int32_t* pdf_ids;     // path mode: pdf of every frame ([N rounded up to 4] int32), weight = path_scale
    float path_scale;
    int Kp, M, C;
    double* acc;
    int n_gtiles;
    int64_t frames_per_cta;
};

struct Barriers {
    uint64_t full[MAX_STAGES], empty[MAX_STAGES];   // shared-memory stages: producers <-> MMA
    uint64_t tfull[2], tempty[2];           // TMEM accumulator buffers: MMA <-> drain
    uint32_t tmem_base;
};

// D4 = D / 4; KF = frames per stage (multiple of 8); ST = shared-memory stages
template <int D4, int KF, int ST>
struct Cfg {
    static constexpr int STAGES = ST;
    static constexpr int D = 4 * D4;
    static constexpr int NB = (2 * D + 1 + 15) / 16 * 16;  // [x | x^2 | 1 0 0 ...], UMMA N % 16 == 0 at M = 128
    static constexpr int KG = KF / 8;            // K-groups (MMAs per pass) per stage
    static constexpr int A_FLOATS = KF * GM;     // one A image (hi or lo)
    static constexpr int B_FLOATS = KF * NB;
    static constexpr int STAGE_FLOATS = 2 * A_FLOATS + 2 * B_FLOATS;
    static constexpr int BU = ((KF / 4) * D + PRODUCERS - 1) / PRODUCERS;  // B units (4 frames x 1 feature) per thread
    static constexpr int NCH = NB / 16;          // 16-column chunks of the accumulator
    static constexpr int MYCH = (NCH + 1) / 2;   // chunks drained by one thread (two warps share a row)
    static constexpr uint32_t TMEM_COLS = 2 * NB <= 64 ? 64 : (2 * NB <= 128 ? 128 : (2 * NB <= 256 ? 256 : 512));
    // running sums across drains: fp64 registers when they fit, fp32 (round-to-nearest) otherwise
    using acc_t = typename std::conditional<(MYCH <= 3), double, float>::type;
    // stages accumulated in TMEM between drains: 64 frames = 24 truncating accumulations (the
    // float -> double conversions of the drain issue on the quarter-rate XU pipe)
    static constexpr int DR = 64 / KF;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_FLOATS * 4 + GM * 4 + sizeof(Barriers) + 1024;
};

template <int D4, int KF, int ST, bool MIX>
__global__ void __launch_bounds__(THREADS, 1) accumulate_tc_kernel(Args a) {
    using C = Cfg<D4, KF, ST>;
    constexpr int D = C::D, NB = C::NB, KG = C::KG, STAGES = ST;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    float* stage_base = reinterpret_cast<float*>(smem_raw);
    int* s_pdf = reinterpret_cast<int*>(stage_base + (size_t)STAGES * C::STAGE_FLOATS);
    Barriers* bars = reinterpret_cast<Barriers*>(s_pdf + GM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gtile = blockIdx.x % a.n_gtiles;
    const int64_t chunk = blockIdx.x / a.n_gtiles;
    const int g0 = gtile * GM;
    const int ng = min(GM, a.M - g0);
    const int64_t f_begin = chunk * a.frames_per_cta;
    const int64_t f_end = min(a.N, f_begin + a.frames_per_cta);
    const int n_tiles = (f_end > f_begin) ? (int)((f_end - f_begin + KF - 1) / KF) : 0;

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&bars->full[i], PRODUCERS);
            mbar_init(&bars->empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->tfull[i], 1);
            mbar_init(&bars->tempty[i], PRODUCERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == PRODUCERS / 32) tmem_alloc(&bars->tmem_base, C::TMEM_COLS);
    // zero every image once: padded Gaussians / features stay zero for the whole kernel
    for (int i = tid; i < STAGES * C::STAGE_FLOATS / 4; i += THREADS)
        reinterpret_cast<float4*>(stage_base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // pdf of every Gaussian of the tile
    for (int g = tid; g < GM; g += THREADS) {
        int k = 0;
        if (g < ng) {
            int j = g0 + g;
            if (a.comp_off == nullptr) {
                k = j / a.C;
            } else {
                int lo = 0, hi = a.Kp;  // last k with comp_off[k] <= j
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (a.comp_off[mid] <= j) lo = mid; else hi = mid;
                }
                k = lo;
            }
        }
        s_pdf[g] = k;
    }
    __syncthreads();
    // the constant-one feature (count column) of the hi images
    for (int i = tid; i < STAGES * KF; i += THREADS) {
        int st = i / KF, f = i - st * KF;
        float* b_hi = stage_base + (size_t)st * C::STAGE_FLOATS + 2 * C::A_FLOATS;
        b_hi[((2 * D) >> 3) * (KF * 8) + (f >> 2) * 32 + ((2 * D) & 7) * 4 + (f & 3)] = 1.f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == PRODUCERS / 32) {
        // ------------------------------ MMA issuer -------------------------------
        if (n_tiles > 0 && elect_one()) {
            // D = f32, A = B = tf32, both K-major (K = frames), N = NB, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) |
                                   ((uint32_t)(GM >> 4) << 24);
            constexpr uint32_t LBO = 128, SBO = (KF / 4) * 128;
            for (int it = 0; it < n_tiles; ++it) {
                const int st = it % STAGES;
                const int grp = it / C::DR, buf = grp & 1;
                const bool first = it % C::DR == 0, last = (it % C::DR == C::DR - 1) || it == n_tiles - 1;
                mbar_wait(&bars->full[st], (it / STAGES) & 1);
                if (first) mbar_wait(&bars->tempty[buf], ((grp >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * NB);
                const uint32_t a_hi = smem_u32(stage_base + (size_t)st * C::STAGE_FLOATS);
                const uint32_t a_lo = a_hi + C::A_FLOATS * 4u;
                const uint32_t b_hi = a_lo + C::A_FLOATS * 4u;
                const uint32_t b_lo = b_hi + C::B_FLOATS * 4u;
#pragma unroll 1
                for (int ks = 0; ks < KG; ++ks) {
                    const uint32_t ko = (uint32_t)ks * 256u;   // 8 frames = two 16-byte core-matrix columns
                    const uint64_t dah = make_desc(a_hi + ko, LBO, SBO), dal = make_desc(a_lo + ko, LBO, SBO);
                    const uint64_t dbh = make_desc(b_hi + ko, LBO, SBO), dbl = make_desc(b_lo + ko, LBO, SBO);
                    // the TMEM accumulation truncates: keep it short (KF frames), sum tiles in registers
                    umma_tf32(d_tmem, dah, dbh, idesc, !(first && ks == 0));
                    umma_tf32(d_tmem, dal, dbh, idesc, 1);
                    umma_tf32(d_tmem, dah, dbl, idesc, 1);
                }
                umma_commit(&bars->empty[st]);
                if (last) umma_commit(&bars->tfull[buf]);
            }
        }
    } else {
        // ------------------------------ producers --------------------------------
        // The operands are frame-major in HBM but K-major (frames contiguous) for the MMA: a
        // thread gathers 4 consecutive frames of ONE Gaussian / feature (each load instruction is
        // coalesced across the warp) and writes them as one 16-byte core-matrix row.
        const int gl = tid & (GM - 1);       // Gaussian (local) of this thread's A cells
        const int fq0 = tid >> 7;            // first frame quad; quads fq0, fq0 + 2, ...
        const bool a_active = gl < ng;
        const int kpdf = s_pdf[gl];
        constexpr int AU = KF / 8;           // A units (frame quads) per thread
        const int a_row = (gl >> 3) * (KF * 8) + (gl & 7) * 4;

        float wa[AU][4];
        float ca[MIX ? AU : 1][4], la[MIX ? AU : 1][4];   // component / pdf llh (mixtures only)
        float xb[C::BU][4];

        // drain: TMEM lane quarter q (row = Gaussian), 16-column chunks half, half + 2, ...
        const int q = warp & 3, half = warp >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        using acc_t = typename C::acc_t;
        acc_t sums[C::MYCH][16];
#pragma unroll
        for (int m = 0; m < C::MYCH; ++m)
#pragma unroll
            for (int i = 0; i < 16; ++i) sums[m][i] = (acc_t)0;
        auto drain = [&](int g) {      // g = drain group (DR stages)
            const int buf = g & 1;
            mbar_wait(&bars->tfull[buf], (g >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int m = 0; m < C::MYCH; ++m) {
                const int ch = half + 2 * m;
                if (ch < C::NCH) {
                    float v[16];
                    tmem_ld16(taddr + (uint32_t)(buf * NB + ch * 16), v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) sums[m][i] += (acc_t)v[i];
                }
            }
            tc_fence_before();
            mbar_arrive(&bars->tempty[buf]);
        };

        // per-thread constants of the gathers (inactive threads read a valid cell and never store it)
        const int gl_c = a_active ? gl : 0;
        const int kpdf_c = a_active ? kpdf : 0;
        int xd[C::BU], xfq[C::BU];
        bool x_active[C::BU];
#pragma unroll
        for (int j = 0; j < C::BU; ++j) {
            const int u = tid + j * PRODUCERS;
            x_active[j] = u < (KF / 4) * D;
            xd[j] = x_active[j] ? u % D : 0;
            xfq[j] = x_active[j] ? u / D : 0;
        }

        // Loads of one stage straight into their registers, nothing consumed here, so that they all
        // stay in flight together (a select right behind each load made ptxas funnel them through
        // one temporary: one DRAM round trip per load).  Full stages carry no predicates at all.
        auto load_tile = [&](int it) {
            const int64_t t0 = f_begin + (int64_t)it * KF;
            if (t0 + KF <= f_end) {
                const float* pp = (a.pdf_post != nullptr) ? a.pdf_post + (size_t)(t0 + fq0 * 4) * a.ld_post + kpdf_c : nullptr;
                const float* pc = MIX ? a.comp_llh + (size_t)(t0 + fq0 * 4) * a.M + g0 + gl_c : nullptr;
                const float* pl = MIX ? a.pdf_llh + (size_t)(t0 + fq0 * 4) * a.ld_pdf + kpdf_c : nullptr;
#pragma unroll
                for (int j = 0; j < AU; ++j) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = 8 * j + i;   // frame offset of quad fq0 + 2j
                        wa[j][i] = (pp != nullptr) ? __ldg(pp + (size_t)r * a.ld_post) : 1.f;
                        if constexpr (MIX) {
                            ca[j][i] = __ldg(pc + (size_t)r * a.M);
                            la[j][i] = __ldg(pl + (size_t)r * a.ld_pdf);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < C::BU; ++j) {
                    const float* px = a.X + (size_t)(t0 + xfq[j] * 4) * D + xd[j];
#pragma unroll
                    for (int i = 0; i < 4; ++i) xb[j][i] = __ldg(px + i * D);
                }
                return;
            }
#pragma unroll
            for (int j = 0; j < AU; ++j) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t t = t0 + (fq0 + 2 * j) * 4 + i;
                    const bool ok = a_active && t < f_end;
                    wa[j][i] = (a.pdf_post != nullptr) ? 0.f : (ok ? 1.f : 0.f);
                    if (ok && a.pdf_post != nullptr) wa[j][i] = __ldg(a.pdf_post + (size_t)t * a.ld_post + kpdf);
                    if constexpr (MIX) {
                        ca[j][i] = 0.f;
                        la[j][i] = 0.f;
                        if (ok) {
                            ca[j][i] = __ldg(a.comp_llh + (size_t)t * a.M + g0 + gl);
                            la[j][i] = __ldg(a.pdf_llh + (size_t)t * a.ld_pdf + kpdf);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < C::BU; ++j) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t t = t0 + xfq[j] * 4 + i;
                    xb[j][i] = 0.f;
                    if (x_active[j] && t < f_end) xb[j][i] = __ldg(a.X + (size_t)t * D + xd[j]);
                }
            }
        };

        if (n_tiles > 0) load_tile(0);
        for (int it = 0; it < n_tiles; ++it) {
            const int st = it % STAGES;
            float* A_hi = stage_base + (size_t)st * C::STAGE_FLOATS;
            float* A_lo = A_hi + C::A_FLOATS;
            float* B_hi = A_lo + C::A_FLOATS;
            float* B_lo = B_hi + C::B_FLOATS;
            mbar_wait(&bars->empty[st], ((it / STAGES) & 1) ^ 1);
            if (a_active) {
#pragma unroll
                for (int j = 0; j < AU; ++j) {
                    float h[4], l[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float w = wa[j][i];
                        if constexpr (MIX) w = (w != 0.f) ? w * __expf(ca[j][i] - la[j][i]) : 0.f;
                        h[i] = tf32_rn(w);
                        l[i] = w - h[i];
                    }
                    const int off = a_row + (fq0 + 2 * j) * 32;
                    *reinterpret_cast<float4*>(A_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4*>(A_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
                }
            }
#pragma unroll
            for (int j = 0; j < C::BU; ++j) {
                if (x_active[j]) {
                    const int d = xd[j], fq = xfq[j];
                    float xh[4], xl[4], qh[4], ql[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float x = xb[j][i], q = x * x;
                        xh[i] = tf32_rn(x);
                        xl[i] = x - xh[i];
                        qh[i] = tf32_rn(q);
                        ql[i] = q - qh[i];
                    }
                    const int offx = (d >> 3) * (KF * 8) + fq * 32 + (d & 7) * 4;
                    const int offq = ((D + d) >> 3) * (KF * 8) + fq * 32 + ((D + d) & 7) * 4;
                    *reinterpret_cast<float4*>(B_hi + offx) = make_float4(xh[0], xh[1], xh[2], xh[3]);
                    *reinterpret_cast<float4*>(B_lo + offx) = make_float4(xl[0], xl[1], xl[2], xl[3]);
                    *reinterpret_cast<float4*>(B_hi + offq) = make_float4(qh[0], qh[1], qh[2], qh[3]);
                    *reinterpret_cast<float4*>(B_lo + offq) = make_float4(ql[0], ql[1], ql[2], ql[3]);
                }
            }
            fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core
            mbar_arrive(&bars->full[st]);
            if (it + 1 < n_tiles) load_tile(it + 1);
            // the previous group's MMAs have had a whole stage to finish
            if (it % C::DR == 0 && it > 0) drain(it / C::DR - 1);
        }

        // ------------------------------ epilogue ---------------------------------
        if (n_tiles > 0) {
            drain((n_tiles - 1) / C::DR);
            const int g = q * 32 + lane;
            if (g < ng) {
                const int Q = 2 * D + 2;
                double* row = a.acc + (size_t)(g0 + g) * Q;
#pragma unroll
                for (int m = 0; m < C::MYCH; ++m) {
                    const int ch = half + 2 * m;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = ch * 16 + i;
                        const double v = (double)sums[m][i];
                        if (c < 2 * D) {
                            if (v != 0.0) atomicAdd(row + c, c < D ? v : -0.5 * v);
                        } else if (c == 2 * D) {
                            if (v != 0.0) {
                                atomicAdd(row + 2 * D, -0.5 * v);
                                atomicAdd(row + 2 * D + 1, 0.5 * v);
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == PRODUCERS / 32) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

template <int D4, int KF, int ST>
static int launch(const Args& a0, cudaStream_t st) {
    using C = Cfg<D4, KF, ST>;
    static_assert(C::SMEM <= 227 * 1024, "stage too large");
    static_assert(C::NB <= 256 && C::NB % 8 == 0, "UMMA N");
    static bool attr_set = false;
    if (!attr_set) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(accumulate_tc_kernel<D4, KF, ST, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        BEER_CUDA_TRY(cudaFuncSetAttribute(accumulate_tc_kernel<D4, KF, ST, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        attr_set = true;
    }
    Args a = a0;
    a.n_gtiles = (a.M + GM - 1) / GM;
    // one CTA per SM (shared memory bound): ~3 waves when there are many Gaussian tiles
    int64_t chunks = (a.n_gtiles >= kNumSMs) ? 1 : ((a.n_gtiles > kNumSMs / 3 ? 3 * kNumSMs : kNumSMs) / a.n_gtiles);
    int64_t max_chunks = (a.N + KF - 1) / KF;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    int64_t fpc = (a.N + chunks - 1) / chunks;
    fpc = (fpc + KF - 1) / KF * KF;
    chunks = (a.N + fpc - 1) / fpc;
    a.frames_per_cta = fpc;
    int grid = (int)(chunks * a.n_gtiles);
    if (a.comp_llh != nullptr)
        accumulate_tc_kernel<D4, KF, ST, true><<<grid, THREADS, C::SMEM, st>>>(a);
    else
        accumulate_tc_kernel<D4, KF, ST, false><<<grid, THREADS, C::SMEM, st>>>(a);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

// ---------------------------------------------------------------------------
// Bulk-staged variant: one Gaussian tile (M <= 128), no mixtures, posteriors stored densely
// ([N, M], the layout KB writes).  The frames x M posterior block and the frames x D feature
// block of a stage are each ONE contiguous range in HBM, so a loader thread moves them with two
// cp.async.bulk (TMA) copies into a deep shared-memory ring: ~100 KB in flight per SM instead
// of one stage of per-thread loads.  Producers transpose / split from that ring.
//
// MIX = true: mixtures with C | 128 components per pdf and any number of Gaussian tiles.  A stage of a
// tile is 32 rows of 128 per-Gaussian llhs + 32 rows of 128 / C pdf posteriors and pdf llhs (one bulk row
// copy per frame and array, issued by the 32 lanes of the loader warp) + the contiguous feature block; the
// producers form w = post * exp(comp_llh - pdf_llh) from shared memory (mixtureset.py:100-112).
// ---------------------------------------------------------------------------
constexpr int RAW_KF = 32;          // frames per stage
constexpr int RAW_OPS = 2;          // operand (MMA) stages
constexpr int RAW_MAX = 8;          // raw ring depth (upper bound)
constexpr int RAW_PRODUCERS = 512;  // 16 producer warps = 4 per scheduler (2 per scheduler were latency-bound)
constexpr int RAW_THREADS = RAW_PRODUCERS + 64;   // + MMA warp + loader warp
constexpr int MIX_LOADERS = 96;                   // mixtures: three loader warps (one warp's issue latency bound the ring;
                                                  // 20 warps keep the 96-register budget of the producers)
constexpr int MIX_THREADS = RAW_PRODUCERS + 32 + MIX_LOADERS;

struct RawBarriers {
    uint64_t raw_full[RAW_MAX], raw_empty[RAW_MAX];
    uint64_t full[RAW_OPS], empty[RAW_OPS];
    uint64_t tfull[2], tempty[2];
    uint32_t tmem_base;
    uint32_t pad[3];
};

template <int D4>
struct RawCfg {
    static constexpr int D = 4 * D4, KF = RAW_KF;
    static constexpr int NB = (2 * D + 1 + 15) / 16 * 16;
    static constexpr int KG = KF / 8;
    static constexpr int A_FLOATS = KF * GM, B_FLOATS = KF * NB;
    static constexpr int STAGE_FLOATS = 2 * A_FLOATS + 2 * B_FLOATS;
    static constexpr int BU = ((KF / 4) * D + RAW_PRODUCERS - 1) / RAW_PRODUCERS;
    static constexpr int NCH = NB / 16, MYCH = (NCH + 3) / 4;   // four warps share a TMEM lane quarter
    static constexpr uint32_t TMEM_COLS = 2 * NB <= 64 ? 64 : (2 * NB <= 128 ? 128 : (2 * NB <= 256 ? 256 : 512));
    static constexpr int DR = 128 / KF;    // stages per drain: 48 truncating TMEM accumulations (~3e-7 low)
    static constexpr size_t FIXED = (size_t)RAW_OPS * STAGE_FLOATS * 4 + sizeof(RawBarriers) + 1024;
    // floats of one raw stage; w = width of the posterior block (M, or for mixtures GM + 2 * GM / C)
    static size_t raw_stage_bytes(int w) { return (size_t)KF * (w + D) * 4; }
    static int raw_stages(int w) {
        size_t left = 227 * 1024 - FIXED;
        int n = (int)(left / raw_stage_bytes(w));
        return n > RAW_MAX ? RAW_MAX : n;
    }
};

// MODE: 0 = dense posteriors of one Gaussian tile, 1 = mixtures, 2 = PATH: the posteriors are the one-hot rows of a
// state path (Viterbi training, hmm.py:42-58), never materialised: a stage brings the 32 pdf ids of its frames
// (128 bytes instead of 32 x M floats) and the producers form w = (pdf id == Gaussian) ? scale : 0.
template <int D4, int MODE>
__global__ void __launch_bounds__(MODE == 1 ? MIX_THREADS : RAW_THREADS, 1) accumulate_tc_raw_kernel(Args a, int RS) {
    constexpr bool MIX = MODE == 1, PATH = MODE == 2;
    using C = RawCfg<D4>;
    constexpr int D = C::D, NB = C::NB, KG = C::KG, KF = C::KF, STAGES = RAW_OPS;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    float* stage_base = reinterpret_cast<float*>(smem_raw);
    RawBarriers* bars = reinterpret_cast<RawBarriers*>(stage_base + (size_t)STAGES * C::STAGE_FLOATS);
    float* ring = reinterpret_cast<float*>(bars + 1);          // RS x [KF x M posteriors | KF x D features]
    // mixtures: RS x [KF x GM comp llh | KF x nk posteriors | KF x nk pdf llh | KF x D features]
    const int gtile = MIX ? (int)(blockIdx.x % a.n_gtiles) : 0;
    const int g0 = gtile * GM;
    const int M = MIX ? min(GM, a.M - g0) : a.M;                // Gaussians of this tile
    const int nk = MIX ? GM / a.C : 0, k0 = MIX ? g0 / a.C : 0;
    const int PW = MIX ? GM + 2 * nk : (PATH ? 1 : M);         // floats per frame ahead of the features
    const int raw_floats = KF * (PW + D);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t f_begin = (int64_t)(MIX ? blockIdx.x / a.n_gtiles : blockIdx.x) * a.frames_per_cta;
    const int64_t f_end = min(a.N, f_begin + a.frames_per_cta);
    const int n_tiles = (f_end > f_begin) ? (int)((f_end - f_begin + KF - 1) / KF) : 0;

    if (tid == 0) {
        for (int i = 0; i < RAW_MAX; ++i) {
            mbar_init(&bars->raw_full[i], MIX ? MIX_LOADERS : 1);
            mbar_init(&bars->raw_empty[i], RAW_PRODUCERS);
        }
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&bars->full[i], RAW_PRODUCERS);
            mbar_init(&bars->empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->tfull[i], 1);
            mbar_init(&bars->tempty[i], RAW_PRODUCERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == RAW_PRODUCERS / 32) tmem_alloc(&bars->tmem_base, C::TMEM_COLS);
    for (int i = tid; i < STAGES * C::STAGE_FLOATS / 4; i += (int)blockDim.x)
        reinterpret_cast<float4*>(stage_base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int i = tid; i < STAGES * KF; i += (int)blockDim.x) {
        int st = i / KF, f = i - st * KF;
        float* b_hi = stage_base + (size_t)st * C::STAGE_FLOATS + 2 * C::A_FLOATS;
        b_hi[((2 * D) >> 3) * (KF * 8) + (f >> 2) * 32 + ((2 * D) & 7) * 4 + (f & 3)] = 1.f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp > RAW_PRODUCERS / 32) {
        // ------------------------------- loader -----------------------------------
        if constexpr (MIX) {
            // Row pieces of 64 - 512 bytes: 16-byte cp.async (LDGSTS) per lane, completion counted on the stage's
            // mbarrier by one asynchronous arrive per lane (a bulk copy per row piece cost ~65 cycles of TMA issue
            // each, 97 per stage: slower than the per-thread gathers it replaced).
            const int cq = M >> 2, kq = (M / a.C) >> 2;           // 16-byte pieces per row: comp llh, posteriors
            const int lw = warp - (RAW_PRODUCERS / 32 + 1);       // loader warp: rows lw, lw + NLW, ...
            const int lt = lw * 32 + lane;
            // posteriors / pdf llhs: nk / 4 (a power of two) pieces per row, thread lt -> (row, piece) of element lt, lt + 128, ...
            const int nkq = nk >> 2, nkq_sh = 31 - __clz(nkq);
            int rs = 0;
            uint32_t rph = 0;
            const float* c_src = a.comp_llh + (size_t)(f_begin + lw) * a.M + g0 + 4 * lane;
            constexpr int NLW = MIX_LOADERS / 32;
            const size_t c_step = (size_t)NLW * a.M, c_adv = (size_t)KF * a.M;
            for (int it = 0; it < n_tiles; ++it, c_src += c_adv) {
                const int64_t t0 = f_begin + (int64_t)it * KF;
                const int rows = (int)min((int64_t)KF, f_end - t0);
                mbar_wait(&bars->raw_empty[rs], rph ^ 1);
                float* dst = ring + (size_t)rs * raw_floats;
                if (lane < cq) {
                    const float* src = c_src;
                    float* d = dst + lw * GM + 4 * lane;
#pragma unroll 4
                    for (int r = lw; r < rows; r += NLW, src += c_step, d += NLW * GM) cp_async16(d, src);
                }
                for (int e = lt; e < (rows << nkq_sh); e += MIX_LOADERS) {
                    const int r = e >> nkq_sh, p = e & (nkq - 1);
                    if (p < kq) {
                        cp_async16(dst + KF * GM + r * nk + 4 * p, a.pdf_post + (size_t)(t0 + r) * a.ld_post + k0 + 4 * p);
                        cp_async16(dst + KF * (GM + nk) + r * nk + 4 * p,
                                   a.pdf_llh + (size_t)(t0 + r) * a.ld_pdf + k0 + 4 * p);
                    }
                }
                for (int e = lt; e < rows * (D / 4); e += MIX_LOADERS)
                    cp_async16(dst + KF * PW + 4 * e, a.X + (size_t)t0 * D + 4 * e);
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars->raw_full[rs]))
                             : "memory");
                if (++rs == RS) {
                    rs = 0;
                    rph ^= 1;
                }
            }
        } else if (elect_one()) {
            int rs = 0;
            uint32_t rph = 0;                      // ring position / pass parity (no runtime division)
            for (int it = 0; it < n_tiles; ++it) {
                const int64_t t0 = f_begin + (int64_t)it * KF;
                const uint32_t rows = (uint32_t)min((int64_t)KF, f_end - t0);
                mbar_wait(&bars->raw_empty[rs], rph ^ 1);
                float* dst = ring + (size_t)rs * raw_floats;
                if constexpr (PATH) {
                    const uint32_t idb = ((rows + 3u) & ~3u) * 4u;       // whole 16-byte pieces (the id array is padded)
                    mbar_arrive_expect_tx(&bars->raw_full[rs], idb + rows * (uint32_t)D * 4u);
                    bulk_g2s(dst, a.pdf_ids + t0, idb, &bars->raw_full[rs]);
                } else {
                    mbar_arrive_expect_tx(&bars->raw_full[rs], rows * (uint32_t)(M + D) * 4u);
                    bulk_g2s(dst, a.pdf_post + (size_t)t0 * M, rows * (uint32_t)M * 4u, &bars->raw_full[rs]);
                }
                bulk_g2s(dst + KF * PW, a.X + (size_t)t0 * D, rows * (uint32_t)D * 4u, &bars->raw_full[rs]);
                if (++rs == RS) {
                    rs = 0;
                    rph ^= 1;
                }
            }
        }
    } else if (warp == RAW_PRODUCERS / 32) {
        // ------------------------------ MMA issuer -------------------------------
        if (n_tiles > 0 && elect_one()) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) |
                                   ((uint32_t)(GM >> 4) << 24);
            constexpr uint32_t LBO = 128, SBO = (KF / 4) * 128;
            for (int it = 0; it < n_tiles; ++it) {
                const int st = it % STAGES;
                const int grp = it / C::DR, buf = grp & 1;
                const bool first = it % C::DR == 0, last = (it % C::DR == C::DR - 1) || it == n_tiles - 1;
                mbar_wait(&bars->full[st], (it / STAGES) & 1);
                if (first) mbar_wait(&bars->tempty[buf], ((grp >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * NB);
                const uint32_t a_hi = smem_u32(stage_base + (size_t)st * C::STAGE_FLOATS);
                const uint32_t a_lo = a_hi + C::A_FLOATS * 4u;
                const uint32_t b_hi = a_lo + C::A_FLOATS * 4u;
                const uint32_t b_lo = b_hi + C::B_FLOATS * 4u;
#pragma unroll 1
                for (int ks = 0; ks < KG; ++ks) {
                    const uint32_t ko = (uint32_t)ks * 256u;
                    const uint64_t dah = make_desc(a_hi + ko, LBO, SBO), dal = make_desc(a_lo + ko, LBO, SBO);
                    const uint64_t dbh = make_desc(b_hi + ko, LBO, SBO), dbl = make_desc(b_lo + ko, LBO, SBO);
                    umma_tf32(d_tmem, dah, dbh, idesc, !(first && ks == 0));
                    umma_tf32(d_tmem, dal, dbh, idesc, 1);
                    umma_tf32(d_tmem, dah, dbl, idesc, 1);
                }
                umma_commit(&bars->empty[st]);
                if (last) umma_commit(&bars->tfull[buf]);
            }
        }
    } else {
        // ------------------------------ producers --------------------------------
        const int gl = tid & (GM - 1);
        const int fq0 = tid >> 7;            // quads fq0, fq0 + 4, ...
        const bool a_active = gl < M;
        const int kl = MIX ? gl / a.C : 0;   // pdf (local) of this thread's Gaussian
        constexpr int AU = KF / 16;
        const int a_row = (gl >> 3) * (KF * 8) + (gl & 7) * 4;
        int xd[C::BU], xfq[C::BU];
        bool x_active[C::BU];
#pragma unroll
        for (int j = 0; j < C::BU; ++j) {
            const int u = tid + j * RAW_PRODUCERS;
            x_active[j] = u < (KF / 4) * D;
            xd[j] = x_active[j] ? u % D : 0;
            xfq[j] = x_active[j] ? u / D : 0;
        }
        // drain: lane quarter q, 16-column chunks part, part + 4, ...; Kahan-compensated fp32 sums
        // (fp64 adds + float->double conversions throttled the fp64 / XU pipes)
        const int q = warp & 3, part = warp >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        float sums[C::MYCH][16], comp[C::MYCH][16];
#pragma unroll
        for (int m = 0; m < C::MYCH; ++m)
#pragma unroll
            for (int i = 0; i < 16; ++i) sums[m][i] = comp[m][i] = 0.f;
        auto drain = [&](int g) {
            const int buf = g & 1;
            mbar_wait(&bars->tfull[buf], (g >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int m = 0; m < C::MYCH; ++m) {
                const int ch = part + 4 * m;
                if (ch < C::NCH) {
                    float v[16];
                    tmem_ld16(taddr + (uint32_t)(buf * NB + ch * 16), v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float y = v[i] - comp[m][i];
                        const float t = sums[m][i] + y;
                        comp[m][i] = (t - sums[m][i]) - y;
                        sums[m][i] = t;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&bars->tempty[buf]);
        };

        int rs = 0;
        uint32_t rph = 0;
        for (int it = 0; it < n_tiles; ++it) {
            const int st = it % STAGES;
            const int rows = (int)min((int64_t)KF, f_end - (f_begin + (int64_t)it * KF));
            const float* rp = ring + (size_t)rs * raw_floats;     // posteriors [KF][M] (mixtures: comp llh [KF][GM])
            const float* rx = rp + KF * PW;                        // features   [KF][D]
            float* A_hi = stage_base + (size_t)st * C::STAGE_FLOATS;
            float* A_lo = A_hi + C::A_FLOATS;
            float* B_hi = A_lo + C::A_FLOATS;
            float* B_lo = B_hi + C::B_FLOATS;
            mbar_wait(&bars->raw_full[rs], rph);
            mbar_wait(&bars->empty[st], ((it / STAGES) & 1) ^ 1);
            if (rows < KF) {
                // ragged last stage of the range: stale posterior and feature rows -> 0 (w = 0 and x = 0 for those
                // frames), so that the loads below carry no per-element row predicate
                float* r0 = const_cast<float*>(rp);
                if constexpr (MIX) {
                    for (int e = rows * nk + tid; e < KF * nk; e += RAW_PRODUCERS) r0[KF * GM + e] = 0.f;
                } else if constexpr (PATH) {
                    for (int e = rows + tid; e < KF; e += RAW_PRODUCERS) reinterpret_cast<int*>(r0)[e] = -1;
                } else {
                    for (int e = rows * M + tid; e < KF * M; e += RAW_PRODUCERS) r0[e] = 0.f;
                }
                for (int e = rows * D + tid; e < KF * D; e += RAW_PRODUCERS) r0[KF * PW + e] = 0.f;
                asm volatile("bar.sync 1, %0;" ::"n"(RAW_PRODUCERS) : "memory");
            }
            if (a_active) {
#pragma unroll
                for (int j = 0; j < AU; ++j) {
                    const int f0 = (fq0 + 4 * j) * 4;
                    float h[4], l[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float w;
                        if constexpr (MIX) {
                            // log2 domain: one FFMA feeds the ex2; rows past the end of a ragged stage hold zeros
                            // (cleared below), posteriors that are exactly zero stay zero whatever the llhs are
                            const float* rq = rp + KF * GM + (f0 + i) * nk + kl;
                            const float post = rq[0];
                            const float e = ex2(1.4426950408889634f * (rp[(f0 + i) * GM + gl] - rq[KF * nk]));
                            w = (post != 0.f) ? post * e : 0.f;
                        } else if constexpr (PATH) {
                            w = (reinterpret_cast<const int*>(rp)[f0 + i] == gl) ? a.path_scale : 0.f;
                        } else {
                            w = rp[(f0 + i) * M + gl];
                        }
                        h[i] = tf32_rn(w);
                        l[i] = w - h[i];
                    }
                    const int off = a_row + (fq0 + 4 * j) * 32;
                    *reinterpret_cast<float4*>(A_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4*>(A_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
                }
            }
#pragma unroll
            for (int j = 0; j < C::BU; ++j) {
                if (x_active[j]) {
                    const int d = xd[j], fq = xfq[j];
                    float xh[4], xl[4], qh[4], ql[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float x = rx[(fq * 4 + i) * D + d];
                        const float qq = x * x;
                        xh[i] = tf32_rn(x);
                        xl[i] = x - xh[i];
                        qh[i] = tf32_rn(qq);
                        ql[i] = qq - qh[i];
                    }
                    const int offx = (d >> 3) * (KF * 8) + fq * 32 + (d & 7) * 4;
                    const int offq = ((D + d) >> 3) * (KF * 8) + fq * 32 + ((D + d) & 7) * 4;
                    *reinterpret_cast<float4*>(B_hi + offx) = make_float4(xh[0], xh[1], xh[2], xh[3]);
                    *reinterpret_cast<float4*>(B_lo + offx) = make_float4(xl[0], xl[1], xl[2], xl[3]);
                    *reinterpret_cast<float4*>(B_hi + offq) = make_float4(qh[0], qh[1], qh[2], qh[3]);
                    *reinterpret_cast<float4*>(B_lo + offq) = make_float4(ql[0], ql[1], ql[2], ql[3]);
                }
            }
            fence_proxy_async();
            mbar_arrive(&bars->full[st]);
            mbar_arrive(&bars->raw_empty[rs]);
            if (++rs == RS) {
                rs = 0;
                rph ^= 1;
            }
            // drain the previous group one stage late: its last MMAs have certainly retired by then
            if (it % C::DR == 1 && it > C::DR) drain(it / C::DR - 1);
        }
        if (n_tiles > 0) {
            const int last_g = (n_tiles - 1) / C::DR;
            // groups not drained inside the loop: the last one, and the one before it when the loop ended
            // before that group's (late) drain slot
            if (last_g >= 1 && (n_tiles - 1) < last_g * C::DR + 1) drain(last_g - 1);
            drain(last_g);
            const int g = q * 32 + lane;
            if (g < M) {
                const int Q = 2 * D + 2;
                double* row = a.acc + (size_t)(g0 + g) * Q;
#pragma unroll
                for (int m = 0; m < C::MYCH; ++m) {
                    const int ch = part + 4 * m;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = ch * 16 + i;
                        const double v = (double)sums[m][i] - (double)comp[m][i];
                        if (c < 2 * D) {
                            if (v != 0.0) atomicAdd(row + c, c < D ? v : -0.5 * v);
                        } else if (c == 2 * D) {
                            if (v != 0.0) {
                                atomicAdd(row + 2 * D, -0.5 * v);
                                atomicAdd(row + 2 * D + 1, 0.5 * v);
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == RAW_PRODUCERS / 32) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

template <int D4, int MODE>
static int launch_raw(const Args& a0, cudaStream_t st) {
    constexpr bool MIX = MODE == 1;
    using C = RawCfg<D4>;
    Args a = a0;
    const int pw = MIX ? GM + 2 * (GM / a.C) : (MODE == 2 ? 1 : a.M);
    const int RS = C::raw_stages(pw);
    if (RS < 3) return BEER_ERR_UNSUPPORTED;
    const size_t smem = C::FIXED + (size_t)RS * C::raw_stage_bytes(pw);
    static bool attr_set = false;
    if (!attr_set) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(accumulate_tc_raw_kernel<D4, MODE>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    a.n_gtiles = MIX ? (a.M + GM - 1) / GM : 1;
    // one CTA per SM; with several Gaussian tiles: the largest grid of whole tile sets within 3 waves
    int64_t chunks = MIX ? (a.n_gtiles >= 3 * kNumSMs ? 1 : 3 * kNumSMs / a.n_gtiles) : kNumSMs;
    int64_t max_chunks = (a.N + RAW_KF - 1) / RAW_KF;
    if (chunks > max_chunks) chunks = max_chunks;
    int64_t fpc = (a.N + chunks - 1) / chunks;
    fpc = (fpc + RAW_KF - 1) / RAW_KF * RAW_KF;
    chunks = (a.N + fpc - 1) / fpc;
    a.frames_per_cta = fpc;
    accumulate_tc_raw_kernel<D4, MODE><<<(int)(chunks * a.n_gtiles), MIX ? MIX_THREADS : RAW_THREADS, smem, st>>>(a, RS);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

}  // namespace kctc
}  // namespace beer

using namespace beer;

extern "C" {

int beer_accumulate_tc_supported(int M, int D) {
    if (M <= 0 || D <= 0 || D % 4 != 0) return 0;
    int d4 = D / 4;
    return (d4 == 5 || d4 == 10 || d4 == 16 || d4 == 20) ? 1 : 0;
}

int beer_accumulate_stats_tc(const float* X, int64_t N, int D, const float* pdf_post, int64_t ld_post,
                             const float* pdf_llh, int64_t ld_pdf, const float* comp_llh, const int32_t* comp_off,
                             int Kp, int M, double* acc_normal, void* stream) {
    if (!X || !acc_normal || N < 0 || D <= 0 || M <= 0 || Kp <= 0) return BEER_ERR_ARG;
    if (comp_llh != nullptr && pdf_llh == nullptr) return BEER_ERR_ARG;
    if (comp_off == nullptr && M % Kp != 0) return BEER_ERR_ARG;
    if (comp_llh == nullptr && M != Kp) return BEER_ERR_ARG;
    if (pdf_post != nullptr && ld_post < Kp) return BEER_ERR_ARG;
    if (!beer_accumulate_tc_supported(M, D)) return BEER_ERR_UNSUPPORTED;
    if (((uintptr_t)X & 15) != 0) return BEER_ERR_ARG;
    if (N == 0) return BEER_OK;
    kctc::Args a;
    a.X = X; a.N = N; a.pdf_post = pdf_post; a.ld_post = ld_post; a.pdf_llh = pdf_llh; a.ld_pdf = ld_pdf;
    a.comp_llh = comp_llh; a.comp_off = comp_off; a.Kp = Kp; a.M = M; a.C = M / Kp; a.acc = acc_normal;
    a.pdf_ids = nullptr; a.path_scale = 0.f;
    a.n_gtiles = 0; a.frames_per_cta = 0;
    cudaStream_t st = (cudaStream_t)stream;
    // dense posteriors of a single Gaussian tile: bulk-staged kernel (every stage is two contiguous copies)
    if (comp_llh == nullptr && pdf_post != nullptr && M <= kctc::GM && ld_post == M && M % 4 == 0 &&
        ((uintptr_t)pdf_post & 15) == 0 && getenv("BEER_B200_KC_NO_BULK") == nullptr) {
        int rc = BEER_ERR_UNSUPPORTED;
        switch (D / 4) {
            case 5: rc = kctc::launch_raw<5, 0>(a, st); break;      // wider D: the drain's register sums
            case 10: rc = kctc::launch_raw<10, 0>(a, st); break;    // do not fit 18 warps per SM
        }
        if (rc != BEER_ERR_UNSUPPORTED) return rc;
    }
    // mixtures, C | 128 components per pdf: bulk-staged rows (16-byte aligned row pieces)
    if (comp_llh != nullptr && pdf_post != nullptr && comp_off == nullptr && a.C >= 1 && a.C <= 32 &&
        (a.C & (a.C - 1)) == 0 && M % 4 == 0 && Kp % 4 == 0 && ld_post % 4 == 0 && ld_pdf % 4 == 0 &&
        ((uintptr_t)pdf_post & 15) == 0 && ((uintptr_t)pdf_llh & 15) == 0 && ((uintptr_t)comp_llh & 15) == 0 &&
        getenv("BEER_B200_KC_NO_BULK") == nullptr) {
        int rc = BEER_ERR_UNSUPPORTED;
        switch (D / 4) {
            case 5: rc = kctc::launch_raw<5, 1>(a, st); break;
            case 10: rc = kctc::launch_raw<10, 1>(a, st); break;
        }
        if (rc != BEER_ERR_UNSUPPORTED) return rc;
    }
    switch (D / 4) {
        case 5: return kctc::launch<5, 64, 2>(a, st);
        case 10: return kctc::launch<10, 32, 4>(a, st);
        case 16: return kctc::launch<16, 32, 2>(a, st);
        case 20: return kctc::launch<20, 32, 2>(a, st);
    }
    return BEER_ERR_UNSUPPORTED;
}

int beer_accumulate_stats_path(const float* X, int64_t N, int D, const int32_t* pdf_ids, float scale, int M,
                               double* acc_normal, void* stream) {
    if (!X || !pdf_ids || !acc_normal || N < 0 || D <= 0 || M <= 0) return BEER_ERR_ARG;
    if (M > kctc::GM || !(D == 20 || D == 40)) return BEER_ERR_UNSUPPORTED;
    if (((uintptr_t)X & 15) != 0 || ((uintptr_t)pdf_ids & 15) != 0) return BEER_ERR_ARG;
    if (N == 0) return BEER_OK;
    kctc::Args a;
    a.X = X; a.N = N; a.pdf_post = nullptr; a.ld_post = 0; a.pdf_llh = nullptr; a.ld_pdf = 0; a.comp_llh = nullptr;
    a.comp_off = nullptr; a.Kp = M; a.M = M; a.C = 1; a.acc = acc_normal; a.n_gtiles = 0; a.frames_per_cta = 0;
    a.pdf_ids = pdf_ids; a.path_scale = scale;
    cudaStream_t st = (cudaStream_t)stream;
    return D == 20 ? kctc::launch_raw<5, 2>(a, st) : kctc::launch_raw<10, 2>(a, st);
}

}  // extern "C"
