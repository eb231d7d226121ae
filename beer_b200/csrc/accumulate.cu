// KC -- posterior-weighted sufficient statistics (0th / 1st / 2nd moments).
//
//   acc[j] += sum_t w_tj [x_t, -x_t^2/2, -1/2, 1/2],   w_tj = g_t,pdf(j) * r_tj
//
// a tall-skinny (M x T).(T x 2D) contraction whose reduction dimension is time.
// Each CTA owns a tile of 128 Gaussians and a contiguous range of frames, keeps
// the (128 x 2D) partial sums in registers and flushes them once with fp64 atomics.
//
// Reference semantics: beer/models/mixtureset.py:100-112,
// beer/models/normalset.py:121-123, beer/models/categoricalset.py:54-55.
#include "common.cuh"
#include "../../include/beer_b200.h"

namespace beer {

constexpr int KC_G = 128;       // Gaussians per tile
constexpr int KC_TF = 32;       // frames per smem stage
constexpr int KC_THREADS = 256;

struct KcArgs {
    const float* X;
    int64_t N;
    int D;
    const float* pdf_post;
    int64_t ld_post;
    const float* pdf_llh;
    int64_t ld_pdf;
    const float* comp_llh;
    const int32_t* comp_off;
    int Kp, M, C;  // C: uniform components per pdf when comp_off == NULL
    double* acc;
    int n_gtiles;
    int64_t frames_per_cta;
};

template <int FPT>
__global__ void __launch_bounds__(KC_THREADS, 2) accumulate_simt_kernel(KcArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int D = a.D, D2 = 2 * D;
    const int SLD = 16 * FPT;                 // padded feature stride
    float* ws = smem;                         // [KC_TF][KC_G]
    float* ss = ws + KC_TF * KC_G;            // [KC_TF][SLD]
    int* s_pdf = reinterpret_cast<int*>(ss + KC_TF * SLD);  // [KC_G]

    const int tid = threadIdx.x;
    const int gtile = blockIdx.x % a.n_gtiles;
    const int64_t chunk = blockIdx.x / a.n_gtiles;
    const int g0 = gtile * KC_G;
    const int ng = min(KC_G, a.M - g0);
    const int64_t f_begin = chunk * a.frames_per_cta;
    const int64_t f_end = min(a.N, f_begin + a.frames_per_cta);
    if (f_begin >= f_end) return;

    // pdf of every Gaussian of the tile
    for (int g = tid; g < KC_G; g += KC_THREADS) {
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
    // zero the padded feature columns once
    for (int e = tid; e < KC_TF * SLD; e += KC_THREADS) ss[e] = 0.f;

    const int tf = tid & 15, tg = tid >> 4;
    float acc[8][FPT];
    float cnt[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        cnt[i] = 0.f;
#pragma unroll
        for (int f = 0; f < FPT; ++f) acc[i][f] = 0.f;
    }

    for (int64_t t0 = f_begin; t0 < f_end; t0 += KC_TF) {
        const int nf = (int)min((int64_t)KC_TF, f_end - t0);
        __syncthreads();
        // weights tile
        for (int e = tid; e < KC_TF * KC_G; e += KC_THREADS) {
            int f = e / KC_G, g = e - f * KC_G;
            float w = 0.f;
            if (f < nf && g < ng) {
                int64_t t = t0 + f;
                int k = s_pdf[g];
                w = (a.pdf_post != nullptr) ? a.pdf_post[(size_t)t * a.ld_post + k] : 1.f;
                if (a.comp_llh != nullptr && w != 0.f)
                    w *= __expf(a.comp_llh[(size_t)t * a.M + g0 + g] - a.pdf_llh[(size_t)t * a.ld_pdf + k]);
            }
            ws[e] = w;
        }
        // statistics tile [x, x^2]
        for (int e = tid; e < KC_TF * D; e += KC_THREADS) {
            int f = e / D, d = e - f * D;
            float x = (f < nf) ? a.X[(size_t)t0 * D + e] : 0.f;
            ss[f * SLD + d] = x;
            ss[f * SLD + D + d] = x * x;
        }
        __syncthreads();
#pragma unroll 4
        for (int f = 0; f < KC_TF; ++f) {
            float4 w0 = *reinterpret_cast<const float4*>(&ws[f * KC_G + tg * 8]);
            float4 w1 = *reinterpret_cast<const float4*>(&ws[f * KC_G + tg * 8 + 4]);
            float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            float sv[FPT];
#pragma unroll
            for (int q = 0; q < FPT; ++q) sv[q] = ss[f * SLD + tf + 16 * q];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int q = 0; q < FPT; ++q) acc[i][q] = fmaf(wv[i], sv[q], acc[i][q]);
            }
            if (tf == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) cnt[i] += wv[i];
            }
        }
    }

    const int Q = D2 + 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int g = tg * 8 + i;
        if (g >= ng) continue;
        double* row = a.acc + (size_t)(g0 + g) * Q;
#pragma unroll
        for (int q = 0; q < FPT; ++q) {
            int c = tf + 16 * q;
            if (c < D2 && acc[i][q] != 0.f) {
                double v = (double)acc[i][q];
                atomicAdd(row + c, c < D ? v : -0.5 * v);
            }
        }
        if (tf == 0 && cnt[i] != 0.f) {
            atomicAdd(row + D2, -0.5 * (double)cnt[i]);
            atomicAdd(row + D2 + 1, 0.5 * (double)cnt[i]);
        }
    }
}

__global__ void mixture_weight_stats_kernel(const double* __restrict__ acc, int M, int D,
                                            const int32_t* __restrict__ comp_off, int Kp, double* __restrict__ out) {
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;     // one warp per pdf
    if (k >= Kp) return;
    int Q = 2 * D + 2;
    int c0, c1;
    if (comp_off == nullptr) {
        int C = M / Kp;
        c0 = k * C; c1 = c0 + C;
    } else {
        c0 = comp_off[k]; c1 = comp_off[k + 1];
    }
    double tot = 0.0;
    for (int j = c0 + lane; j < c1; j += 32) {
        double n = 2.0 * acc[(size_t)j * Q + 2 * D + 1];
        tot += n;
        if (j < c1 - 1) out[j] = n;
    }
    tot = warp_sum(tot);
    if (lane == 0 && c1 > c0) out[c1 - 1] = tot;
}

template <int FPT>
static int launch_kc(const KcArgs& a, int grid, cudaStream_t st) {
    size_t smem = sizeof(float) * (KC_TF * KC_G + KC_TF * 16 * FPT) + sizeof(int) * KC_G;
    accumulate_simt_kernel<FPT><<<grid, KC_THREADS, smem, st>>>(a);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}


// ---------------------------------------------------------------------------------------------
// Statistics of a mixture model along a STATE PATH (Viterbi training, hmm.py:42-58 with viterbi=True; one-hot pdf
// posteriors x the responsibilities inside the chosen pdf, mixtureset.py:100-112): per frame only the C Gaussians of its
// pdf carry weight, so nothing dense is formed.  One warp walks a block of consecutive frames, two dimensions per
// lane; the C weight rows of the current pdf stay in registers while the path stays in it (a state lasts several
// frames), the per-Gaussian sums of the run are flushed with fp64 atomics when the pdf changes.
//   z_c = W_jc . [x, -x^2/2] + bias_jc,  r_c = exp(z_c - lse),  acc[jc] += scale r_c T(x),  frame = scale (lse + ref_t)
// ---------------------------------------------------------------------------------------------
constexpr int PATH_FRAMES = 64;      // frames per warp

template <int C>
__global__ void __launch_bounds__(256) path_mix_stats_kernel(const float* __restrict__ X, int64_t N, int D,
                                                             const int32_t* __restrict__ pdf_ids,
                                                             const float* __restrict__ W, const float* __restrict__ bias,
                                                             const float* __restrict__ frame_ref, float scale,
                                                             double* __restrict__ acc, float* __restrict__ frame_out) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t t_begin = gwarp * PATH_FRAMES, t_end = min(N, t_begin + PATH_FRAMES);
    if (t_begin >= N) return;
    const int d0 = lane, d1 = lane + 32, Q = 2 * D + 2;
    const bool v0 = d0 < D, v1 = d1 < D;
    float w1[C][2], w2[C][2], b[C];          // weight rows of the current pdf: linear / quadratic part, bias
    float s1[C][2], s2[C][2], cnt[C];        // sums of the current run
    int cur = -1;
    auto flush = [&]() {
        if (cur < 0) return;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            double* row = acc + (size_t)(cur * C + c) * Q;
            if (cnt[c] != 0.f) {
                if (v0) {
                    atomicAdd(row + d0, (double)s1[c][0]);
                    atomicAdd(row + D + d0, (double)s2[c][0]);
                }
                if (v1) {
                    atomicAdd(row + d1, (double)s1[c][1]);
                    atomicAdd(row + D + d1, (double)s2[c][1]);
                }
                if (lane == 0) {
                    atomicAdd(row + 2 * D, -0.5 * (double)cnt[c]);
                    atomicAdd(row + 2 * D + 1, 0.5 * (double)cnt[c]);
                }
            }
        }
    };
    for (int64_t t = t_begin; t < t_end; ++t) {
        const int k = __ldg(pdf_ids + t);
        if (k != cur) {
            flush();
            cur = k;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float* wr = W + (size_t)(k * C + c) * 2 * D;
                w1[c][0] = v0 ? __ldg(wr + d0) : 0.f;
                w2[c][0] = v0 ? __ldg(wr + D + d0) : 0.f;
                w1[c][1] = v1 ? __ldg(wr + d1) : 0.f;
                w2[c][1] = v1 ? __ldg(wr + D + d1) : 0.f;
                b[c] = __ldg(bias + k * C + c);
                s1[c][0] = s1[c][1] = s2[c][0] = s2[c][1] = cnt[c] = 0.f;
            }
        }
        const float x0 = v0 ? __ldg(X + (size_t)t * D + d0) : 0.f, x1 = v1 ? __ldg(X + (size_t)t * D + d1) : 0.f;
        const float q0 = -0.5f * x0 * x0, q1 = -0.5f * x1 * x1;
        float z[C];
#pragma unroll
        for (int c = 0; c < C; ++c) z[c] = fmaf(w1[c][0], x0, fmaf(w2[c][0], q0, fmaf(w1[c][1], x1, w2[c][1] * q1)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int c = 0; c < C; ++c) z[c] += __shfl_xor_sync(0xffffffffu, z[c], o);
        float m = kNegInf;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            z[c] += b[c];
            m = fmaxf(m, z[c]);
        }
        float se = 0.f;
        float e[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            e[c] = (m == kNegInf) ? 0.f : __expf(z[c] - m);
            se += e[c];
        }
        const float inv = se > 0.f ? scale / se : 0.f;
        if (lane == 0 && frame_out != nullptr)
            frame_out[t] = scale * ((m == kNegInf ? kNegInf : m + __logf(se)) + (frame_ref ? __ldg(frame_ref + t) : 0.f));
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float w = e[c] * inv;
            s1[c][0] = fmaf(w, x0, s1[c][0]);
            s2[c][0] = fmaf(w, q0, s2[c][0]);
            s1[c][1] = fmaf(w, x1, s1[c][1]);
            s2[c][1] = fmaf(w, q1, s2[c][1]);
            cnt[c] += w;
        }
    }
    flush();
}

template <int C>
static int launch_path_mix(const float* X, int64_t N, int D, const int32_t* pdf_ids, const float* W, const float* bias,
                           const float* frame_ref, float scale, double* acc, float* frame_out, cudaStream_t st) {
    const int64_t warps = (N + PATH_FRAMES - 1) / PATH_FRAMES;
    const int blocks = (int)((warps + 7) / 8);
    path_mix_stats_kernel<C><<<blocks, 256, 0, st>>>(X, N, D, pdf_ids, W, bias, frame_ref, scale, acc, frame_out);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

}  // namespace beer

using namespace beer;

extern "C" {

int beer_accumulate_stats(const float* X, int64_t N, int D, const float* pdf_post, int64_t ld_post,
                          const float* pdf_llh, int64_t ld_pdf, const float* comp_llh, const int32_t* comp_off,
                          int Kp, int M, double* acc_normal, void* stream) {
    if (!X || !acc_normal || N < 0 || D <= 0 || M <= 0 || Kp <= 0) return BEER_ERR_ARG;
    if (comp_llh != nullptr && pdf_llh == nullptr) return BEER_ERR_ARG;
    if (comp_off == nullptr && M % Kp != 0) return BEER_ERR_ARG;
    if (comp_llh == nullptr && M != Kp) return BEER_ERR_ARG;
    if (pdf_post != nullptr && ld_post < Kp) return BEER_ERR_ARG;
    if (2 * D > 128) return BEER_ERR_UNSUPPORTED;
    if (N == 0) return BEER_OK;
    KcArgs a;
    a.X = X; a.N = N; a.D = D; a.pdf_post = pdf_post; a.ld_post = ld_post;
    a.pdf_llh = pdf_llh; a.ld_pdf = ld_pdf; a.comp_llh = comp_llh; a.comp_off = comp_off;
    a.Kp = Kp; a.M = M; a.C = M / Kp; a.acc = acc_normal;
    a.n_gtiles = (M + KC_G - 1) / KC_G;
    int64_t chunks = (2 * kNumSMs + a.n_gtiles - 1) / a.n_gtiles;
    int64_t max_chunks = (N + 4 * KC_TF - 1) / (4 * KC_TF);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    int64_t fpc = (N + chunks - 1) / chunks;
    fpc = (fpc + KC_TF - 1) / KC_TF * KC_TF;
    chunks = (N + fpc - 1) / fpc;
    a.frames_per_cta = fpc;
    int grid = (int)(chunks * a.n_gtiles);
    cudaStream_t st = (cudaStream_t)stream;
    int fpt = (2 * D + 15) / 16;
    if (fpt <= 2) return launch_kc<2>(a, grid, st);
    if (fpt <= 4) return launch_kc<4>(a, grid, st);
    if (fpt == 5) return launch_kc<5>(a, grid, st);
    if (fpt == 6) return launch_kc<6>(a, grid, st);
    return launch_kc<8>(a, grid, st);
}

int beer_mixture_weight_stats(const double* acc_normal, int M, int D, const int32_t* comp_off, int Kp,
                              double* acc_weights, void* stream) {
    if (!acc_normal || !acc_weights || M <= 0 || D <= 0 || Kp <= 0) return BEER_ERR_ARG;
    if (comp_off == nullptr && M % Kp != 0) return BEER_ERR_ARG;
    mixture_weight_stats_kernel<<<(Kp * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(acc_normal, M, D, comp_off, Kp,
                                                                                    acc_weights);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_path_accumulate_mix(const float* X, int64_t N, int D, const int32_t* pdf_ids, const float* W, const float* bias,
                             int M, int C, const float* frame_ref, float scale, double* acc_normal, float* frame_exp_llh,
                             void* stream) {
    if (!X || !pdf_ids || !W || !bias || !acc_normal || N < 0 || M <= 0 || C <= 0 || M % C != 0) return BEER_ERR_ARG;
    if (D <= 0 || D > 64) return BEER_ERR_UNSUPPORTED;
    if (N == 0) return BEER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (C) {
        case 1: return launch_path_mix<1>(X, N, D, pdf_ids, W, bias, frame_ref, scale, acc_normal, frame_exp_llh, st);
        case 2: return launch_path_mix<2>(X, N, D, pdf_ids, W, bias, frame_ref, scale, acc_normal, frame_exp_llh, st);
        case 4: return launch_path_mix<4>(X, N, D, pdf_ids, W, bias, frame_ref, scale, acc_normal, frame_exp_llh, st);
        case 8: return launch_path_mix<8>(X, N, D, pdf_ids, W, bias, frame_ref, scale, acc_normal, frame_exp_llh, st);
        case 16: return launch_path_mix<16>(X, N, D, pdf_ids, W, bias, frame_ref, scale, acc_normal, frame_exp_llh, st);
    }
    return BEER_ERR_UNSUPPORTED;
}

}  // extern "C"
