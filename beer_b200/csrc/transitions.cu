// Transition posteriors of CompiledGraph.posteriors(trans_posteriors=True) (beer/graph.py:308-323)
// and the sub-block of them that BigramPhoneLoop.accumulate keeps (beer/models/phoneloop.py:175-186).
//
//   xi_t[i, j] = exp(la_t[i] + ln A[i, j] + p_{t+1}[j] + lb_{t+1}[j] - lognorm_t),  normalised per step, NaN -> 0
//
// Written without the backward variables: with gamma the state posteriors the scan kernel already
// produced,  p_{t+1,j} beta_{t+1,j} = gamma_{t+1,j} / sum_i alpha_t[i] A[i,j]  up to a per-step constant, so
//
//   xi_t[i, j] = gamma_{t+1}[j] * alpha_t[i] A[i,j] / sum_i' alpha_t[i'] A[i',j]
//
// which sums to one per step like the reference's normalisation (0/0 for an unreachable column -> 0, the
// reference's NaN -> 0).  Only a forward recursion is run here (log domain, renormalised every frame).
// The output is O(T R C) by definition; this is the API-parity / small-model path.  PhoneLoop training
// does not come here: its reduction of xi is fused into the scan kernel (scan.cu).
//
// One CTA per utterance; a thread owns destination states j, j + 256, ... (reads of ln A coalesced in j).
#include "common.cuh"
#include "../../include/beer_b200.h"

namespace beer {
namespace {

constexpr int XI_THREADS = 256;

struct XiArgs {
    const float* pdf_llh;      // [N, ld]
    int64_t ld;
    const int32_t* pdf_map;    // [K] or null (identity)
    float scale;
    const float* log_init;     // [K]
    const float* log_trans;    // [K, K]
    const float* gamma;        // [N, K] state posteriors
    const int64_t* utt_off;    // [n_utts + 1]
    const int32_t* rows;       // [R] source states kept, or null (all K)
    const int32_t* cols;       // [C] destination states kept, or null (all K)
    int K, R, C;
    float* xi;                 // [N - n_utts, R, C]: utterance u starts at row (utt_off[u] - u)
};

__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float m = red[0];
#pragma unroll
    for (int w = 1; w < XI_THREADS / 32; ++w) m = fmaxf(m, red[w]);
    return m;
}

__global__ void __launch_bounds__(XI_THREADS) hmm_xi_kernel(XiArgs a) {
    extern __shared__ float sm[];
    float* alpha = sm;             // [K] ln alpha_t, max = 0
    float* den = alpha + a.K;      // [K] ln sum_i alpha_t[i] A[i, j]
    __shared__ float red[XI_THREADS / 32];
    const int K = a.K, tid = threadIdx.x;
    const int u = blockIdx.x;
    const int64_t t0 = a.utt_off[u];
    const int T = (int)(a.utt_off[u + 1] - t0);
    if (T <= 1) return;
    float* out = a.xi + (size_t)(t0 - u) * a.R * a.C;

    auto llh = [&](int t, int j) {
        const int col = a.pdf_map ? a.pdf_map[j] : j;
        return a.scale * a.pdf_llh[(size_t)(t0 + t) * a.ld + col];
    };
    float m = kNegInf;
    for (int j = tid; j < K; j += XI_THREADS) {
        alpha[j] = llh(0, j) + a.log_init[j];
        m = fmaxf(m, alpha[j]);
    }
    m = block_max(m, red);
    for (int j = tid; j < K; j += XI_THREADS) alpha[j] = (m == kNegInf) ? alpha[j] : alpha[j] - m;
    __syncthreads();

    for (int t = 0; t + 1 < T; ++t) {
        // den_j = ln sum_i alpha_t[i] A[i, j] (streaming log-sum-exp)
        for (int j = tid; j < K; j += XI_THREADS) {
            float mx = kNegInf, s = 0.f;
            for (int i = 0; i < K; ++i) {
                const float v = alpha[i] + a.log_trans[(size_t)i * K + j];
                if (v > mx) {
                    s = s * expf(mx - v) + 1.f;      // mx = -inf: s = 0 * 0 + 1
                    mx = v;
                } else if (v != kNegInf) {
                    s += expf(v - mx);
                }
            }
            den[j] = (mx == kNegInf) ? kNegInf : mx + logf(s);
        }
        __syncthreads();
        // xi_t[r, c] for the kept rows / columns
        const float* g1 = a.gamma + (size_t)(t0 + t + 1) * K;
        float* o = out + (size_t)t * a.R * a.C;
        for (int e = tid; e < a.R * a.C; e += XI_THREADS) {
            const int r = e / a.C, c = e - r * a.C;
            const int i = a.rows ? a.rows[r] : r, j = a.cols ? a.cols[c] : c;
            const float d = den[j], g = g1[j];
            float v = 0.f;
            if (d != kNegInf && g > 0.f) v = g * expf(alpha[i] + a.log_trans[(size_t)i * K + j] - d);
            o[e] = v;
        }
        // alpha_{t+1}
        float mloc = kNegInf;
        float nxt[4];          // K <= 1024
        int n = 0;
        for (int j = tid; j < K; j += XI_THREADS, ++n) {
            nxt[n] = llh(t + 1, j) + den[j];
            mloc = fmaxf(mloc, nxt[n]);
        }
        mloc = block_max(mloc, red);     // (its barriers also order the reads of alpha above with the writes below)
        n = 0;
        for (int j = tid; j < K; j += XI_THREADS, ++n) alpha[j] = (mloc == kNegInf) ? nxt[n] : nxt[n] - mloc;
        __syncthreads();
    }
}

}  // namespace
}  // namespace beer

using namespace beer;

extern "C" int beer_hmm_transition_posteriors(const float* pdf_llh, int64_t ld_pdf, const int32_t* pdf_map, float scale,
                                              const float* log_init, const float* log_trans, int K,
                                              const float* state_post, const int64_t* utt_off, int n_utts,
                                              const int32_t* rows, int n_rows, const int32_t* cols, int n_cols,
                                              float* xi, void* stream) {
    if (!pdf_llh || !log_init || !log_trans || !state_post || !utt_off || !xi) return BEER_ERR_ARG;
    if (K <= 0 || n_utts < 0 || ld_pdf <= 0) return BEER_ERR_ARG;
    if (K > 4 * XI_THREADS) return BEER_ERR_UNSUPPORTED;
    if ((rows == nullptr) != (n_rows == 0) || (cols == nullptr) != (n_cols == 0)) return BEER_ERR_ARG;
    if (n_utts == 0) return BEER_OK;
    XiArgs a;
    a.pdf_llh = pdf_llh; a.ld = ld_pdf; a.pdf_map = pdf_map; a.scale = scale; a.log_init = log_init;
    a.log_trans = log_trans; a.gamma = state_post; a.utt_off = utt_off; a.rows = rows; a.cols = cols;
    a.K = K; a.R = rows ? n_rows : K; a.C = cols ? n_cols : K; a.xi = xi;
    hmm_xi_kernel<<<n_utts, XI_THREADS, 2 * K * sizeof(float), (cudaStream_t)stream>>>(a);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}
