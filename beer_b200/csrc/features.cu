// KF -- fbank front-end: pre-emphasis, framing, Hamming window, |rFFT|, mel filterbank, log(1 + .),
// and the delta regression filter.  One warp per frame: the frame is windowed into shared memory,
// transformed by a radix-2 decimation-in-frequency FFT (fp32, twiddles from an exact table built
// with sincospi), and the filterbank is a [bins x filters] product read coalesced.
//
// Reference semantics: beer/features.py:145-204 (fbank), 82-100 (add_deltas).  Not on the timed
// VB path (the metric uses synthetic fbank frames); SURVEY section 8(f) "next".
#include "common.cuh"
#include "../../include/beer_b200.h"

namespace beer {

constexpr int KF_WARPS = 4;

template <int N>   // FFT length (power of two)
__global__ void __launch_bounds__(KF_WARPS * 32) fbank_kernel(const float* __restrict__ sig, int64_t L, int nframes,
                                                              int flen, int fshift, float preemph,
                                                              const float* __restrict__ window,
                                                              const float* __restrict__ filtT, int nfilt,
                                                              float* __restrict__ out, float dc, int frame_preemph,
                                                              float log_offset) {
    constexpr int LOGN = (N == 256) ? 8 : (N == 512 ? 9 : 10);
    __shared__ float2 tw[N / 2];
    __shared__ float2 xs[KF_WARPS][N];
    __shared__ float mag[KF_WARPS][N / 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < N / 2; k += blockDim.x) {
        float s, c;
        sincospif(2.f * (float)k / (float)N, &s, &c);
        tw[k] = make_float2(c, -s);
    }
    __syncthreads();
    float2* x = xs[warp];
    for (int f = blockIdx.x * KF_WARPS + warp; f < nframes; f += gridDim.x * KF_WARPS) {
        const int64_t s0 = (int64_t)f * fshift;
        // pre-emphasis runs over the whole signal: sample 0 is filtered against itself (features.py:181-182)
        for (int i = lane; i < N; i += 32) {
            float v = 0.f;
            if (i < flen) {
                const int64_t t = s0 + i;
                // fbank(): pre-emphasis over the whole signal; short_term_mspec(): DC removed, pre-emphasis inside
                // the frame (its first sample against itself, features.py:131-132)
                const float cur = sig[t] - dc;
                const float prev = frame_preemph ? (i > 0 ? sig[t - 1] - dc : cur) : sig[t > 0 ? t - 1 : 0] - dc;
                v = (cur - preemph * prev) * window[i];
            }
            x[i] = make_float2(v, 0.f);
        }
        __syncwarp();
#pragma unroll 1
        for (int half = N / 2; half >= 1; half >>= 1) {
            const int tstep = (N / 2) / half;
            for (int b = lane; b < N / 2; b += 32) {
                const int r = b & (half - 1);
                const int i = ((b - r) << 1) + r, j = i + half;
                const float2 a = x[i], c = x[j], w = tw[r * tstep];
                x[i] = make_float2(a.x + c.x, a.y + c.y);
                const float dx = a.x - c.x, dy = a.y - c.y;
                x[j] = make_float2(dx * w.x - dy * w.y, dx * w.y + dy * w.x);
            }
            __syncwarp();
        }
        for (int k = lane; k < N / 2; k += 32) {
            const int rk = (int)(__brev((unsigned)k) >> (32 - LOGN));   // DIF leaves the spectrum bit-reversed
            const float2 v = x[rk];
            mag[warp][k] = sqrtf(v.x * v.x + v.y * v.y);
        }
        __syncwarp();
        if (filtT == nullptr) {          // magnitude spectrum itself (short_term_mspec)
            for (int k = lane; k < N / 2; k += 32) out[(size_t)f * (N / 2) + k] = mag[warp][k];
        } else {
            for (int m = lane; m < nfilt; m += 32) {
                float acc = 0.f;
                for (int k = 0; k < N / 2; ++k) acc = fmaf(mag[warp][k], __ldg(filtT + (size_t)k * nfilt + m), acc);
                out[(size_t)f * nfilt + m] = (log_offset == 1.f) ? log1pf(acc) : logf(log_offset + acc);
            }
        }
        __syncwarp();
    }
}

// y[t] = sum_{k=-w..w} k / (2 sum k^2) * x[clamp(t + k)]  (features.py:93-99)
__global__ void deltas_kernel(const float* __restrict__ fea, int T, int F, int wlen, float* __restrict__ out) {
    float norm = 0.f;
    for (int k = 1; k <= wlen; ++k) norm += 2.f * k * k;
    norm = 1.f / (2.f * norm);
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < (int64_t)T * F;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(e / F), d = (int)(e - (int64_t)t * F);
        float acc = 0.f;
        for (int k = 1; k <= wlen; ++k) {
            const int tp = min(t + k, T - 1), tm = max(t - k, 0);
            acc += (float)k * (fea[(size_t)tp * F + d] - fea[(size_t)tm * F + d]);
        }
        out[e] = acc * norm;
    }
}

}  // namespace beer

using namespace beer;

extern "C" {

static int launch_fbank(const float* signal, int64_t n_samples, int frame_len, int frame_shift, float preemph,
                        const float* window, const float* filters_t, int fft_len, int n_filters, float* out, float dc,
                        int frame_preemph, float log_offset, void* stream) {
    if (!signal || !window || !out || frame_len <= 0 || frame_shift <= 0) return BEER_ERR_ARG;
    if (filters_t != nullptr && n_filters <= 0) return BEER_ERR_ARG;
    if (frame_len > fft_len) return BEER_ERR_ARG;
    if (n_samples < frame_len) return BEER_OK;
    const int nframes = (int)((n_samples - frame_len) / frame_shift + 1);
    int blocks = (nframes + KF_WARPS - 1) / KF_WARPS;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    cudaStream_t st = (cudaStream_t)stream;
#define BEER_FBANK_CASE(n)                                                                                             \
    case n:                                                                                                            \
        fbank_kernel<n><<<blocks, KF_WARPS * 32, 0, st>>>(signal, n_samples, nframes, frame_len, frame_shift, preemph, \
                                                          window, filters_t, n_filters, out, dc, frame_preemph,       \
                                                          log_offset);                                                 \
        break;
    switch (fft_len) {
        BEER_FBANK_CASE(256)
        BEER_FBANK_CASE(512)
        BEER_FBANK_CASE(1024)
        default: return BEER_ERR_UNSUPPORTED;
    }
#undef BEER_FBANK_CASE
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_fbank(const float* signal, int64_t n_samples, int frame_len, int frame_shift, float preemph,
               const float* window, const float* filters_t, int fft_len, int n_filters, float* out, void* stream) {
    if (!filters_t) return BEER_ERR_ARG;
    return launch_fbank(signal, n_samples, frame_len, frame_shift, preemph, window, filters_t, fft_len, n_filters, out,
                        0.f, 0, 1.f, stream);
}

int beer_short_term_mspec(const float* signal, int64_t n_samples, int frame_len, int frame_shift, float preemph,
                          float dc_offset, const float* window, const float* filters_t, int fft_len, int n_filters,
                          float log_offset, float* out, void* stream) {
    if (filters_t != nullptr && !(log_offset > 0.f)) return BEER_ERR_ARG;
    return launch_fbank(signal, n_samples, frame_len, frame_shift, preemph, window, filters_t, fft_len, n_filters, out,
                        dc_offset, 1, log_offset, stream);
}

int beer_add_deltas(const float* fea, int n_frames, int dim, int wlen, float* out, void* stream) {
    if (!fea || !out || n_frames < 0 || dim <= 0 || wlen <= 0) return BEER_ERR_ARG;
    if (n_frames == 0) return BEER_OK;
    int64_t total = (int64_t)n_frames * dim;
    int blocks = (int)((total + 255) / 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    deltas_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(fea, n_frames, dim, wlen, out);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

}  // extern "C"
