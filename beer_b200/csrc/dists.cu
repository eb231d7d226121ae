// Conjugate exponential-family parameter math on the device: expected natural
// statistics, emission weights, natural-gradient M-step and KL divergences.
// All of it is O(M*D) per VB iteration (< 0.1 % of the E-step), so it is computed
// in fp64 from the fp32 standard parameters and rounded once.
//
// Reference semantics: beer/dists/normalgamma.py, beer/dists/dirichlet.py,
// beer/dists/basedist.py:243-263, beer/models/parameters.py:134-141.
#include <algorithm>
#include "common.cuh"
#include "../../include/beer_b200.h"

namespace beer {

constexpr double kLog2Pi = 1.8378770664093453;

// ---- Normal-Gamma ----------------------------------------------------------

// One warp per Gaussian.  ets row = [lam*m (D), lam (D), D/k + sum lam m^2, sum psi(a) - ln b]
__global__ void ng_expected_stats_kernel(const float* __restrict__ mean, const float* __restrict__ scale,
                                         const float* __restrict__ shape, const float* __restrict__ rates,
                                         int M, int D, float* __restrict__ ets) {
    int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (j >= M) return;
    double a = shape[j], k = scale[j];
    double psi_a = digamma_d(a);
    double pqm = 0.0, logdet = 0.0;
    int Q = 2 * D + 2;
    for (int d = lane; d < D; d += 32) {
        double m = mean[(size_t)j * D + d], b = rates[(size_t)j * D + d];
        double lam = a / b;
        ets[(size_t)j * Q + d] = (float)(lam * m);
        ets[(size_t)j * Q + D + d] = (float)lam;
        pqm += lam * m * m;
        logdet += psi_a - log(b);
    }
    pqm = warp_sum(pqm);
    logdet = warp_sum(logdet);
    if (lane == 0) {
        ets[(size_t)j * Q + 2 * D] = (float)(pqm + D / k);
        ets[(size_t)j * Q + 2 * D + 1] = (float)logdet;
    }
}

// Per-Gaussian bias before centring (fp64), one warp per Gaussian.
__device__ __forceinline__ double ng_bias_row(const float* mean, const float* scale, const float* shape,
                                              const float* rates, const float* logw, int j, int D, int lane) {
    double a = shape[j], k = scale[j];
    double psi_a = digamma_d(a);
    double acc = 0.0;
    for (int d = lane; d < D; d += 32) {
        double m = mean[(size_t)j * D + d], b = rates[(size_t)j * D + d];
        acc += -0.5 * (a / b) * m * m + 0.5 * (psi_a - log(b));
    }
    acc = warp_sum(acc);
    acc += -0.5 * D / k - 0.5 * D * kLog2Pi;
    if (logw != nullptr) acc += (double)logw[j];
    return acc;
}

// Single block: ref[d] = mean_j lam_jd, ref[D] = mean_j bias_j.
__global__ void emission_ref_kernel(const float* __restrict__ mean, const float* __restrict__ scale,
                                    const float* __restrict__ shape, const float* __restrict__ rates,
                                    const float* __restrict__ logw, int M, int D, float* __restrict__ ref) {
    __shared__ double s_bias[32];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        double acc = 0.0;
        for (int j = 0; j < M; ++j) acc += (double)shape[j] / (double)rates[(size_t)j * D + d];
        ref[d] = (float)(acc / M);
    }
    double b = 0.0;
    for (int j = warp; j < M; j += nwarp) b += ng_bias_row(mean, scale, shape, rates, logw, j, D, lane);
    if (lane == 0) s_bias[warp] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < nwarp; ++w) t += s_bias[w];
        ref[D] = (float)(t / M);
    }
}

// Same reference for many Gaussians (the single block above takes 2.4 ms at M = 8000): per-block partial
// sums in a fixed order into a scratch area, then one block adds the partials in a fixed order, so the
// result does not depend on scheduling.  part [n_blocks][D + 1] (fp64).
__global__ void emission_ref_partial_kernel(const float* __restrict__ mean, const float* __restrict__ scale,
                                            const float* __restrict__ shape, const float* __restrict__ rates,
                                            const float* __restrict__ logw, int M, int D, double* __restrict__ part) {
    __shared__ double s_bias[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int per = (M + gridDim.x - 1) / gridDim.x, j0 = blockIdx.x * per, j1 = min(M, j0 + per);
    double* out = part + (size_t)blockIdx.x * (D + 1);
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        double acc = 0.0;
        for (int j = j0; j < j1; ++j) acc += (double)shape[j] / (double)rates[(size_t)j * D + d];
        out[d] = acc;
    }
    double b = 0.0;
    for (int j = j0 + warp; j < j1; j += nwarp) b += ng_bias_row(mean, scale, shape, rates, logw, j, D, lane);
    if (lane == 0) s_bias[warp] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < nwarp; ++w) t += s_bias[w];
        out[D] = t;
    }
}

__global__ void emission_ref_finish_kernel(const double* __restrict__ part, int n_blocks, int M, int D,
                                           float* __restrict__ ref) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d > D) return;
    double t = 0.0;
    for (int b = 0; b < n_blocks; ++b) t += part[(size_t)b * (D + 1) + d];
    ref[d] = (float)(t / M);
}

__global__ void emission_weights_kernel(const float* __restrict__ mean, const float* __restrict__ scale,
                                        const float* __restrict__ shape, const float* __restrict__ rates,
                                        const float* __restrict__ logw, int M, int D,
                                        const float* __restrict__ ref, float* __restrict__ W,
                                        float* __restrict__ bias) {
    int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (j >= M) return;
    double a = shape[j];
    for (int d = lane; d < D; d += 32) {
        double m = mean[(size_t)j * D + d], b = rates[(size_t)j * D + d];
        double lam = a / b;
        W[(size_t)j * 2 * D + d] = (float)(lam * m);
        W[(size_t)j * 2 * D + D + d] = (float)(lam - (double)ref[d]);
    }
    double br = ng_bias_row(mean, scale, shape, rates, logw, j, D, lane);
    if (lane == 0) bias[j] = (float)(br - (double)ref[D]);
}

// eta <- eta + lr (eta0 + s*acc - eta) per Gaussian, then back to standard form.
__global__ void ng_update_kernel(const float* __restrict__ pm, const float* __restrict__ pk,
                                 const float* __restrict__ pa, const float* __restrict__ pb, float* mean,
                                 float* scale, float* shape, float* rates, const double* __restrict__ acc,
                                 double s, double lr, int M, int D) {
    int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (j >= M) return;
    int Q = 2 * D + 2;
    const double* st = acc + (size_t)j * Q;
    double k0 = pk[j], a0 = pa[j], k = scale[j], a = shape[j];
    double e3 = -0.5 * k, e4 = a - 0.5;
    e3 += lr * (-0.5 * k0 + s * st[2 * D] - e3);
    e4 += lr * (a0 - 0.5 + s * st[2 * D + 1] - e4);
    double kn = -2.0 * e3, an = e4 + 0.5;
    for (int d = lane; d < D; d += 32) {
        size_t i = (size_t)j * D + d;
        double m0 = pm[i], b0 = pb[i], m = mean[i], b = rates[i];
        double e1 = k * m, e2 = -0.5 * k * m * m - b;
        e1 += lr * (k0 * m0 + s * st[d] - e1);
        e2 += lr * (-0.5 * k0 * m0 * m0 - b0 + s * st[D + d] - e2);
        double mn = e1 / kn;
        mean[i] = (float)mn;
        rates[i] = (float)(-e2 - 0.5 * kn * mn * mn);
    }
    __syncwarp();
    if (lane == 0) {
        scale[j] = (float)kn;
        shape[j] = (float)an;
    }
}

// KL(q || p) = A(p) - A(q) - <E_q[T], eta_p - eta_q>, summed over Gaussians.
__global__ void ng_kl_kernel(const float* __restrict__ pm, const float* __restrict__ pk,
                             const float* __restrict__ pa, const float* __restrict__ pb,
                             const float* __restrict__ mean, const float* __restrict__ scale,
                             const float* __restrict__ shape, const float* __restrict__ rates, int M, int D,
                             double* kl) {
    int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    double val = 0.0;
    if (j < M) {
        double k0 = pk[j], a0 = pa[j], k = scale[j], a = shape[j];
        double psi_a = digamma_d(a);
        double sum_lnb = 0.0, sum_lnb0 = 0.0, pqm = 0.0, logdet = 0.0, dot = 0.0;
        for (int d = lane; d < D; d += 32) {
            size_t i = (size_t)j * D + d;
            double m0 = pm[i], b0 = pb[i], m = mean[i], b = rates[i];
            double lam = a / b;
            sum_lnb += log(b);
            sum_lnb0 += log(b0);
            pqm += lam * m * m;
            logdet += psi_a - log(b);
            // E[T]_1 * (eta_p1 - eta_q1) + E[T]_2 * (eta_p2 - eta_q2)
            dot += lam * m * (k0 * m0 - k * m) + lam * ((-0.5 * k0 * m0 * m0 - b0) - (-0.5 * k * m * m - b));
        }
        sum_lnb = warp_sum(sum_lnb);
        sum_lnb0 = warp_sum(sum_lnb0);
        pqm = warp_sum(pqm);
        logdet = warp_sum(logdet);
        dot = warp_sum(dot);
        double Aq = D * lgamma(a) - a * sum_lnb - 0.5 * D * log(k);
        double Ap = D * lgamma(a0) - a0 * sum_lnb0 - 0.5 * D * log(k0);
        dot += (pqm + D / k) * (-0.5 * k0 + 0.5 * k) + logdet * (a0 - a);
        val = Ap - Aq - dot;
    }
    // block reduction -> one atomic per block
    __shared__ double s_val[32];
    int warp = threadIdx.x >> 5;
    if (lane == 0) s_val[warp] = val;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_val[w];
        atomicAdd(kl, t);
    }
}

// ---- Dirichlet -------------------------------------------------------------

// One thread per row (K rows, C small).
// One WARP per Dirichlet (row): a thread per row made a 512-component GMM (K = 1) a single-thread loop over 512
// digamma / lgamma evaluations in fp64 (0.6 ms).
__global__ void dir_logw_kernel(const float* __restrict__ conc, int K, int C, float* __restrict__ logw) {
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= K) return;
    double tot = 0.0;
    for (int c = lane; c < C; c += 32) tot += conc[(size_t)k * C + c];
    tot = warp_sum(tot);
    const double psi_tot = digamma_d(tot);
    for (int c = lane; c < C; c += 32)
        logw[(size_t)k * C + c] = (float)(digamma_d((double)conc[(size_t)k * C + c]) - psi_tot);
}

__global__ void dir_update_kernel(const float* __restrict__ prior, float* conc, const double* __restrict__ acc,
                                  double s, double lr, int K, int C) {
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= K) return;
    const float* p = prior + (size_t)k * C;
    float* q = conc + (size_t)k * C;
    const double* st = acc + (size_t)k * C;
    double sum_p = 0.0, sum_q = 0.0, sum_new = 0.0;
    for (int c = lane; c < C; c += 32) {
        sum_p += (double)p[c] - 1.0;
        sum_q += (double)q[c] - 1.0;
    }
    sum_p = warp_sum(sum_p);
    sum_q = warp_sum(sum_q);        // every lane has read the old row before anyone writes it
    const double e_last = sum_q + lr * (sum_p + s * st[C - 1] - sum_q);
    for (int c = lane; c < C - 1; c += 32) {
        double e = (double)q[c] - 1.0;
        e += lr * ((double)p[c] - 1.0 + s * st[c] - e);
        sum_new += e;
        q[c] = (float)(e + 1.0);
    }
    sum_new = warp_sum(sum_new);
    if (lane == 0) q[C - 1] = (float)(e_last - sum_new + 1.0);
}

__global__ void dir_kl_kernel(const float* __restrict__ prior, const float* __restrict__ conc, int K, int C,
                              double* kl) {
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= K) return;
    const float* p = prior + (size_t)k * C;
    const float* q = conc + (size_t)k * C;
    double sp = 0.0, sq = 0.0, A = 0.0;
    for (int c = lane; c < C; c += 32) {
        sp += p[c];
        sq += q[c];
        A += lgamma((double)p[c]) - lgamma((double)q[c]);
    }
    sp = warp_sum(sp);
    sq = warp_sum(sq);
    const double psi_last = digamma_d((double)q[C - 1]), psi_tot = digamma_d(sq);
    double dot = 0.0;
    for (int c = lane; c < C - 1; c += 32) dot += (digamma_d((double)q[c]) - psi_last) * ((double)p[c] - (double)q[c]);
    double val = warp_sum(A - dot);
    val += -lgamma(sp) + lgamma(sq) - (psi_last - psi_tot) * ((sp - C) - (sq - C));
    if (lane == 0 && val != 0.0) atomicAdd(kl, val);
}


// ---- API-completeness kernels (beer.dists accessors; not on the VB iteration's hot path) -------

// [x, -x^2/2, -1/2, 1/2] per frame (beer/dists/normalgamma.py:19-27).
__global__ void normal_suff_stats_kernel(const float* __restrict__ X, int64_t N, int D, float* __restrict__ out) {
    const int Q = 2 * D + 2;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < N * Q; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = e / Q;
        const int c = (int)(e - t * Q);
        float v;
        if (c < D) v = X[t * D + c];
        else if (c < 2 * D) { const float x = X[t * D + c - D]; v = -0.5f * x * x; }
        else v = (c == 2 * D) ? -0.5f : 0.5f;
        out[e] = v;
    }
}

// eta = [k m, -k m^2 / 2 - b, -k / 2, a - 1/2] (normalgamma.py:163-180), one warp per Gaussian.
__global__ void ng_natural_kernel(const float* __restrict__ mean, const float* __restrict__ scale,
                                  const float* __restrict__ shape, const float* __restrict__ rates, int M, int D,
                                  float* __restrict__ out) {
    int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= M) return;
    const int Q = 2 * D + 2;
    const double k = scale[j], a = shape[j];
    for (int d = lane; d < D; d += 32) {
        const double m = mean[(size_t)j * D + d], b = rates[(size_t)j * D + d];
        out[(size_t)j * Q + d] = (float)(k * m);
        out[(size_t)j * Q + D + d] = (float)(-0.5 * k * m * m - b);
    }
    if (lane == 0) {
        out[(size_t)j * Q + 2 * D] = (float)(-0.5 * k);
        out[(size_t)j * Q + 2 * D + 1] = (float)(a - 0.5);
    }
}

// inverse map (normalgamma.py:76-94)
__global__ void ng_from_natural_kernel(const float* __restrict__ nat, int M, int D, float* __restrict__ mean,
                                       float* __restrict__ scale, float* __restrict__ shape,
                                       float* __restrict__ rates) {
    int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= M) return;
    const int Q = 2 * D + 2;
    const double k = -2.0 * (double)nat[(size_t)j * Q + 2 * D];
    for (int d = lane; d < D; d += 32) {
        const double m = (double)nat[(size_t)j * Q + d] / k;
        mean[(size_t)j * D + d] = (float)m;
        rates[(size_t)j * D + d] = (float)(-(double)nat[(size_t)j * Q + D + d] - 0.5 * k * m * m);
    }
    if (lane == 0) {
        scale[j] = (float)k;
        shape[j] = (float)((double)nat[(size_t)j * Q + 2 * D + 1] + 0.5);
    }
}

// A(eta) = D lgamma(a) - a sum ln b - D/2 ln k (normalgamma.py:151-157)
__global__ void ng_log_norm_kernel(const float* __restrict__ scale, const float* __restrict__ shape,
                                   const float* __restrict__ rates, int M, int D, double* __restrict__ out) {
    int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= M) return;
    double s = 0.0;
    for (int d = lane; d < D; d += 32) s += log((double)rates[(size_t)j * D + d]);
    s = warp_sum(s);
    const double a = shape[j], k = scale[j];
    if (lane == 0) out[j] = D * lgamma(a) - a * s - 0.5 * D * log(k);
}

// per-Gaussian KL (the vector kl_div returns, basedist.py:243-263): same arithmetic as ng_kl_kernel
__global__ void dir_rows_kernel(const float* __restrict__ conc, int K, int C, int what, float* __restrict__ outf,
                                double* __restrict__ outd) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const float* q = conc + (size_t)k * C;
    double tot = 0.0;
    for (int c = 0; c < C; ++c) tot += q[c];
    if (what == 0) {            // natural parameters [a_i - 1 (i < C-1), sum_i (a_i - 1)] (dirichlet.py:144-159)
        for (int c = 0; c < C - 1; ++c) outf[(size_t)k * C + c] = (float)((double)q[c] - 1.0);
        outf[(size_t)k * C + C - 1] = (float)(tot - C);
    } else if (what == 1) {     // expected statistics, log-odds form (dirichlet.py:106-128)
        const double psi_last = digamma_d((double)q[C - 1]);
        for (int c = 0; c < C - 1; ++c) outf[(size_t)k * C + c] = (float)(digamma_d((double)q[c]) - psi_last);
        outf[(size_t)k * C + C - 1] = (float)(psi_last - digamma_d(tot));
    } else if (what == 2) {     // log-normaliser (dirichlet.py:135-138)
        double A = 0.0;
        for (int c = 0; c < C; ++c) A += lgamma((double)q[c]);
        outd[k] = A - lgamma(tot);
    } else {                    // from natural parameters, in: conc holds eta (dirichlet.py:70-81)
        double s = 0.0;
        for (int c = 0; c < C - 1; ++c) {
            s += q[c];
            outf[(size_t)k * C + c] = (float)((double)q[c] + 1.0);
        }
        outf[(size_t)k * C + C - 1] = (float)((double)q[C - 1] - s + 1.0);
    }
}

// pdf_llh[t, k] = logsumexp_{j in pdf k} comp_llh[t, j]: mixtures wider than one emission tile
__global__ void segment_lse_kernel(const float* __restrict__ comp, int64_t N, int M,
                                   const int32_t* __restrict__ comp_off, int Kp, float* __restrict__ out,
                                   int64_t ld) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = gw; e < N * Kp; e += nw) {
        const int64_t t = e / Kp;
        const int k = (int)(e - t * Kp);
        int c0, c1;
        if (comp_off == nullptr) { const int C = M / Kp; c0 = k * C; c1 = c0 + C; }
        else { c0 = comp_off[k]; c1 = comp_off[k + 1]; }
        const float* row = comp + (size_t)t * M;
        float m = kNegInf;
        for (int c = c0 + lane; c < c1; c += 32) m = fmaxf(m, row[c]);
        m = warp_max(m);
        const float ms = (m == kNegInf) ? 0.f : m;
        float s = 0.f;
        for (int c = c0 + lane; c < c1; c += 32) s += __expf(row[c] - ms);
        s = warp_sum(s);
        if (lane == 0) out[(size_t)t * ld + k] = ms + __logf(s);
    }
}

// Posteriors of a given state path (Viterbi training / forced path, beer/models/hmm.py:42-58, 87):
// pdf_post[t, map[path_t]] = scale, frame_exp_llh[t] = scale * (pdf_llh[t, map[path_t]] + frame_ref[t]).
__global__ void path_posteriors_kernel(const int32_t* __restrict__ path, int64_t N, const int32_t* __restrict__ map,
                                       float scale, const float* __restrict__ pdf_llh, int64_t ld_pdf,
                                       const float* __restrict__ frame_ref, float* __restrict__ pdf_post,
                                       int64_t ld_post, int Kp, float* __restrict__ frame_exp_llh) {
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    if (frame_exp_llh != nullptr) {
        for (int64_t t = tid; t < N; t += nthr) {
            const int s = path[t];
            const int k = (map != nullptr) ? map[s] : s;
            frame_exp_llh[t] = scale * (pdf_llh[(size_t)t * ld_pdf + k] + (frame_ref != nullptr ? frame_ref[t] : 0.f));
        }
    }
    if (pdf_post == nullptr) return;
    // one-hot rows, written coalesced: consecutive threads take consecutive (groups of four) columns of a row
    if (ld_post == Kp && (Kp & 3) == 0 && ((uintptr_t)pdf_post & 15) == 0) {
        const int q = Kp >> 2;
        for (int64_t i = tid; i < N * q; i += nthr) {
            const int64_t t = i / q;
            const int c0 = (int)(i - t * q) * 4;
            const int s = path[t];
            const int k = ((map != nullptr) ? map[s] : s) - c0;
            reinterpret_cast<float4*>(pdf_post)[i] = make_float4(k == 0 ? scale : 0.f, k == 1 ? scale : 0.f,
                                                                  k == 2 ? scale : 0.f, k == 3 ? scale : 0.f);
        }
    } else {
        for (int64_t i = tid; i < N * Kp; i += nthr) {
            const int64_t t = i / Kp;
            const int c = (int)(i - t * Kp);
            const int s = path[t];
            pdf_post[(size_t)t * ld_post + c] = (c == ((map != nullptr) ? map[s] : s)) ? scale : 0.f;
        }
    }
}

}  // namespace beer

using namespace beer;

extern "C" {

int beer_b200_version(void) { return 100; }

int beer_normalgamma_expected_stats(const float* mean, const float* scale, const float* shape,
                                    const float* rates, int M, int D, float* ets, void* stream) {
    if (M <= 0 || D <= 0) return BEER_ERR_ARG;
    int threads = 128, blocks = (M * 32 + threads - 1) / threads;
    ng_expected_stats_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(mean, scale, shape, rates, M, D, ets);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_dirichlet_expected_logw(const float* conc, int K, int C, float* logw, void* stream) {
    if (K <= 0 || C <= 0) return BEER_ERR_ARG;
    dir_logw_kernel<<<(K * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(conc, K, C, logw);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_emission_prepare(const float* mean, const float* scale, const float* shape, const float* rates,
                          const float* logw, int M, int D, float* W, float* bias, float* ref, void* stream) {
    if (M <= 0 || D <= 0) return BEER_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    // many Gaussians: two-stage reduction with W (written only after ref is final) as fp64 scratch
    const int n_blocks = 64;
    if (M >= 256 && (size_t)M * 2 * D * sizeof(float) >= (size_t)n_blocks * (D + 1) * sizeof(double) &&
        ((uintptr_t)W & 7) == 0) {
        double* part = reinterpret_cast<double*>(W);
        emission_ref_partial_kernel<<<n_blocks, 256, 0, st>>>(mean, scale, shape, rates, logw, M, D, part);
        BEER_LAUNCH_CHECK();
        emission_ref_finish_kernel<<<(D + 1 + 127) / 128, 128, 0, st>>>(part, n_blocks, M, D, ref);
        BEER_LAUNCH_CHECK();
    } else {
        emission_ref_kernel<<<1, 1024, 0, st>>>(mean, scale, shape, rates, logw, M, D, ref);
        BEER_LAUNCH_CHECK();
    }
    int threads = 128, blocks = (M * 32 + threads - 1) / threads;
    emission_weights_kernel<<<blocks, threads, 0, st>>>(mean, scale, shape, rates, logw, M, D, ref, W, bias);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_normalgamma_update(const float* prior_mean, const float* prior_scale, const float* prior_shape,
                            const float* prior_rates, float* mean, float* scale, float* shape, float* rates,
                            const double* acc, double stats_scale, double lrate, int M, int D, void* stream) {
    if (M <= 0 || D <= 0) return BEER_ERR_ARG;
    int threads = 128, blocks = (M * 32 + threads - 1) / threads;
    ng_update_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(prior_mean, prior_scale, prior_shape,
                                                                   prior_rates, mean, scale, shape, rates, acc,
                                                                   stats_scale, lrate, M, D);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_normalgamma_kl(const float* prior_mean, const float* prior_scale, const float* prior_shape,
                        const float* prior_rates, const float* mean, const float* scale, const float* shape,
                        const float* rates, int M, int D, double* kl, void* stream) {
    if (M <= 0 || D <= 0) return BEER_ERR_ARG;
    int threads = 128, blocks = (M * 32 + threads - 1) / threads;
    ng_kl_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(prior_mean, prior_scale, prior_shape, prior_rates,
                                                               mean, scale, shape, rates, M, D, kl);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_dirichlet_update(const float* prior_conc, float* conc, const double* acc, double stats_scale,
                          double lrate, int K, int C, void* stream) {
    if (K <= 0 || C <= 0) return BEER_ERR_ARG;
    dir_update_kernel<<<(K * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(prior_conc, conc, acc, stats_scale, lrate,
                                                                         K, C);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_dirichlet_kl(const float* prior_conc, const float* conc, int K, int C, double* kl, void* stream) {
    if (K <= 0 || C <= 0) return BEER_ERR_ARG;
    dir_kl_kernel<<<(K * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(prior_conc, conc, K, C, kl);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_normal_sufficient_statistics(const float* X, int64_t N, int D, float* out, void* stream) {
    if (!X || !out || N < 0 || D <= 0) return BEER_ERR_ARG;
    if (N == 0) return BEER_OK;
    int64_t total = N * (2 * D + 2);
    int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    normal_suff_stats_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(X, N, D, out);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_normalgamma_natural_params(const float* mean, const float* scale, const float* shape, const float* rates,
                                    int M, int D, float* nat, void* stream) {
    if (M <= 0 || D <= 0) return BEER_ERR_ARG;
    ng_natural_kernel<<<(M * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mean, scale, shape, rates, M, D, nat);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_normalgamma_from_natural(const float* nat, int M, int D, float* mean, float* scale, float* shape,
                                  float* rates, void* stream) {
    if (M <= 0 || D <= 0) return BEER_ERR_ARG;
    ng_from_natural_kernel<<<(M * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(nat, M, D, mean, scale, shape,
                                                                                   rates);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_normalgamma_log_norm(const float* scale, const float* shape, const float* rates, int M, int D,
                              double* out, void* stream) {
    if (M <= 0 || D <= 0) return BEER_ERR_ARG;
    ng_log_norm_kernel<<<(M * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(scale, shape, rates, M, D, out);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

static int dir_rows(const float* in, int K, int C, int what, float* outf, double* outd, void* stream) {
    if (!in || K <= 0 || C <= 0) return BEER_ERR_ARG;
    dir_rows_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(in, K, C, what, outf, outd);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}
int beer_dirichlet_natural_params(const float* conc, int K, int C, float* nat, void* stream) {
    return dir_rows(conc, K, C, 0, nat, nullptr, stream);
}
int beer_dirichlet_expected_stats(const float* conc, int K, int C, float* ets, void* stream) {
    return dir_rows(conc, K, C, 1, ets, nullptr, stream);
}
int beer_dirichlet_log_norm(const float* conc, int K, int C, double* out, void* stream) {
    return dir_rows(conc, K, C, 2, nullptr, out, stream);
}
int beer_dirichlet_from_natural(const float* nat, int K, int C, float* conc, void* stream) {
    return dir_rows(nat, K, C, 3, conc, nullptr, stream);
}

int beer_segment_logsumexp(const float* comp_llh, int64_t N, int M, const int32_t* comp_off, int Kp,
                           float* pdf_llh, int64_t ld_pdf, void* stream) {
    if (!comp_llh || !pdf_llh || N < 0 || M <= 0 || Kp <= 0 || ld_pdf < Kp) return BEER_ERR_ARG;
    if (comp_off == nullptr && M % Kp != 0) return BEER_ERR_ARG;
    if (N == 0) return BEER_OK;
    int64_t warps = N * Kp;
    int blocks = (int)std::min<int64_t>((warps + 7) / 8, 148 * 16);
    segment_lse_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(comp_llh, N, M, comp_off, Kp, pdf_llh, ld_pdf);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

int beer_path_posteriors(const int32_t* path, int64_t N, const int32_t* pdf_map, float scale, const float* pdf_llh,
                         int64_t ld_pdf, const float* frame_ref, float* pdf_post, int64_t ld_post, int Kp,
                         float* frame_exp_llh, void* stream) {
    if (!path || N < 0 || Kp <= 0) return BEER_ERR_ARG;
    if (frame_exp_llh != nullptr && pdf_llh == nullptr) return BEER_ERR_ARG;
    if (N == 0) return BEER_OK;
    int blocks = (int)std::min<int64_t>((N + 255) / 256, 148 * 8);
    path_posteriors_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(path, N, pdf_map, scale, pdf_llh, ld_pdf,
                                                                     frame_ref, pdf_post, ld_post, Kp, frame_exp_llh);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

}  // extern "C"
