// KA -- per-frame, per-pdf expected log-likelihood of diagonal Gaussians
// (NormalSet / MixtureSet), offset form.
//
//   llh_tj - r_t = [x_t, -x_t^2/2] . W_j + bias_j          (GEMM-shaped, K-dim = 2D)
//   pdf_llh[t,k] = logsumexp_{j in pdf k} (llh_tj - r_t)    (mixture epilogue)
//
// Reference semantics: beer/dists/normalgamma.py:19-27, 55-59;
// beer/models/normalset.py:117-119; beer/models/mixtureset.py:85-98.
//
// This file holds the fp32 SIMT register-tiled version: a persistent CTA keeps its
// 128-Gaussian weight tile in shared memory and streams 64-frame tiles of X
// through it.
#include "common.cuh"
#include "../../include/beer_b200.h"

namespace beer {

constexpr int KA_FR = 64;      // frames per tile
constexpr int KA_G = 128;      // Gaussians per tile
constexpr int KA_THREADS = 256;
constexpr int KA_AS_LD = 68;   // padded frame stride of the statistics tile
constexpr int KA_CS_LD = 129;  // padded Gaussian stride of the epilogue tile

struct KaTile {
    int g0, g1;  // Gaussian range
    int k0, k1;  // pdf range
};

// Greedy packing of whole pdfs into tiles of <= KA_G Gaussians.
__device__ inline KaTile ka_find_tile(const int32_t* comp_off, int Kp, int M, int tile) {
    KaTile t;
    if (comp_off == nullptr) {
        int C = M / Kp;
        int per = KA_G / C;
        t.k0 = tile * per;
        t.k1 = min(Kp, t.k0 + per);
        t.g0 = t.k0 * C;
        t.g1 = t.k1 * C;
        return t;
    }
    int k = 0, cur = 0;
    while (true) {
        int k0 = k, g0 = comp_off[k];
        while (k < Kp && comp_off[k + 1] - g0 <= KA_G) ++k;
        if (cur == tile || k >= Kp) {
            t.k0 = k0; t.k1 = k; t.g0 = g0; t.g1 = comp_off[k];
            return t;
        }
        ++cur;
    }
}

__global__ void __launch_bounds__(KA_THREADS, 2)
emission_llh_simt_kernel(const float* __restrict__ X, int64_t N, int D, const float* __restrict__ W,
                         const float* __restrict__ bias, const float* __restrict__ ref, int M,
                         const int32_t* __restrict__ comp_off, int Kp, int n_gtiles,
                         float* __restrict__ pdf_llh, int64_t ld_pdf, float* __restrict__ comp_llh,
                         float* __restrict__ frame_ref) {
    extern __shared__ __align__(16) float smem[];
    const int D2 = 2 * D;
    float* Bs = smem;                          // [D2][KA_G]
    float* As = Bs + (size_t)D2 * KA_G;        // [D2][KA_AS_LD]
    float* Cs = As + (size_t)D2 * KA_AS_LD;    // [KA_FR][KA_CS_LD]
    float* s_bias = Cs + KA_FR * KA_CS_LD;     // [KA_G]
    float* s_ref = s_bias + KA_G;              // [D + 1]
    __shared__ KaTile s_tile;

    const int tid = threadIdx.x;
    const int gtile = blockIdx.x % n_gtiles;
    const int fslot = blockIdx.x / n_gtiles;
    const int fstride = gridDim.x / n_gtiles;
    if (fslot >= fstride) return;  // leftover CTAs when gridDim is not a multiple of n_gtiles

    if (tid == 0) s_tile = ka_find_tile(comp_off, Kp, M, gtile);
    __syncthreads();
    const KaTile tile = s_tile;
    const int ng = tile.g1 - tile.g0;

    // weight tile, transposed to [feature][gaussian]
    for (int e = tid; e < KA_G * D2; e += KA_THREADS) {
        int g = e % KA_G, kk = e / KA_G;
        Bs[kk * KA_G + g] = (g < ng) ? W[(size_t)(tile.g0 + g) * D2 + kk] : 0.f;
    }
    for (int g = tid; g < KA_G; g += KA_THREADS) s_bias[g] = (g < ng) ? bias[tile.g0 + g] : 0.f;
    for (int d = tid; d <= D; d += KA_THREADS) s_ref[d] = ref[d];

    const int tx = tid & 15, ty = tid >> 4;
    const int64_t n_ftiles = (N + KA_FR - 1) / KA_FR;

    for (int64_t ft = fslot; ft < n_ftiles; ft += fstride) {
        const int64_t t0 = ft * KA_FR;
        const int nf = (int)min((int64_t)KA_FR, N - t0);
        __syncthreads();  // previous tile's readers of As / Cs are done
        for (int e = tid; e < KA_FR * D; e += KA_THREADS) {
            int f = e / D, d = e - f * D;
            float x = (f < nf) ? X[(size_t)t0 * D + e] : 0.f;
            As[d * KA_AS_LD + f] = x;
            As[(D + d) * KA_AS_LD + f] = -0.5f * x * x;
        }
        __syncthreads();

        if (gtile == 0 && frame_ref != nullptr && tid < nf) {
            float r = 0.f;
            for (int d = 0; d < D; ++d) r = fmaf(As[(D + d) * KA_AS_LD + tid], s_ref[d], r);
            frame_ref[t0 + tid] = r + s_ref[D];
        }

        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

#pragma unroll 4
        for (int kk = 0; kk < D2; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&As[kk * KA_AS_LD + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk * KA_G + tx * 8]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk * KA_G + tx * 8 + 4]);
            float av[4] = {a.x, a.y, a.z, a.w};
            float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }

#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                Cs[(ty * 4 + i) * KA_CS_LD + tx * 8 + j] = acc[i][j] + s_bias[tx * 8 + j];
        __syncthreads();

        if (comp_llh != nullptr) {
            for (int e = tid; e < nf * ng; e += KA_THREADS) {
                int f = e / ng, g = e - f * ng;
                comp_llh[(size_t)(t0 + f) * M + tile.g0 + g] = Cs[f * KA_CS_LD + g];
            }
        }
        const int nk = tile.k1 - tile.k0;
        for (int e = tid; e < nf * nk; e += KA_THREADS) {
            int f = e / nk, k = e - f * nk;
            int c0, c1;
            if (comp_off == nullptr) {
                int C = M / Kp;
                c0 = k * C; c1 = c0 + C;
            } else {
                c0 = comp_off[tile.k0 + k] - tile.g0;
                c1 = comp_off[tile.k0 + k + 1] - tile.g0;
            }
            const float* row = Cs + f * KA_CS_LD;
            float m = row[c0];
            for (int c = c0 + 1; c < c1; ++c) m = fmaxf(m, row[c]);
            float out = m;
            if (c1 - c0 > 1) {
                float s = 0.f;
                for (int c = c0; c < c1; ++c) s += __expf(row[c] - m);
                out = m + __logf(s);
            }
            pdf_llh[(size_t)(t0 + f) * ld_pdf + tile.k0 + k] = out;
        }
    }
}

// Host-side mirror of ka_find_tile's tile count.
static int ka_count_tiles(const int32_t* comp_off_host, int Kp, int M) {
    if (comp_off_host == nullptr) {
        int C = M / Kp;
        int per = KA_G / C;
        return (Kp + per - 1) / per;
    }
    int k = 0, n = 0;
    while (k < Kp) {
        int g0 = comp_off_host[k];
        int k_start = k;
        while (k < Kp && comp_off_host[k + 1] - g0 <= KA_G) ++k;
        if (k == k_start) return -1;  // a pdf with more than KA_G components
        ++n;
    }
    return n;
}

}  // namespace beer

using namespace beer;

extern "C" int beer_emission_llh(const float* X, int64_t N, int D, const float* W, const float* bias,
                                 const float* ref, int M, const int32_t* comp_off, int Kp, float* pdf_llh,
                                 int64_t ld_pdf, float* comp_llh, float* frame_ref, void* stream) {
    if (N < 0 || D <= 0 || M <= 0 || Kp <= 0 || ld_pdf < Kp) return BEER_ERR_ARG;
    if (N == 0) return BEER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int n_gtiles;
    if (comp_off == nullptr) {
        if (M % Kp != 0) return BEER_ERR_ARG;
        if (M / Kp > KA_G) return BEER_ERR_UNSUPPORTED;
        n_gtiles = ka_count_tiles(nullptr, Kp, M);
    } else {
        // comp_off lives on the device; the tile count needs it on the host.  It is a
        // per-model constant of Kp+1 ints, so a synchronous copy here is acceptable.
        int32_t* h = (int32_t*)malloc(sizeof(int32_t) * (Kp + 1));
        if (!h) return BEER_ERR_ALLOC;
        cudaError_t e = cudaMemcpyAsync(h, comp_off, sizeof(int32_t) * (Kp + 1), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { free(h); return (int)e; }
        n_gtiles = (h[0] == 0 && h[Kp] == M) ? ka_count_tiles(h, Kp, M) : -2;
        free(h);
        if (n_gtiles == -1) return BEER_ERR_UNSUPPORTED;
        if (n_gtiles < 0) return BEER_ERR_ARG;
    }
    size_t smem = sizeof(float) * ((size_t)2 * D * KA_G + (size_t)2 * D * KA_AS_LD + KA_FR * KA_CS_LD + KA_G + D + 1);
    // two blocks per SM up to D = 50; wider frames (a 64-d latent space: BASELINE configs[3]) run one block per SM
    if (smem > 227 * 1024) return BEER_ERR_UNSUPPORTED;  // D too large for this tiling
    static size_t attr_set = 0;
    if (smem > attr_set) {       // (the kernel has static shared memory too: ask for what is needed, not for the maximum)
        const size_t want = smem > 110 * 1024 ? smem : 110 * 1024;
        BEER_CUDA_TRY(cudaFuncSetAttribute(emission_llh_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)want));
        attr_set = want;
    }
    int64_t n_ftiles = (N + KA_FR - 1) / KA_FR;
    int64_t per_g = (2 * kNumSMs + n_gtiles - 1) / n_gtiles;
    if (per_g > n_ftiles) per_g = n_ftiles;
    if (per_g < 1) per_g = 1;
    int grid = (int)(per_g * n_gtiles);
    emission_llh_simt_kernel<<<grid, KA_THREADS, smem, st>>>(X, N, D, W, bias, ref, M, comp_off, Kp, n_gtiles,
                                                            pdf_llh, ld_pdf, comp_llh, frame_ref);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}
