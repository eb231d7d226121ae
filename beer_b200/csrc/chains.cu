// KB for per-utterance alignment graphs (aligned training: `beer hmm accumulate --alis`,
// beer/cli/subcommands/hmm/accumulate.py:47-57 -> HMM.expected_log_likelihood(inference_graph=...)
// beer/models/hmm.py:73-92).  An alignment graph (mkaligraph.py:18-39) compiles to a left-to-right
// CHAIN: state j has a self loop and one arc to state j + 1, the sequence starts in state 0 and
// ends in the last state.  Every utterance of a batch brings its own chain (own length, own pdf
// ids, own weights), stored back to back:
//
//   chain_off [n_utts + 1], and per state: pdf id, ln a(j,j), ln a(j,j+1) (last state: the final weight)
//
// One warp per utterance, lane l owns states l*S .. l*S+S-1 in registers; the only exchange between
// lanes is the neighbour's edge state (one shuffle per frame and direction).  log2 domain; every LANE keeps its values
// relative to its own integer offset (value = v + off, off an int32, re-based every frame so that the
// lane's maximum is in [0, 1)): a chain moves its probability mass along the states like a wave, and a state on the
// best path can sit hundreds of bits below the frame's maximum in alpha AND in beta -- relative to a per-FRAME
// maximum its fp32 log value then carries 2^-24 x that distance (3e-5 on posteriors with i.i.d.-noise llhs on a
// 300-state chain); relative to its lane's maximum it is within S states of a value near zero.  Offsets are integers,
// so differences of offsets are exact.
// S (4, 8, 16 or 32 states per lane) is picked PER UTTERANCE from its chain length: the batch is launched once per
// class and a warp skips the utterances of the other classes (a batch-wide S = 16 for one long chain cost every
// utterance 168 registers).  Rows in shared memory and in the alpha workspace are stored chunk-major (16-byte
// chunk v of lane l at float offset (32 v + l) * 4): conflict-free 16-byte accesses for every S (lane-major rows
// were 4-way conflicted at S = 16) and 512-byte coalesced global rows.
// The llhs are gathered through the chain's pdf ids by 4-byte cp.async into a per-warp ring; the
// posteriors are scatter-added onto pdf ids (modelset.py:148-154: a pdf may occur several times in a chain).
#include "common.cuh"
#include "../../include/beer_b200.h"

namespace beer {
namespace {

constexpr int CH_WARPS = 4;

struct ChainArgs {
    const float* pl;
    int64_t ld;
    const float* frame_ref;
    const int64_t* utt_off;
    int n_utts;
    float scale;
    const int64_t* chain_off;
    const int32_t* pdf;
    const float* lself;
    const float* lnext;
    const float* linit;      // [n_utts] ln weight of entering state 0
    float* la_ws;            // [N, Kw + 32]: lane-relative log2 alphas | the 32 lane offsets of the row
    int Kw;
    int kpad;                // ROWS variant: floats of one staged llh row (Kp rounded up to 4)
    float* state_post;       // [N, Kw] or null (row stride Kw: chains differ in length)
    float* pdf_post;         // [N, ld_post], zeroed by the caller
    int64_t ld_post;
    float* frame_exp_llh;
    double* utt_exp_llh;
    double* utt_logz;
};

__host__ __device__ __forceinline__ int chain_class(int len) {
    return len <= 128 ? 4 : (len <= 256 ? 8 : (len <= 512 ? 16 : (len <= 1024 ? 32 : 0)));
}

__device__ __forceinline__ float lse2c(float a, float b) {
    // min - max = -|a - b|: one subtraction, |.| and the sign are operand modifiers of the clamp (NaN of -inf - -inf and
    // -inf itself are absorbed by fmaxf)
    const float d = fmaxf(-fabsf(a - b), -1000.f);
    return fmaxf(a, b) + lg2(1.f + ex2(d));
}

// ROWS: the llh row of a frame is staged WHOLE in shared memory (coalesced 16-byte copies) and the states pick their pdf's
// value from there; otherwise every state gathers its value from global memory by a 4-byte cp.async (rows too wide to
// stage).  The gathers lean on the L1 cache -- with more resident blocks (less L1 next to their shared memory) the
// kernel got slower, not faster.
template <int S, bool ROWS>
__global__ void __launch_bounds__(CH_WARPS * 32) hmm_fb_chain_kernel(ChainArgs a) {
    static_assert(S % 4 == 0, "vector rows");
    constexpr int PF = (S <= 4) ? 6 : (S <= 8 ? 4 : 3);
    constexpr int ROW = 32 * S;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int PROW = ROWS ? a.kpad : ROW;          // floats of one llh slot
    float* ring_p = smem + (size_t)warp * (PF * PROW + PF * ROW + PF * 32);
    float* ring_a = ring_p + PF * PROW;
    float* ring_o = ring_a + PF * ROW;             // [PF][32] lane offsets of the alpha rows
    constexpr int kOffEmpty = -(1 << 29);          // offset of a lane without probability mass (differences of two never overflow)
    const float p_scale = a.scale * kLog2e;
    const int gwarp = blockIdx.x * CH_WARPS + warp, nwarps = gridDim.x * CH_WARPS;

    for (int u = gwarp; u < a.n_utts; u += nwarps) {
        const int64_t t0 = a.utt_off[u];
        const int T = (int)(a.utt_off[u + 1] - t0);
        const int64_t c0 = a.chain_off[u];
        const int L = (int)(a.chain_off[u + 1] - c0);
        if (T <= 0 || L <= 0) {
            if (lane == 0 && S == 4) {
                a.utt_exp_llh[u] = 0.0;
                if (a.utt_logz) a.utt_logz[u] = 0.0;
            }
            continue;
        }
        if (chain_class(L) != S) continue;         // another launch of this batch owns this utterance
        // this lane's states
        int pdf[S];
        float w_self[S], w_in[S];
#pragma unroll
        for (int j = 0; j < S; ++j) {
            const int k = lane * S + j;
            pdf[j] = (k < L) ? __ldg(a.pdf + c0 + k) : 0;
            w_self[j] = (k < L) ? __ldg(a.lself + c0 + k) * kLog2e : kNegInf;
            w_in[j] = (k < L && k > 0) ? __ldg(a.lnext + c0 + k - 1) * kLog2e : kNegInf;
        }
        const bool own = lane * S < L;
        for (int i = lane; i < PF * PROW + PF * ROW + PF * 32; i += 32) ring_p[i] = 0.f;   // states past L stay finite
        __syncwarp();
        const float* pl_u = a.pl + (size_t)t0 * a.ld;
        const int Kws = a.Kw + 32;                 // workspace row: alphas | lane offsets
        float* la_u = a.la_ws + (size_t)t0 * Kws;
        float* lo_u = la_u + a.Kw;

        auto gather = [&](float* slot, const float* row) {
            if constexpr (ROWS) {
                for (int c = lane; c < a.kpad / 4; c += 32) cp_async16(slot + 4 * c, row + 4 * c);
                return;
            }
            if (!own) return;
#pragma unroll
            for (int j = 0; j < S; ++j)
                if (lane * S + j < L) cp_async4(slot + ((j >> 2) * 32 + lane) * 4 + (j & 3), row + pdf[j]);
        };
        // the llhs of this lane's states from a landed slot
        auto read_llh = [&](const float* slot, float* out) {
            if constexpr (ROWS) {
#pragma unroll
                for (int j = 0; j < S; ++j) out[j] = (lane * S + j < L) ? slot[pdf[j]] : 0.f;
            } else {
#pragma unroll
                for (int v = 0; v < S / 4; ++v) {
                    const float4 q = reinterpret_cast<const float4*>(slot)[v * 32 + lane];
                    out[4 * v] = q.x; out[4 * v + 1] = q.y; out[4 * v + 2] = q.z; out[4 * v + 3] = q.w;
                }
            }
        };
        auto copy_row = [&](float* slot, const float* row) {
            if (!own) return;
#pragma unroll
            for (int v = 0; v < S / 4; ++v)
                if (lane * S + 4 * v < L) cp_async16(slot + (v * 32 + lane) * 4, row + (v * 32 + lane) * 4);
        };
        auto read_row = [&](const float* slot, float* out) {
#pragma unroll
            for (int v = 0; v < S / 4; ++v) {
                const float4 q = reinterpret_cast<const float4*>(slot)[v * 32 + lane];
                out[4 * v] = q.x; out[4 * v + 1] = q.y; out[4 * v + 2] = q.z; out[4 * v + 3] = q.w;
            }
        };
        // chunk-major (alpha workspace) or natural state order (state posteriors handed to the caller)
        auto write_row = [&](float* row, const float* v, bool natural) {
            if (!own) return;
#pragma unroll
            for (int q = 0; q < S / 4; ++q)
                if (lane * S + 4 * q < L)
                    reinterpret_cast<float4*>(row)[natural ? (lane * S) / 4 + q : q * 32 + lane] =
                        make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        };

        // ------------------------------ forward ------------------------------
        for (int r = 0; r < PF; ++r) {
            if (r < T) gather(ring_p + r * PROW, pl_u + (size_t)r * a.ld);
            cp_async_commit();
        }
        float cur[S];
        int off = kOffEmpty;
        int slot = 0;
        for (int t = 0; t < T; ++t) {
            cp_async_wait<PF - 1>();
            float p[S];
            float* ring_slot = ring_p + slot * PROW;
            read_llh(ring_slot, p);
            if constexpr (ROWS) __syncwarp();      // every lane has read the slot before it is refilled
            if (t + PF < T) gather(ring_slot, pl_u + (size_t)(t + PF) * a.ld);
            cp_async_commit();
            slot = (slot + 1 == PF) ? 0 : slot + 1;
            int base = kOffEmpty;                   // common offset of this lane's and its left neighbour's values
            if (t == 0) {
#pragma unroll
                for (int j = 0; j < S; ++j)
                    cur[j] = (lane == 0 && j == 0) ? fmaf(p[0], p_scale, __ldg(a.linit + u) * kLog2e) : kNegInf;
                base = 0;
            } else {
                float up = __shfl_up_sync(0xffffffffu, cur[S - 1], 1);
                int up_off = __shfl_up_sync(0xffffffffu, off, 1);
                if (lane == 0) {
                    up = kNegInf;
                    up_off = kOffEmpty;
                }
                base = max(off, up_off);
                const float sh = (float)(off - base);      // <= 0 (about -5e8 for an empty lane: -inf stays -inf)
                up += (float)(up_off - base);
#pragma unroll
                for (int j = 0; j < S; ++j) cur[j] += sh;
#pragma unroll
                for (int j = S - 1; j >= 0; --j) {
                    const float prev = (j == 0) ? up : cur[j == 0 ? 0 : j - 1];
                    cur[j] = fmaf(p[j], p_scale, lse2c(cur[j] + w_self[j], prev + w_in[j]));
                }
            }
            float mx = cur[0];
#pragma unroll
            for (int j = 1; j < S; ++j) mx = fmaxf(mx, cur[j]);
            if (mx == kNegInf) {
                off = kOffEmpty;
            } else {
                const int r = __float2int_rd(mx);
                const float rf = (float)r;
#pragma unroll
                for (int j = 0; j < S; ++j) cur[j] -= rf;
                off = base + r;
            }
            write_row(la_u + (size_t)t * Kws, cur, false);
            lo_u[(size_t)t * Kws + lane] = __int_as_float(off);
        }
        cp_async_wait<0>();
        // final weight: the "next" arc of the last state
        float b_start[S];
#pragma unroll
        for (int j = 0; j < S; ++j)
            b_start[j] = (lane * S + j == L - 1) ? __ldg(a.lnext + c0 + L - 1) * kLog2e : kNegInf;
        if (a.utt_logz != nullptr) {
            // the chain ends in its last state: log Z = its alpha + the final weight
            double z = -INFINITY;
#pragma unroll
            for (int j = 0; j < S; ++j)
                if (lane * S + j == L - 1) z = (double)cur[j] + (double)off + (double)b_start[j];
            for (int o = 16; o > 0; o >>= 1) z = fmax(z, __shfl_xor_sync(0xffffffffu, z, o));
            double rs = 0.0;
            if (a.frame_ref != nullptr)
                for (int t = lane; t < T; t += 32) rs += (double)a.frame_ref[t0 + t];
            rs = warp_sum(rs);
            if (lane == 0) a.utt_logz[u] = z * (double)kLn2 + (double)a.scale * rs;
        }

        // ------------------------------ backward -----------------------------
        __threadfence_block();   // this warp's la stores -> visible to its own async copies
        __syncwarp();
        for (int r = 0; r < PF; ++r) {
            const int t = T - 1 - r;
            if (t >= 0) {
                gather(ring_p + r * PROW, pl_u + (size_t)t * a.ld);
                copy_row(ring_a + r * ROW, la_u + (size_t)t * Kws);
                cp_async4(ring_o + r * 32 + lane, lo_u + (size_t)t * Kws + lane);
            }
            cp_async_commit();
        }
        float lb[S];
        int boff = kOffEmpty;                         // lane offset of the betas
#pragma unroll
        for (int j = 0; j < S; ++j) {
            lb[j] = b_start[j];
            if (lane * S + j == L - 1) boff = 0;
        }
        float ell = 0.f;
        double ell_d = 0.0;
        slot = 0;
        for (int i = 0; i < T; ++i) {
            const int t = T - 1 - i;
            cp_async_wait<PF - 1>();
            float p[S], la[S];
            read_llh(ring_p + slot * PROW, p);
            read_row(ring_a + slot * ROW, la);
            const int aoff = __float_as_int(ring_o[slot * 32 + lane]);
            if constexpr (ROWS) __syncwarp();
            if (t - PF >= 0) {
                gather(ring_p + slot * PROW, pl_u + (size_t)(t - PF) * a.ld);
                copy_row(ring_a + slot * ROW, la_u + (size_t)(t - PF) * Kws);
                cp_async4(ring_o + slot * 32 + lane, lo_u + (size_t)(t - PF) * Kws + lane);
            }
            cp_async_commit();
            slot = (slot + 1 == PF) ? 0 : slot + 1;

            // posteriors: alpha + beta of a state = (la + lb) + (aoff + boff); the lanes' integer totals relative to
            // the largest one (exact), so that the states that matter stay near zero
            const int tot = aoff + boff;
            const float dl = (float)(tot - __reduce_max_sync(0xffffffffu, tot));
            float v[S], m = kNegInf;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                v[j] = (lane * S + j < L) ? (la[j] + lb[j]) + dl : kNegInf;
                m = fmaxf(m, v[j]);
            }
            m = warp_max(m);
            const float ms = (m == kNegInf) ? 0.f : m;
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                v[j] = ex2(v[j] - ms);
                sum += v[j];
            }
            sum = warp_sum(sum);
            const float inv = (sum > 0.f) ? __fdividef(1.f, sum) : 0.f;
            float fe = 0.f;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                v[j] *= inv;
                fe = fmaf(p[j], v[j], fe);
            }
            ell += fe;
            if ((i & 31) == 31) {
                ell_d += (double)ell;
                ell = 0.f;
            }
            if (a.frame_exp_llh != nullptr) {
                const float f = warp_sum(fe);
                if (lane == 0) {
                    const float r = (a.frame_ref != nullptr) ? a.scale * a.frame_ref[t0 + t] : 0.f;
                    a.frame_exp_llh[t0 + t] = f * p_scale * kLn2 + r;
                }
            }
            if (a.state_post != nullptr) write_row(a.state_post + (size_t)(t0 + t) * a.Kw, v, true);
            if (a.pdf_post != nullptr) {
                float* prow = a.pdf_post + (size_t)(t0 + t) * a.ld_post;
#pragma unroll
                for (int j = 0; j < S; ++j)
                    if (v[j] != 0.f) atomicAdd(prow + pdf[j], a.scale * v[j]);
            }
            if (t == 0) break;
            float delta[S];
#pragma unroll
            for (int j = 0; j < S; ++j) delta[j] = fmaf(p[j], p_scale, lb[j]);
            float dn = __shfl_down_sync(0xffffffffu, delta[0] + w_in[0], 1);   // into the next lane's first state
            int dn_off = __shfl_down_sync(0xffffffffu, boff, 1);
            if (lane == 31) {
                dn = kNegInf;
                dn_off = kOffEmpty;
            }
            const int base = max(boff, dn_off);
            const float sh = (float)(boff - base);
            dn += (float)(dn_off - base);
#pragma unroll
            for (int j = 0; j < S; ++j) delta[j] += sh;
            float mb = kNegInf;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const float nxt = (j == S - 1) ? dn : delta[j == S - 1 ? j : j + 1] + w_in[j == S - 1 ? j : j + 1];
                lb[j] = lse2c(delta[j] + w_self[j], nxt);
                mb = fmaxf(mb, lb[j]);
            }
            if (mb == kNegInf) {
                boff = kOffEmpty;
            } else {
                const int r = __float2int_rd(mb);
                const float rf = (float)r;
#pragma unroll
                for (int j = 0; j < S; ++j) lb[j] -= rf;
                boff = base + r;
            }
        }
        cp_async_wait<0>();
        ell_d += (double)ell;
        ell_d = warp_sum(ell_d);
        double rs = 0.0;
        if (a.frame_ref != nullptr)
            for (int t = lane; t < T; t += 32) rs += (double)a.frame_ref[t0 + t];
        rs = warp_sum(rs);
        if (lane == 0) a.utt_exp_llh[u] = ell_d * (double)p_scale * (double)kLn2 + (double)a.scale * rs;
        __syncwarp();
    }
}

template <int S, bool ROWS>
int launch_chain_v(const ChainArgs& a, cudaStream_t st) {
    constexpr int PF = (S <= 4) ? 6 : (S <= 8 ? 4 : 3);
    const size_t smem = sizeof(float) * (size_t)CH_WARPS * (PF * (ROWS ? a.kpad : 32 * S) + PF * 32 * S + PF * 32);
    static size_t attr_set = 0;
    if (smem > attr_set) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(hmm_fb_chain_kernel<S, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        attr_set = smem;
    }
    int blocks = (a.n_utts + CH_WARPS - 1) / CH_WARPS;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    hmm_fb_chain_kernel<S, ROWS><<<blocks, CH_WARPS * 32, smem, st>>>(a);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

template <int S>
int launch_chain(const ChainArgs& a, cudaStream_t st) {
    // whole rows are staged when they are 16-byte aligned and no wider than the gathered slots would be x 2
    // (Kp <= 64 S: a row of 100 pdfs for chains of 250 states is 400 B against 8 scattered 4-byte copies per lane)
    const bool rows = a.kpad > 0 && a.kpad <= 64 * S && a.ld % 4 == 0 && ((uintptr_t)a.pl & 15) == 0;
    return rows ? launch_chain_v<S, true>(a, st) : launch_chain_v<S, false>(a, st);
}

int chain_S(int max_len) { return chain_class(max_len); }

}  // namespace
}  // namespace beer

using namespace beer;

extern "C" {

int beer_hmm_chain_row_stride(int max_chain_len) {
    const int s = chain_S(max_chain_len);
    return s ? 32 * s : BEER_ERR_UNSUPPORTED;
}

int64_t beer_hmm_chain_workspace_bytes(int max_chain_len, int64_t N) {
    const int s = chain_S(max_chain_len);
    return s ? (int64_t)N * (32 * s + 32) * (int64_t)sizeof(float) : BEER_ERR_UNSUPPORTED;     // alpha rows + lane offsets
}

int beer_hmm_forward_backward_chains(const float* pdf_llh, int64_t ld_pdf, const float* frame_ref, const int64_t* utt_off,
                                     int n_utts, const int64_t* chain_off, const int32_t* chain_pdf,
                                     const float* chain_log_self, const float* chain_log_next,
                                     const float* chain_log_init, int max_chain_len, float scale, float* state_post,
                                     float* pdf_post, int64_t ld_post, float* frame_exp_llh, double* utt_exp_llh,
                                     double* utt_logz, void* workspace, void* stream) {
    if (!pdf_llh || !utt_off || !chain_off || !chain_pdf || !chain_log_self || !chain_log_next || !chain_log_init ||
        !utt_exp_llh || !workspace || n_utts < 0 || max_chain_len <= 0)
        return BEER_ERR_ARG;
    const int S = chain_S(max_chain_len);
    if (S == 0) return BEER_ERR_UNSUPPORTED;
    if (n_utts == 0) return BEER_OK;
    ChainArgs a;
    a.pl = pdf_llh; a.ld = ld_pdf; a.frame_ref = frame_ref; a.utt_off = utt_off; a.n_utts = n_utts; a.scale = scale;
    a.chain_off = chain_off; a.pdf = chain_pdf; a.lself = chain_log_self; a.lnext = chain_log_next;
    a.linit = chain_log_init; a.la_ws = (float*)workspace; a.Kw = 32 * S; a.state_post = state_post;
    a.kpad = (ld_pdf % 4 == 0 && ld_pdf <= 8192) ? (int)ld_pdf : 0;      // whole rows of ld_pdf floats are staged
    a.pdf_post = pdf_post; a.ld_post = ld_post; a.frame_exp_llh = frame_exp_llh; a.utt_exp_llh = utt_exp_llh;
    a.utt_logz = utt_logz;
    cudaStream_t st = (cudaStream_t)stream;
    // One launch per length class up to the longest chain's; a warp skips the utterances of the other classes.  The
    // classes are independent (disjoint utterances, scatter-adds into disjoint rows), so they run side by side on
    // forked streams: the longer classes are register-limited to 12 - 17 warps per SM and leave room for the others
    // (fork / join with events: stream-ordered for the caller and capturable in a CUDA graph).
    if (S == 4) return launch_chain<4>(a, st);
    static cudaStream_t aux[3] = {nullptr, nullptr, nullptr};
    static cudaEvent_t fork = nullptr, join[3] = {nullptr, nullptr, nullptr};
    if (fork == nullptr) {
        BEER_CUDA_TRY(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        for (int i = 0; i < 3; ++i) {
            BEER_CUDA_TRY(cudaStreamCreateWithFlags(&aux[i], cudaStreamNonBlocking));
            BEER_CUDA_TRY(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming));
        }
    }
    BEER_CUDA_TRY(cudaEventRecord(fork, st));
    const int n_aux = S >= 32 ? 3 : (S >= 16 ? 2 : 1);
    int rc = BEER_OK;
    for (int i = 0; i < n_aux && rc == BEER_OK; ++i) {
        BEER_CUDA_TRY(cudaStreamWaitEvent(aux[i], fork, 0));
        rc = i == 0 ? launch_chain<8>(a, aux[0]) : (i == 1 ? launch_chain<16>(a, aux[1]) : launch_chain<32>(a, aux[2]));
        BEER_CUDA_TRY(cudaEventRecord(join[i], aux[i]));
    }
    if (rc == BEER_OK) rc = launch_chain<4>(a, st);
    for (int i = 0; i < n_aux; ++i) BEER_CUDA_TRY(cudaStreamWaitEvent(st, join[i], 0));
    return rc;
}

}  // extern "C"
