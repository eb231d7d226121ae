// KB / KV -- HMM forward-backward and Viterbi over a compiled graph.
//
// One warp owns one utterance; lane l owns the S consecutive states
// [l*S, (l+1)*S).  The recursion runs in the log2 domain, renormalised every
// frame (max = 0) so that fp32 stays accurate over thousands of frames -- the
// reference (beer/graph.py:270-287) never renormalises and loses ~1e-2 on the
// posteriors in fp32.  Transitions are a sparse "lane-major ELL" list; rank-1
// blocks (unit-end -> unit-start of a phone loop) go through junction nodes.
//
// Reference semantics: beer/graph.py:270-344, beer/models/hmm.py:79-100,
// beer/models/modelset.py:140-154.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "common.cuh"
#include "../../include/beer_b200.h"

namespace beer {

struct ScanLists {
    const int* st_cnt;   // [S]  arcs per state slot (max over lanes)
    const int* st_off;   // [S]  first ELL row of the slot
    const int* jn_cnt;   // [J]  ELL rows per junction
    const int* jn_off;   // [J]
    const int* src;      // ELL sources: row r, lane l at r*32 + l
    const float* lw;     // ELL weights (log2 domain for fwd/bwd, natural log for Viterbi)
    const float* start;  // [K] initial (fwd, Viterbi) or final (bwd) log-probabilities
};

}  // namespace beer

struct beer_graph_plan {
    int K = 0, Kp = 0, J = 0, S = 0;
    int n_direct = 0, n_jin = 0, n_jout = 0, dense_nnz = 0;
    int map_identity = 0;
    int fast_ok = 0;   // every state has <= 2 in/out arcs after factoring and there is <= 1 junction
    int jrows = 0;     // ELL rows of that junction (max over directions)
    int lr_su = 0, lr_u = 0;          // aligned left-to-right loop: states per unit, units per lane (0 = not)
    const float* lr_w = nullptr;      // device [3][32 * lr_su * lr_u]: self, incoming, end->junction (log2)
    // the same loop for Viterbi, dense natural-log weights: [self K | previous state K | unit end -> any unit start P];
    // vlr_ok: every unit start sees the SAME weight from a given unit end (bitwise), so that the first maximum over the
    // unit ends is shared by all starts (register-resident kernel for more than 32 units)
    const float* vlr = nullptr;
    int vlr_ok = 0;
    beer::ScanLists fwd{}, bwd{}, vit{};
    const int* map = nullptr;        // device [K]
    const float* vit_final = nullptr;  // device [K] natural log
    void* dev_blob = nullptr;
};

namespace beer {

// ---------------------------------------------------------------------------
// host side: plan construction
// ---------------------------------------------------------------------------
struct Arc { int src; double lw; };

struct EllBuilder {
    std::vector<int> src;
    std::vector<float> lw;
    int rows = 0;
    // append `n` rows, return first row
    int add_rows(int n) {
        int r = rows;
        rows += n;
        src.resize((size_t)rows * 32, 0);
        lw.resize((size_t)rows * 32, -INFINITY);
        return r;
    }
    void set(int row, int lane, int s, float w) {
        src[(size_t)row * 32 + lane] = s;
        lw[(size_t)row * 32 + lane] = w;
    }
};

struct HostLists {
    std::vector<int> st_cnt, st_off, jn_cnt, jn_off;
    EllBuilder ell;
    std::vector<float> start;
};

// state_lists[k] = arcs feeding state k; junction_lists[n] = arcs feeding junction n.
static void build_lists(int K, int S, const std::vector<std::vector<Arc>>& state_lists,
                        const std::vector<std::vector<Arc>>& junction_lists, double unit, HostLists& out) {
    out.st_cnt.assign(S, 0);
    out.st_off.assign(S, 0);
    for (int s = 0; s < S; ++s) {
        int cnt = 0;
        for (int lane = 0; lane < 32; ++lane) {
            int k = lane * S + s;
            if (k < K) cnt = std::max(cnt, (int)state_lists[k].size());
        }
        int row0 = out.ell.add_rows(cnt);
        out.st_cnt[s] = cnt;
        out.st_off[s] = row0;
        for (int lane = 0; lane < 32; ++lane) {
            int k = lane * S + s;
            if (k >= K) continue;
            for (size_t a = 0; a < state_lists[k].size(); ++a)
                out.ell.set(row0 + (int)a, lane, state_lists[k][a].src, (float)(state_lists[k][a].lw * unit));
        }
    }
    int J = (int)junction_lists.size();
    out.jn_cnt.assign(std::max(J, 1), 0);
    out.jn_off.assign(std::max(J, 1), 0);
    for (int n = 0; n < J; ++n) {
        int cnt = ((int)junction_lists[n].size() + 31) / 32;
        int row0 = out.ell.add_rows(cnt);
        out.jn_cnt[n] = cnt;
        out.jn_off[n] = row0;
        for (size_t m = 0; m < junction_lists[n].size(); ++m)
            out.ell.set(row0 + (int)(m / 32), (int)(m % 32), junction_lists[n][m].src,
                        (float)(junction_lists[n][m].lw * unit));
    }
}

struct BlobWriter {
    std::vector<char> bytes;
    size_t add(const void* p, size_t n) {
        size_t off = (bytes.size() + 15) & ~(size_t)15;
        bytes.resize(off + std::max<size_t>(n, 4), 0);
        if (n) memcpy(bytes.data() + off, p, n);
        return off;
    }
};

struct ListOffsets { size_t st_cnt, st_off, jn_cnt, jn_off, src, lw, start; };

static ListOffsets write_lists(BlobWriter& w, const HostLists& h) {
    ListOffsets o;
    o.st_cnt = w.add(h.st_cnt.data(), h.st_cnt.size() * 4);
    o.st_off = w.add(h.st_off.data(), h.st_off.size() * 4);
    o.jn_cnt = w.add(h.jn_cnt.data(), h.jn_cnt.size() * 4);
    o.jn_off = w.add(h.jn_off.data(), h.jn_off.size() * 4);
    o.src = w.add(h.ell.src.data(), h.ell.src.size() * 4);
    o.lw = w.add(h.ell.lw.data(), h.ell.lw.size() * 4);
    o.start = w.add(h.start.data(), h.start.size() * 4);
    return o;
}

static ScanLists bind_lists(const char* base, const ListOffsets& o) {
    ScanLists l;
    l.st_cnt = (const int*)(base + o.st_cnt);
    l.st_off = (const int*)(base + o.st_off);
    l.jn_cnt = (const int*)(base + o.jn_cnt);
    l.jn_off = (const int*)(base + o.jn_off);
    l.src = (const int*)(base + o.src);
    l.lw = (const float*)(base + o.lw);
    l.start = (const float*)(base + o.start);
    return l;
}

// ---------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------

// log2-domain logsumexp of one ELL list as seen by this lane.
__device__ __forceinline__ float lane_list_lse(const int* __restrict__ src, const float* __restrict__ lw,
                                               int row0, int cnt, int lane, const float* buf) {
    if (cnt == 1) {
        int i = row0 * 32 + lane;
        return buf[__ldg(src + i)] + __ldg(lw + i);
    }
    if (cnt == 2) {
        int i = row0 * 32 + lane;
        float a = buf[__ldg(src + i)] + __ldg(lw + i);
        float b = buf[__ldg(src + i + 32)] + __ldg(lw + i + 32);
        float mx = fmaxf(a, b), mn = fminf(a, b);
        float d = (mn == kNegInf) ? kNegInf : mn - mx;
        return mx + lg2(1.f + ex2(d));
    }
    float m = kNegInf;
    for (int a = 0; a < cnt; ++a) {
        int i = (row0 + a) * 32 + lane;
        m = fmaxf(m, buf[__ldg(src + i)] + __ldg(lw + i));
    }
    float ms = (m == kNegInf) ? 0.f : m;
    float s = 0.f;
    for (int a = 0; a < cnt; ++a) {
        int i = (row0 + a) * 32 + lane;
        s += ex2(buf[__ldg(src + i)] + __ldg(lw + i) - ms);
    }
    return ms + lg2(s);
}

// Junction value: logsumexp over a list spread over the whole warp.
__device__ __forceinline__ float warp_list_lse(const int* __restrict__ src, const float* __restrict__ lw,
                                               int row0, int cnt, int lane, const float* buf) {
    float m = kNegInf;
    for (int a = 0; a < cnt; ++a) {
        int i = (row0 + a) * 32 + lane;
        m = fmaxf(m, buf[__ldg(src + i)] + __ldg(lw + i));
    }
    m = warp_max(m);
    float ms = (m == kNegInf) ? 0.f : m;
    float s = 0.f;
    for (int a = 0; a < cnt; ++a) {
        int i = (row0 + a) * 32 + lane;
        s += ex2(buf[__ldg(src + i)] + __ldg(lw + i) - ms);
    }
    s = warp_sum(s);
    return ms + lg2(s);
}

struct FbArgs {
    ScanLists fwd, bwd;
    int K, J, Kp;
    const int* map;
    int map_identity;
    const float* pl;
    int64_t ld;
    const float* frame_ref;
    const int64_t* utt_off;
    int n_utts;
    float scale;
    float* la_ws;  // [N, Kw] normalised log2 alphas
    int Kw;
    float* state_post;
    float* pdf_post;
    int64_t ld_post;
    float* frame_exp_llh;
    double* utt_exp_llh;
    double* utt_logz;
    int vec;  // 16-byte row copies are legal
    const float* lr_w;  // aligned left-to-right loop weights (hmm_fb_lr_kernel)
    int lr_row;         // stride between the three weight arrays
    double* unit_counts;  // [K / SU] or NULL: sum_t xi_t(unit ends -> unit start) + gamma_0(start)
    float* pdf_lpost;     // [N, ld_lpost] or NULL: log2(scale * gamma) per pdf (identity pdf maps, loop kernels only)
    int64_t ld_lpost;
    float llh_mul;        // log2(e) for llhs in nats, 1 for llhs already in log2 units (the kernels work in log2)
    int lpost_rel;        // pdf_lpost = log2(scale * gamma) - log2 llh (what beer_mix16_accumulate adds to z), not log2(scale * gamma)
    // activity map for the statistics kernel: blk_active[(frame / 64) * blk_ld + pdf / blk_ppb] = 1 wherever a pdf posterior
    // of the block reaches 2^blk_thr (caller-zeroed; below that the weights round to zero in the kernel's fp16 operands)
    uint8_t* blk_active;
    int64_t blk_ld;
    int blk_ppb;
    float blk_thr;
};

// A lane's part of the activity map: the blocks g0 .. g1 of its pdfs (identity pdf map: state = pdf), and the tile it
// marked last (a lane with mass marks once per tile of 64 frames, not once per frame).
struct BlockMarker {
    int g0, ng;      // first block and number of blocks of this lane's pdfs (ng = 1 when the block width is a multiple of the unit)
    float tmax;      // largest log2 posterior of this lane's pdfs over the frames of the current tile
    __device__ __forceinline__ void init(const FbArgs& a, int first, int n) {
        g0 = 0;
        ng = 0;
        tmax = kNegInf;
        if (a.blk_active != nullptr && n > 0) {
            g0 = first / a.blk_ppb;
            ng = (first + n - 1) / a.blk_ppb - g0 + 1;
        }
    }
    // per frame: one max; at the first frame of a tile (the sweep runs backwards) or of the utterance -- a warp-uniform
    // test -- the lanes with mass in the tile store their byte(s)
    __device__ __forceinline__ void frame(const FbArgs& a, float lpost_max, int64_t f, bool first_of_utt) {
        tmax = fmaxf(tmax, lpost_max);
        if ((f & 63) == 0 || first_of_utt) {
            if (ng > 0 && tmax >= a.blk_thr) {
                uint8_t* row = a.blk_active + (f >> 6) * a.blk_ld + g0;
                row[0] = 1;
#pragma unroll 1
                for (int g = 1; g < ng; ++g) row[g] = 1;
            }
            tmax = kNegInf;
        }
    }
};

constexpr int FB_WARPS = 4;

template <int S>
struct FbCfg {
    static constexpr int PF = (S <= 4) ? 6 : (S <= 8 ? 4 : 2);  // prefetch depth (rows)
};

// Issue the async copy of this lane's S values of one row into a ring slot.
template <int S>
__device__ __forceinline__ void prefetch_row(float* ring_slot, const float* row, int lane, int K, bool vec,
                                             const int* __restrict__ map, bool identity) {
    float* dst = ring_slot + lane * S;
    if (vec) {
        if constexpr (S % 4 == 0) {
#pragma unroll
            for (int v = 0; v < S / 4; ++v)
                if (lane * S + 4 * v < K) cp_async16(dst + 4 * v, row + lane * S + 4 * v);
        }
    } else {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            int k = lane * S + s;
            if (k < K) cp_async4(dst + s, row + (identity ? k : __ldg(map + k)));
        }
    }
}

template <int S>
__global__ void __launch_bounds__(FB_WARPS * 32) hmm_fb_kernel(FbArgs a) {
    constexpr int PF = FbCfg<S>::PF;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nbuf = (a.K + a.J + 3) & ~3;
    const int per_warp = nbuf + 2 * PF * 32 * S;
    float* buf = smem + (size_t)warp * per_warp;
    float* ring_p = buf + nbuf;            // [PF][32*S]
    float* ring_a = ring_p + PF * 32 * S;  // [PF][32*S]
    const int K = a.K, J = a.J;
    const bool vec = a.vec != 0, ident = a.map_identity != 0;
    const float p_scale = a.scale * a.llh_mul;
    const int gwarp = blockIdx.x * FB_WARPS + warp, nwarps = gridDim.x * FB_WARPS;

    for (int u = gwarp; u < a.n_utts; u += nwarps) {
        const int64_t t0 = a.utt_off[u];
        const int T = (int)(a.utt_off[u + 1] - t0);
        if (T <= 0) {
            if (lane == 0) {
                a.utt_exp_llh[u] = 0.0;
                if (a.utt_logz) a.utt_logz[u] = 0.0;
            }
            continue;
        }
        const float* pl_u = a.pl + (size_t)t0 * a.ld;
        float* la_u = a.la_ws + (size_t)t0 * a.Kw;
        // identity-mapped vector copies of la rows use the same helper (map unused)
        double logz2 = 0.0;

        // ------------------------------ forward ------------------------------
        for (int r = 0; r < PF; ++r) {
            if (r < T) prefetch_row<S>(ring_p + r * 32 * S, pl_u + (size_t)r * a.ld, lane, K, vec, a.map, ident);
            cp_async_commit();
        }
        float cur[S];
        for (int t = 0; t < T; ++t) {
            cp_async_wait<PF - 1>();
            float p[S];
            const float* slot = ring_p + (t % PF) * 32 * S + lane * S;
#pragma unroll
            for (int s = 0; s < S; ++s) p[s] = (lane * S + s < K) ? slot[s] * p_scale : kNegInf;
            if (t + PF < T)
                prefetch_row<S>(ring_p + (t % PF) * 32 * S, pl_u + (size_t)(t + PF) * a.ld, lane, K, vec, a.map,
                                ident);
            cp_async_commit();

            if (t == 0) {
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    int k = lane * S + s;
                    cur[s] = (k < K) ? p[s] + __ldg(a.fwd.start + k) : kNegInf;
                }
            } else {
                for (int n = 0; n < J; ++n) {
                    float v = warp_list_lse(a.fwd.src, a.fwd.lw, __ldg(a.fwd.jn_off + n), __ldg(a.fwd.jn_cnt + n),
                                            lane, buf);
                    if (lane == 0) buf[K + n] = v;
                }
                if (J > 0) __syncwarp();
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    int k = lane * S + s;
                    float v = lane_list_lse(a.fwd.src, a.fwd.lw, __ldg(a.fwd.st_off + s), __ldg(a.fwd.st_cnt + s),
                                            lane, buf);
                    cur[s] = (k < K) ? p[s] + v : kNegInf;
                }
            }
            float mx = cur[0];
#pragma unroll
            for (int s = 1; s < S; ++s) mx = fmaxf(mx, cur[s]);
            mx = warp_max(mx);
            const float mxs = (mx == kNegInf) ? 0.f : mx;
            logz2 += (double)mxs;
            __syncwarp();  // every lane has finished reading buf
#pragma unroll
            for (int s = 0; s < S; ++s) {
                cur[s] -= mxs;
                if (lane * S + s < K) buf[lane * S + s] = cur[s];
            }
            __syncwarp();
            // store normalised log-alphas
            float* la_row = la_u + (size_t)t * a.Kw;
            if constexpr (S % 4 == 0) {
                if ((a.Kw & 3) == 0) {
#pragma unroll
                    for (int v = 0; v < S / 4; ++v)
                        if (lane * S + 4 * v < K)
                            *reinterpret_cast<float4*>(la_row + lane * S + 4 * v) =
                                make_float4(cur[4 * v], cur[4 * v + 1], cur[4 * v + 2], cur[4 * v + 3]);
                } else {
#pragma unroll
                    for (int s = 0; s < S; ++s)
                        if (lane * S + s < K) la_row[lane * S + s] = cur[s];
                }
            } else {
#pragma unroll
                for (int s = 0; s < S; ++s)
                    if (lane * S + s < K) la_row[lane * S + s] = cur[s];
            }
        }
        cp_async_wait<0>();

        // log evidence: sum_t normalisers + LSE_k(la_{T-1,k} + final_k)
        if (a.utt_logz != nullptr) {
            float m = kNegInf;
            float v[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                int k = lane * S + s;
                v[s] = (k < K) ? cur[s] + __ldg(a.bwd.start + k) : kNegInf;
                m = fmaxf(m, v[s]);
            }
            m = warp_max(m);
            float ms = (m == kNegInf) ? 0.f : m;
            float sum = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) sum += ex2(v[s] - ms);
            sum = warp_sum(sum);
            double z = (logz2 + (double)ms + (double)lg2(sum)) * (double)kLn2;
            double rs = 0.0;
            if (a.frame_ref != nullptr)
                for (int t = lane; t < T; t += 32) rs += (double)a.frame_ref[t0 + t];
            rs = warp_sum(rs);
            if (lane == 0) a.utt_logz[u] = z + (double)a.scale * rs;
        }

        // ------------------------------ backward -----------------------------
        // make this warp's la stores visible to its own async copies
        __threadfence_block();
        __syncwarp();
        const bool la_vec = (S % 4 == 0) && ((a.Kw & 3) == 0);
        for (int r = 0; r < PF; ++r) {
            int t = T - 1 - r;
            if (t >= 0) {
                prefetch_row<S>(ring_p + r * 32 * S, pl_u + (size_t)t * a.ld, lane, K, vec, a.map, ident);
                prefetch_row<S>(ring_a + r * 32 * S, la_u + (size_t)t * a.Kw, lane, K, la_vec, nullptr, true);
            }
            cp_async_commit();
        }
        float lb[S];
        {
            float m = kNegInf;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                int k = lane * S + s;
                lb[s] = (k < K) ? __ldg(a.bwd.start + k) : kNegInf;
                m = fmaxf(m, lb[s]);
            }
            m = warp_max(m);
            float ms = (m == kNegInf) ? 0.f : m;
#pragma unroll
            for (int s = 0; s < S; ++s) lb[s] -= ms;
        }
        double ell = 0.0;  // this lane's share of sum_t sum_k p2_tk gamma_tk
        for (int i = 0; i < T; ++i) {
            const int t = T - 1 - i;
            cp_async_wait<PF - 1>();
            float p[S], la[S];
            const float* sp = ring_p + (i % PF) * 32 * S + lane * S;
            const float* sa = ring_a + (i % PF) * 32 * S + lane * S;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                bool ok = lane * S + s < K;
                p[s] = ok ? sp[s] * p_scale : kNegInf;
                la[s] = ok ? sa[s] : kNegInf;
            }
            if (t - PF >= 0) {
                prefetch_row<S>(ring_p + (i % PF) * 32 * S, pl_u + (size_t)(t - PF) * a.ld, lane, K, vec, a.map,
                                ident);
                prefetch_row<S>(ring_a + (i % PF) * 32 * S, la_u + (size_t)(t - PF) * a.Kw, lane, K, la_vec,
                                nullptr, true);
            }
            cp_async_commit();

            // gamma_t
            float v[S], m = kNegInf;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                v[s] = la[s] + lb[s];
                m = fmaxf(m, v[s]);
            }
            m = warp_max(m);
            const float ms = (m == kNegInf) ? 0.f : m;
            float sum = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                v[s] = ex2(v[s] - ms);
                sum += v[s];
            }
            sum = warp_sum(sum);
            const float inv = (sum > 0.f) ? 1.f / sum : 0.f;
            float fe = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                v[s] *= inv;
                if (v[s] > 0.f) fe = fmaf(p[s], v[s], fe);
            }
            ell += (double)fe;
            if (a.frame_exp_llh != nullptr) {
                float f = warp_sum(fe);
                if (lane == 0) {
                    float r = (a.frame_ref != nullptr) ? a.scale * a.frame_ref[t0 + t] : 0.f;
                    a.frame_exp_llh[t0 + t] = f * kLn2 + r;
                }
            }
            if (a.state_post != nullptr) {
                float* row = a.state_post + (size_t)(t0 + t) * K;
#pragma unroll
                for (int s = 0; s < S; ++s)
                    if (lane * S + s < K) row[lane * S + s] = v[s];
            }
            if (a.pdf_post != nullptr) {
                float* row = a.pdf_post + (size_t)(t0 + t) * a.ld_post;
                if (ident) {
                    if constexpr (S % 4 == 0) {
                        if ((a.ld_post & 3) == 0 && (K & 3) == 0) {
#pragma unroll
                            for (int q = 0; q < S / 4; ++q)
                                if (lane * S + 4 * q < K)
                                    *reinterpret_cast<float4*>(row + lane * S + 4 * q) =
                                        make_float4(a.scale * v[4 * q], a.scale * v[4 * q + 1],
                                                    a.scale * v[4 * q + 2], a.scale * v[4 * q + 3]);
                        } else {
#pragma unroll
                            for (int s = 0; s < S; ++s)
                                if (lane * S + s < K) row[lane * S + s] = a.scale * v[s];
                        }
                    } else {
#pragma unroll
                        for (int s = 0; s < S; ++s)
                            if (lane * S + s < K) row[lane * S + s] = a.scale * v[s];
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        int k = lane * S + s;
                        if (k < K && v[s] != 0.f) atomicAdd(row + __ldg(a.map + k), a.scale * v[s]);
                    }
                }
            }

            if (t == 0) break;
            // beta_{t-1}: delta_j = p_tj + lb_tj, then the transposed recursion
            __syncwarp();
#pragma unroll
            for (int s = 0; s < S; ++s)
                if (lane * S + s < K) buf[lane * S + s] = p[s] + lb[s];
            __syncwarp();
            for (int n = 0; n < J; ++n) {
                float vj = warp_list_lse(a.bwd.src, a.bwd.lw, __ldg(a.bwd.jn_off + n), __ldg(a.bwd.jn_cnt + n), lane,
                                         buf);
                if (lane == 0) buf[K + n] = vj;
            }
            if (J > 0) __syncwarp();
            float mb = kNegInf;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                int k = lane * S + s;
                float vb = lane_list_lse(a.bwd.src, a.bwd.lw, __ldg(a.bwd.st_off + s), __ldg(a.bwd.st_cnt + s), lane,
                                         buf);
                lb[s] = (k < K) ? vb : kNegInf;
                mb = fmaxf(mb, lb[s]);
            }
            mb = warp_max(mb);
            const float mbs = (mb == kNegInf) ? 0.f : mb;
#pragma unroll
            for (int s = 0; s < S; ++s) lb[s] -= mbs;
        }
        cp_async_wait<0>();
        ell = warp_sum(ell);
        double rs = 0.0;
        if (a.frame_ref != nullptr)
            for (int t = lane; t < T; t += 32) rs += (double)a.frame_ref[t0 + t];
        rs = warp_sum(rs);
        if (lane == 0) a.utt_exp_llh[u] = ell * (double)kLn2 + (double)a.scale * rs;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Fast path: graphs whose states have at most two incoming and two outgoing arcs once the
// rank-1 blocks are routed through (at most one) junction -- phone loops of left-to-right
// units (beer/cli/subcommands/hmm/mkphoneloopgraph.py) and alignment graphs (mkaligraph.py).
// The arc lists of a lane live in registers for the whole kernel, the junction value is
// reduced into a register of every lane (no shared-memory round trip) and selected into the
// arcs that leave it.  Same arithmetic as the generic kernel.
// ---------------------------------------------------------------------------
template <int S, int JR>
struct LaneLists {
    int src[S][2];
    float w[S][2];
    unsigned jmask;   // bit 2s+c: arc c of slot s leaves the junction
    int jsrc[JR];
    float jw[JR];
};

template <int S, int JR>
__device__ __forceinline__ void load_lane_lists(LaneLists<S, JR>& L, const ScanLists& l, int K, int J, int lane) {
    L.jmask = 0u;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int row0 = __ldg(l.st_off + s), cnt = __ldg(l.st_cnt + s);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            int src = 0;
            float w = kNegInf;
            if (c < cnt) {
                const int i = (row0 + c) * 32 + lane;
                src = __ldg(l.src + i);
                w = __ldg(l.lw + i);
            }
            if (src >= K) {
                L.jmask |= 1u << (2 * s + c);
                src = 0;
            }
            L.src[s][c] = (src % S) * 32 + src / S;   // slot-major position in buf (conflict-free reads)
            L.w[s][c] = w;
        }
    }
    const int jrow0 = (J > 0) ? __ldg(l.jn_off) : 0, jcnt = (J > 0) ? __ldg(l.jn_cnt) : 0;
#pragma unroll
    for (int r = 0; r < JR; ++r) {
        L.jsrc[r] = 0;
        L.jw[r] = kNegInf;
        if (r < jcnt) {
            const int i = (jrow0 + r) * 32 + lane;
            const int src = __ldg(l.src + i);
            L.jsrc[r] = (src % S) * 32 + src / S;
            L.jw[r] = __ldg(l.lw + i);
        }
    }
}

// log2-domain log(2^a + 2^b); -inf safe (NaN of -inf - -inf is absorbed by fmaxf)
__device__ __forceinline__ float lse2(float a, float b) {
    // min - max = -|a - b|: one subtraction, |.| and the sign are operand modifiers of the clamp (NaN of -inf - -inf and
    // -inf itself are absorbed by fmaxf)
    const float d = fmaxf(-fabsf(a - b), -1000.f);
    return fmaxf(a, b) + lg2(1.f + ex2(d));
}

// value of the junction from the published per-state values in buf
template <int S, int JR>
__device__ __forceinline__ float junction_value(const LaneLists<S, JR>& L, const float* buf) {
    float v[JR], m = kNegInf;
#pragma unroll
    for (int r = 0; r < JR; ++r) {
        v[r] = buf[L.jsrc[r]] + L.jw[r];
        m = fmaxf(m, v[r]);
    }
    m = warp_max(m);
    const float ms = (m == kNegInf) ? 0.f : m;
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < JR; ++r) sum += ex2(v[r] - ms);
    sum = warp_sum(sum);
    return ms + lg2(sum);
}

template <int S, int JR>
__device__ __forceinline__ void lane_states(const LaneLists<S, JR>& L, const float* buf, float jv, float* out) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
        float x0 = buf[L.src[s][0]], x1 = buf[L.src[s][1]];
        if (L.jmask & (1u << (2 * s))) x0 = jv;
        if (L.jmask & (1u << (2 * s + 1))) x1 = jv;
        out[s] = lse2(x0 + L.w[s][0], x1 + L.w[s][1]);
    }
}

template <int S, int JR>
__global__ void __launch_bounds__(FB_WARPS * 32) hmm_fb_fast_kernel(FbArgs a) {
    static_assert(S % 4 == 0, "vector rows");
    constexpr int PF = FbCfg<S>::PF;
    constexpr int ROW = 32 * S;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int per_warp = ROW + 2 * PF * ROW;
    float* buf = smem + (size_t)warp * per_warp;   // [S][32] published per-state values, slot-major
    float* ring_p = buf + ROW;                     // [PF][32 * S]
    float* ring_a = ring_p + PF * ROW;             // [PF][32 * S]
    const int K = a.K, J = a.J;
    const float p_scale = a.scale * a.llh_mul;
    const int gwarp = blockIdx.x * FB_WARPS + warp, nwarps = gridDim.x * FB_WARPS;
    const bool own = lane * S < K;                 // this lane holds real states (K % 4 == 0)
    // lanes past K never receive async copies: keep their ring slots finite
    for (int i = lane; i < 2 * PF * ROW; i += 32) ring_p[i] = 0.f;
    __syncwarp();

    LaneLists<S, JR> F, B;
    load_lane_lists(F, a.fwd, K, J, lane);
    load_lane_lists(B, a.bwd, K, J, lane);
    float f_start[S], b_start[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int k = lane * S + s;
        f_start[s] = (k < K) ? __ldg(a.fwd.start + k) : kNegInf;
        b_start[s] = (k < K) ? __ldg(a.bwd.start + k) : kNegInf;
    }

    auto prefetch = [&](float* slot, const float* row) {
        if (own) {
#pragma unroll
            for (int v = 0; v < S / 4; ++v)
                if (lane * S + 4 * v < K) cp_async16(slot + lane * S + 4 * v, row + lane * S + 4 * v);
        }
    };

    for (int u = gwarp; u < a.n_utts; u += nwarps) {
        const int64_t t0 = a.utt_off[u];
        const int T = (int)(a.utt_off[u + 1] - t0);
        if (T <= 0) {
            if (lane == 0) {
                a.utt_exp_llh[u] = 0.0;
                if (a.utt_logz) a.utt_logz[u] = 0.0;
            }
            continue;
        }
        const float* pl_u = a.pl + (size_t)t0 * a.ld;
        float* la_u = a.la_ws + (size_t)t0 * a.Kw;
        double logz2 = 0.0;

        // ------------------------------ forward ------------------------------
        for (int r = 0; r < PF; ++r) {
            if (r < T) prefetch(ring_p + r * ROW, pl_u + (size_t)r * a.ld);
            cp_async_commit();
        }
        float cur[S];
        for (int t = 0; t < T; ++t) {
            cp_async_wait<PF - 1>();
            const float4* slot = reinterpret_cast<const float4*>(ring_p + (t % PF) * ROW + lane * S);
            float p[S];
#pragma unroll
            for (int v = 0; v < S / 4; ++v) {
                const float4 q = slot[v];
                p[4 * v] = q.x * p_scale; p[4 * v + 1] = q.y * p_scale;
                p[4 * v + 2] = q.z * p_scale; p[4 * v + 3] = q.w * p_scale;
            }
            if (t + PF < T) prefetch(ring_p + (t % PF) * ROW, pl_u + (size_t)(t + PF) * a.ld);
            cp_async_commit();

            if (t == 0) {
#pragma unroll
                for (int s = 0; s < S; ++s) cur[s] = p[s] + f_start[s];
            } else {
                const float jv = (J > 0) ? junction_value(F, buf) : kNegInf;
                float v[S];
                lane_states(F, buf, jv, v);
#pragma unroll
                for (int s = 0; s < S; ++s) cur[s] = p[s] + v[s];
            }
            float mx = cur[0];
#pragma unroll
            for (int s = 1; s < S; ++s) mx = fmaxf(mx, cur[s]);
            mx = warp_max(mx);
            const float mxs = (mx == kNegInf) ? 0.f : mx;
            logz2 += (double)mxs;
#pragma unroll
            for (int s = 0; s < S; ++s) cur[s] -= mxs;
            __syncwarp();  // every lane has finished reading buf
#pragma unroll
            for (int s = 0; s < S; ++s) buf[s * 32 + lane] = cur[s];
            __syncwarp();
            if (own) {
                float* la_row = la_u + (size_t)t * a.Kw + lane * S;
#pragma unroll
                for (int v = 0; v < S / 4; ++v)
                    if (lane * S + 4 * v < K)
                        reinterpret_cast<float4*>(la_row)[v] = make_float4(cur[4 * v], cur[4 * v + 1], cur[4 * v + 2], cur[4 * v + 3]);
            }
        }
        cp_async_wait<0>();

        if (a.utt_logz != nullptr) {
            float m = kNegInf, v[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                v[s] = cur[s] + b_start[s];
                m = fmaxf(m, v[s]);
            }
            m = warp_max(m);
            const float ms = (m == kNegInf) ? 0.f : m;
            float sum = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) sum += ex2(v[s] - ms);
            sum = warp_sum(sum);
            const double z = (logz2 + (double)ms + (double)lg2(sum)) * (double)kLn2;
            double rs = 0.0;
            if (a.frame_ref != nullptr)
                for (int t = lane; t < T; t += 32) rs += (double)a.frame_ref[t0 + t];
            rs = warp_sum(rs);
            if (lane == 0) a.utt_logz[u] = z + (double)a.scale * rs;
        }

        // ------------------------------ backward -----------------------------
        __threadfence_block();   // this warp's la stores -> visible to its own async copies
        __syncwarp();
        for (int r = 0; r < PF; ++r) {
            const int t = T - 1 - r;
            if (t >= 0) {
                prefetch(ring_p + r * ROW, pl_u + (size_t)t * a.ld);
                prefetch(ring_a + r * ROW, la_u + (size_t)t * a.Kw);
            }
            cp_async_commit();
        }
        float lb[S];
        {
            float m = kNegInf;
#pragma unroll
            for (int s = 0; s < S; ++s) m = fmaxf(m, b_start[s]);
            m = warp_max(m);
            const float ms = (m == kNegInf) ? 0.f : m;
#pragma unroll
            for (int s = 0; s < S; ++s) lb[s] = b_start[s] - ms;
        }
        float ell = 0.f;        // this lane's share of sum_t sum_k p2_tk gamma_tk (flushed every 32 frames)
        double ell_d = 0.0;
        for (int i = 0; i < T; ++i) {
            const int t = T - 1 - i;
            cp_async_wait<PF - 1>();
            const float4* sp = reinterpret_cast<const float4*>(ring_p + (i % PF) * ROW + lane * S);
            const float4* sa = reinterpret_cast<const float4*>(ring_a + (i % PF) * ROW + lane * S);
            float p[S], la[S];
#pragma unroll
            for (int v = 0; v < S / 4; ++v) {
                const float4 q = sp[v], r = sa[v];
                p[4 * v] = q.x * p_scale; p[4 * v + 1] = q.y * p_scale;
                p[4 * v + 2] = q.z * p_scale; p[4 * v + 3] = q.w * p_scale;
                la[4 * v] = r.x; la[4 * v + 1] = r.y; la[4 * v + 2] = r.z; la[4 * v + 3] = r.w;
            }
            if (t - PF >= 0) {
                prefetch(ring_p + (i % PF) * ROW, pl_u + (size_t)(t - PF) * a.ld);
                prefetch(ring_a + (i % PF) * ROW, la_u + (size_t)(t - PF) * a.Kw);
            }
            cp_async_commit();

            // gamma_t (lanes past K: la = 0, lb = -inf -> 0)
            float v[S], m = kNegInf;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                v[s] = la[s] + lb[s];
                m = fmaxf(m, v[s]);
            }
            m = warp_max(m);
            const float ms = (m == kNegInf) ? 0.f : m;
            float sum = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                v[s] = ex2(v[s] - ms);
                sum += v[s];
            }
            sum = warp_sum(sum);
            const float inv = (sum > 0.f) ? __fdividef(1.f, sum) : 0.f;
            float fe = 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                v[s] *= inv;
                if (v[s] > 0.f) fe = fmaf(p[s], v[s], fe);
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
                    a.frame_exp_llh[t0 + t] = f * kLn2 + r;
                }
            }
            if (own) {
                if (a.state_post != nullptr) {
                    float* row = a.state_post + (size_t)(t0 + t) * K + lane * S;
#pragma unroll
                    for (int q = 0; q < S / 4; ++q)
                        if (lane * S + 4 * q < K)
                            reinterpret_cast<float4*>(row)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
                if (a.pdf_post != nullptr) {
                    float* row = a.pdf_post + (size_t)(t0 + t) * a.ld_post + lane * S;
#pragma unroll
                    for (int q = 0; q < S / 4; ++q)
                        if (lane * S + 4 * q < K)
                            reinterpret_cast<float4*>(row)[q] = make_float4(a.scale * v[4 * q], a.scale * v[4 * q + 1],
                                                                           a.scale * v[4 * q + 2], a.scale * v[4 * q + 3]);
                }
            }
            if (t == 0) break;
            // beta_{t-1}: delta_j = p_tj + lb_tj published, then the transposed recursion
            __syncwarp();
#pragma unroll
            for (int s = 0; s < S; ++s) buf[s * 32 + lane] = p[s] + lb[s];
            __syncwarp();
            const float jv = (J > 0) ? junction_value(B, buf) : kNegInf;
            lane_states(B, buf, jv, lb);
            float mb = lb[0];
#pragma unroll
            for (int s = 1; s < S; ++s) mb = fmaxf(mb, lb[s]);
            mb = warp_max(mb);
            const float mbs = (mb == kNegInf) ? 0.f : mb;
#pragma unroll
            for (int s = 0; s < S; ++s) lb[s] -= mbs;
        }
        cp_async_wait<0>();
        ell_d += (double)ell;
        ell_d = warp_sum(ell_d);
        double rs = 0.0;
        if (a.frame_ref != nullptr)
            for (int t = lane; t < T; t += 32) rs += (double)a.frame_ref[t0 + t];
        rs = warp_sum(rs);
        if (lane == 0) a.utt_exp_llh[u] = ell_d * (double)kLn2 + (double)a.scale * rs;
        __syncwarp();
    }
}

template <int S, int JR>
static int launch_fb_fast(const FbArgs& a, int n_utts, cudaStream_t st) {
    constexpr int PF = FbCfg<S>::PF;
    size_t smem = sizeof(float) * (size_t)FB_WARPS * (32 * S + 2 * PF * 32 * S);
    static bool attr_set = false;
    if (!attr_set) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(hmm_fb_fast_kernel<S, JR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        attr_set = true;
    }
    int blocks = (n_utts + FB_WARPS - 1) / FB_WARPS;
    int max_blocks = kNumSMs * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    hmm_fb_fast_kernel<S, JR><<<blocks, FB_WARPS * 32, smem, st>>>(a);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

// ---------------------------------------------------------------------------
// Aligned left-to-right loop: P units of SU states each (self-loop + next), unit ends feeding
// one junction that feeds the unit starts -- the phone loop of mkphoneloopgraph.py with one
// topology for all units.  A lane owns U whole units, so every arc of the recursion connects
// two registers of the same lane or goes through the junction, which is a warp reduction:
// no shared-memory traffic, no warp barriers.  Weights per slot j (log2): w_self[j],
// w_in[j] (from slot j-1, or from the junction for a unit start), w_jout[j] (unit end ->
// junction, -inf elsewhere); the backward recursion uses the same three transposed.
// ---------------------------------------------------------------------------
template <int SU, int U, bool LP>     // LP: also write log2 posteriors (kept out of the plain kernel: registers)
__global__ void __launch_bounds__(FB_WARPS * 32) hmm_fb_lr_kernel(FbArgs a) {
    constexpr int S = SU * U;
    constexpr bool VEC = (S % 4 == 0);
    constexpr int PF = (S <= 4) ? 6 : (S <= 8 ? 4 : 3);
    constexpr int ROW = 32 * S;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ring_p = smem + (size_t)warp * (2 * PF * ROW);   // [PF][32 * S]
    float* ring_a = ring_p + PF * ROW;                      // [PF][32 * S]
    const int K = a.K;
    const float p_scale = a.scale * a.llh_mul;
    const float lp_rel = a.lpost_rel ? -a.llh_mul : 0.f;
    const int gwarp = blockIdx.x * FB_WARPS + warp, nwarps = gridDim.x * FB_WARPS;
    const bool own = lane * S < K;
    for (int i = lane; i < 2 * PF * ROW; i += 32) ring_p[i] = 0.f;   // lanes past K stay finite
    __syncwarp();
    BlockMarker marker;
    marker.init(a, lane * S, own ? min(S, K - lane * S) : 0);

    float w_self[S], w_in[S], w_jout[S], f_start[S], b_start[S];
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const int k = lane * S + j;
        w_self[j] = __ldg(a.lr_w + k);
        w_in[j] = __ldg(a.lr_w + ROW + k);
        w_jout[j] = __ldg(a.lr_w + 2 * ROW + k);
        f_start[j] = (k < K) ? __ldg(a.fwd.start + k) : kNegInf;
        b_start[j] = (k < K) ? __ldg(a.bwd.start + k) : kNegInf;
    }

    auto prefetch = [&](float* slot, const float* row) {
        if (!own) return;
        if constexpr (VEC) {
#pragma unroll
            for (int v = 0; v < S / 4; ++v)
                if (lane * S + 4 * v < K) cp_async16(slot + lane * S + 4 * v, row + lane * S + 4 * v);
        } else {
#pragma unroll
            for (int j = 0; j < S; ++j)
                if (lane * S + j < K) cp_async4(slot + lane * S + j, row + lane * S + j);
        }
    };
    auto read_row = [&](const float* slot, float* out) {
        if constexpr (VEC) {
#pragma unroll
            for (int v = 0; v < S / 4; ++v) {
                const float4 q = reinterpret_cast<const float4*>(slot + lane * S)[v];
                out[4 * v] = q.x; out[4 * v + 1] = q.y; out[4 * v + 2] = q.z; out[4 * v + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < S; ++j) out[j] = slot[lane * S + j];
        }
    };
    auto write_row = [&](float* row, const float* v, float mul) {
        if (!own) return;
        if constexpr (VEC) {
#pragma unroll
            for (int q = 0; q < S / 4; ++q)
                if (lane * S + 4 * q < K)
                    reinterpret_cast<float4*>(row + lane * S)[q] =
                        make_float4(mul * v[4 * q], mul * v[4 * q + 1], mul * v[4 * q + 2], mul * v[4 * q + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < S; ++j)
                if (lane * S + j < K) row[lane * S + j] = mul * v[j];
        }
    };
    // log2-sum-exp2 over the warp of U values per lane
    auto warp_lse = [&](const float* v) {
        float m = v[0];
#pragma unroll
        for (int u = 1; u < U; ++u) m = fmaxf(m, v[u]);
        m = warp_max(m);
        const float ms = (m == kNegInf) ? 0.f : m;
        float sum = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) sum += ex2(v[u] - ms);
        sum = warp_sum(sum);
        return ms + lg2(sum);
    };

    for (int u = gwarp; u < a.n_utts; u += nwarps) {
        const int64_t t0 = a.utt_off[u];
        const int T = (int)(a.utt_off[u + 1] - t0);
        if (T <= 0) {
            if (lane == 0) {
                a.utt_exp_llh[u] = 0.0;
                if (a.utt_logz) a.utt_logz[u] = 0.0;
            }
            continue;
        }
        const float* pl_u = a.pl + (size_t)t0 * a.ld;
        float* la_u = a.la_ws + (size_t)t0 * a.Kw;
        double logz2 = 0.0;

        // ------------------------------ forward ------------------------------
        for (int r = 0; r < PF; ++r) {
            if (r < T) prefetch(ring_p + r * ROW, pl_u + (size_t)r * a.ld);
            cp_async_commit();
        }
        float cur[S];
        const bool want_logz = a.utt_logz != nullptr;
        float lz = 0.f;                                   // normalisers, folded into fp64 every 16 frames
        int slot = 0;                                     // ring slot of frame t
        const float* pf_row = pl_u + (size_t)PF * a.ld;   // row to prefetch next
        float* la_row = la_u;
        for (int t = 0; t < T; ++t) {
            cp_async_wait<PF - 1>();
            float p[S];
            float* ring_slot = ring_p + slot * ROW;
            read_row(ring_slot, p);
            if (t + PF < T) prefetch(ring_slot, pf_row);
            cp_async_commit();
            pf_row += a.ld;
            slot = (slot + 1 == PF) ? 0 : slot + 1;
            if (t == 0) {
#pragma unroll
                for (int j = 0; j < S; ++j) cur[j] = fmaf(p[j], p_scale, f_start[j]);
            } else {
                float ends[U];
#pragma unroll
                for (int q = 0; q < U; ++q) ends[q] = cur[q * SU + SU - 1] + w_jout[q * SU + SU - 1];
                const float jv = warp_lse(ends);
                float v[S];
#pragma unroll
                for (int j = 0; j < S; ++j)
                    v[j] = lse2(cur[j] + w_self[j], ((j % SU == 0) ? jv : cur[j - (j % SU == 0 ? 0 : 1)]) + w_in[j]);
#pragma unroll
                for (int j = 0; j < S; ++j) cur[j] = fmaf(p[j], p_scale, v[j]);
            }
            float mx = cur[0];
#pragma unroll
            for (int j = 1; j < S; ++j) mx = fmaxf(mx, cur[j]);
            mx = warp_max(mx);
            const float mxs = (mx == kNegInf) ? 0.f : mx;
            if (want_logz) {
                lz += mxs;
                if ((t & 15) == 15) {
                    logz2 += (double)lz;
                    lz = 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < S; ++j) cur[j] -= mxs;
            write_row(la_row, cur, 1.f);
            la_row += a.Kw;
        }
        cp_async_wait<0>();
        logz2 += (double)lz;

        if (a.utt_logz != nullptr) {
            float m = kNegInf, v[S];
#pragma unroll
            for (int j = 0; j < S; ++j) {
                v[j] = cur[j] + b_start[j];
                m = fmaxf(m, v[j]);
            }
            m = warp_max(m);
            const float ms = (m == kNegInf) ? 0.f : m;
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < S; ++j) sum += ex2(v[j] - ms);
            sum = warp_sum(sum);
            const double z = (logz2 + (double)ms + (double)lg2(sum)) * (double)kLn2;
            double rs = 0.0;
            if (a.frame_ref != nullptr)
                for (int t = lane; t < T; t += 32) rs += (double)a.frame_ref[t0 + t];
            rs = warp_sum(rs);
            if (lane == 0) a.utt_logz[u] = z + (double)a.scale * rs;
        }

        // ------------------------------ backward -----------------------------
        __threadfence_block();   // this warp's la stores -> visible to its own async copies
        __syncwarp();
        for (int r = 0; r < PF; ++r) {
            const int t = T - 1 - r;
            if (t >= 0) {
                prefetch(ring_p + r * ROW, pl_u + (size_t)t * a.ld);
                prefetch(ring_a + r * ROW, la_u + (size_t)t * a.Kw);
            }
            cp_async_commit();
        }
        float lb[S];
        {
            float m = kNegInf;
#pragma unroll
            for (int j = 0; j < S; ++j) m = fmaxf(m, b_start[j]);
            m = warp_max(m);
            const float ms = (m == kNegInf) ? 0.f : m;
#pragma unroll
            for (int j = 0; j < S; ++j) lb[j] = b_start[j] - ms;
        }
        float ell = 0.f;
        double ell_d = 0.0;
        const bool units = a.unit_counts != nullptr;
        float cnt[U], dstart[U], mbs_prev = 0.f;   // unit counts, delta_{t+1} of the unit starts
#pragma unroll
        for (int q = 0; q < U; ++q) cnt[q] = dstart[q] = 0.f;
        slot = 0;
        const float* pfb_p = pl_u + (ptrdiff_t)(T - 1 - PF) * (ptrdiff_t)a.ld;      // rows to prefetch next (t - PF)
        const float* pfb_a = la_u + (ptrdiff_t)(T - 1 - PF) * (ptrdiff_t)a.Kw;
        for (int i = 0; i < T; ++i) {
            const int t = T - 1 - i;
            cp_async_wait<PF - 1>();
            float p[S], la[S];     // p: RAW llh (the scale is folded into the FMAs below)
            read_row(ring_p + slot * ROW, p);
            read_row(ring_a + slot * ROW, la);
            if (t - PF >= 0) {
                prefetch(ring_p + slot * ROW, pfb_p);
                prefetch(ring_a + slot * ROW, pfb_a);
            }
            cp_async_commit();
            pfb_p -= a.ld;
            pfb_a -= a.Kw;
            slot = (slot + 1 == PF) ? 0 : slot + 1;

            float v[S], m = kNegInf;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                v[j] = la[j] + lb[j];
                m = fmaxf(m, v[j]);
            }
            const float ml = m;              // this lane's largest log posterior (unnormalised)
            m = warp_max(m);
            const float ms = (m == kNegInf) ? 0.f : m;
            float sum = 0.f;
            float vlog[LP ? S : 1];
#pragma unroll
            for (int j = 0; j < S; ++j) {
                if constexpr (LP) vlog[j] = fmaf(p[j], lp_rel, v[j]);       // (relative form: minus the log2 llh)
                v[j] = ex2(v[j] - ms);
                sum += v[j];
            }
            sum = warp_sum(sum);
            const float inv = (sum > 0.f) ? __fdividef(1.f, sum) : 0.f;
            if constexpr (LP) {
                // log2(scale gamma) = log value - log2 sum + log2 scale; an impossible frame (sum = 0) gives -inf, not NaN
                const float lnorm = (sum > 0.f) ? lg2(a.scale) - lg2(sum) - ms : kNegInf;
#pragma unroll
                for (int j = 0; j < S; ++j) vlog[j] += lnorm;
                write_row(a.pdf_lpost + (size_t)(t0 + t) * a.ld_lpost, vlog, 1.f);
                if (a.blk_active != nullptr) marker.frame(a, ml + lnorm, t0 + t, t == 0);
            }
            float fe = 0.f;      // sum_k llh_k gamma_k in raw-llh units (x p_scale at the end); llh is finite
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
            if (units && sum > 0.f) {
                // transition posteriors through the junction, reduced over the unit ends (graph.py:308-323,
                // phoneloop.py:88-97): xi_t(ends -> start_q) = 2^(jv_t + w_in + delta_{t+1}(start_q) - mbs_t - Zg_t)
                if (i > 0) {
                    float ends[U];
#pragma unroll
                    for (int q = 0; q < U; ++q) ends[q] = la[q * SU + SU - 1] + w_jout[q * SU + SU - 1];
                    const float base = warp_lse(ends) - mbs_prev - (ms + lg2(sum));
#pragma unroll
                    for (int q = 0; q < U; ++q) cnt[q] += ex2(base + w_in[q * SU] + dstart[q]);
                }
                if (t == 0) {
#pragma unroll
                    for (int q = 0; q < U; ++q) cnt[q] += v[q * SU];
                }
            }
            if (a.state_post != nullptr) write_row(a.state_post + (size_t)(t0 + t) * K, v, 1.f);
            if (a.pdf_post != nullptr) write_row(a.pdf_post + (size_t)(t0 + t) * a.ld_post, v, a.scale);
            if (t == 0) break;
            // beta_{t-1}: delta_j = p_tj + lb_tj, transposed recursion on registers
            float delta[S], starts[U];
#pragma unroll
            for (int j = 0; j < S; ++j) delta[j] = fmaf(p[j], p_scale, lb[j]);
#pragma unroll
            for (int q = 0; q < U; ++q) starts[q] = delta[q * SU] + w_in[q * SU];
            const float jb = warp_lse(starts);
            float mb = kNegInf;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const bool end = (j % SU == SU - 1);
                const float nxt = end ? jb + w_jout[j] : delta[end ? j : j + 1] + w_in[end ? j : j + 1];
                lb[j] = lse2(delta[j] + w_self[j], nxt);
                mb = fmaxf(mb, lb[j]);
            }
            mb = warp_max(mb);
            const float mbs = (mb == kNegInf) ? 0.f : mb;
#pragma unroll
            for (int j = 0; j < S; ++j) lb[j] -= mbs;
            mbs_prev = mbs;
#pragma unroll
            for (int q = 0; q < U; ++q) dstart[q] = delta[q * SU];
        }
        cp_async_wait<0>();
        if (units) {
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int unit = lane * U + q;
                if (unit * SU < K && cnt[q] != 0.f) atomicAdd(a.unit_counts + unit, (double)cnt[q]);
            }
        }
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

template <int SU, int U, bool LP = false>
static int launch_fb_lr(const FbArgs& a, int n_utts, cudaStream_t st) {
    if (!LP && a.pdf_lpost != nullptr) return launch_fb_lr<SU, U, true>(a, n_utts, st);
    constexpr int S = SU * U;
    constexpr int PF = (S <= 4) ? 6 : (S <= 8 ? 4 : 3);
    size_t smem = sizeof(float) * (size_t)FB_WARPS * (2 * PF * 32 * S);
    static bool attr_set = false;
    if (!attr_set) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(hmm_fb_lr_kernel<SU, U, LP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        attr_set = true;
    }
    int blocks = (n_utts + FB_WARPS - 1) / FB_WARPS;
    int max_blocks = kNumSMs * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    hmm_fb_lr_kernel<SU, U, LP><<<blocks, FB_WARPS * 32, smem, st>>>(a);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

// ---------------------------------------------------------------------------
// The same aligned left-to-right loop for MANY units (P > 128, e.g. 250 units x 4 states = the
// 1000-state HMM of BASELINE configs[2]): W warps share one utterance, every lane owns ONE unit,
// and the two warp reductions of a step (normaliser, junction) become one exchange through shared
// memory + one block barrier (partials are published relative to the warp's own maximum and
// recombined, so a single round suffices; slots are double-buffered, so one barrier per exchange).
// Forward: 1 exchange per frame; backward: 2 (posterior normaliser, junction of the beta recursion).
// ---------------------------------------------------------------------------
template <int SU, int W, bool LP>
__global__ void __launch_bounds__(W * 32, 4) hmm_fb_lrb_kernel(FbArgs a) {
    constexpr int S = SU;
    constexpr bool VEC = (S % 4 == 0);
    constexpr int PF = 4;
    constexpr int ROW = 32 * S;                 // one warp's slice of a row
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ring_p = smem + (size_t)warp * (2 * PF * ROW);
    float* ring_a = ring_p + PF * ROW;
    float* xch = smem + (size_t)W * (2 * PF * ROW);     // [2][W][8] exchange slots
    float* ml_park = xch + 2 * W * 8 + threadIdx.x;     // this thread's lane maximum, parked across the block exchange
    const int K = a.K;
    const float p_scale = a.scale * a.llh_mul;
    const float lp_rel_s = a.lpost_rel ? -1.f / a.scale : 0.f;      // p is scaled below: p / scale = the llh in log2 units
    const int gl = warp * 32 + lane;            // unit owned by this lane
    const int k0 = gl * S;                      // its first state
    const bool own = k0 < K;
    for (int i = lane; i < 2 * PF * ROW; i += 32) ring_p[i] = 0.f;
    __syncwarp();
    BlockMarker marker;
    marker.init(a, k0, own ? min(S, K - k0) : 0);

    float w_self[S], w_in[S], w_jout, f_start[S], b_start[S];
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const int k = k0 + j;
        w_self[j] = own ? __ldg(a.lr_w + k) : kNegInf;
        w_in[j] = own ? __ldg(a.lr_w + a.lr_row + k) : kNegInf;
        f_start[j] = (k < K) ? __ldg(a.fwd.start + k) : kNegInf;
        b_start[j] = (k < K) ? __ldg(a.bwd.start + k) : kNegInf;
    }
    w_jout = own ? __ldg(a.lr_w + 2 * a.lr_row + k0 + S - 1) : kNegInf;

    auto prefetch = [&](float* slot, const float* row) {
        if (!own) return;
        if constexpr (VEC) {
#pragma unroll
            for (int v = 0; v < S / 4; ++v) cp_async16(slot + lane * S + 4 * v, row + k0 + 4 * v);
        } else {
#pragma unroll
            for (int j = 0; j < S; ++j)
                if (k0 + j < K) cp_async4(slot + lane * S + j, row + k0 + j);
        }
    };
    auto read_row = [&](const float* slot, float* out) {
#pragma unroll
        for (int j = 0; j < S; ++j) out[j] = slot[lane * S + j];
    };
    auto write_row = [&](float* row, const float* v, float mul) {
        if (!own) return;
        if constexpr (VEC) {
#pragma unroll
            for (int q = 0; q < S / 4; ++q)
                reinterpret_cast<float4*>(row + k0)[q] =
                    make_float4(mul * v[4 * q], mul * v[4 * q + 1], mul * v[4 * q + 2], mul * v[4 * q + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < S; ++j)
                if (k0 + j < K) row[k0 + j] = mul * v[j];
        }
    };
    // One exchange: every warp publishes up to two (maximum m, sum s of 2^(x - m)) pairs, the first with an extra sum e
    // scaled the same way; afterwards every thread holds the block maxima M, S = sum_w s_w 2^(m_w - M) and E.  The
    // recombination runs ONCE per warp across lanes (lane l takes warp l % W's slot: one exp2 per pair and warp instead
    // of W per thread -- the W-fold version was 46 % of the kernel's special-function work).
    static_assert((W & (W - 1)) == 0 && W <= 32, "W must be a power of two");
    int xn = 0;
    auto sum_w = [&](float v) {        // sum over the W distinct slots (every group of W lanes holds all of them)
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
    auto exchange2 = [&](float m1, float s1, float e1, float m2, float s2, float& M1, float& S1, float& E1, float& M2,
                         float& S2, bool second) {
        float4* slot = reinterpret_cast<float4*>(xch + (xn & 1) * (W * 8));
        ++xn;
        if (lane == 0) {
            slot[warp * 2] = make_float4(m1, s1, e1, m2);
            if (second) slot[warp * 2 + 1] = make_float4(s2, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        const float4 x = slot[(lane & (W - 1)) * 2];
        const float mx1 = warp_max(x.x);
        M1 = (mx1 == kNegInf) ? 0.f : mx1;
        const float f1 = ex2(x.x - M1);           // 2^(-inf) = 0 for warps without reachable states
        S1 = sum_w(x.y * f1);
        E1 = sum_w(x.z * f1);
        M2 = 0.f;
        S2 = 0.f;
        if (second) {
            const float y = slot[(lane & (W - 1)) * 2 + 1].x;
            const float mx2 = warp_max(x.w);
            M2 = (mx2 == kNegInf) ? 0.f : mx2;
            S2 = sum_w(y * ex2(x.w - M2));
        }
    };
    auto exchange = [&](float m, float s, float e, float& M, float& Ssum, float& E) {
        float M2, S2;
        exchange2(m, s, e, kNegInf, 0.f, M, Ssum, E, M2, S2, false);
    };

    for (int u = blockIdx.x; u < a.n_utts; u += gridDim.x) {
        const int64_t t0 = a.utt_off[u];
        const int T = (int)(a.utt_off[u + 1] - t0);
        if (T <= 0) {
            if (threadIdx.x == 0) {
                a.utt_exp_llh[u] = 0.0;
                if (a.utt_logz) a.utt_logz[u] = 0.0;
            }
            continue;
        }
        const float* pl_u = a.pl + (size_t)t0 * a.ld;
        float* la_u = a.la_ws + (size_t)t0 * a.Kw;
        double logz2 = 0.0;

        // ------------------------------ forward ------------------------------
        for (int r = 0; r < PF; ++r) {
            if (r < T) prefetch(ring_p + r * ROW, pl_u + (size_t)r * a.ld);
            cp_async_commit();
        }
        float cur[S], jv = kNegInf;
        for (int t = 0; t < T; ++t) {
            cp_async_wait<PF - 1>();
            float p[S];
            read_row(ring_p + (t % PF) * ROW, p);
            if (t + PF < T) prefetch(ring_p + (t % PF) * ROW, pl_u + (size_t)(t + PF) * a.ld);
            cp_async_commit();
            if (t == 0) {
#pragma unroll
                for (int j = 0; j < S; ++j) cur[j] = fmaf(p[j], p_scale, f_start[j]);
            } else {
                float v[S];
#pragma unroll
                for (int j = 0; j < S; ++j)
                    v[j] = lse2(cur[j] + w_self[j], ((j == 0) ? jv : cur[j == 0 ? 0 : j - 1]) + w_in[j]);
#pragma unroll
                for (int j = 0; j < S; ++j) cur[j] = fmaf(p[j], p_scale, v[j]);
            }
            float ml = cur[0];
#pragma unroll
            for (int j = 1; j < S; ++j) ml = fmaxf(ml, cur[j]);
            ml = warp_max(ml);
            const float mls = (ml == kNegInf) ? 0.f : ml;
            const float sl = warp_sum(ex2(cur[S - 1] + w_jout - mls));
            float mx, js, unused;
            exchange(ml, sl, 0.f, mx, js, unused);
            jv = lg2(js);                         // junction of the NORMALISED values
            logz2 += (double)mx;
#pragma unroll
            for (int j = 0; j < S; ++j) cur[j] -= mx;
            write_row(la_u + (size_t)t * a.Kw, cur, 1.f);
        }
        cp_async_wait<0>();

        if (a.utt_logz != nullptr) {
            float m = kNegInf, v[S];
#pragma unroll
            for (int j = 0; j < S; ++j) {
                v[j] = cur[j] + b_start[j];
                m = fmaxf(m, v[j]);
            }
            m = warp_max(m);
            const float ms = (m == kNegInf) ? 0.f : m;
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < S; ++j) sum += ex2(v[j] - ms);
            sum = warp_sum(sum);
            float M, Ssum, unused;
            exchange(m, sum, 0.f, M, Ssum, unused);
            double rs = 0.0;
            if (a.frame_ref != nullptr && warp == 0)
                for (int t = lane; t < T; t += 32) rs += (double)a.frame_ref[t0 + t];
            rs = warp_sum(rs);
            if (threadIdx.x == 0)
                a.utt_logz[u] = (logz2 + (double)M + (double)lg2(Ssum)) * (double)kLn2 + (double)a.scale * rs;
        }

        // ------------------------------ backward -----------------------------
        __threadfence_block();
        __syncthreads();         // every warp's la stores are visible to every warp's async copies
        for (int r = 0; r < PF; ++r) {
            const int t = T - 1 - r;
            if (t >= 0) {
                prefetch(ring_p + r * ROW, pl_u + (size_t)t * a.ld);
                prefetch(ring_a + r * ROW, la_u + (size_t)t * a.Kw);
            }
            cp_async_commit();
        }
        float lb[S];
        {
            float m = kNegInf;
#pragma unroll
            for (int j = 0; j < S; ++j) m = fmaxf(m, b_start[j]);
            m = warp_max(m);
            float M, s0, s1;
            exchange(m, 0.f, 0.f, M, s0, s1);
#pragma unroll
            for (int j = 0; j < S; ++j) lb[j] = b_start[j] - M;
        }
        float ell = 0.f;
        double ell_d = 0.0;
        for (int i = 0; i < T; ++i) {
            const int t = T - 1 - i;
            cp_async_wait<PF - 1>();
            float p[S], la[S];
            read_row(ring_p + (i % PF) * ROW, p);
            read_row(ring_a + (i % PF) * ROW, la);
#pragma unroll
            for (int j = 0; j < S; ++j) p[j] *= p_scale;
            if (t - PF >= 0) {
                prefetch(ring_p + (i % PF) * ROW, pl_u + (size_t)(t - PF) * a.ld);
                prefetch(ring_a + (i % PF) * ROW, la_u + (size_t)(t - PF) * a.Kw);
            }
            cp_async_commit();

            // ONE exchange per frame: the posterior normaliser of frame t (max, sum, sum of p * 2^(v - max)) and the
            // junction of the beta recursion (max of delta_t = p_t + lb_t, partial sum over the unit starts)
            float v[S], m = kNegInf;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                v[j] = own ? la[j] + lb[j] : kNegInf;
                m = fmaxf(m, v[j]);
            }
            if (a.blk_active != nullptr) *ml_park = m;       // this lane's largest log posterior (unnormalised): no register across the exchange
            m = warp_max(m);
            const float mls = (m == kNegInf) ? 0.f : m;
            float sl = 0.f, pe = 0.f;
            float vlog[LP ? S : 1];
#pragma unroll
            for (int j = 0; j < S; ++j) {
                if constexpr (LP) vlog[j] = fmaf(p[j], lp_rel_s, v[j]);     // (relative form: minus the log2 llh, here where p is live)
                v[j] = ex2(v[j] - mls);
                sl += v[j];
                pe = fmaf(p[j], v[j], pe);
            }
            float delta[S];
            float md = kNegInf;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                delta[j] = p[j] + lb[j];
                md = fmaxf(md, delta[j]);
            }
            md = warp_max(md);
            const float mds = (md == kNegInf) ? 0.f : md;
            float sj = ex2(delta[0] + w_in[0] - mds);
            // three sums in one butterfly
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sl += __shfl_xor_sync(0xffffffffu, sl, o);
                pe += __shfl_xor_sync(0xffffffffu, pe, o);
                sj += __shfl_xor_sync(0xffffffffu, sj, o);
            }
            float ms, sum, pes, Md, Sj;
            exchange2(m, sl, pe, md, sj, ms, sum, pes, Md, Sj, true);
            const float inv = (sum > 0.f) ? __fdividef(1.f, sum) : 0.f;
            const float resc = ex2(mls - ms) * inv;            // this warp's values -> block normalisation
            if constexpr (LP) {
                const float lnorm = (sum > 0.f) ? lg2(a.scale) - ms - lg2(sum) : kNegInf;
#pragma unroll
                for (int j = 0; j < S; ++j) vlog[j] += lnorm;
                write_row(a.pdf_lpost + (size_t)(t0 + t) * a.ld_lpost, vlog, 1.f);
                if (a.blk_active != nullptr) marker.frame(a, *ml_park + lnorm, t0 + t, t == 0);
            }
            if (a.state_post != nullptr || a.pdf_post != nullptr) {
#pragma unroll
                for (int j = 0; j < S; ++j) v[j] *= resc;
                if (a.state_post != nullptr) write_row(a.state_post + (size_t)(t0 + t) * K, v, 1.f);
                if (a.pdf_post != nullptr) write_row(a.pdf_post + (size_t)(t0 + t) * a.ld_post, v, a.scale);
            }
            if (threadIdx.x == 0) {
                ell += pes * inv;
                if ((i & 31) == 31) {
                    ell_d += (double)ell;
                    ell = 0.f;
                }
                if (a.frame_exp_llh != nullptr) {
                    const float r = (a.frame_ref != nullptr) ? a.scale * a.frame_ref[t0 + t] : 0.f;
                    a.frame_exp_llh[t0 + t] = pes * inv * kLn2 + r;
                }
            }
            if (t == 0) break;
            // beta_{t-1} from delta_t and the junction value
            const float jb = Md + lg2(Sj);
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const bool end = (j == S - 1);
                const float nxt = end ? jb + w_jout : delta[end ? j : j + 1] + w_in[end ? j : j + 1];
                lb[j] = lse2(delta[j] + w_self[j], nxt) - Md;      // normalised by max(delta): <= 1
            }
        }
        cp_async_wait<0>();
        if (threadIdx.x == 0) {
            ell_d += (double)ell;
            double rs = 0.0;
            a.utt_exp_llh[u] = ell_d * (double)kLn2;
            (void)rs;
        }
        if (warp == 0) {
            double rs = 0.0;
            if (a.frame_ref != nullptr)
                for (int t = lane; t < T; t += 32) rs += (double)a.frame_ref[t0 + t];
            rs = warp_sum(rs);
            if (lane == 0) a.utt_exp_llh[u] += (double)a.scale * rs;
        }
        __syncthreads();
    }
}

template <int SU, int W, bool LP = false>
static int launch_fb_lrb(const FbArgs& a, int n_utts, cudaStream_t st) {
    if (!LP && a.pdf_lpost != nullptr) return launch_fb_lrb<SU, W, true>(a, n_utts, st);
    constexpr int PF = 4;
    size_t smem = sizeof(float) * ((size_t)W * (2 * PF * 32 * SU) + 2 * W * 8 + W * 32);
    static bool attr_set = false;
    if (!attr_set) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(hmm_fb_lrb_kernel<SU, W, LP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        attr_set = true;
    }
    int blocks = n_utts < kNumSMs * 8 ? n_utts : kNumSMs * 8;
    // (fewer blocks than 8 per SM -- 3 or 2 resident utterances per SM in equal rounds instead of 4 + 4 + a sparse third
    // round at 1250 utterances -- measured slower: 5.51 / 6.15 ms against 5.10 - 5.25)
    hmm_fb_lrb_kernel<SU, W, LP><<<blocks, W * 32, smem, st>>>(a);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

// ---------------------------------------------------------------------------
// The multi-warp loop kernel with U units per lane (hmm_fb_lrc_kernel<SU, W, U>): W x 32 x U >= P units with HALF the
// threads of the one-unit-per-lane kernel, so that more utterances are resident per SM.  The scan is latency bound
// (three block exchanges per frame), its throughput is the number of utterances in flight: 1250 utterances of the
// 1000-state graph (BASELINE configs[2]) over 148 SMs need 8.45 resident utterances per SM to finish in ONE wave;
// the one-unit kernel fits 4 (64 registers x 256 threads) = three waves.  Per-state weights are re-read through the
// read-only cache every step instead of living in registers (56 registers per thread at 9 blocks of 128 threads).
// ---------------------------------------------------------------------------
template <int SU, int W, int U, bool LP>
__global__ void __launch_bounds__(W * 32, 9) hmm_fb_lrc_kernel(FbArgs a) {
    constexpr int S = SU * U;
    static_assert(S % 4 == 0, "float4 rows");
    constexpr int PF = 3;
    constexpr int ROW = 32 * S;                 // one warp's slice of a row
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ring_p = smem + (size_t)warp * (2 * PF * ROW);
    float* ring_a = ring_p + PF * ROW;
    float* xch = smem + (size_t)W * (2 * PF * ROW);     // [2][W][4] exchange slots
    const int K = a.K;
    const float p_scale = a.scale * a.llh_mul;
    const float lp_rel_s = a.lpost_rel ? -1.f / a.scale : 0.f;      // p is scaled below: p / scale = the llh in log2 units
    const int k0 = (warp * 32 + lane) * S;      // first state of this lane
    const bool own = k0 < K;
    for (int i = lane; i < 2 * PF * ROW; i += 32) ring_p[i] = 0.f;
    __syncwarp();
    const float* w_self_p = a.lr_w + k0;
    const float* w_in_p = a.lr_w + a.lr_row + k0;
    float w_jout[U];
#pragma unroll
    for (int q = 0; q < U; ++q) w_jout[q] = own ? __ldg(a.lr_w + 2 * a.lr_row + k0 + q * SU + SU - 1) : kNegInf;

    auto load_w = [&](const float* src, float* out) {       // rows are padded with -inf up to lr_row >= K
#pragma unroll
        for (int v = 0; v < S / 4; ++v) {
            const float4 x = own ? __ldg(reinterpret_cast<const float4*>(src) + v)
                                 : make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
            out[4 * v] = x.x; out[4 * v + 1] = x.y; out[4 * v + 2] = x.z; out[4 * v + 3] = x.w;
        }
    };
    auto prefetch = [&](float* slot, const float* row) {
#pragma unroll
        for (int v = 0; v < S / 4; ++v)
            if (k0 + 4 * v < K) cp_async16(slot + lane * S + 4 * v, row + k0 + 4 * v);
    };
    auto read_row = [&](const float* slot, float* out) {
#pragma unroll
        for (int v = 0; v < S / 4; ++v) {
            const float4 x = reinterpret_cast<const float4*>(slot + lane * S)[v];
            out[4 * v] = x.x; out[4 * v + 1] = x.y; out[4 * v + 2] = x.z; out[4 * v + 3] = x.w;
        }
    };
    auto write_row = [&](float* row, const float* v, float mul) {
#pragma unroll
        for (int q = 0; q < S / 4; ++q)
            if (k0 + 4 * q < K)
                reinterpret_cast<float4*>(row + k0)[q] =
                    make_float4(mul * v[4 * q], mul * v[4 * q + 1], mul * v[4 * q + 2], mul * v[4 * q + 3]);
    };
    int xn = 0;
    auto exchange = [&](float m, float s, float e, float& M, float& Ssum, float& E) {
        float* slot = xch + (xn & 1) * (W * 4);
        ++xn;
        if (lane == 0) {
            slot[warp * 4] = m;
            slot[warp * 4 + 1] = s;
            slot[warp * 4 + 2] = e;
        }
        __syncthreads();
        float mm[W];
        M = kNegInf;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            mm[w] = slot[w * 4];
            M = fmaxf(M, mm[w]);
        }
        const float Ms = (M == kNegInf) ? 0.f : M;
        Ssum = 0.f;
        E = 0.f;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const float f = ex2(mm[w] - Ms);
            Ssum = fmaf(slot[w * 4 + 1], f, Ssum);
            E = fmaf(slot[w * 4 + 2], f, E);
        }
        M = Ms;
    };

    for (int u = blockIdx.x; u < a.n_utts; u += gridDim.x) {
        const int64_t t0 = a.utt_off[u];
        const int T = (int)(a.utt_off[u + 1] - t0);
        if (T <= 0) {
            if (threadIdx.x == 0) {
                a.utt_exp_llh[u] = 0.0;
                if (a.utt_logz) a.utt_logz[u] = 0.0;
            }
            continue;
        }
        const float* pl_u = a.pl + (size_t)t0 * a.ld;
        float* la_u = a.la_ws + (size_t)t0 * a.Kw;
        double logz2 = 0.0;

        // ------------------------------ forward ------------------------------
        for (int r = 0; r < PF; ++r) {
            if (r < T) prefetch(ring_p + r * ROW, pl_u + (size_t)r * a.ld);
            cp_async_commit();
        }
        float cur[S], jv = kNegInf;
        int slot = 0;
        for (int t = 0; t < T; ++t) {
            cp_async_wait<PF - 1>();
            float p[S];
            read_row(ring_p + slot * ROW, p);
            if (t + PF < T) prefetch(ring_p + slot * ROW, pl_u + (size_t)(t + PF) * a.ld);
            cp_async_commit();
            slot = (slot + 1 == PF) ? 0 : slot + 1;
            if (t == 0) {
#pragma unroll
                for (int j = 0; j < S; ++j)
                    cur[j] = fmaf(p[j], p_scale, (k0 + j < K) ? __ldg(a.fwd.start + k0 + j) : kNegInf);
            } else {
                float ws[S], wi[S], v[S];
                load_w(w_self_p, ws);
                load_w(w_in_p, wi);
#pragma unroll
                for (int j = 0; j < S; ++j)
                    v[j] = lse2(cur[j] + ws[j], ((j % SU == 0) ? jv : cur[j - (j % SU == 0 ? 0 : 1)]) + wi[j]);
#pragma unroll
                for (int j = 0; j < S; ++j) cur[j] = fmaf(p[j], p_scale, v[j]);
            }
            float ml = cur[0];
#pragma unroll
            for (int j = 1; j < S; ++j) ml = fmaxf(ml, cur[j]);
            ml = warp_max(ml);
            const float mls = (ml == kNegInf) ? 0.f : ml;
            float se = 0.f;
#pragma unroll
            for (int q = 0; q < U; ++q) se += ex2(cur[q * SU + SU - 1] + w_jout[q] - mls);
            const float sl = warp_sum(se);
            float mx, js, unused;
            exchange(ml, sl, 0.f, mx, js, unused);
            jv = lg2(js);                         // junction of the NORMALISED values
            logz2 += (double)mx;
#pragma unroll
            for (int j = 0; j < S; ++j) cur[j] -= mx;
            write_row(la_u + (size_t)t * a.Kw, cur, 1.f);
        }
        cp_async_wait<0>();

        if (a.utt_logz != nullptr) {
            float m = kNegInf, v[S];
#pragma unroll
            for (int j = 0; j < S; ++j) {
                v[j] = cur[j] + ((k0 + j < K) ? __ldg(a.bwd.start + k0 + j) : kNegInf);
                m = fmaxf(m, v[j]);
            }
            m = warp_max(m);
            const float ms = (m == kNegInf) ? 0.f : m;
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < S; ++j) sum += ex2(v[j] - ms);
            sum = warp_sum(sum);
            float M, Ssum, unused;
            exchange(m, sum, 0.f, M, Ssum, unused);
            double rs = 0.0;
            if (a.frame_ref != nullptr && warp == 0)
                for (int t = lane; t < T; t += 32) rs += (double)a.frame_ref[t0 + t];
            rs = warp_sum(rs);
            if (threadIdx.x == 0)
                a.utt_logz[u] = (logz2 + (double)M + (double)lg2(Ssum)) * (double)kLn2 + (double)a.scale * rs;
        }

        // ------------------------------ backward -----------------------------
        __threadfence_block();
        __syncthreads();         // every warp's la stores are visible to every warp's async copies
        for (int r = 0; r < PF; ++r) {
            const int t = T - 1 - r;
            if (t >= 0) {
                prefetch(ring_p + r * ROW, pl_u + (size_t)t * a.ld);
                prefetch(ring_a + r * ROW, la_u + (size_t)t * a.Kw);
            }
            cp_async_commit();
        }
        float lb[S];
        {
            float m = kNegInf;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                lb[j] = (k0 + j < K) ? __ldg(a.bwd.start + k0 + j) : kNegInf;
                m = fmaxf(m, lb[j]);
            }
            m = warp_max(m);
            float M, s0, s1;
            exchange(m, 0.f, 0.f, M, s0, s1);
#pragma unroll
            for (int j = 0; j < S; ++j) lb[j] -= M;
        }
        float ell = 0.f;
        double ell_d = 0.0;
        slot = 0;
        for (int i = 0; i < T; ++i) {
            const int t = T - 1 - i;
            cp_async_wait<PF - 1>();
            float p[S], v[S];
            read_row(ring_p + slot * ROW, p);
            read_row(ring_a + slot * ROW, v);
#pragma unroll
            for (int j = 0; j < S; ++j) p[j] *= p_scale;
            if (t - PF >= 0) {
                prefetch(ring_p + slot * ROW, pl_u + (size_t)(t - PF) * a.ld);
                prefetch(ring_a + slot * ROW, la_u + (size_t)(t - PF) * a.Kw);
            }
            cp_async_commit();
            slot = (slot + 1 == PF) ? 0 : slot + 1;

            // gamma_t: exchange (max, sum, sum of p * 2^(v - max))
            float m = kNegInf;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                v[j] = (k0 + j < K) ? v[j] + lb[j] : kNegInf;
                m = fmaxf(m, v[j]);
            }
            m = warp_max(m);
            const float mls = (m == kNegInf) ? 0.f : m;
            float sl = 0.f, pe = 0.f;
            float vlog[LP ? S : 1];
#pragma unroll
            for (int j = 0; j < S; ++j) {
                if constexpr (LP) vlog[j] = fmaf(p[j], lp_rel_s, v[j]);     // (relative form: minus the log2 llh, here where p is live)
                v[j] = ex2(v[j] - mls);
                sl += v[j];
                pe = fmaf(p[j], v[j], pe);
            }
            sl = warp_sum(sl);
            pe = warp_sum(pe);
            float ms, sum, pes;
            exchange(m, sl, pe, ms, sum, pes);
            const float inv = (sum > 0.f) ? __fdividef(1.f, sum) : 0.f;
            const float resc = ex2(mls - ms) * inv;            // this warp's values -> block normalisation
            if constexpr (LP) {
                const float lnorm = (sum > 0.f) ? lg2(a.scale) - ms - lg2(sum) : kNegInf;
#pragma unroll
                for (int j = 0; j < S; ++j) vlog[j] += lnorm;
                write_row(a.pdf_lpost + (size_t)(t0 + t) * a.ld_lpost, vlog, 1.f);
            }
            if (a.state_post != nullptr || a.pdf_post != nullptr) {
#pragma unroll
                for (int j = 0; j < S; ++j) v[j] *= resc;
                if (a.state_post != nullptr) write_row(a.state_post + (size_t)(t0 + t) * K, v, 1.f);
                if (a.pdf_post != nullptr) write_row(a.pdf_post + (size_t)(t0 + t) * a.ld_post, v, a.scale);
            }
            if (threadIdx.x == 0) {
                ell += pes * inv;
                if ((i & 31) == 31) {
                    ell_d += (double)ell;
                    ell = 0.f;
                }
                if (a.frame_exp_llh != nullptr) {
                    const float r = (a.frame_ref != nullptr) ? a.scale * a.frame_ref[t0 + t] : 0.f;
                    a.frame_exp_llh[t0 + t] = pes * inv * kLn2 + r;
                }
            }
            if (t == 0) break;
            // beta_{t-1}: exchange (max of delta, junction partial over the unit starts); delta overwrites p
            float ws[S], wi[S];
            load_w(w_self_p, ws);
            load_w(w_in_p, wi);
            float md = kNegInf;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                p[j] += lb[j];
                md = fmaxf(md, p[j]);
            }
            md = warp_max(md);
            const float mds = (md == kNegInf) ? 0.f : md;
            float sj = 0.f;
#pragma unroll
            for (int q = 0; q < U; ++q) sj += ex2(p[q * SU] + wi[q * SU] - mds);
            sj = warp_sum(sj);
            float Md, Sj, unused;
            exchange(md, sj, 0.f, Md, Sj, unused);
            const float jb = Md + lg2(Sj);
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const bool end = (j % SU == SU - 1);
                const float nxt = end ? jb + w_jout[j / SU] : p[end ? j : j + 1] + wi[end ? j : j + 1];
                lb[j] = lse2(p[j] + ws[j], nxt) - Md;      // normalised by max(delta): <= 1
            }
        }
        cp_async_wait<0>();
        if (threadIdx.x == 0) {
            ell_d += (double)ell;
            a.utt_exp_llh[u] = ell_d * (double)kLn2;
        }
        if (warp == 0) {
            double rs = 0.0;
            if (a.frame_ref != nullptr)
                for (int t = lane; t < T; t += 32) rs += (double)a.frame_ref[t0 + t];
            rs = warp_sum(rs);
            __syncwarp();
            if (lane == 0) a.utt_exp_llh[u] += (double)a.scale * rs;
        }
        __syncthreads();
    }
}

template <int SU, int W, int U, bool LP = false>
static int launch_fb_lrc(const FbArgs& a, int n_utts, cudaStream_t st) {
    if (!LP && a.pdf_lpost != nullptr) return launch_fb_lrc<SU, W, U, true>(a, n_utts, st);
    constexpr int PF = 3;
    size_t smem = sizeof(float) * ((size_t)W * (2 * PF * 32 * SU * U) + 2 * W * 4);
    static bool attr_set = false;
    if (!attr_set) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(hmm_fb_lrc_kernel<SU, W, U, LP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        attr_set = true;
    }
    int blocks = n_utts < kNumSMs * 9 ? n_utts : kNumSMs * 9;
    hmm_fb_lrc_kernel<SU, W, U, LP><<<blocks, W * 32, smem, st>>>(a);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

// ---------------------------------------------------------------------------
// Aligned left-to-right loop with MANY units on ONE warp per utterance (hmm_fb_lrw_kernel<SU, U>): a lane owns U = 8
// whole units (S = 32 states for the 250 x 4 = 1000-state graph of BASELINE configs[2]).  The multi-warp kernels above
// spend most of their ~4300 warp-instructions per frame on what they replicate per warp (ring bookkeeping, block
// exchanges recombined by every warp, reductions): with the whole utterance on one warp the junction and the
// normaliser are warp reductions again, a frame costs ~1100 warp-instructions and 32 independent log-add chains per
// lane hide each other's latency.  Registers hold only the recursion state (alpha / beta row slice, llh slice); the
// per-state weights live in shared memory, shared by the 12 warps of the block, rows of the llh / alpha rings and the
// weights are stored float4-interleaved ([vector][lane]) so that 16-byte accesses of a warp are conflict-free.
// ---------------------------------------------------------------------------
constexpr int LRW_WARPS = 12;

template <int SU, int U, bool LP>
__global__ void __launch_bounds__(LRW_WARPS * 32, 1) hmm_fb_lrw_kernel(FbArgs a) {
    constexpr int S = SU * U, V = S / 4;
    static_assert(S % 4 == 0, "float4 rows");
    constexpr int PF = 2;
    constexpr int ROW = 32 * S;
    extern __shared__ __align__(16) float smem[];
    float* s_wself = smem;                      // [V][32] float4, interleaved
    float* s_win = smem + ROW;
    float* s_wjout = smem + 2 * ROW;            // [32 * U] per unit (end state -> junction)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ring_p = smem + 2 * ROW + 32 * U + (size_t)warp * (2 * PF * ROW);   // [PF][ROW]
    float* ring_a = ring_p + PF * ROW;
    const int K = a.K;
    const float p_scale = a.scale * a.llh_mul;
    const float lp_rel = a.lpost_rel ? -a.llh_mul : 0.f;
    const int gwarp = blockIdx.x * LRW_WARPS + warp, nwarps = gridDim.x * LRW_WARPS;
    const int k0 = lane * S;

    for (int i = threadIdx.x; i < ROW; i += blockDim.x) {
        // state k = l * S + 4 v + e  ->  interleaved position (v * 32 + l) * 4 + e
        const int l = i / S, r = i - l * S, pos = ((r >> 2) * 32 + l) * 4 + (r & 3);
        s_wself[pos] = __ldg(a.lr_w + i);
        s_win[pos] = __ldg(a.lr_w + a.lr_row + i);
    }
    for (int i = threadIdx.x; i < 32 * U; i += blockDim.x) s_wjout[i] = __ldg(a.lr_w + 2 * a.lr_row + i * SU + SU - 1);
    for (int i = lane; i < 2 * PF * ROW; i += 32) ring_p[i] = 0.f;
    __syncthreads();

    auto prefetch = [&](float* slot, const float* row) {
#pragma unroll
        for (int v = 0; v < V; ++v)
            if (k0 + 4 * v < K) cp_async16(slot + (v * 32 + lane) * 4, row + k0 + 4 * v);
    };
    auto read_row = [&](const float* slot, float* out) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float4 q = reinterpret_cast<const float4*>(slot)[v * 32 + lane];
            out[4 * v] = q.x; out[4 * v + 1] = q.y; out[4 * v + 2] = q.z; out[4 * v + 3] = q.w;
        }
    };
    auto write_row = [&](float* row, const float* v, float mul) {
#pragma unroll
        for (int q = 0; q < V; ++q)
            if (k0 + 4 * q < K)
                reinterpret_cast<float4*>(row + k0)[q] =
                    make_float4(mul * v[4 * q], mul * v[4 * q + 1], mul * v[4 * q + 2], mul * v[4 * q + 3]);
    };
    // log2-sum-exp2 over the warp of U values per lane
    auto warp_lse = [&](const float* v) {
        float m = v[0];
#pragma unroll
        for (int u = 1; u < U; ++u) m = fmaxf(m, v[u]);
        m = warp_max(m);
        const float ms = (m == kNegInf) ? 0.f : m;
        float sum = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) sum += ex2(v[u] - ms);
        sum = warp_sum(sum);
        return ms + lg2(sum);
    };
    auto row_max = [&](const float* v) {
        float m = v[0];
#pragma unroll
        for (int j = 1; j < S; ++j) m = fmaxf(m, v[j]);
        return warp_max(m);
    };

    for (int u = gwarp; u < a.n_utts; u += nwarps) {
        const int64_t t0 = a.utt_off[u];
        const int T = (int)(a.utt_off[u + 1] - t0);
        if (T <= 0) {
            if (lane == 0) {
                a.utt_exp_llh[u] = 0.0;
                if (a.utt_logz) a.utt_logz[u] = 0.0;
            }
            continue;
        }
        const float* pl_u = a.pl + (size_t)t0 * a.ld;
        float* la_u = a.la_ws + (size_t)t0 * a.Kw;
        double logz2 = 0.0;

        // ------------------------------ forward ------------------------------
        for (int r = 0; r < PF; ++r) {
            if (r < T) prefetch(ring_p + r * ROW, pl_u + (size_t)r * a.ld);
            cp_async_commit();
        }
        float cur[S];
        const bool want_logz = a.utt_logz != nullptr;
        float lz = 0.f;
        for (int t = 0; t < T; ++t) {
            cp_async_wait<PF - 1>();
            float p[S];
            float* ring_slot = ring_p + (t & (PF - 1)) * ROW;
            read_row(ring_slot, p);
            if (t + PF < T) prefetch(ring_slot, pl_u + (size_t)(t + PF) * a.ld);
            cp_async_commit();
            if (t == 0) {
#pragma unroll
                for (int j = 0; j < S; ++j)
                    cur[j] = fmaf(p[j], p_scale, (k0 + j < K) ? __ldg(a.fwd.start + k0 + j) : kNegInf);
            } else {
                float ends[U];
#pragma unroll
                for (int q = 0; q < U; ++q) ends[q] = cur[q * SU + SU - 1] + s_wjout[lane * U + q];
                const float jv = warp_lse(ends);
                // in place, last state first: state j needs the OLD value of state j - 1
#pragma unroll
                for (int v = V - 1; v >= 0; --v) {
                    const float4 ws4 = reinterpret_cast<const float4*>(s_wself)[v * 32 + lane];
                    const float4 wi4 = reinterpret_cast<const float4*>(s_win)[v * 32 + lane];
                    const float ws[4] = {ws4.x, ws4.y, ws4.z, ws4.w}, wi[4] = {wi4.x, wi4.y, wi4.z, wi4.w};
#pragma unroll
                    for (int e = 3; e >= 0; --e) {
                        const int j = 4 * v + e;
                        const float prev = (j % SU == 0) ? jv : cur[j == 0 ? 0 : j - 1];
                        cur[j] = fmaf(p[j], p_scale, lse2(cur[j] + ws[e], prev + wi[e]));
                    }
                }
            }
            const float mx = row_max(cur);
            const float mxs = (mx == kNegInf) ? 0.f : mx;
            if (want_logz) {
                lz += mxs;
                if ((t & 15) == 15) {
                    logz2 += (double)lz;
                    lz = 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < S; ++j) cur[j] -= mxs;
            write_row(la_u + (size_t)t * a.Kw, cur, 1.f);
        }
        cp_async_wait<0>();
        logz2 += (double)lz;

        if (a.utt_logz != nullptr) {
            float m = kNegInf, v[S];
#pragma unroll
            for (int j = 0; j < S; ++j) {
                v[j] = cur[j] + ((k0 + j < K) ? __ldg(a.bwd.start + k0 + j) : kNegInf);
                m = fmaxf(m, v[j]);
            }
            m = warp_max(m);
            const float ms = (m == kNegInf) ? 0.f : m;
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < S; ++j) sum += ex2(v[j] - ms);
            sum = warp_sum(sum);
            const double z = (logz2 + (double)ms + (double)lg2(sum)) * (double)kLn2;
            double rs = 0.0;
            if (a.frame_ref != nullptr)
                for (int t = lane; t < T; t += 32) rs += (double)a.frame_ref[t0 + t];
            rs = warp_sum(rs);
            if (lane == 0) a.utt_logz[u] = z + (double)a.scale * rs;
        }

        // ------------------------------ backward -----------------------------
        __threadfence_block();   // this warp's la stores -> visible to its own async copies
        __syncwarp();
        for (int r = 0; r < PF; ++r) {
            const int t = T - 1 - r;
            if (t >= 0) {
                prefetch(ring_p + r * ROW, pl_u + (size_t)t * a.ld);
                prefetch(ring_a + r * ROW, la_u + (size_t)t * a.Kw);
            }
            cp_async_commit();
        }
        float lb[S];
        {
#pragma unroll
            for (int j = 0; j < S; ++j) lb[j] = (k0 + j < K) ? __ldg(a.bwd.start + k0 + j) : kNegInf;
            const float m = row_max(lb);
            const float ms = (m == kNegInf) ? 0.f : m;
#pragma unroll
            for (int j = 0; j < S; ++j) lb[j] -= ms;
        }
        float ell = 0.f;
        double ell_d = 0.0;
        const bool units = a.unit_counts != nullptr;
        float cnt[U], dstart[U], mbs_prev = 0.f;   // unit counts, delta_{t+1} of the unit starts
#pragma unroll
        for (int q = 0; q < U; ++q) cnt[q] = dstart[q] = 0.f;
        for (int i = 0; i < T; ++i) {
            const int t = T - 1 - i;
            cp_async_wait<PF - 1>();
            float p[S], la[S];     // p: RAW llh (the scale is folded into the FMAs below)
            read_row(ring_p + (i & (PF - 1)) * ROW, p);
            read_row(ring_a + (i & (PF - 1)) * ROW, la);
            if (t - PF >= 0) {
                prefetch(ring_p + (i & (PF - 1)) * ROW, pl_u + (size_t)(t - PF) * a.ld);
                prefetch(ring_a + (i & (PF - 1)) * ROW, la_u + (size_t)(t - PF) * a.Kw);
            }
            cp_async_commit();

            float jvf = 0.f;       // forward junction value of frame t (unit counts only)
            if (units && i > 0) {
                float ends[U];
#pragma unroll
                for (int q = 0; q < U; ++q) ends[q] = la[q * SU + SU - 1] + s_wjout[lane * U + q];
                jvf = warp_lse(ends);
            }
            // la := log2 of the unnormalised posteriors relative to their maximum
#pragma unroll
            for (int j = 0; j < S; ++j) la[j] += lb[j];
            const float m = row_max(la);
            const float ms = (m == kNegInf) ? 0.f : m;
            float sum = 0.f, fe = 0.f;      // fe: sum_k llh_k gamma_k in raw-llh units (x p_scale at the end)
            if constexpr (LP) {
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    la[j] -= ms;
                    const float e = ex2(la[j]);
                    sum += e;
                    fe = fmaf(p[j], e, fe);
                }
            } else {
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    la[j] = ex2(la[j] - ms);
                    sum += la[j];
                    fe = fmaf(p[j], la[j], fe);
                }
            }
            sum = warp_sum(sum);
            const float inv = (sum > 0.f) ? __fdividef(1.f, sum) : 0.f;
            fe *= inv;
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
            if (units && sum > 0.f) {
                // transition posteriors through the junction, reduced over the unit ends (graph.py:308-323,
                // phoneloop.py:88-97): xi_t(ends -> start_q) = 2^(jv_t + w_in + delta_{t+1}(start_q) - mbs_t - Zg_t)
                if (i > 0) {
                    const float base = jvf - mbs_prev - (ms + lg2(sum));
#pragma unroll
                    for (int q = 0; q < U; ++q) {
                        const int j = q * SU;
                        cnt[q] += ex2(base + s_win[((j >> 2) * 32 + lane) * 4 + (j & 3)] + dstart[q]);
                    }
                }
                if (t == 0) {
#pragma unroll
                    for (int q = 0; q < U; ++q) cnt[q] += (LP ? ex2(la[q * SU]) : la[q * SU]) * inv;
                }
            }
            if constexpr (LP) {
                // log2(scale gamma) = log value - log2 sum + log2 scale; an impossible frame (sum = 0) gives -inf
                const float lnorm = (sum > 0.f) ? lg2(a.scale) - lg2(sum) : kNegInf;
#pragma unroll
                for (int j = 0; j < S; ++j) la[j] = fmaf(p[j], lp_rel, la[j] + lnorm);
                write_row(a.pdf_lpost + (size_t)(t0 + t) * a.ld_lpost, la, 1.f);
            } else {
                if (a.state_post != nullptr) write_row(a.state_post + (size_t)(t0 + t) * K, la, inv);
                if (a.pdf_post != nullptr) write_row(a.pdf_post + (size_t)(t0 + t) * a.ld_post, la, a.scale * inv);
            }
            if (t == 0) break;
            // beta_{t-1}: lb := delta_j = p_tj + lb_tj, then the transposed recursion in place, first state first
#pragma unroll
            for (int j = 0; j < S; ++j) lb[j] = fmaf(p[j], p_scale, lb[j]);
            float starts[U];
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int j = q * SU;
                dstart[q] = lb[j];
                starts[q] = lb[j] + s_win[((j >> 2) * 32 + lane) * 4 + (j & 3)];
            }
            const float jb = warp_lse(starts);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 ws4 = reinterpret_cast<const float4*>(s_wself)[v * 32 + lane];
                // w_in of states 4v+1 .. 4v+4 (the successor of each state of the group)
                const float4 wi4 = reinterpret_cast<const float4*>(s_win)[v * 32 + lane];
                const float wi_next = (v + 1 < V) ? s_win[((v + 1) * 32 + lane) * 4] : kNegInf;
                const float ws[4] = {ws4.x, ws4.y, ws4.z, ws4.w}, wn[4] = {wi4.y, wi4.z, wi4.w, wi_next};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = 4 * v + e;
                    const bool end = (j % SU == SU - 1);
                    const float nxt = end ? jb + s_wjout[lane * U + j / SU] : lb[j + 1 < S ? j + 1 : j] + wn[e];
                    lb[j] = lse2(lb[j] + ws[e], nxt);
                }
            }
            const float mb = row_max(lb);
            const float mbs = (mb == kNegInf) ? 0.f : mb;
#pragma unroll
            for (int j = 0; j < S; ++j) lb[j] -= mbs;
            mbs_prev = mbs;
        }
        cp_async_wait<0>();
        if (units) {
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int unit = lane * U + q;
                if (unit * SU < K && cnt[q] != 0.f) atomicAdd(a.unit_counts + unit, (double)cnt[q]);
            }
        }
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

template <int SU, int U, bool LP = false>
static int launch_fb_lrw(const FbArgs& a, int n_utts, cudaStream_t st) {
    if (!LP && a.pdf_lpost != nullptr) return launch_fb_lrw<SU, U, true>(a, n_utts, st);
    constexpr int S = SU * U, PF = 2;
    size_t smem = sizeof(float) * ((size_t)2 * 32 * S + 32 * U + (size_t)LRW_WARPS * (2 * PF * 32 * S));
    static bool attr_set = false;
    if (!attr_set) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(hmm_fb_lrw_kernel<SU, U, LP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        attr_set = true;
    }
    int blocks = (n_utts + LRW_WARPS - 1) / LRW_WARPS;
    if (blocks > kNumSMs) blocks = kNumSMs;
    hmm_fb_lrw_kernel<SU, U, LP><<<blocks, LRW_WARPS * 32, smem, st>>>(a);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

// ------------------------------- Viterbi -----------------------------------
struct VitArgs {
    ScanLists vit;
    const float* vit_final;
    int K;
    const int* map;
    int map_identity;
    const float* pl;
    int64_t ld;
    const int64_t* utt_off;
    int n_utts;
    float scale;
    uint16_t* bt;  // [N, K]
    int32_t* path;
};

template <int S>
__global__ void __launch_bounds__(FB_WARPS * 32) hmm_viterbi_kernel(VitArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = a.K;
    const int nbuf = (K + 3) & ~3;
    float* buf = smem + (size_t)warp * nbuf;
    const int gwarp = blockIdx.x * FB_WARPS + warp, nwarps = gridDim.x * FB_WARPS;
    const bool ident = a.map_identity != 0;

    for (int u = gwarp; u < a.n_utts; u += nwarps) {
        const int64_t t0 = a.utt_off[u];
        const int T = (int)(a.utt_off[u + 1] - t0);
        if (T <= 0) continue;
        float om[S];
        for (int t = 0; t < T; ++t) {
            const float* row = a.pl + (size_t)(t0 + t) * a.ld;
            float p[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                int k = lane * S + s;
                p[s] = (k < K) ? a.scale * __ldg(row + (ident ? k : __ldg(a.map + k))) : kNegInf;
            }
            if (t == 0) {
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    int k = lane * S + s;
                    om[s] = (k < K) ? p[s] + __ldg(a.vit.start + k) : kNegInf;
                }
            } else {
                uint16_t* bt_row = a.bt + (size_t)(t0 + t) * K;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    int k = lane * S + s;
                    const int row0 = __ldg(a.vit.st_off + s), cnt = __ldg(a.vit.st_cnt + s);
                    float best = kNegInf;
                    int arg = 0;
                    for (int q = 0; q < cnt; ++q) {
                        int i = (row0 + q) * 32 + lane;
                        int sidx = __ldg(a.vit.src + i);
                        float v = buf[sidx] + __ldg(a.vit.lw + i);
                        if (v > best) { best = v; arg = sidx; }
                    }
                    if (k < K) {
                        om[s] = p[s] + best;
                        bt_row[k] = (uint16_t)arg;
                    } else {
                        om[s] = kNegInf;
                    }
                }
            }
            float mx = om[0];
#pragma unroll
            for (int s = 1; s < S; ++s) mx = fmaxf(mx, om[s]);
            mx = warp_max(mx);
            const float mxs = (mx == kNegInf) ? 0.f : mx;
            __syncwarp();
#pragma unroll
            for (int s = 0; s < S; ++s) {
                om[s] -= mxs;
                if (lane * S + s < K) buf[lane * S + s] = om[s];
            }
            __syncwarp();
        }
        // last state: first maximal index of omega + final
        float best = kNegInf;
        int arg = 0x7fffffff;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            int k = lane * S + s;
            if (k < K) {
                float v = om[s] + __ldg(a.vit_final + k);
                if (arg == 0x7fffffff || v > best) { best = v; arg = k; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        // torch.argmax of an all -inf row is 0; states of higher lanes never win ties
        __threadfence_block();
        __syncwarp();
        if (lane == 0) {
            int k = arg;
            a.path[t0 + T - 1] = k;
            for (int t = T - 1; t >= 1; --t) {
                k = a.bt[(size_t)(t0 + t) * K + k];
                a.path[t0 + t - 1] = k;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Viterbi over an aligned left-to-right loop with one unit per lane (P <= 32 units of SU states: the phone loop of
// BASELINE configs[1]).  Same arithmetic and the same first-max tie-breaking as the generic kernel (the candidates of
// a state are compared in ascending source order with a strict >, torch.argmax of graph.py:329-344), but the
// recursion lives in registers: a non-start state has two candidates (previous state, itself), a unit start has the
// 32 unit ends -- fetched by shuffles and weighted with the DENSE ln A[end, start] values, not the factored ones,
// so that near-ties resolve exactly as in the reference -- and itself.  llh rows come through a cp.async ring,
// back-pointers are one uint16 per lane and frame (2 bits per inner state, 6 bits for the unit start: 64 bytes per
// frame instead of 2 K), and the backtrack walks them through shared memory 32 frames at a time.
// ---------------------------------------------------------------------------
template <int SU>
__global__ void __launch_bounds__(FB_WARPS * 32) hmm_viterbi_lr_kernel(VitArgs a) {
    constexpr int S = SU, PF = 6, ROW = 32 * S;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ring = smem + (size_t)warp * (PF * ROW + 512);           // [PF][32 * S] llh rows
    uint16_t* bt_s = reinterpret_cast<uint16_t*>(ring + PF * ROW);   // [32 frames][32 lanes]
    const int K = a.K, P = K / SU;
    const bool own = lane < P;
    const int gwarp = blockIdx.x * FB_WARPS + warp, nwarps = gridDim.x * FB_WARPS;
    uint16_t* bt16 = a.bt;                                           // [N][32]

    // weights of this lane's unit (natural log, dense lists of the plan)
    float w_self[S], w_prev[S], wend[32], start[S], fin[S];
#pragma unroll
    for (int v = 0; v < 32; ++v) wend[v] = kNegInf;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int k = lane * S + s;
        w_self[s] = w_prev[s] = kNegInf;
        start[s] = (k < K) ? __ldg(a.vit.start + k) : kNegInf;
        fin[s] = (k < K) ? __ldg(a.vit_final + k) : kNegInf;
        const int row0 = __ldg(a.vit.st_off + s), cnt = __ldg(a.vit.st_cnt + s);
        for (int q = 0; q < cnt; ++q) {
            const int i = (row0 + q) * 32 + lane;
            const int src = __ldg(a.vit.src + i);
            const float lw = __ldg(a.vit.lw + i);
            if (k >= K || lw == kNegInf) continue;
            if (src == k) {
                w_self[s] = lw;
            } else if (s > 0) {
                if (src == k - 1) w_prev[s] = lw;
            } else {
#pragma unroll
                for (int v = 0; v < 32; ++v)
                    if (src == v * SU + SU - 1) wend[v] = lw;
            }
        }
    }
    for (int i = lane; i < PF * ROW; i += 32) ring[i] = 0.f;
    __syncwarp();

    auto prefetch = [&](float* slot, const float* row) {
        if (!own) return;
        if constexpr (SU == 4) {
            cp_async16(slot + lane * 4, row + lane * 4);
        } else {
#pragma unroll
            for (int j = 0; j < S; ++j) cp_async4(slot + lane * S + j, row + lane * S + j);
        }
    };

    for (int u = gwarp; u < a.n_utts; u += nwarps) {
        const int64_t t0 = a.utt_off[u];
        const int T = (int)(a.utt_off[u + 1] - t0);
        if (T <= 0) continue;
        const float* pl_u = a.pl + (size_t)t0 * a.ld;
        for (int r = 0; r < PF; ++r) {
            if (r < T) prefetch(ring + r * ROW, pl_u + (size_t)r * a.ld);
            cp_async_commit();
        }
        float om[S];
        int slot = 0;
        for (int t = 0; t < T; ++t) {
            cp_async_wait<PF - 1>();
            float p[S];
#pragma unroll
            for (int s = 0; s < S; ++s) p[s] = own ? a.scale * ring[slot * ROW + lane * S + s] : kNegInf;
            if (t + PF < T) prefetch(ring + slot * ROW, pl_u + (size_t)(t + PF) * a.ld);
            cp_async_commit();
            slot = (slot + 1 == PF) ? 0 : slot + 1;
            if (t == 0) {
#pragma unroll
                for (int s = 0; s < S; ++s) om[s] = p[s] + start[s];
            } else {
                // unit start: first maximum over the unit ends (ascending source), then this state itself
                const float e = om[S - 1];
                float best = kNegInf;
                int code = 33;                                   // 33: every candidate is -inf (argmax -> state 0)
#pragma unroll
                for (int v = 0; v < 32; ++v) {
                    const float c = __shfl_sync(0xffffffffu, e, v) + wend[v];
                    if (c > best) {
                        best = c;
                        code = v;
                    }
                }
                const float self0 = om[0] + w_self[0];
                // source of the self arc is SU * lane: it precedes the end of unit v iff lane <= v
                if (self0 > best || (self0 == best && self0 != kNegInf && lane <= code)) {
                    best = self0;
                    code = 32;
                }
                unsigned packed = (unsigned)code << 8;
                float nw[S];
                nw[0] = p[0] + best;
#pragma unroll
                for (int s = 1; s < S; ++s) {
                    const float prev = om[s - 1] + w_prev[s], self = om[s] + w_self[s];
                    float b = prev;
                    unsigned c = (prev == kNegInf) ? 2u : 1u;    // 1: previous state, 0: itself, 2: all -inf
                    if (self > b) {
                        b = self;
                        c = 0u;
                    }
                    packed |= c << (2 * (s - 1));
                    nw[s] = p[s] + b;
                }
#pragma unroll
                for (int s = 0; s < S; ++s) om[s] = own ? nw[s] : kNegInf;
                bt16[(size_t)(t0 + t) * 32 + lane] = (uint16_t)packed;
            }
            float mx = om[0];
#pragma unroll
            for (int s = 1; s < S; ++s) mx = fmaxf(mx, om[s]);
            mx = warp_max(mx);
            const float mxs = (mx == kNegInf) ? 0.f : mx;
#pragma unroll
            for (int s = 0; s < S; ++s) om[s] -= mxs;
        }
        cp_async_wait<0>();
        // last state: first maximal index of omega + final
        float best = kNegInf;
        int arg = 0x7fffffff;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int k = lane * S + s;
            if (k < K) {
                const float v = om[s] + fin[s];
                if (arg == 0x7fffffff || v > best) { best = v; arg = k; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        __threadfence_block();
        __syncwarp();
        // backtrack, 32 frames of back-pointers at a time through shared memory
        int k = arg;
        if (lane == 0) a.path[t0 + T - 1] = k;
        for (int tb = T - 1; tb >= 1; tb -= 32) {
            const int t = tb - lane;
            if (t >= 1) {
                const uint4* src = reinterpret_cast<const uint4*>(bt16 + (size_t)(t0 + t) * 32);
                uint4* dst = reinterpret_cast<uint4*>(bt_s + lane * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[q] = src[q];
            }
            __syncwarp();
            if (lane == 0) {
                for (int i = 0; i < 32 && tb - i >= 1; ++i) {
                    const unsigned w = bt_s[i * 32 + k / SU];
                    const int s = k % SU;
                    int kp;
                    if (s == 0) {
                        const int c = (int)(w >> 8);
                        kp = c < 32 ? c * SU + SU - 1 : (c == 32 ? k : 0);
                    } else {
                        const unsigned c = (w >> (2 * (s - 1))) & 3u;
                        kp = c == 1u ? k - 1 : (c == 0u ? k : 0);
                    }
                    a.path[t0 + tb - i - 1] = kp;
                    k = kp;
                }
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------
// The same recursion for MORE than 32 units (BASELINE configs[2]: 250 units x 4 states): one warp per utterance, U units
// of SU states per lane, everything in registers.  UNI: a loop whose unit starts all see the same weight from a given
// unit end (plan->vlr_ok: the uniform phone loops of mkphoneloopgraph.py; bitwise equal columns of ln A[ends, starts]):
// the first maximum over the P unit ends -- ascending source order, strict > -- is then ONE (value, index) pair per
// frame, found with a lexicographic butterfly, and every start only compares it with its own self loop (which precedes
// the end of unit v in source order iff its unit index <= v).  Learned unit weights (!UNI): see the comment at `vend`.  Back-pointers: one uint16 per unit and frame
// (2 bits per inner state, 10 bits for the unit start), the backtrack stages 32 frames of them in shared memory.
// The generic kernel walks 250 candidates for each of the 250 starts from shared memory: 340 ms per cfg3 batch
// against ~2 ms here (same arithmetic, same tie-breaking: tests/test_kernels_gpu.py compares the paths exactly).
// ---------------------------------------------------------------------------
constexpr int VLM_WARPS = 4;

template <int SU, int U, bool UNI>
__global__ void __launch_bounds__(VLM_WARPS * 32, 3) hmm_viterbi_lrm_kernel(VitArgs a, const float* __restrict__ vlr) {
    constexpr int S = SU * U, PF = 4, ROW = 32 * S, BTLD = 32 * U;
    constexpr int WARP_FLOATS = (PF * ROW * 4 > 32 * BTLD * 2) ? PF * ROW : 32 * BTLD / 2;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ring = smem + (size_t)warp * WARP_FLOATS;                  // [PF][32 * S] llh rows, in state order
    uint16_t* bt_s = reinterpret_cast<uint16_t*>(ring);               // backtrack: [32 frames][32 * U] (after the sweep)
    const int K = a.K, P = K / SU, NONE = P + 1;
    const int gwarp = blockIdx.x * VLM_WARPS + warp, nwarps = gridDim.x * VLM_WARPS;
    // lane l owns the units l, l + 32, ..., l + 32 (U - 1): consecutive lanes read consecutive unit blocks of a row
    // (a blocked assignment -- 32 consecutive floats per lane -- made every shared-memory read a 32-way bank conflict:
    // 718 conflict cycles per frame, profiles/r02_viterbi.md)
    const bool vec_ok = (SU == 4) && (a.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.pl) & 15) == 0);

    // UNI: ln A[end v, any start] of this lane's unit ends.  Otherwise (learned unit weights, phoneloop.py:53-65): the
    // factored weight lv[v] of this lane's unit ends, used only to RANK the ends; the candidates that can still be a
    // first maximum for some start (within the factorisation error of the best one) are then compared exactly, per
    // start, with the dense weights vend[v][u], in ascending order of v.
    const float* vend = vlr + 2 * K + 2 * P;
    float w_self[U][SU], w_prev[U][SU], wend[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int unit = lane + 32 * u;
        const bool own = unit < P;
        wend[u] = own ? __ldg(vlr + 2 * K + (UNI ? 0 : P) + unit) : kNegInf;
#pragma unroll
        for (int s = 0; s < SU; ++s) {
            const int k = unit * SU + s;
            w_self[u][s] = own ? __ldg(vlr + k) : kNegInf;
            w_prev[u][s] = (own && s > 0) ? __ldg(vlr + K + k) : kNegInf;
        }
    }

    auto prefetch = [&](float* slot, const float* row) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int unit = lane + 32 * u;
            if (unit >= P) continue;
            if (vec_ok) {
                cp_async16(slot + unit * SU, row + unit * SU);
            } else {
#pragma unroll
                for (int j = 0; j < SU; ++j) cp_async4(slot + unit * SU + j, row + unit * SU + j);
            }
        }
    };

    for (int utt = gwarp; utt < a.n_utts; utt += nwarps) {
        const int64_t t0 = a.utt_off[utt];
        const int T = (int)(a.utt_off[utt + 1] - t0);
        if (T <= 0) continue;
        const float* pl_u = a.pl + (size_t)t0 * a.ld;
        uint16_t* bt_u = a.bt + (size_t)t0 * BTLD;
        __syncwarp();                                    // the previous utterance's backtrack is done with the ring
        for (int r = 0; r < PF; ++r) {
            if (r < T) prefetch(ring + r * ROW, pl_u + (size_t)r * a.ld);
            cp_async_commit();
        }
        float om[U][SU];
        int slot = 0;
        for (int t = 0; t < T; ++t) {
            cp_async_wait<PF - 1>();
            float p[U][SU];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int unit = lane + 32 * u;
                if constexpr (SU == 4) {
                    const float4 v = (unit < P) ? *reinterpret_cast<const float4*>(ring + slot * ROW + unit * 4)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                    p[u][0] = v.x; p[u][1] = v.y; p[u][2] = v.z; p[u][3] = v.w;
                } else {
#pragma unroll
                    for (int s = 0; s < SU; ++s) p[u][s] = (unit < P) ? ring[slot * ROW + unit * SU + s] : 0.f;
                }
#pragma unroll
                for (int s = 0; s < SU; ++s) p[u][s] = (unit < P) ? a.scale * p[u][s] : kNegInf;
            }
            if (t + PF < T) prefetch(ring + slot * ROW, pl_u + (size_t)(t + PF) * a.ld);
            cp_async_commit();
            slot = (slot + 1 == PF) ? 0 : slot + 1;
            if (t == 0) {
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int s = 0; s < SU; ++s) {
                        const int unit = lane + 32 * u;
                        om[u][s] = (unit < P) ? p[u][s] + __ldg(a.vit.start + unit * SU + s) : kNegInf;
                    }
            } else {
                // first maximum over all unit ends: own units in ascending order, then across lanes (value, index)
                float best = kNegInf;
                int code = NONE;                         // NONE: every candidate is -inf (argmax -> state 0)
                float bests[U];
                int codes[U];
                if constexpr (UNI) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const float c = om[u][SU - 1] + wend[u];
                        if (c > best) {
                            best = c;
                            code = lane + 32 * u;
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oc = __shfl_xor_sync(0xffffffffu, code, o);
                        if (ob > best || (ob == best && oc < code)) {
                            best = ob;
                            code = oc;
                        }
                    }
                } else {
                    float sc[U], m = kNegInf;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        sc[u] = om[u][SU - 1] + wend[u];
                        m = fmaxf(m, sc[u]);
                        bests[u] = kNegInf;
                        codes[u] = NONE;
                    }
                    m = warp_max(m);
                    // |ln A - (lv + lw)| <= 4e-6 (plan), three roundings of values of magnitude <= |m| + |weights|
                    const float thr = m - (1.2e-5f + 1e-6f * fabsf(m));
#pragma unroll
                    for (int uu = 0; uu < U; ++uu) {     // ascending v = src + 32 uu
                        unsigned lanes = __ballot_sync(0xffffffffu, sc[uu] >= thr && sc[uu] > kNegInf);
                        while (lanes != 0u) {
                            const int src = __ffs(lanes) - 1;
                            lanes &= lanes - 1;
                            const float e = __shfl_sync(0xffffffffu, om[uu][SU - 1], src);
                            const int v = src + 32 * uu;
                            const float* row = vend + (size_t)v * P;
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                const int unit = lane + 32 * u;
                                const float c = e + ((unit < P) ? __ldg(row + unit) : kNegInf);
                                if (c > bests[u]) {
                                    bests[u] = c;
                                    codes[u] = v;
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int unit = lane + 32 * u;
                    const bool own = unit < P;
                    // the self arc (source SU * unit) precedes the end of unit v in source order iff unit <= v
                    const float self0 = om[u][0] + w_self[u][0];
                    float b = UNI ? best : bests[u];
                    int c = UNI ? code : codes[u];
                    if (self0 > b || (self0 == b && self0 != kNegInf && unit <= c)) {
                        b = self0;
                        c = P;                           // P: the state itself
                    }
                    unsigned pk = (unsigned)c << 6;
                    float nw[SU];
                    nw[0] = p[u][0] + b;
#pragma unroll
                    for (int s = 1; s < SU; ++s) {
                        const float prev = om[u][s - 1] + w_prev[u][s], self = om[u][s] + w_self[u][s];
                        float bb = prev;
                        unsigned cc = (prev == kNegInf) ? 2u : 1u;     // 1: previous state, 0: itself, 2: all -inf
                        if (self > bb) {
                            bb = self;
                            cc = 0u;
                        }
                        pk |= cc << (2 * (s - 1));
                        nw[s] = p[u][s] + bb;
                    }
#pragma unroll
                    for (int s = 0; s < SU; ++s) om[u][s] = own ? nw[s] : kNegInf;
                    bt_u[(size_t)t * BTLD + unit] = (uint16_t)pk;
                }
            }
            float mxu[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                mxu[u] = om[u][0];
#pragma unroll
                for (int s = 1; s < SU; ++s) mxu[u] = fmaxf(mxu[u], om[u][s]);
            }
#pragma unroll
            for (int w = 1; w < U; w *= 2)
#pragma unroll
                for (int u = 0; u + w < U; u += 2 * w) mxu[u] = fmaxf(mxu[u], mxu[u + w]);
            const float mx = warp_max(mxu[0]);
            const float mxs = (mx == kNegInf) ? 0.f : mx;
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int s = 0; s < SU; ++s) om[u][s] -= mxs;
        }
        cp_async_wait<0>();
        // last state: first maximal index of omega + final
        float best = kNegInf;
        int arg = 0x7fffffff;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int s = 0; s < SU; ++s) {
                const int k = (lane + 32 * u) * SU + s;
                if (k < K) {
                    const float v = om[u][s] + __ldg(a.vit_final + k);
                    if (arg == 0x7fffffff || v > best) { best = v; arg = k; }
                }
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        __threadfence_block();
        __syncwarp();
        // backtrack, 32 frames of back-pointers at a time through shared memory (the llh ring is free now)
        int k = arg;
        if (lane == 0) a.path[t0 + T - 1] = k;
        for (int tb = T - 1; tb >= 1; tb -= 32) {
            // row i of the stage = frame tb - i, copied by the whole warp 16 bytes per lane at a time
            for (int e = lane; e < 32 * (BTLD / 8); e += 32) {
                const int i = e / (BTLD / 8), q = e % (BTLD / 8);
                if (tb - i >= 1)
                    reinterpret_cast<uint4*>(bt_s + i * BTLD)[q] =
                        reinterpret_cast<const uint4*>(bt_u + (size_t)(tb - i) * BTLD)[q];
            }
            __syncwarp();
            if (lane == 0) {
                for (int i = 0; i < 32 && tb - i >= 1; ++i) {
                    const unsigned w = bt_s[i * BTLD + k / SU];
                    const int s = k % SU;
                    int kp;
                    if (s == 0) {
                        const int c = (int)(w >> 6);
                        kp = c < P ? c * SU + SU - 1 : (c == P ? k : 0);
                    } else {
                        const unsigned c = (w >> (2 * (s - 1))) & 3u;
                        kp = c == 1u ? k - 1 : (c == 0u ? k : 0);
                    }
                    a.path[t0 + tb - i - 1] = kp;
                    k = kp;
                }
            }
            __syncwarp();
        }
    }
}

template <int SU, int U, bool UNI>
static int launch_vit_lrm(const VitArgs& a, const float* vlr, int n_utts, cudaStream_t st) {
    constexpr int S = SU * U;
    constexpr int WARP_FLOATS = (4 * 32 * S * 4 > 32 * 32 * U * 2) ? 4 * 32 * S : 32 * 32 * U / 2;
    const size_t smem = sizeof(float) * (size_t)VLM_WARPS * WARP_FLOATS;
    static bool attr_set = false;
    if (!attr_set) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(hmm_viterbi_lrm_kernel<SU, U, UNI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        attr_set = true;
    }
    int blocks = (n_utts + VLM_WARPS - 1) / VLM_WARPS;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    hmm_viterbi_lrm_kernel<SU, U, UNI><<<blocks, VLM_WARPS * 32, smem, st>>>(a, vlr);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

template <int SU>
static int launch_vit_lr(const VitArgs& a, int n_utts, cudaStream_t st) {
    const size_t smem = sizeof(float) * (size_t)FB_WARPS * (6 * 32 * SU + 512);
    int blocks = (n_utts + FB_WARPS - 1) / FB_WARPS;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    hmm_viterbi_lr_kernel<SU><<<blocks, FB_WARPS * 32, smem, st>>>(a);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

template <int S>
static int launch_fb(const FbArgs& a, int n_utts, cudaStream_t st) {
    constexpr int PF = FbCfg<S>::PF;
    int nbuf = (a.K + a.J + 3) & ~3;
    size_t smem = sizeof(float) * (size_t)FB_WARPS * (nbuf + 2 * PF * 32 * S);
    if (smem > 200 * 1024) return BEER_ERR_UNSUPPORTED;
    static bool attr_set = false;
    if (!attr_set) {
        BEER_CUDA_TRY(cudaFuncSetAttribute(hmm_fb_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    int blocks = (n_utts + FB_WARPS - 1) / FB_WARPS;
    int max_blocks = kNumSMs * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    hmm_fb_kernel<S><<<blocks, FB_WARPS * 32, smem, st>>>(a);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

template <int S>
static int launch_vit(const VitArgs& a, int n_utts, cudaStream_t st) {
    int nbuf = (a.K + 3) & ~3;
    size_t smem = sizeof(float) * (size_t)FB_WARPS * nbuf;
    int blocks = (n_utts + FB_WARPS - 1) / FB_WARPS;
    int max_blocks = kNumSMs * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    hmm_viterbi_kernel<S><<<blocks, FB_WARPS * 32, smem, st>>>(a);
    BEER_LAUNCH_CHECK();
    return BEER_OK;
}

static int pick_S(int K) {
    int S = 1;
    while (S * 32 < K) S *= 2;
    return S;
}

}  // namespace beer

using namespace beer;

extern "C" {

int beer_graph_plan_create(const float* init_log, const float* final_log, const float* trans_log,
                           const int32_t* pdf_map, int K, int Kp, int factorize, beer_graph_plan** plan_out) {
    if (!init_log || !final_log || !trans_log || !pdf_map || !plan_out || K <= 0 || Kp <= 0) return BEER_ERR_ARG;
    if (K > 1024) return BEER_ERR_UNSUPPORTED;
    for (int k = 0; k < K; ++k)
        if (pdf_map[k] < 0 || pdf_map[k] >= Kp) return BEER_ERR_ARG;
    const int S = pick_S(K);

    // dense -> sparse rows
    std::vector<std::vector<Arc>> out_arcs(K);  // out_arcs[i] = (j, lw)
    int nnz = 0;
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) {
            float w = trans_log[(size_t)i * K + j];
            if (w > -INFINITY && w == w) {
                out_arcs[i].push_back({j, (double)w});
                ++nnz;
            }
        }

    // rank-1 block detection
    struct Junction { std::vector<int> rows; std::vector<double> lv; std::vector<int> cols; std::vector<double> lw; };
    std::vector<Junction> junctions;
    std::vector<char> row_factored(K, 0);
    if (factorize) {
        std::map<std::vector<int>, std::vector<int>> groups;
        for (int i = 0; i < K; ++i) {
            std::vector<int> key;
            for (const Arc& a : out_arcs[i])
                if (a.src != i) key.push_back(a.src);
            if (key.size() >= 2) groups[key].push_back(i);
        }
        for (auto& kv : groups) {
            const std::vector<int>& cols = kv.first;
            const std::vector<int>& rows = kv.second;
            if (rows.size() < 2) continue;
            auto row_vals = [&](int i) {
                std::vector<double> v;
                for (const Arc& a : out_arcs[i])
                    if (a.src != i) v.push_back(a.lw);
                return v;
            };
            std::vector<double> base = row_vals(rows[0]);
            double mx = base[0];
            for (double b : base) mx = std::max(mx, b);
            double se = 0.0;
            for (double b : base) se += exp(b - mx);
            double lse = mx + log(se);
            Junction jn;
            jn.cols = cols;
            for (double b : base) jn.lw.push_back(b - lse);
            for (int i : rows) {
                std::vector<double> v = row_vals(i);
                double mean = 0.0;
                for (size_t c = 0; c < v.size(); ++c) mean += v[c] - jn.lw[c];
                mean /= (double)v.size();
                double dev = 0.0;
                for (size_t c = 0; c < v.size(); ++c) dev = std::max(dev, fabs(v[c] - jn.lw[c] - mean));
                if (dev <= 4e-6) {
                    jn.rows.push_back(i);
                    jn.lv.push_back(mean);
                }
            }
            if (jn.rows.size() >= 2 && jn.rows.size() * cols.size() > jn.rows.size() + cols.size()) {
                for (int i : jn.rows) row_factored[i] = 1;
                junctions.push_back(std::move(jn));
            }
        }
    }
    const int J = (int)junctions.size();
    if (K + J > 8192) return BEER_ERR_UNSUPPORTED;

    // forward / backward / Viterbi lists
    std::vector<std::vector<Arc>> f_state(K), b_state(K), v_state(K), f_jn(J), b_jn(J);
    int n_direct = 0, n_jin = 0, n_jout = 0;
    for (int i = 0; i < K; ++i)
        for (const Arc& a : out_arcs[i]) {
            v_state[a.src].push_back({i, a.lw});
            if (row_factored[i] && a.src != i) continue;
            f_state[a.src].push_back({i, a.lw});
            b_state[i].push_back({a.src, a.lw});
            ++n_direct;
        }
    for (int n = 0; n < J; ++n) {
        const Junction& jn = junctions[n];
        for (size_t r = 0; r < jn.rows.size(); ++r) {
            f_jn[n].push_back({jn.rows[r], jn.lv[r]});
            b_state[jn.rows[r]].push_back({K + n, jn.lv[r]});
            ++n_jin;
        }
        for (size_t c = 0; c < jn.cols.size(); ++c) {
            f_state[jn.cols[c]].push_back({K + n, jn.lw[c]});
            b_jn[n].push_back({jn.cols[c], jn.lw[c]});
            ++n_jout;
        }
    }
    auto by_src = [](const Arc& x, const Arc& y) { return x.src < y.src; };
    for (int k = 0; k < K; ++k) {
        std::sort(f_state[k].begin(), f_state[k].end(), by_src);
        std::sort(b_state[k].begin(), b_state[k].end(), by_src);
        std::sort(v_state[k].begin(), v_state[k].end(), by_src);
    }

    const double L2E = 1.4426950408889634;
    HostLists hf, hb, hv;
    build_lists(K, S, f_state, f_jn, L2E, hf);
    build_lists(K, S, b_state, b_jn, L2E, hb);
    build_lists(K, S, v_state, {}, 1.0, hv);
    hf.start.resize(K);
    hb.start.resize(K);
    hv.start.resize(K);
    std::vector<float> vfinal(K);
    for (int k = 0; k < K; ++k) {
        hf.start[k] = (float)((double)init_log[k] * L2E);
        hb.start[k] = (float)((double)final_log[k] * L2E);
        hv.start[k] = init_log[k];
        vfinal[k] = final_log[k];
    }

    // aligned left-to-right loop?  (unit starts = junction destinations 0, SU, 2 SU, ...)
    int lr_su = 0, lr_u = 0;
    std::vector<float> lr_w;
    if (J == 1 && junctions[0].cols.size() >= 2) {
        const Junction& jn = junctions[0];
        std::vector<int> cols = jn.cols;   // sorted (built from sorted keys)
        const int su = cols[1] - cols[0];
        bool ok = su >= 1 && su <= 8 && K % su == 0 && (int)cols.size() == K / su;
        for (size_t c = 0; ok && c < cols.size(); ++c) ok = cols[c] == (int)c * su;
        for (int i = 0; ok && i < K; ++i) {
            if (row_factored[i] && i % su != su - 1) ok = false;
            for (const Arc& a : out_arcs[i]) {
                if (a.src == i) continue;                                  // self loop
                if (row_factored[i]) continue;                             // through the junction
                if (!(a.src == i + 1 && (i + 1) % su != 0)) ok = false;    // only "next state of my unit"
            }
        }
        const int P = ok ? K / su : 0;
        int u = (P + 31) / 32;
        u = (u <= 1) ? 1 : (u <= 2 ? 2 : (u <= 4 ? 4 : (u <= 8 ? 8 : 0)));   // 8: block kernel, 8 warps x 1 unit per lane
        if (ok && u > 0 && (su == 3 || su == 4)) {
            lr_su = su; lr_u = u;
            const int row = 32 * su * u;
            lr_w.assign((size_t)3 * row, -INFINITY);
            for (int i = 0; i < K; ++i)
                for (const Arc& a : out_arcs[i]) {
                    if (a.src == i) lr_w[i] = (float)(a.lw * L2E);
                    else if (!row_factored[i]) lr_w[row + a.src] = (float)(a.lw * L2E);
                }
            for (size_t c = 0; c < jn.cols.size(); ++c) lr_w[row + jn.cols[c]] = (float)(jn.lw[c] * L2E);
            for (size_t r = 0; r < jn.rows.size(); ++r) lr_w[2 * row + jn.rows[r]] = (float)(jn.lv[r] * L2E);
        }
    }

    std::vector<float> vlr;
    int vlr_ok = 0;
    if (lr_su) {
        const int su = lr_su, P = K / su;
        auto val = [&](int i, int j) {
            const float w = trans_log[(size_t)i * K + j];
            return (w > -INFINITY && w == w) ? w : -INFINITY;
        };
        vlr.assign((size_t)2 * K + 2 * P + (size_t)P * P, -INFINITY);
        for (int k = 0; k < K; ++k) {
            vlr[k] = val(k, k);
            if (k % su) vlr[K + k] = val(k - 1, k);
        }
        vlr_ok = 1;
        for (int v = 0; v < P; ++v) {
            const int e = v * su + su - 1;
            const float w0 = val(e, 0);
            for (int u = 0; u < P; ++u) {
                const float wu = val(e, u * su);
                if (memcmp(&wu, &w0, sizeof(float)) != 0) vlr_ok = 0;
                vlr[(size_t)2 * K + 2 * P + (size_t)v * P + u] = wu;       // dense ln A[end v, start u]
            }
            vlr[(size_t)2 * K + v] = w0;
        }
        // the factored weight of every unit end (ln A[end v, start u] = lv[v] + lw[u] to 4e-6): ranks the ends
        const Junction& jn = junctions[0];
        for (size_t r = 0; r < jn.rows.size(); ++r) vlr[(size_t)2 * K + P + jn.rows[r] / su] = (float)jn.lv[r];
    }

    BlobWriter w;
    ListOffsets of = write_lists(w, hf), ob = write_lists(w, hb), ov = write_lists(w, hv);
    size_t o_lrw = w.add(lr_w.data(), lr_w.size() * 4);
    size_t o_vlr = w.add(vlr.data(), vlr.size() * 4);
    size_t o_map = w.add(pdf_map, (size_t)K * 4);
    size_t o_vfinal = w.add(vfinal.data(), (size_t)K * 4);

    beer_graph_plan* p = new (std::nothrow) beer_graph_plan();
    if (!p) return BEER_ERR_ALLOC;
    cudaError_t e = cudaMalloc(&p->dev_blob, w.bytes.size());
    if (e != cudaSuccess) { delete p; return (int)e; }
    e = cudaMemcpy(p->dev_blob, w.bytes.data(), w.bytes.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(p->dev_blob); delete p; return (int)e; }
    const char* base = (const char*)p->dev_blob;
    p->K = K; p->Kp = Kp; p->J = J; p->S = S;
    p->n_direct = n_direct; p->n_jin = n_jin; p->n_jout = n_jout; p->dense_nnz = nnz;
    p->fwd = bind_lists(base, of);
    p->bwd = bind_lists(base, ob);
    p->vit = bind_lists(base, ov);
    p->map = (const int*)(base + o_map);
    p->lr_su = lr_su; p->lr_u = lr_u;
    p->lr_w = lr_su ? (const float*)(base + o_lrw) : nullptr;
    p->vlr = lr_su ? (const float*)(base + o_vlr) : nullptr;
    p->vlr_ok = vlr_ok;
    p->vit_final = (const float*)(base + o_vfinal);
    p->map_identity = 1;
    for (int k = 0; k < K; ++k)
        if (pdf_map[k] != k) p->map_identity = 0;
    {
        int max_cnt = 0, jr = 0;
        for (int v : hf.st_cnt) max_cnt = std::max(max_cnt, v);
        for (int v : hb.st_cnt) max_cnt = std::max(max_cnt, v);
        if (J > 0) jr = std::max(hf.jn_cnt[0], hb.jn_cnt[0]);
        p->jrows = jr;
        p->fast_ok = (max_cnt <= 2 && J <= 1 && jr <= 2 && (S == 4 || S == 8) && K % 4 == 0) ? 1 : 0;
    }
    *plan_out = p;
    return BEER_OK;
}

void beer_graph_plan_destroy(beer_graph_plan* plan) {
    if (!plan) return;
    if (plan->dev_blob) cudaFree(plan->dev_blob);
    delete plan;
}

int beer_graph_plan_info(const beer_graph_plan* p, int32_t* info) {
    if (!p || !info) return BEER_ERR_ARG;
    info[0] = p->K; info[1] = p->J; info[2] = p->n_direct; info[3] = p->n_jin;
    info[4] = p->n_jout; info[5] = p->S; info[6] = p->map_identity; info[7] = p->dense_nnz;
    return BEER_OK;
}

static int fb_row_stride(const beer_graph_plan* p) { return (p->K + 3) & ~3; }

int64_t beer_hmm_workspace_bytes(const beer_graph_plan* plan, int64_t N) {
    if (!plan || N < 0) return BEER_ERR_ARG;
    return (N * fb_row_stride(plan) + 64) * (int64_t)sizeof(float);
}

// Units whose counts the forward-backward kernels can reduce themselves: only the one-warp left-to-right loop
// kernels (<= 128 units of 3 or 4 states, <= 256 units of 4 states) with an identity pdf map carry the fused counts.  Everything else
// reports 0 and the caller reduces the ends x starts block of beer_hmm_transition_posteriors instead.
int beer_hmm_unit_count_size(const beer_graph_plan* plan) {
    if (!plan) return BEER_ERR_ARG;
    if (!plan->lr_su || !plan->map_identity) return 0;
    if (plan->lr_u > 4 && !(plan->lr_u == 8 && plan->lr_su == 4 && plan->K % 4 == 0)) return 0;
    return plan->K / plan->lr_su;
}

// 1 when beer_hmm_forward_backward_ex can write log2 posteriors for this graph (left-to-right loop kernels)
int beer_hmm_lpost_supported(const beer_graph_plan* plan) {
    if (!plan) return 0;
    return (plan->lr_su && plan->map_identity && plan->lr_u > 0) ? 1 : 0;
}

int beer_hmm_forward_backward(const beer_graph_plan* plan, const float* pdf_llh, int64_t ld_pdf,
                              const float* frame_ref, const int64_t* utt_off, int n_utts, float scale,
                              float* state_post, float* pdf_post, int64_t ld_post, float* frame_exp_llh,
                              double* utt_exp_llh, double* utt_logz, void* workspace, void* stream) {
    return beer_hmm_forward_backward_units(plan, pdf_llh, ld_pdf, frame_ref, utt_off, n_utts, scale, state_post,
                                           pdf_post, ld_post, frame_exp_llh, utt_exp_llh, utt_logz, nullptr,
                                           workspace, stream);
}

int beer_hmm_forward_backward_units(const beer_graph_plan* plan, const float* pdf_llh, int64_t ld_pdf,
                                    const float* frame_ref, const int64_t* utt_off, int n_utts, float scale,
                                    float* state_post, float* pdf_post, int64_t ld_post, float* frame_exp_llh,
                                    double* utt_exp_llh, double* utt_logz, double* unit_counts, void* workspace,
                                    void* stream) {
    return beer_hmm_forward_backward_ex(plan, pdf_llh, ld_pdf, frame_ref, utt_off, n_utts, scale, state_post, pdf_post,
                                        ld_post, frame_exp_llh, utt_exp_llh, utt_logz, unit_counts, 0, nullptr, 0, workspace, stream);
}

int beer_hmm_block_activity_supported(const beer_graph_plan* plan, int with_unit_counts) {
    if (!plan || !beer_hmm_lpost_supported(plan)) return 0;
    // the one-warp kernel for more than 128 units (the only one that reduces unit counts there) does not mark blocks
    if (plan->lr_u == 8) return with_unit_counts ? 0 : 1;
    return (plan->lr_u == 1 || plan->lr_u == 2 || plan->lr_u == 4) ? 1 : 0;
}

int beer_hmm_forward_backward_ex(const beer_graph_plan* plan, const float* pdf_llh, int64_t ld_pdf,
                                 const float* frame_ref, const int64_t* utt_off, int n_utts, float scale,
                                 float* state_post, float* pdf_post, int64_t ld_post, float* frame_exp_llh,
                                 double* utt_exp_llh, double* utt_logz, double* unit_counts, int flags,
                                 float* pdf_lpost, int64_t ld_lpost, void* workspace, void* stream) {
    return beer_hmm_forward_backward_blocks(plan, pdf_llh, ld_pdf, frame_ref, utt_off, n_utts, scale, state_post, pdf_post,
                                            ld_post, frame_exp_llh, utt_exp_llh, utt_logz, unit_counts, flags, pdf_lpost,
                                            ld_lpost, nullptr, 0, 0, workspace, stream);
}

int beer_hmm_forward_backward_blocks(const beer_graph_plan* plan, const float* pdf_llh, int64_t ld_pdf,
                                     const float* frame_ref, const int64_t* utt_off, int n_utts, float scale,
                                     float* state_post, float* pdf_post, int64_t ld_post, float* frame_exp_llh,
                                     double* utt_exp_llh, double* utt_logz, double* unit_counts, int flags,
                                     float* pdf_lpost, int64_t ld_lpost, uint8_t* block_active, int64_t ld_active,
                                     int pdfs_per_block, void* workspace, void* stream) {
    if (!plan || !pdf_llh || !utt_off || !utt_exp_llh || !workspace || n_utts < 0) return BEER_ERR_ARG;
    if (block_active != nullptr) {
        if (pdf_lpost == nullptr || pdfs_per_block <= 0 || ld_active < (plan->Kp + pdfs_per_block - 1) / pdfs_per_block)
            return BEER_ERR_ARG;
        if (!beer_hmm_block_activity_supported(plan, unit_counts != nullptr)) return BEER_ERR_UNSUPPORTED;
    }
    const bool llh_log2 = (flags & BEER_FB_LLH_LOG2) != 0;
    if (ld_pdf < plan->Kp || (pdf_post && ld_post < plan->Kp)) return BEER_ERR_ARG;
    if (n_utts == 0) return BEER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    FbArgs a;
    a.fwd = plan->fwd; a.bwd = plan->bwd;
    a.K = plan->K; a.J = plan->J; a.Kp = plan->Kp;
    a.map = plan->map; a.map_identity = plan->map_identity;
    a.pl = pdf_llh; a.ld = ld_pdf; a.frame_ref = frame_ref; a.utt_off = utt_off; a.n_utts = n_utts;
    a.scale = scale;
    a.la_ws = (float*)workspace; a.Kw = fb_row_stride(plan);
    a.state_post = state_post; a.pdf_post = pdf_post; a.ld_post = ld_post;
    a.frame_exp_llh = frame_exp_llh; a.utt_exp_llh = utt_exp_llh; a.utt_logz = utt_logz;
    a.vec = (plan->map_identity && plan->S % 4 == 0 && plan->K % 4 == 0 && ld_pdf % 4 == 0 &&
             ((uintptr_t)pdf_llh & 15) == 0) ? 1 : 0;
    a.lr_w = plan->lr_w;
    a.lr_row = 32 * plan->lr_su * plan->lr_u;
    a.unit_counts = unit_counts;
    a.llh_mul = llh_log2 ? 1.f : kLog2e;
    a.pdf_lpost = pdf_lpost; a.ld_lpost = ld_lpost;
    a.lpost_rel = (flags & BEER_FB_LPOST_RELATIVE) ? 1 : 0;
    a.blk_active = block_active; a.blk_ld = ld_active; a.blk_ppb = pdfs_per_block;
    // a weight below 2^-25 of the statistics kernel's fp16 range (it carries w 2^wexp, wexp as in beer_mix16_accumulate) is
    // zero in both halves of its operand; two more bits of margin for the roundings on the way
    a.blk_thr = -(25.f + (float)beer_mix16_weight_exponent(scale) + 2.f);
    if (pdf_lpost != nullptr && (ld_lpost < plan->Kp || ld_lpost % 4 != 0 || ((uintptr_t)pdf_lpost & 15) != 0)) return BEER_ERR_ARG;
    if (unit_counts != nullptr && beer_hmm_unit_count_size(plan) <= 0) return BEER_ERR_UNSUPPORTED;
    const bool post_ok = (pdf_post == nullptr || (ld_post % 4 == 0 && ((uintptr_t)pdf_post & 15) == 0)) &&
                         (state_post == nullptr || (plan->K % 4 == 0 && ((uintptr_t)state_post & 15) == 0));
    const char* force = getenv("BEER_B200_SCAN");   // debug: "generic" | "fast" | unset (best available)
    if (unit_counts != nullptr || pdf_lpost != nullptr) force = nullptr;    // both live in the loop kernels only
    const bool lr_vec = (plan->lr_su * plan->lr_u) % 4 == 0;   // the kernel moves whole float4 rows
    // (its float4 rows need 16-byte aligned llh / posterior rows; a.vec is about the generic kernels' lane layout)
    const bool lr_rows_ok = ld_pdf % 4 == 0 && ((uintptr_t)pdf_llh & 15) == 0 && post_ok;
    if (plan->lr_su && plan->map_identity && (force == nullptr || force[0] == 'l') && (!lr_vec || lr_rows_ok)) {
        const int u = plan->lr_u;
        // 129 .. 256 units.  Measured on BASELINE configs[2] (1250 utterances x 1000 frames, 250 units x 4 states): eight
        // warps x one unit per lane 7.3 ms, four warps x two units 7.5 ms, ONE warp x eight units 12.0 ms (a frame is then
        // one long dependent instruction stream per warp at 12 warps per SM).  The one-warp kernel is the only one of
        // the three that reduces unit counts, so it runs when they are asked for.
        const char* lrc = getenv("BEER_B200_SCAN_LRC");       // debug: "1" = four warps x two units, "w" = one warp
        if (a.blk_active != nullptr) lrc = nullptr;            // (those two variants do not write the activity map)
        if (u == 8 && plan->lr_su == 4 && plan->K % 4 == 0 && (unit_counts != nullptr || (lrc != nullptr && lrc[0] == 'w')))
            return launch_fb_lrw<4, 8>(a, n_utts, st);
        if (u == 8 && unit_counts == nullptr) {
            if (plan->lr_su == 4 && plan->K <= 4 * 32 * 2 * 4 && plan->K % 4 == 0 && lrc != nullptr && lrc[0] == '1')
                return launch_fb_lrc<4, 4, 2>(a, n_utts, st);
            if (plan->lr_su == 4) return launch_fb_lrb<4, 8>(a, n_utts, st);
            if (plan->lr_su == 3) return launch_fb_lrb<3, 8>(a, n_utts, st);
        }
        if (plan->lr_su == 4) {
            if (u == 1) return launch_fb_lr<4, 1>(a, n_utts, st);
            if (u == 2) return launch_fb_lr<4, 2>(a, n_utts, st);
            if (u == 4) return launch_fb_lr<4, 4>(a, n_utts, st);
        } else if (plan->lr_su == 3) {
            if (u == 1) return launch_fb_lr<3, 1>(a, n_utts, st);
            if (u == 2) return launch_fb_lr<3, 2>(a, n_utts, st);
            if (u == 4) return launch_fb_lr<3, 4>(a, n_utts, st);
        }
    }
    // only the left-to-right loop kernels accumulate unit counts / write log2 posteriors: never drop them silently
    if (unit_counts != nullptr || pdf_lpost != nullptr) return BEER_ERR_UNSUPPORTED;
    if (force != nullptr && force[0] == 'g') goto generic;
    if (plan->fast_ok && a.vec && (state_post == nullptr || plan->K % 4 == 0) &&
        (pdf_post == nullptr || (ld_post % 4 == 0 && ((uintptr_t)pdf_post & 15) == 0)) &&
        (state_post == nullptr || ((uintptr_t)state_post & 15) == 0)) {
        const int jr = plan->jrows <= 1 ? 1 : 2;
        if (plan->S == 4) return jr == 1 ? launch_fb_fast<4, 1>(a, n_utts, st) : launch_fb_fast<4, 2>(a, n_utts, st);
        if (plan->S == 8) return jr == 1 ? launch_fb_fast<8, 1>(a, n_utts, st) : launch_fb_fast<8, 2>(a, n_utts, st);
    }
generic:
    switch (plan->S) {
        case 1: return launch_fb<1>(a, n_utts, st);
        case 2: return launch_fb<2>(a, n_utts, st);
        case 4: return launch_fb<4>(a, n_utts, st);
        case 8: return launch_fb<8>(a, n_utts, st);
        case 16: return launch_fb<16>(a, n_utts, st);
        case 32: return launch_fb<32>(a, n_utts, st);
    }
    return BEER_ERR_UNSUPPORTED;
}

int beer_hmm_viterbi(const beer_graph_plan* plan, const float* pdf_llh, int64_t ld_pdf, const int64_t* utt_off,
                     int n_utts, float scale, int32_t* path, void* workspace, void* stream) {
    if (!plan || !pdf_llh || !utt_off || !path || !workspace || n_utts < 0) return BEER_ERR_ARG;
    if (ld_pdf < plan->Kp) return BEER_ERR_ARG;
    if (n_utts == 0) return BEER_OK;
    VitArgs a;
    a.vit = plan->vit; a.vit_final = plan->vit_final; a.K = plan->K;
    a.map = plan->map; a.map_identity = plan->map_identity;
    a.pl = pdf_llh; a.ld = ld_pdf; a.utt_off = utt_off; a.n_utts = n_utts; a.scale = scale;
    a.bt = (uint16_t*)workspace; a.path = path;
    cudaStream_t st = (cudaStream_t)stream;
    // aligned left-to-right loop, one unit per lane: register-resident kernel (its back-pointers need 64 bytes per
    // frame of the N * K * 2 byte workspace, hence K >= 32)
    {
        const char* force = getenv("BEER_B200_SCAN");
        if (plan->lr_su && plan->lr_u == 1 && plan->map_identity && plan->S == plan->lr_su && plan->K >= 32 &&
            plan->K == plan->lr_su * (plan->K / plan->lr_su) && plan->K / plan->lr_su <= 32 &&
            ((uintptr_t)workspace & 15) == 0 && (force == nullptr || force[0] == 'l')) {
            if (plan->lr_su == 4 && ld_pdf % 4 == 0 && ((uintptr_t)pdf_llh & 15) == 0)
                return launch_vit_lr<4>(a, n_utts, st);
            if (plan->lr_su == 3) return launch_vit_lr<3>(a, n_utts, st);
        }
    }
    {
        // more than 32 units of a uniform loop: U units per lane (back-pointers: 32 U uint16 per frame <= K of them)
        const char* force = getenv("BEER_B200_SCAN");
        const int su = plan->lr_su, P = su ? plan->K / su : 0;
        if (su && plan->map_identity && P > 32 && plan->K == su * P && ((uintptr_t)workspace & 15) == 0 &&
            (force == nullptr || force[0] == 'l')) {
            const int u = (P + 31) / 32;
            const bool uni = plan->vlr_ok != 0 && getenv("BEER_B200_VIT_DENSE") == nullptr;
#define BEER_VLM_CASE(su_, u_)                                                                   \
    if (su == su_ && u <= u_)                                                                    \
        return uni ? launch_vit_lrm<su_, u_, true>(a, plan->vlr, n_utts, st)                     \
                   : launch_vit_lrm<su_, u_, false>(a, plan->vlr, n_utts, st);
            BEER_VLM_CASE(4, 2) BEER_VLM_CASE(4, 4) BEER_VLM_CASE(4, 8)
            BEER_VLM_CASE(3, 2) BEER_VLM_CASE(3, 4) BEER_VLM_CASE(3, 8)
#undef BEER_VLM_CASE
        }
    }
    switch (plan->S) {
        case 1: return launch_vit<1>(a, n_utts, st);
        case 2: return launch_vit<2>(a, n_utts, st);
        case 4: return launch_vit<4>(a, n_utts, st);
        case 8: return launch_vit<8>(a, n_utts, st);
        case 16: return launch_vit<16>(a, n_utts, st);
        case 32: return launch_vit<32>(a, n_utts, st);
    }
    return BEER_ERR_UNSUPPORTED;
}

}  // extern "C"
