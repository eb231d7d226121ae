/* beer_b200 -- C ABI of the B200-native VB-EM hot path (libbeer_b200.so).
 *
 * Drop-in boundary for ONE path of beer-asr/beer (reference @ d53d2a1): what
 * `beer.evidence_lower_bound(model, X, ...)` runs on HMM / GMM models
 * (beer/inference/objectives.py:119-190).  The reference has no FFI; each entry
 * point below names the reference Python op site it replaces (file:line,
 * relative to the reference checkout).  INTEGRATION.md shows the ctypes stub a
 * maintainer would add on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - the caller owns every buffer, kernels never allocate (the only objects the
 *     library owns are graph plans, created/destroyed explicitly);
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*,
 *     NULL = legacy default stream), performs no device synchronisation and is
 *     re-entrant across streams;
 *   - return value: 0 = ok, > 0 = cudaError_t, < 0 = BEER_ERR_* below;
 *   - frames of all utterances are concatenated ("ragged batch"): X is [N, D]
 *     row-major fp32, utt_off[n_utts + 1] (int64) holds the first frame of each
 *     utterance and N at the end;
 *   - float data is fp32, accumulated statistics and ELBO terms are fp64.
 */
#ifndef BEER_B200_H_
#define BEER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BEER_API __attribute__((visibility("default")))
#else
#define BEER_API
#endif

#define BEER_ERR_ARG (-1)
#define BEER_ERR_UNSUPPORTED (-2)
#define BEER_ERR_ALLOC (-3)

/* Library / ABI version (major * 100 + minor). */
BEER_API int beer_b200_version(void);

/* ------------------------------------------------------------------------
 * Conjugate exponential-family parameter math (beer/dists)
 * ---------------------------------------------------------------------- */

/* E_q[T(theta)] of M Normal-Gamma posteriors, [M, 2D+2] =
 * [a/b*m, a/b, D/k + sum a/b*m^2, sum psi(a) - ln b].
 * Replaces NormalGamma.expected_sufficient_statistics (beer/dists/normalgamma.py:118-146).
 * mean [M,D], scale [M], shape [M], rates [M,D]. */
BEER_API int beer_normalgamma_expected_stats(const float* mean, const float* scale, const float* shape,
                                    const float* rates, int M, int D, float* ets, void* stream);

/* E[ln pi] of K Dirichlet posteriors with C categories, [K, C] =
 * psi(a_kc) - psi(sum_c a_kc).  Replaces the eye(C) evaluation of
 * MixtureSet._log_weights / Mixture._log_weights (beer/models/mixtureset.py:64-67,
 * mixture.py:45-48) through Dirichlet.expected_sufficient_statistics
 * (beer/dists/dirichlet.py:106-128). */
BEER_API int beer_dirichlet_expected_logw(const float* conc, int K, int C, float* logw, void* stream);

/* Emission weights for the per-frame expected log-likelihood kernel.
 * From the Normal-Gamma posteriors of M diagonal Gaussians (and optional
 * per-Gaussian mixture log-weights logw[M], NULL for none) builds
 *   W[M, 2D]  = [a/b*m | a/b - lam_ref]            (fp32)
 *   bias[M]   = -1/2 (D/k + sum a/b m^2) + 1/2 sum(psi(a) - ln b)
 *               - D/2 ln 2pi + logw - bias_ref      (fp32, computed in fp64)
 *   ref[D+1]  = [lam_ref (D) | bias_ref]            (mean over the M Gaussians)
 * so that  llh_tj = [x_t, -x_t^2/2] . W_j + bias_j + r_t  with the per-frame
 * constant  r_t = -1/2 sum_d lam_ref_d x_td^2 + bias_ref.  Splitting r_t off keeps
 * the per-state values small so that fp32 resolves posteriors to ~1e-6.
 * Same contraction as NormalDiagonalLikelihood.__call__
 * (beer/dists/normalgamma.py:55-59) on the statistics of normalgamma.py:19-27. */
BEER_API int beer_emission_prepare(const float* mean, const float* scale, const float* shape,
                          const float* rates, const float* logw, int M, int D, float* W,
                          float* bias, float* ref, void* stream);

/* ------------------------------------------------------------------------
 * E-step kernels
 * ---------------------------------------------------------------------- */

/* KA: per-frame per-pdf expected log-likelihood of a NormalSet / MixtureSet.
 * Replaces NormalSet.expected_log_likelihood (beer/models/normalset.py:117-119)
 * and MixtureSet.expected_log_likelihood (beer/models/mixtureset.py:85-98).
 *   X [N,D]; W, bias, ref from beer_emission_prepare;
 *   comp_off[Kp+1] (int32): pdf k owns Gaussians comp_off[k]..comp_off[k+1]-1
 *     (NULL: one Gaussian per pdf, Kp == M);
 *   pdf_llh [N, ld_pdf] <- logsumexp_c(llh_t,kc) - r_t     (offset form)
 *   comp_llh [N, M] or NULL <- llh_tj - r_t                (kept for responsibilities)
 *   frame_ref [N] <- r_t. */
BEER_API int beer_emission_llh(const float* X, int64_t N, int D, const float* W, const float* bias,
                      const float* ref, int M, const int32_t* comp_off, int Kp, float* pdf_llh,
                      int64_t ld_pdf, float* comp_llh, float* frame_ref, void* stream);

/* KA on the tensor cores (tcgen05 + TMEM, 3xTF32 split, fp32 accumulate): same result
 * contract as beer_emission_llh for models with a uniform number C of Gaussians per pdf
 * (C in {1,2,4,8,16}) and D in {20, 40}.  The weights are packed once per VB iteration:
 *   beer_emission_tc_supported(M, D, C)      -> 1 if this shape has a tensor-core path
 *   beer_emission_tc_image_floats(M, D, C)   -> floats of the packed image buffer
 *   beer_emission_tc_pack(W, bias, ...)      -> image  (W, bias from beer_emission_prepare)
 *   beer_emission_llh_tc(X, ..., image, ref) -> pdf_llh / comp_llh / frame_ref as above. */
BEER_API int beer_emission_tc_supported(int M, int D, int C);
BEER_API int64_t beer_emission_tc_image_floats(int M, int D, int C);
BEER_API int beer_emission_tc_pack(const float* W, const float* bias, int M, int D, int C, float* image,
                                   void* stream);
BEER_API int beer_emission_llh_tc(const float* X, int64_t N, int D, const float* image, const float* ref, int M,
                                  int C, float* pdf_llh, int64_t ld_pdf, float* comp_llh, float* frame_ref,
                                  void* stream);

/* Statistics of a mixture model along a state path (Viterbi training: hmm.py:42-58 with viterbi=True gives one-hot pdf
 * posteriors; the responsibilities inside the chosen pdf are those of mixtureset.py:100-112), sparse: per frame only the
 * C Gaussians of pdf_ids[t] are evaluated (from W / bias of beer_emission_prepare, fp32 SIMT) and accumulated:
 *   acc_normal [M, 2D+2] (fp64) += scale r_tc T(x_t);  frame_exp_llh [N] (optional) = scale (log sum_c exp z_tc + frame_ref[t]).
 * C in {1, 2, 4, 8, 16}, D <= 64. */
BEER_API int beer_path_accumulate_mix(const float* X, int64_t N, int D, const int32_t* pdf_ids, const float* W,
                                      const float* bias, int M, int C, const float* frame_ref, float scale,
                                      double* acc_normal, float* frame_exp_llh, void* stream);

/* KA backward: gradient of sum_t grad_out[t] * sum_k pdf_post[t,k] llh_k(x_t) w.r.t. the frames, posteriors held fixed
 * (beer/models/hmm.py:79-87 and mixtureset.py:85-98 run the inference on detached llhs; beer/models/vae.py:63-89 is the
 * caller that back-propagates through prior.expected_log_likelihood):
 *   grad_X[t] = grad_out[t] * ( sum_j w_tj E[lambda_j mu_j] - x_t o sum_j w_tj E[lambda_j] ),
 *   w_tj = pdf_post[t, pdf(j)] * exp(comp_llh[t,j] - pdf_llh[t, pdf(j)])        (mixtures; comp_llh = NULL: w = pdf_post)
 * One tcgen05 kernel (fp16 3-pass split, w as the tensor-memory A operand): w [N, M] is never stored.
 *   beer_emission_bwd_pack(exp_stats [M, ld] = E[T(theta)] of beer_normalgamma_expected_stats, ...) -> image, inv_scale[2D]
 *     (image: beer_emission_bwd_image_bytes(M, D) bytes; colmax_scratch: 2D uint32)
 *   pdf_of [M] = pdf id of every Gaussian (mixtures only); grad_out [N] or NULL (= ones); scale = the factor pdf_post carries. */
BEER_API int beer_emission_bwd_supported(int M, int D);
BEER_API int64_t beer_emission_bwd_image_bytes(int M, int D);
BEER_API int beer_emission_bwd_pack(const float* exp_stats, int M, int D, int64_t ld, void* image, float* inv_scale,
                                    uint32_t* colmax_scratch, void* stream);
BEER_API int beer_emission_llh_bwd(const float* X, int64_t N, int D, const void* image, const float* inv_scale, int M,
                                   const float* pdf_post, int64_t ld_post, const float* comp_llh, int64_t ld_comp,
                                   const float* pdf_llh, int64_t ld_pdf, const int* pdf_of, const float* grad_out,
                                   float scale, float* grad_X, void* stream);

/* Graph plan: device-resident sparse form of a CompiledGraph
 * (beer/graph.py:243-268: init_log_probs[K], final_log_probs[K], dense
 * trans_log_probs[K,K], pdf_id_mapping[K]).  Host pointers in, opaque handle out.
 * Arcs with log-probability -inf are dropped; with `factorize` != 0, groups of
 * rows that share one off-diagonal support set with proportional weights (the
 * unit-end -> unit-start block of a phone loop, beer/models/phoneloop.py:53-65)
 * are routed through one non-emitting junction node, as they were in the
 * uncompiled Graph (beer/graph.py:185-240). */
typedef struct beer_graph_plan beer_graph_plan;
BEER_API int beer_graph_plan_create(const float* init_log_host, const float* final_log_host,
                           const float* trans_log_host, const int32_t* pdf_map_host, int K, int Kp,
                           int factorize, beer_graph_plan** plan_out);
BEER_API void beer_graph_plan_destroy(beer_graph_plan* plan);
/* info[0]=K, [1]=#junctions, [2]=#direct arcs, [3]=#junction in-arcs, [4]=#junction
 * out-arcs, [5]=states per lane, [6]=1 if pdf map is the identity, [7]=dense nnz. */
BEER_API int beer_graph_plan_info(const beer_graph_plan* plan, int32_t* info8_host);

/* Bytes of workspace beer_hmm_forward_backward needs for N frames. */
BEER_API int64_t beer_hmm_workspace_bytes(const beer_graph_plan* plan, int64_t N);

/* KB: forward-backward over one graph for a ragged batch of utterances.
 * Replaces CompiledGraph._baum_welch_forward/_backward/posteriors
 * (beer/graph.py:270-326) and the expected value of HMM.expected_log_likelihood
 * (beer/models/hmm.py:79-92), p_tk = scale * pdf_llh[t, map[k]].
 *   state_post [N,K] or NULL  <- gamma_tk (per-frame normalised, graph.py:306-307)
 *   pdf_post  [N,ld_post] or NULL <- scale * sum_{k: map[k]=pdf} gamma_tk
 *                                   (hmm.py:94-95 + modelset.py:148-154)
 *   frame_exp_llh [N] or NULL <- sum_k p_tk gamma_tk + scale*frame_ref[t]  (hmm.py:87)
 *   utt_exp_llh [n_utts] (fp64) <- sum_t of the above
 *   utt_logz [n_utts] (fp64) or NULL <- log evidence of the scaled llhs (natural log,
 *                                   includes scale*frame_ref)
 *   workspace: beer_hmm_workspace_bytes(plan, N) bytes. */
BEER_API int beer_hmm_forward_backward(const beer_graph_plan* plan, const float* pdf_llh, int64_t ld_pdf,
                              const float* frame_ref, const int64_t* utt_off, int n_utts,
                              float scale, float* state_post, float* pdf_post, int64_t ld_post,
                              float* frame_exp_llh, double* utt_exp_llh, double* utt_logz,
                              void* workspace, void* stream);

/* KB + the transition-posterior reductions a phone loop needs (beer/graph.py:308-323 reduced as in
 * PhoneLoop.accumulate, beer/models/phoneloop.py:88-97): same as beer_hmm_forward_backward, and
 *   unit_counts [P] (fp64, += ) <- sum_t sum_{e in unit ends} xi_t[e, start_u] + gamma_0[start_u]
 * for every unit u of an aligned left-to-right loop (P = beer_hmm_unit_count_size(plan) units of
 * K / P states each; 0 = the graph is not such a loop, the call then returns BEER_ERR_UNSUPPORTED).
 * The (T-1) x K x K tensor of the reference is never formed. */
BEER_API int beer_hmm_unit_count_size(const beer_graph_plan* plan);
BEER_API int beer_hmm_forward_backward_units(const beer_graph_plan* plan, const float* pdf_llh, int64_t ld_pdf,
                                    const float* frame_ref, const int64_t* utt_off, int n_utts, float scale,
                                    float* state_post, float* pdf_post, int64_t ld_post, float* frame_exp_llh,
                                    double* utt_exp_llh, double* utt_logz, double* unit_counts, void* workspace,
                                    void* stream);

/* KV: Viterbi best path (first-max tie-breaking) for a ragged batch.
 * Replaces CompiledGraph.best_path (beer/graph.py:329-344).
 *   path [N] (int32) <- state ids; workspace: N*K*sizeof(uint16_t) bytes. */
BEER_API int beer_hmm_viterbi(const beer_graph_plan* plan, const float* pdf_llh, int64_t ld_pdf,
                     const int64_t* utt_off, int n_utts, float scale, int32_t* path,
                     void* workspace, void* stream);

/* KB for per-utterance alignment graphs (aligned training: accumulate.py:47-57 -> hmm.py:73-92 with
 * inference_graph=).  An alignment graph (mkaligraph.py:18-39) compiles to a left-to-right chain: state j has a
 * self loop and one arc to j + 1, the path starts in state 0 and ends in the last state.  Utterance u owns states
 * chain_off[u] .. chain_off[u+1]-1 of the per-state arrays (device): chain_pdf (pdf id), chain_log_self = ln a(j,j),
 * chain_log_next = ln a(j,j+1) (last state: the final weight); chain_log_init [n_utts] = ln weight of entering state 0.
 * Outputs as beer_hmm_forward_backward, except: pdf_post must be ZEROED by the caller (scatter-add, a pdf may occur
 * several times in a chain) and state_post, if given, has row stride beer_hmm_chain_row_stride(max_chain_len).
 * max_chain_len <= 1024.  workspace: beer_hmm_chain_workspace_bytes(max_chain_len, N) bytes.  pdf_llh is an [N, ld_pdf]
 * allocation: when ld_pdf is a multiple of 4 (and the base 16-byte aligned) whole rows of ld_pdf floats are staged in
 * shared memory, so the padding columns of a row, the last one included, must be readable. */
BEER_API int beer_hmm_chain_row_stride(int max_chain_len);
BEER_API int64_t beer_hmm_chain_workspace_bytes(int max_chain_len, int64_t N);
BEER_API int beer_hmm_forward_backward_chains(const float* pdf_llh, int64_t ld_pdf, const float* frame_ref,
                                     const int64_t* utt_off, int n_utts, const int64_t* chain_off,
                                     const int32_t* chain_pdf, const float* chain_log_self,
                                     const float* chain_log_next, const float* chain_log_init, int max_chain_len,
                                     float scale, float* state_post, float* pdf_post, int64_t ld_post,
                                     float* frame_exp_llh, double* utt_exp_llh, double* utt_logz, void* workspace,
                                     void* stream);

/* Transition posteriors, per-step normalised with NaN -> 0: CompiledGraph.posteriors(trans_posteriors=True)
 * (beer/graph.py:308-323), or only the rows / columns a caller keeps: BigramPhoneLoop.accumulate reads
 * xi[:, ends, starts] (beer/models/phoneloop.py:175-186).  Dense O(T R C) output by definition: the API-parity
 * path for small graphs (K <= 1024); PhoneLoop training uses the fused unit counts of
 * beer_hmm_forward_backward_units instead.
 *   pdf_llh [N, ld_pdf], pdf_map [K] int32 or NULL (identity), log_init [K], log_trans [K, K] (natural log),
 *   state_post [N, K] from beer_hmm_forward_backward*, rows [n_rows] / cols [n_cols] int32 state ids or NULL
 *   (all K), xi [N - n_utts, R, C] <- utterance u starts at row utt_off[u] - u. */
BEER_API int beer_hmm_transition_posteriors(const float* pdf_llh, int64_t ld_pdf, const int32_t* pdf_map, float scale,
                                   const float* log_init, const float* log_trans, int K, const float* state_post,
                                   const int64_t* utt_off, int n_utts, const int32_t* rows, int n_rows,
                                   const int32_t* cols, int n_cols, float* xi, void* stream);

/* KC: posterior-weighted sufficient statistics, accumulated (+=) into a caller-zeroed
 * fp64 buffer.  Replaces the joint responsibilities of MixtureSet.accumulate
 * (beer/models/mixtureset.py:100-112) and NormalSet.accumulate
 * (beer/models/normalset.py:121-123):  w_tj = pdf_post[t, pdf(j)] * r_tj,
 * r_tj = exp(comp_llh_tj - pdf_llh_t,pdf(j)).
 *   pdf_post [N, ld_post] or NULL (all ones: plain Mixture, mixture.py:95-102)
 *   pdf_llh / comp_llh / comp_off as produced by beer_emission_llh
 *     (comp_llh NULL: one Gaussian per pdf, w = pdf_post)
 *   acc_normal [M, 2D+2] += [sum w x, -1/2 sum w x^2, -1/2 sum w, 1/2 sum w]. */
BEER_API int beer_accumulate_stats(const float* X, int64_t N, int D, const float* pdf_post, int64_t ld_post,
                          const float* pdf_llh, int64_t ld_pdf, const float* comp_llh,
                          const int32_t* comp_off, int Kp, int M, double* acc_normal, void* stream);

/* KC on the tensor cores (tcgen05 + TMEM, 3xTF32 split, fp32 accumulation in TMEM, one fp64
 * atomic flush per CTA): same contract as beer_accumulate_stats for D in {20, 40, 64, 80}.
 *   beer_accumulate_tc_supported(M, D) -> 1 if this shape has a tensor-core path. */
BEER_API int beer_accumulate_tc_supported(int M, int D);
BEER_API int beer_accumulate_stats_tc(const float* X, int64_t N, int D, const float* pdf_post, int64_t ld_post,
                             const float* pdf_llh, int64_t ld_pdf, const float* comp_llh,
                             const int32_t* comp_off, int Kp, int M, double* acc_normal, void* stream);

/* Categorical statistics of the mixture weights from the accumulated Normal statistics
 * (sum_t w_tj = 2 * acc_normal[j, 2D+1]):  per pdf [n_c (c < C-1), sum_c n_c], the layout
 * of CategoricalLikelihood.sufficient_statistics (beer/dists/dirichlet.py:18-21) summed
 * over frames as CategoricalSet.accumulate_from_jointresps does
 * (beer/models/categoricalset.py:54-55).  acc_weights [M] is overwritten. */
BEER_API int beer_mixture_weight_stats(const double* acc_normal, int M, int D, const int32_t* comp_off, int Kp,
                              double* acc_weights, void* stream);

/* ------------------------------------------------------------------------
 * M-step and KL (beer/models/parameters.py:134-141, beer/dists/basedist.py:243-263)
 * ---------------------------------------------------------------------- */

/* Natural-gradient step of M Normal-Gamma posteriors, in place:
 *   eta <- eta + lrate * (eta_prior + stats_scale * acc - eta), then back to
 *   (mean, scale, shape, rates) (beer/dists/normalgamma.py:76-94, 163-180). */
BEER_API int beer_normalgamma_update(const float* prior_mean, const float* prior_scale,
                            const float* prior_shape, const float* prior_rates, float* mean,
                            float* scale, float* shape, float* rates, const double* acc,
                            double stats_scale, double lrate, int M, int D, void* stream);
/* kl[0] += sum_j KL(q_j || p_j) (fp64; normalgamma.py:151-157 log-normaliser). */
BEER_API int beer_normalgamma_kl(const float* prior_mean, const float* prior_scale, const float* prior_shape,
                        const float* prior_rates, const float* mean, const float* scale,
                        const float* shape, const float* rates, int M, int D, double* kl,
                        void* stream);
/* Dirichlet counterparts (beer/dists/dirichlet.py:70-81, 135-159); conc [K,C]. */
BEER_API int beer_dirichlet_update(const float* prior_conc, float* conc, const double* acc,
                          double stats_scale, double lrate, int K, int C, void* stream);
BEER_API int beer_dirichlet_kl(const float* prior_conc, const float* conc, int K, int C, double* kl,
                      void* stream);


/* ------------------------------------------------------------------------
 * beer.dists accessors (API completeness; none of them is on the per-iteration hot path)
 * ---------------------------------------------------------------------- */

/* T(x) = [x, -x^2/2, -1/2, 1/2], out [N, 2D+2].  Replaces
 * NormalDiagonalLikelihood.sufficient_statistics (beer/dists/normalgamma.py:19-27). */
BEER_API int beer_normal_sufficient_statistics(const float* X, int64_t N, int D, float* out, void* stream);
/* NormalGamma.natural_parameters (normalgamma.py:163-180), nat [M, 2D+2]. */
BEER_API int beer_normalgamma_natural_params(const float* mean, const float* scale, const float* shape,
                                    const float* rates, int M, int D, float* nat, void* stream);
/* NormalGammaStdParams.from_natural_parameters (normalgamma.py:76-94). */
BEER_API int beer_normalgamma_from_natural(const float* nat, int M, int D, float* mean, float* scale, float* shape,
                                  float* rates, void* stream);
/* NormalGamma.log_norm (normalgamma.py:151-157), out [M] fp64. */
BEER_API int beer_normalgamma_log_norm(const float* scale, const float* shape, const float* rates, int M, int D,
                              double* out, void* stream);
/* Dirichlet.natural_parameters / expected_sufficient_statistics (log-odds form) / log_norm /
 * DirichletStdParams.from_natural_parameters (beer/dists/dirichlet.py:144-159, 106-128, 135-138, 70-81). */
BEER_API int beer_dirichlet_natural_params(const float* conc, int K, int C, float* nat, void* stream);
BEER_API int beer_dirichlet_expected_stats(const float* conc, int K, int C, float* ets, void* stream);
BEER_API int beer_dirichlet_log_norm(const float* conc, int K, int C, double* out, void* stream);
BEER_API int beer_dirichlet_from_natural(const float* nat, int K, int C, float* conc, void* stream);

/* pdf_llh[t, k] = logsumexp over the Gaussians of pdf k of comp_llh[t, :] (mixtures with more
 * Gaussians than one emission tile holds: Mixture.expected_log_likelihood, beer/models/mixture.py:79-83). */
BEER_API int beer_segment_logsumexp(const float* comp_llh, int64_t N, int M, const int32_t* comp_off, int Kp,
                           float* pdf_llh, int64_t ld_pdf, void* stream);

/* KC for a GIVEN path (Viterbi training / forced alignment, hmm.py:42-58 + hmm.py:94-100 + normalset.py:121-123): the
 * posteriors are the one-hot rows scale * onehot(pdf_ids[t]); they are never materialised (a stage of the tensor-core
 * kernel brings 32 pdf ids instead of 32 x M floats).  acc_normal [M, 2D+2] fp64 += as beer_accumulate_stats.
 *   pdf_ids: int32, 16-byte aligned, readable up to N rounded up to a multiple of 4; M <= 128, D in {20, 40}
 *   (BEER_ERR_UNSUPPORTED otherwise: use beer_path_posteriors + beer_accumulate_stats). */
BEER_API int beer_accumulate_stats_path(const float* X, int64_t N, int D, const int32_t* pdf_ids, float scale, int M,
                               double* acc_normal, void* stream);

/* Posteriors of a GIVEN state path (viterbi=True / state_path=..., beer/models/hmm.py:42-58, 87):
 * pdf_post[t, :] = scale * onehot(map[path_t]) (overwritten), frame_exp_llh[t] = scale * llh of that pdf. */
BEER_API int beer_path_posteriors(const int32_t* path, int64_t N, const int32_t* pdf_map, float scale,
                         const float* pdf_llh, int64_t ld_pdf, const float* frame_ref, float* pdf_post,
                         int64_t ld_post, int Kp, float* frame_exp_llh, void* stream);

/* ------------------------------------------------------------------------
 * fbank front-end (beer/features.py; not on the timed VB path)
 * ---------------------------------------------------------------------- */

/* log(1 + mel filterbank energies of |rFFT| of the pre-emphasised, windowed frames).  Replaces
 * beer.features.fbank (beer/features.py:145-204).
 *   signal [n_samples] fp32; window [frame_len]; filters_t [fft_len/2, n_filters] (the transposed matrix of
 *   create_fbank, features.py:47-79); fft_len in {256, 512, 1024};
 *   out [(n_samples - frame_len) / frame_shift + 1, n_filters]. */
BEER_API int beer_fbank(const float* signal, int64_t n_samples, int frame_len, int frame_shift, float preemph,
               const float* window, const float* filters_t, int fft_len, int n_filters, float* out, void* stream);
/* The front-end `beer features extract` runs (beer/features.py:102-143 short_term_mspec +
 * beer/cli/subcommands/features/extract.py:107-127): DC offset removed (dc_offset = mean of the signal), pre-emphasis
 * INSIDE every frame (first sample against itself), window, |rFFT|.  filters_t == NULL: out [n_frames, fft_len/2] =
 * the magnitude spectrum; else out [n_frames, n_filters] = log(log_offset + mspec @ filters) (extract.py: 1e-6). */
BEER_API int beer_short_term_mspec(const float* signal, int64_t n_samples, int frame_len, int frame_shift, float preemph,
                          float dc_offset, const float* window, const float* filters_t, int fft_len, int n_filters,
                          float log_offset, float* out, void* stream);
/* One order of the delta regression filter with replicated edges (beer/features.py:82-100):
 * out[t] = sum_{k=1..wlen} k (fea[t+k] - fea[t-k]) / (2 sum k^2). */
BEER_API int beer_add_deltas(const float* fea, int n_frames, int dim, int wlen, float* out, void* stream);

/* 1 when beer_hmm_forward_backward_ex can write pdf_lpost for this graph. */
BEER_API int beer_hmm_lpost_supported(const beer_graph_plan* plan);
/* Same as beer_hmm_forward_backward_units with the domain of the llhs stated: flags & BEER_FB_LLH_LOG2 = pdf_llh holds
 * log2 values (what beer_mix16_emission writes; frame_ref stays in nats), else nats.  pdf_lpost (optional, [N, ld_lpost],
 * 16-byte aligned rows) = log2(scale * posterior) per pdf, -inf for zero: only the left-to-right loop kernels write it
 * (BEER_ERR_UNSUPPORTED otherwise: take pdf_post and beer_mix16_log2_posteriors).  flags & BEER_FB_LPOST_RELATIVE:
 * pdf_lpost = log2(scale * posterior) - log2 llh of the pdf instead, the ONE array beer_mix16_accumulate (llh2 = NULL)
 * adds to its z to weight a Gaussian (gamma_tk r_tkc = 2^(z_tkc + lpost_tk - llh2_tk), mixtureset.py:100-112); its
 * rounding error is half an ulp of |llh2| in the exponent (offset-form llhs of a few hundred: <= 1e-5 of the weight,
 * zero-mean over frames), where the two-array form subtracts z - llh2 exactly. */
#define BEER_FB_LLH_LOG2 1
#define BEER_FB_LPOST_RELATIVE 2
BEER_API int beer_hmm_forward_backward_ex(const beer_graph_plan* plan, const float* pdf_llh, int64_t ld_pdf,
                                 const float* frame_ref, const int64_t* utt_off, int n_utts, float scale,
                                 float* state_post, float* pdf_post, int64_t ld_post, float* frame_exp_llh,
                                 double* utt_exp_llh, double* utt_logz, double* unit_counts, int flags,
                                 float* pdf_lpost, int64_t ld_lpost, void* workspace, void* stream);
/* The same with an ACTIVITY MAP for beer_mix16_accumulate_blocks: block_active [ceil(N / 64), ld_active] bytes (zeroed
 * by the caller) gets a 1 for every (tile of 64 frames, block of pdfs_per_block consecutive pdfs) in which some pdf
 * posterior is large enough to be non-zero in the statistics kernel's fp16 weight operands (2^-(25 + wexp + 2), wexp =
 * beer_mix16_weight_exponent(scale)).  Everywhere else gamma_tk r_tkc x_t contributes EXACTLY zero to the first and second
 * moments there, so the statistics kernel may skip those (tile, block) pairs (mixtureset.py:100-112 has no such notion:
 * the reference multiplies the zeros).  Needs pdf_lpost; beer_hmm_block_activity_supported(plan, unit_counts != NULL). */
BEER_API int beer_hmm_block_activity_supported(const beer_graph_plan* plan, int with_unit_counts);
BEER_API int beer_hmm_forward_backward_blocks(const beer_graph_plan* plan, const float* pdf_llh, int64_t ld_pdf,
                                     const float* frame_ref, const int64_t* utt_off, int n_utts, float scale,
                                     float* state_post, float* pdf_post, int64_t ld_post, float* frame_exp_llh,
                                     double* utt_exp_llh, double* utt_logz, double* unit_counts, int flags,
                                     float* pdf_lpost, int64_t ld_lpost, uint8_t* block_active, int64_t ld_active,
                                     int pdfs_per_block, void* workspace, void* stream);
/* The power of two the statistics kernel scales its weights by (w 2^wexp fills the fp16 range; posteriors carry `scale`). */
BEER_API int beer_mix16_weight_exponent(float scale);

/* ------------------------------------------------------------------------
 * Mixture path without per-Gaussian llhs in HBM (csrc/mix16.cu): MixtureSet.expected_log_likelihood +
 * MixtureSet.accumulate + NormalSet.accumulate (beer/models/mixtureset.py:85-112, normalset.py:117-123) for
 * pdfs of C in {4, 8, 16} Gaussians, D in {20, 40}, on tcgen05 kind::f16 with a 3-pass fp16 split.
 * ---------------------------------------------------------------------- */

BEER_API int beer_mix16_supported(int M, int D, int C);
/* Debug: 256 x 8 uint64 time stamps (clock64 of CTA 0) per chunk / tile of the following mix16 launches; NULL = off. */
BEER_API void beer_mix16_set_trace(void* dev_buf);
/* sizes_host[6] = {halfs of img1 (= of img2) for N frames, halfs of wimg, 32-bit words of wtm, floats of k12,
 * Gaussians per emission chunk, padded statistics width KP}. */
BEER_API int beer_mix16_geometry(int M, int D, int C, int64_t N, int64_t* sizes_host);
/* Feature images of a resident run of frames, built once: alpha [2D] (out) = per-dimension power-of-two scales of the
 * statistics [x | -x^2/2] (NormalDiagonalLikelihood.sufficient_statistics, normalgamma.py:19-27), img1 / img2 (out,
 * 128-byte aligned) = their fp16 hi / lo split as UMMA operand tiles of 64 frames (frame-major / statistic-major).
 * absmax_scratch: D words. */
BEER_API int beer_mix16_feature_images(const float* X, int64_t N, int D, float* alpha, uint32_t* absmax_scratch, void* img1,
                              void* img2, void* stream);
/* Per VB iteration: (W [M, 2D], bias [M]) of beer_emission_prepare -> packed fp16 weights for the two kernels
 * (wimg: emission, wtm: statistics) and the per-Gaussian affine z = S k1 + k2 (log2 domain), k12 = (k1, k2) pairs. */
BEER_API int beer_mix16_pack(const float* W, const float* bias, const float* alpha, int M, int D, int C, void* wimg,
                    uint32_t* wtm, float* k12, void* stream);
/* frame_ref [N] = the per-frame constant of the offset form (see beer_emission_prepare). */
BEER_API int beer_mix16_frame_ref(const float* X, int64_t N, int D, const float* ref, float* frame_ref, void* stream);
/* llh2 [N, ld] = log2 sum_c 2^z (offset form, log2 units): MixtureSet.expected_log_likelihood without the
 * per-Gaussian llhs (mixtureset.py:85-98). */
BEER_API int beer_mix16_emission(const void* img1, int64_t N, int D, const void* wimg, const float* k12, int M, int C,
                        float* llh2, int64_t ld, void* stream);
/* acc_normal [M, 2D+2] (fp64) += sum_t 2^pdf_lpost[t, pdf(j)] r_tj T(x_t) with the responsibilities r = 2^(z - llh2)
 * recomputed on chip (mixtureset.py:100-112, normalset.py:121-123).  pdf_lpost [N, ld] = log2 of the pdf posteriors
 * (times `scale`, -inf = zero): written by beer_hmm_forward_backward_ex, or beer_mix16_log2_posteriors of pdf_post.
 * llh2 = NULL (C > 1): pdf_lpost is in the relative form of BEER_FB_LPOST_RELATIVE (already minus llh2): one array
 * streamed instead of two and a third fewer epilogue instructions per (frame, Gaussian).
 * C = 1 (single-Gaussian pdfs, NormalSet.accumulate alone): pdf_lpost holds the posteriors THEMSELVES (linear, as
 * beer_hmm_forward_backward writes pdf_post) and img1 / wtm / k12 / llh2 are not read (may be NULL). */
BEER_API int beer_mix16_accumulate(const void* img1, const void* img2, int64_t N, int D, const uint32_t* wtm, const float* k12,
                          const float* alpha, int M, int C, const float* pdf_lpost, int64_t ld_lpost,
                          const float* llh2, int64_t ld_llh, float scale, double* acc_normal, void* stream);
/* The same with the activity map of beer_hmm_forward_backward_blocks (pdfs_per_block = 128 / C, the pdfs of one tile of
 * 128 Gaussians): a CTA works only on the tiles of 64 frames in which the pdfs of its Gaussian tile are marked.  The
 * skipped pairs contribute exactly zero to the first and second moments (their weights are zero in both fp16 halves of
 * the operand); the remaining products are the same, grouped differently into the kernel's fp32 partial sums (equal to
 * summation order, ~1e-7 relative); the counts differ by the fp32 sum of the skipped weights (< 2^-41 each).  block_active == NULL: the dense call. */
BEER_API int beer_mix16_accumulate_blocks(const void* img1, const void* img2, int64_t N, int D, const uint32_t* wtm,
                                 const float* k12, const float* alpha, int M, int C, const float* pdf_lpost,
                                 int64_t ld_lpost, const float* llh2, int64_t ld_llh, float scale,
                                 const uint8_t* block_active, int64_t ld_active, double* acc_normal, void* stream);
/* GMM without an HMM (Mixture.expected_log_likelihood, beer/models/mixture.py:70-93) on the same kernels: the M
 * components count as Kp pseudo-pdfs of C; after beer_mix16_emission this finishes the softmax over the frame:
 * pdf_lpost[t, k] = llh2[t, k] - log2 sum_k 2^llh2[t, k] + log2(scale), frame_exp_llh[t] (optional) = scale * LSE over
 * all components (nats, frame_ref added), utt_exp_llh[u] (optional, fp64, caller-zeroed) += their sums per utterance. */
BEER_API int beer_mix16_gmm_posteriors(const float* llh2, int64_t N, int Kp, int64_t ld_llh, const float* frame_ref,
                              const int64_t* utt_off, int n_utts, float scale, float* pdf_lpost, int64_t ld_lpost,
                              float* frame_exp_llh, double* utt_exp_llh, void* stream);
BEER_API int beer_mix16_log2_posteriors(const float* pdf_post, int64_t N, int Kp, int64_t ld_post, float* pdf_lpost,
                               int64_t ld_lpost, void* stream);

/* ------------------------------------------------------------------------
 * Roofline probes (measurement only: bench.py / tools/microbench.py time them with CUDA events)
 * ---------------------------------------------------------------------- */

/* n_mma back-to-back tcgen05.mma (M = 128, N = 256, one k-step) per SM on resident operands: the dispatch-limited
 * tensor-pipe peak of MMA kind 0 = tf32 (K = 8) or 1 = f16 (K = 16).  *flops_out_host = flops issued by the launch. */
BEER_API int beer_probe_mma(int kind, int n_mma, double* flops_out_host, void* stream);
/* kind::f16 MMAs of width N (M = 128, one k-step each) rotating over n_acc accumulators and n_buf operand buffers, A in
 * shared memory or (a_tmem != 0) tensor memory, issued under elect.sync (elect != 0) or under `lane == 0`: how tile width, accumulation chains and operand fetch pace the pipe. */
BEER_API int beer_probe_mma_shape(int n_mma, int N, int n_acc, int n_buf, int a_tmem, int elect, double* flops_out_host,
                         void* stream);
/* Write-only stream over `bytes` of dst (128-byte aligned): mode 0 = float4 stores, 1 = 32 KB bulk copies
 * shared -> global (cp.async.bulk).  The DRAM write ceiling of an llh-producing kernel. */
BEER_API int beer_probe_fill(float* dst, int64_t bytes, int mode, void* stream);
/* Global -> shared copy-engine rate: one CTA per SM streams copies_per_sm chunks of chunk_bytes (multiple of 256) from
 * src (src_bytes, meant to fit in L2) through `stages` shared-memory slots, issued by `issuers` threads (one per warp);
 * mode 0 = cp.async.bulk, 1 = 2-D tensor map; shared_walk != 0 = every SM reads the SAME chunks in the same order
 * (what CTAs streaming a common operand do) instead of its own part of the buffer. */
BEER_API int beer_probe_tma(const float* src, int64_t src_bytes, int mode, int chunk_bytes, int stages, int copies_per_sm,
                   int issuers, int shared_walk, void* stream);
/* Read-only stream over `bytes` of src (float4 loads). */
BEER_API int beer_probe_read(const float* src, int64_t bytes, float* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BEER_B200_H_ */
