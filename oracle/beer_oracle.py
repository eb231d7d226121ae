"""CPU oracle for the beer VB-EM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy restatement (fp64 unless the caller passes fp32 arrays) of the
reference algorithm behind ``beer.evidence_lower_bound`` on HMM / GMM models.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; nothing under ``beer_b200/``
does.  Every function cites the reference file:line it restates (paths are
relative to the reference checkout, beer-asr/beer @ d53d2a1).

Parity pinning: the reference's own test-suite does not exercise this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the live
reference run in fp64 in the build container: ``tests/golden/make_goldens.py``
generated ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks
every function below against them.
"""

import numpy as np
from scipy.special import digamma, gammaln

LOG_2PI = float(np.log(2.0 * np.pi))


# --------------------------------------------------------------------------
# helpers (beer/utils.py)
# --------------------------------------------------------------------------

def logsumexp(x, axis):
    """Stable log-sum-exp returning -inf/+inf when the max is infinite
    (beer/utils.py:105-123; torch.logsumexp has the same convention)."""
    x = np.asarray(x)
    m = np.max(x, axis=axis, keepdims=True)
    safe = np.where(np.isfinite(m), m, 0.0)
    with np.errstate(divide='ignore', invalid='ignore'):
        out = safe + np.log(np.sum(np.exp(x - safe), axis=axis, keepdims=True))
    out = np.where(np.isfinite(m), out, m)
    return np.squeeze(out, axis=axis)


def onehot(labels, n, dtype=np.float64):
    """beer/utils.py:84-102."""
    labels = np.asarray(labels, dtype=np.int64)
    out = np.zeros((len(labels), n), dtype=dtype)
    out[np.arange(len(labels)), labels] = 1
    return out


# --------------------------------------------------------------------------
# Normal-Gamma  <->  diagonal Normal likelihood (beer/dists/normalgamma.py)
# --------------------------------------------------------------------------

def normal_diag_sufficient_statistics(X):
    """[x, -x^2/2, -1/2, 1/2]  (normalgamma.py:19-27)."""
    X = np.asarray(X)
    ones = np.ones((len(X), 1), dtype=X.dtype)
    return np.concatenate([X, -.5 * X ** 2, -.5 * ones, .5 * ones], axis=-1)


def normalgamma_natural_parameters(mean, scale, shape, rates):
    """[k m, -k m^2/2 - b, -k/2, a - 1/2]  (normalgamma.py:163-180).
    mean (M,D), scale (M,1), shape (M,1), rates (M,D)."""
    return np.concatenate([scale * mean, -.5 * scale * mean ** 2 - rates,
                           -.5 * scale, shape - .5], axis=-1)


def normalgamma_from_natural_parameters(eta):
    """Inverse map (normalgamma.py:76-94)."""
    dim = (eta.shape[-1] - 2) // 2
    np1, np2 = eta[:, :dim], eta[:, dim:2 * dim]
    np3, np4 = eta[:, -2], eta[:, -1]
    scale = -2 * np3
    shape = np4 + .5
    mean = np1 / scale[:, None]
    rates = -np2 - .5 * scale[:, None] * mean ** 2
    return mean, scale[:, None], shape[:, None], rates


def normalgamma_expected_sufficient_statistics(mean, scale, shape, rates):
    """E_q[T(theta)] = [a/b m, a/b, D/k + sum a/b m^2, sum psi(a) - ln b]
    (normalgamma.py:118-146)."""
    dim = mean.shape[-1]
    prec = shape / rates
    pqm = (prec * mean ** 2).sum(axis=-1, keepdims=True) + dim / scale
    logdet = np.sum(digamma(shape) - np.log(rates), axis=-1, keepdims=True)
    return np.concatenate([prec * mean, prec, pqm, logdet], axis=-1)


def normalgamma_log_norm(mean, scale, shape, rates):
    """normalgamma.py:151-157."""
    dim = rates.shape[-1]
    return (dim * gammaln(shape)
            - shape * np.log(rates).sum(axis=-1, keepdims=True)
            - .5 * dim * np.log(scale)).sum(axis=-1)


def normal_diag_llh(stats, ets, dim):
    """stats @ E[T].T - D/2 ln 2pi  (normalgamma.py:55-59, normalset.py:117-119)."""
    return stats @ ets.T - .5 * dim * LOG_2PI


# --------------------------------------------------------------------------
# Dirichlet  <->  Categorical likelihood (beer/dists/dirichlet.py)
# --------------------------------------------------------------------------

def categorical_sufficient_statistics(data):
    """Last column replaced by the row sum (dirichlet.py:18-21)."""
    out = np.array(data, copy=True)
    flat = out.reshape(-1, out.shape[-1])
    flat[:, -1] = flat.sum(axis=-1)
    return flat.reshape(out.shape)


def dirichlet_natural_parameters(conc):
    """[a_i - 1 (i<C-1), sum_i (a_i - 1)]  (dirichlet.py:144-159)."""
    c = np.atleast_2d(conc)
    out = c - 1
    out[:, -1] = (c - 1).sum(axis=-1)
    return out.reshape(np.shape(conc))


def dirichlet_from_natural_parameters(eta):
    """dirichlet.py:70-81."""
    e = np.atleast_2d(eta)
    conc = e + 1
    conc[:, -1] = e[:, -1] - e[:, :-1].sum(axis=-1) + 1
    return conc.reshape(np.shape(eta))


def dirichlet_expected_sufficient_statistics(conc):
    """Log-odds parameterisation (dirichlet.py:106-128)."""
    c = np.atleast_2d(conc)
    out = np.zeros_like(c)
    psi = digamma(c[:, -1])
    out[:, :-1] = digamma(c[:, :-1]) - psi[:, None]
    out[:, -1] = psi - digamma(c.sum(axis=-1))
    return out.reshape(np.shape(conc))


def dirichlet_log_norm(conc):
    """dirichlet.py:135-138."""
    return gammaln(conc).sum(axis=-1) - gammaln(np.sum(conc, axis=-1))


def categorical_log_weights(conc):
    """E[ln pi] through the eye(C) trick of mixtureset.py:64-67 /
    mixture.py:45-48 (= psi(a_c) - psi(sum a))."""
    c = np.atleast_2d(conc)
    eye = np.eye(c.shape[-1], dtype=c.dtype)
    stats = categorical_sufficient_statistics(eye)
    ets = dirichlet_expected_sufficient_statistics(c)
    out = stats @ ets.T          # (C, K)   dirichlet.py:60-62
    return out.T.reshape(np.shape(conc))


def kl_div(lognorm_q, lognorm_p, exp_stats_q, eta_q, eta_p):
    """KL(q || p) = A(p) - A(q) - <E_q[T], eta_p - eta_q>  (basedist.py:243-263)."""
    return lognorm_p - lognorm_q - np.sum(exp_stats_q * (eta_p - eta_q), axis=-1)


def normalgamma_kl(post, prior):
    """post/prior: tuples (mean, scale, shape, rates)."""
    return kl_div(normalgamma_log_norm(*post), normalgamma_log_norm(*prior),
                  normalgamma_expected_sufficient_statistics(*post),
                  normalgamma_natural_parameters(*post),
                  normalgamma_natural_parameters(*prior))


def dirichlet_kl(post, prior):
    return kl_div(dirichlet_log_norm(post), dirichlet_log_norm(prior),
                  dirichlet_expected_sufficient_statistics(post),
                  dirichlet_natural_parameters(post),
                  dirichlet_natural_parameters(prior))


# --------------------------------------------------------------------------
# emission models
# --------------------------------------------------------------------------

def mixtureset_llh(pc_llh, logw):
    """Per-state llh of K mixtures x C comps (mixtureset.py:85-98).
    pc_llh (T, K*C), logw (K, C) -> (log_norm (T,K), resps (T,K,C))."""
    K, C = logw.shape
    w = pc_llh.reshape(-1, K, C) + logw[None]
    log_norm = logsumexp(w, axis=-1)
    resps = np.exp(w - log_norm[:, :, None])
    return log_norm, resps


def mixture_expected_llh(pc_llh, logw, labels=None):
    """Mixture.expected_log_likelihood (mixture.py:70-93): returns
    (exp_llh (T,), resps (T,K))."""
    if labels is None:
        w = pc_llh + logw[None]
        lnorm = logsumexp(w, axis=1)[:, None]
        log_resps = w - lnorm
        resps = np.exp(log_resps)
        local_kl = np.sum(resps * (log_resps - logw[None]), axis=-1)
    else:
        resps = onehot(labels, pc_llh.shape[1], dtype=pc_llh.dtype)
        local_kl = 0.
    return (pc_llh * resps).sum(axis=-1) - local_kl, resps


# --------------------------------------------------------------------------
# HMM inference over a compiled graph (beer/graph.py)
# --------------------------------------------------------------------------

def forward(llhs, init_log, trans_log):
    """graph.py:270-278."""
    T, K = llhs.shape
    la = np.full_like(llhs, -np.inf)
    la[0] = llhs[0] + init_log
    At = trans_log.T
    for t in range(1, T):
        la[t] = llhs[t] + logsumexp(la[t - 1] + At, axis=1)
    return la


def backward(llhs, final_log, trans_log):
    """graph.py:280-287."""
    T, K = llhs.shape
    lb = np.full_like(llhs, -np.inf)
    lb[-1] = final_log
    for t in reversed(range(T - 1)):
        lb[t] = logsumexp(trans_log + llhs[t + 1] + lb[t + 1], axis=1)
    return lb


def posteriors(llhs, init_log, final_log, trans_log, trans_posteriors=False):
    """CompiledGraph.posteriors (graph.py:289-326): per-frame normalised state
    posteriors, optional per-step normalised transition posteriors (NaN -> 0),
    and lognorm.mean()."""
    la = forward(llhs, init_log, trans_log)
    lb = backward(llhs, final_log, trans_log)
    lognorm = logsumexp(la + lb, axis=1)
    with np.errstate(invalid='ignore'):
        gamma = np.exp(la + lb - lognorm[:, None])
    if not trans_posteriors:
        return gamma, lognorm.mean()
    K = llhs.shape[1]
    with np.errstate(invalid='ignore'):
        lxi = la[:-1, :, None] + trans_log[None] + (llhs + lb)[1:, None, :]
        lxi = lxi.reshape(-1, K * K)
        ln = logsumexp(lxi, axis=1)
        xi = np.exp(lxi - ln[:, None])
    xi = np.where(xi != xi, 0.0, xi).reshape(-1, K, K)
    return (gamma, xi), lognorm.mean()


def best_path(llhs, init_log, final_log, trans_log):
    """Viterbi with first-max tie-breaking (graph.py:329-344)."""
    T, K = llhs.shape
    bt = np.zeros((T, K), dtype=np.int64)
    omega = llhs[0] + init_log
    At = trans_log.T
    for t in range(1, T):
        hyp = omega + At                       # hyp[j, i] = omega_i + logA_ij
        bt[t] = np.argmax(hyp, axis=1)
        omega = llhs[t] + hyp[np.arange(K), bt[t]]
    path = [int(np.argmax(omega + final_log))]
    for t in reversed(range(1, T)):
        path.insert(0, int(bt[t, path[0]]))
    return np.asarray(path, dtype=np.int64)


# --------------------------------------------------------------------------
# graph compilation (beer/graph.py:103-240) -- restated on plain containers
# --------------------------------------------------------------------------

class OracleGraph:
    """Mutable graph with the reference's add_state / add_arc / normalize /
    replace_state / compile semantics (graph.py:60-240)."""

    def __init__(self):
        self.n = 0
        self.pdf = {}            # state id -> pdf id or None (insertion ordered)
        self.arcs = {}           # (start, end) -> weight   (Arc hash = start,end)
        self.start_state = None
        self.end_state = None

    def add_state(self, pdf_id=None):
        sid = self.n
        self.n += 1
        self.pdf[sid] = pdf_id
        return sid

    def add_arc(self, start, end, weight=1.0):
        # set.add of an equal (start, end) Arc keeps the first one (graph.py:24-29,110-113)
        self.arcs.setdefault((start, end), weight)

    def out_arcs(self, s):
        return [(k, w) for k, w in self.arcs.items() if k[0] == s]

    def in_arcs(self, s):
        return [(k, w) for k, w in self.arcs.items() if k[1] == s]

    def normalize(self):
        """graph.py:115-121."""
        for s in list(self.pdf):
            out = self.out_arcs(s)
            tot = 0.
            for _, w in out:
                tot += w
            for k, w in out:
                self.arcs[k] = w / tot

    def replace_state(self, old, g):
        """graph.py:123-156."""
        new = {s: self.add_state(pdf_id=g.pdf[s]) for s in g.pdf}
        for (a, b), w in g.arcs.items():
            self.add_arc(new[a], new[b], w)
        to_del, new_arcs = [], []
        for (a, b), w in self.out_arcs(old):
            to_del.append((a, b))
            new_arcs.append((new[g.end_state], b, w))
        for (a, b), w in self.in_arcs(old):
            to_del.append((a, b))
            new_arcs.append((a, new[g.start_state], w))
        for a, b, w in new_arcs:
            self.add_arc(a, b, w)
        for k in to_del:
            self.arcs.pop(k, None)
        del self.pdf[old]

    def _next(self, start, w0):
        """find_next_pdf_ids (graph.py:158-170)."""
        todo = [(k, w, w0) for k, w in self.out_arcs(start)]
        seen = {start}
        while todo:
            (a, b), aw, w = todo.pop()
            if self.pdf[b] is not None:
                yield b, w * aw
            elif b not in seen:
                todo += [(k, kw, aw * w) for k, kw in self.out_arcs(b)]
                seen.add(b)

    def _prev(self, start, w0):
        """find_previous_pdf_ids (graph.py:172-184)."""
        todo = [(k, w, w0) for k, w in self.in_arcs(start)]
        seen = {start}
        while todo:
            (a, b), aw, w = todo.pop()
            if self.pdf[a] is not None:
                yield a, w * aw
            elif a not in seen:
                todo += [(k, kw, aw * w) for k, kw in self.in_arcs(a)]
                seen.add(a)

    def compile(self, dtype=np.float32):
        """graph.py:185-240.  The reference builds float32 tensors."""
        idx, mapping = {}, []
        for s, p in self.pdf.items():
            if p is not None:
                idx[s] = len(mapping)
                mapping.append(p)
        K = len(mapping)
        init = np.zeros(K, dtype=dtype)
        final = np.zeros(K, dtype=dtype)
        trans = np.zeros((K, K), dtype=dtype)
        for s, w in self._next(self.start_state, 1.0):
            init[idx[s]] += dtype(w)
        init /= init.sum()
        for s, w in self._prev(self.end_state, 1.0):
            final[idx[s]] += dtype(w)
        final /= final.sum()
        for (a, b), w in self.arcs.items():
            if self.pdf[a] is None:
                continue
            if self.pdf[b] is None:
                for s, ww in self._next(b, w):
                    trans[idx[a], idx[s]] += dtype(ww)
            else:
                trans[idx[a], idx[b]] += dtype(w)
        for k in range(K):
            diag = trans[k, k]
            off = trans[k].sum() - diag
            if diag > 0. and off > 0:
                trans[k] /= off / (1 - diag)
                trans[k, k] = diag
        with np.errstate(divide='ignore'):
            return np.log(init), np.log(final), np.log(trans), mapping


# --------------------------------------------------------------------------
# model-level E-step / accumulate / ELBO / M-step
# --------------------------------------------------------------------------

def emission_llh(X, ng_post, dir_post=None):
    """Per-pdf expected log-likelihood of a NormalSet (dir_post None) or a
    MixtureSet over a NormalSet (normalset.py:117-119, mixtureset.py:85-98).
    Returns (pdf_llh (T,Kp), comp_resps (T,Kp,C) or None)."""
    stats = normal_diag_sufficient_statistics(X)
    ets = normalgamma_expected_sufficient_statistics(*ng_post)
    pc = normal_diag_llh(stats, ets, X.shape[1])
    if dir_post is None:
        return pc, None
    return mixtureset_llh(pc, categorical_log_weights(dir_post))


def hmm_estep(X, ng_post, dir_post, graph, scale=1., viterbi=False,
              state_path=None, trans_posteriors=False):
    """HMM.expected_log_likelihood + HMM.accumulate (hmm.py:73-100) with the
    DynamicallyOrderedModelSet gather/scatter (modelset.py:140-154) and the
    MixtureSet / NormalSet accumulation (mixtureset.py:100-112,
    normalset.py:121-123, categoricalset.py:54-55).

    graph = (init_log, final_log, trans_log, pdf_id_mapping).
    Returns dict(exp_llh (T,), gamma (T,K), xi or None, acc_normal (M,Q),
    acc_dirichlet (Kp,C) or None, pdf_llh)."""
    init_log, final_log, trans_log, mapping = graph
    mapping = np.asarray(mapping, dtype=np.int64)
    pdf_llh, comp_resps = emission_llh(X, ng_post, dir_post)
    pc = scale * pdf_llh[:, mapping]
    xi = None
    if viterbi or state_path is not None:
        path = best_path(pc, init_log, final_log, trans_log) \
            if state_path is None else np.asarray(state_path)
        gamma = onehot(path, len(mapping), dtype=pc.dtype)
    else:
        res, _ = posteriors(pc, init_log, final_log, trans_log, trans_posteriors)
        gamma, xi = res if trans_posteriors else (res, None)
    exp_llh = (pc * gamma).sum(axis=-1)
    # accumulate
    g_pdf = np.zeros((len(X), pdf_llh.shape[1]), dtype=pc.dtype)
    for i in range(len(mapping)):          # modelset.py:152-153
        g_pdf[:, mapping[i]] += scale * gamma[:, i]
    stats = normal_diag_sufficient_statistics(X)
    if comp_resps is None:
        acc_normal = g_pdf.T @ stats
        acc_dir = None
    else:
        joint = comp_resps * g_pdf[:, :, None]
        acc_normal = joint.reshape(len(X), -1).T @ stats
        acc_dir = categorical_sufficient_statistics(joint).sum(axis=0)
    return dict(exp_llh=exp_llh, gamma=gamma, xi=xi, acc_normal=acc_normal,
                acc_dirichlet=acc_dir, pdf_llh=pdf_llh, g_pdf=g_pdf)


def phoneloop_counts(gamma, xi, start_idxs, end_idxs):
    """PhoneLoop.accumulate (phoneloop.py:83-101): Categorical stats over units."""
    tr = xi.sum(axis=0)
    ph = tr[:, start_idxs][end_idxs, :].sum(axis=0) + gamma[0][start_idxs]
    return categorical_sufficient_statistics(ph[None, :]).sum(axis=0)


def phoneloop_update_graph(trans_log, conc, start_idxs, end_idxs):
    """PhoneLoop._on_weights_update (phoneloop.py:53-65), in place."""
    logw = categorical_log_weights(conc).astype(trans_log.dtype)
    for e in end_idxs:
        loop = np.exp(trans_log[e, e])
        trans_log[e, start_idxs] = np.log(1 - loop) + logw
    return trans_log


def bigram_counts(xi, start_idxs, end_idxs):
    """BigramPhoneLoop.accumulate (phoneloop.py:175-186): the (T-1, ends, starts) block of the
    transition posteriors -> CategoricalSet statistics, summed over time (categoricalset.py:54-55)."""
    block = xi[:, :, start_idxs][:, end_idxs, :]
    return categorical_sufficient_statistics(block).sum(axis=0)


def bigram_update_graph(trans_log, conc, start_idxs, end_idxs):
    """BigramPhoneLoop._on_weights_update (phoneloop.py:145-157), in place.  Reference quirk kept:
    `expected_log_likelihood(eye(P))` is indexed [class, model], and the reference writes row i of it
    onto the arcs out of unit i's end state, i.e. arc (end_i -> start_m) gets E[ln pi_m(i)] -- the
    transpose of the matrix the statistics of `bigram_counts` (model = end unit, class = start unit)
    are accumulated for."""
    logw = np.stack([categorical_log_weights(c) for c in conc]).astype(trans_log.dtype)   # [model, class]
    for i, e in enumerate(end_idxs):
        loop = np.exp(trans_log[e, e])
        trans_log[e, start_idxs] = np.log(1 - loop) + logw[:, i]
    return trans_log


def sb_log_weights(conc, ordering):
    """SBCategorical.expected_log_likelihood on eye(K) (categorical.py:121-165): E[ln v_k] + sum_{j<k} E[ln(1 - v_j)]
    along the current ordering of the sticks, returned in the original index order."""
    c = conc[ordering]
    s = digamma(c.sum(axis=-1))
    log_v, log_1_v = digamma(c[:, 0]) - s, digamma(c[:, 1]) - s
    lp = log_v.copy()
    lp[1:] += np.cumsum(log_1_v[:-1])
    out = np.empty_like(lp)
    out[ordering] = lp
    return out


def sb_transform_stats(counts):
    """SBCategorical._transform_stats (categorical.py:107-116): sticks re-ordered by decreasing count, Beta
    statistics [n_k, sum_{j>k} n_j] with the Dirichlet convention (last column += the others).
    Returns (stats [K, 2] in the original index order, new ordering)."""
    ordering = np.argsort(-counts, kind='stable')
    s = counts[ordering]
    s2 = np.zeros_like(s)
    s2[:-1] = s[1:]
    s2 = np.cumsum(s2[::-1])[::-1]
    new = np.stack([s, s2 + s], axis=-1)
    out = np.empty_like(new)
    out[ordering] = new
    return out, ordering


def sb_hyper_update(sb_post, ordering, conc_prior):
    """SBCategoricalHyperPrior._on_stickbreaking_update (categorical.py:200-209): Gamma posterior of the
    concentration from E[ln(1 - v_k)] of the updated sticks (natural-gradient step with lrate 1 = prior + statistics,
    gamma.py:148-153).  conc_prior = (shape, rate); returns the posterior (shape, rate); its mean shape / rate is the
    second concentration of every stick's prior from then on (categorical.py:196-198)."""
    c = sb_post[ordering]
    log_1_v = digamma(c[:, 1]) - digamma(c.sum(axis=-1))
    shape0, rate0 = conc_prior
    return shape0 + len(c), rate0 - log_1_v.sum()


def sb_phoneloop_update_graph(trans_log, conc, ordering, start_idxs, end_idxs):
    """PhoneLoop._on_weights_update with stick-breaking weights (phoneloop.py:53-65)."""
    logw = sb_log_weights(conc, ordering).astype(trans_log.dtype)
    for e in end_idxs:
        loop = np.exp(trans_log[e, e])
        trans_log[e, start_idxs] = np.log(1 - loop) + logw
    return trans_log


def gmm_estep(X, ng_post, dir_post, labels=None):
    """Mixture E-step + accumulate (mixture.py:70-102)."""
    stats = normal_diag_sufficient_statistics(X)
    ets = normalgamma_expected_sufficient_statistics(*ng_post)
    pc = normal_diag_llh(stats, ets, X.shape[1])
    exp_llh, resps = mixture_expected_llh(pc, categorical_log_weights(dir_post),
                                          labels)
    acc_dir = categorical_sufficient_statistics(resps).sum(axis=0)
    return dict(exp_llh=exp_llh, resps=resps, acc_normal=resps.T @ stats,
                acc_dirichlet=acc_dir)


def elbo_value(exp_llh, kl, datasize):
    """objectives.py:176-184."""
    n = len(exp_llh)
    if datasize <= 0:
        datasize = n
    return float(datasize / float(n)) * exp_llh.sum() - kl


def natural_grad_update_normalgamma(prior, post, stats, lrate):
    """parameters.py:134-141 + normalgamma.py:76-94."""
    ep = normalgamma_natural_parameters(*prior)
    eq = normalgamma_natural_parameters(*post)
    return normalgamma_from_natural_parameters(eq + lrate * (ep + stats - eq))


def natural_grad_update_dirichlet(prior, post, stats, lrate):
    """parameters.py:134-141 + dirichlet.py:70-81."""
    ep = dirichlet_natural_parameters(prior)
    eq = dirichlet_natural_parameters(post)
    return dirichlet_from_natural_parameters(eq + lrate * (ep + stats - eq))


# --------------------------------------------------------------------------
# whole VB iteration over a list of utterances (hmm/accumulate.py + update.py)
# --------------------------------------------------------------------------

def vb_iteration_hmm(utts, ng_prior, ng_post, dir_prior, dir_post, graph,
                     datasize=None, scale=1., lrate=1., graphs=None, viterbi=False):
    """One data-parallel VB-EM iteration, the way ``beer hmm accumulate`` /
    ``beer hmm update`` compose it (accumulate.py:37-63, update.py:37-62,
    objectives.py:78-107): per-utterance ELBO objects are summed (so the
    global KL is subtracted once per utterance), the statistics are rescaled
    by datasize / sum(T_u) and every parameter takes a natural-gradient step.

    Returns (elbo_sum, new_ng_post, new_dir_post, info)."""
    if datasize is None:
        datasize = sum(len(u) for u in utts)
    kl = normalgamma_kl(ng_post, ng_prior).sum()
    if dir_post is not None:
        kl = kl + dirichlet_kl(dir_post, dir_prior).sum()
    total, frames = 0., 0
    acc_n, acc_d = 0., 0.
    for i, X in enumerate(utts):
        g = graph if graphs is None else graphs[i]
        r = hmm_estep(X, ng_post, dir_post, g, scale=scale, viterbi=viterbi)
        total += elbo_value(r['exp_llh'], kl, datasize)
        acc_n = acc_n + r['acc_normal']
        if dir_post is not None:
            acc_d = acc_d + r['acc_dirichlet']
        frames += len(X)
    s = datasize / frames
    new_ng = natural_grad_update_normalgamma(ng_prior, ng_post, s * acc_n, lrate)
    new_dir = None
    if dir_post is not None:
        new_dir = natural_grad_update_dirichlet(dir_prior, dir_post, s * acc_d, lrate)
    return total, new_ng, new_dir, dict(kl=kl, frames=frames, acc_normal=acc_n,
                                        acc_dirichlet=acc_d)


# --------------------------------------------------------------------------
# synthetic workload shared by tests and bench
# --------------------------------------------------------------------------

def phone_loop_graph(n_units, n_states_per_unit=4, self_loop=0.75):
    """Phone-loop decoding graph: start/end/pivot states, every unit a
    left-to-right HMM (recipes/aud/conf/hmm.yml:37-44 topology), built through
    replace_state / normalize / compile exactly as mkphoneloopgraph.py:28-77 +
    mkdecodegraph.py:50-58 do.  Returns (graph tuple, start_idxs, end_idxs)."""
    g = OracleGraph()
    g.start_state = g.add_state()
    g.end_state = g.add_state()
    pivot = g.add_state()
    unit_states = [g.add_state() for _ in range(n_units)]
    g.add_arc(g.start_state, pivot)
    g.add_arc(pivot, g.end_state)
    for s in unit_states:
        g.add_arc(pivot, s)
        g.add_arc(s, pivot)
    g.normalize()
    pdf = 0
    starts, ends = [], []
    for s in unit_states:
        u = OracleGraph()
        sts = [u.add_state(pdf_id=None)]
        for _ in range(n_states_per_unit):
            sts.append(u.add_state(pdf_id=pdf))
            pdf += 1
        sts.append(u.add_state(pdf_id=None))
        u.start_state, u.end_state = sts[0], sts[-1]
        u.add_arc(sts[0], sts[1], 1.0)
        for a in range(1, n_states_per_unit + 1):
            u.add_arc(sts[a], sts[a], self_loop)
            u.add_arc(sts[a], sts[a + 1], 1 - self_loop)
        starts.append(pdf - n_states_per_unit)
        ends.append(pdf - 1)
        g.replace_state(s, u)
    g.normalize()
    return g.compile(), starts, ends


def sample_utterances(rng, graph, means, n_utts, n_frames, noise=1.0):
    """Sample state paths from the graph and emit x_t = mu_{s_t} + eps
    (SURVEY.md section 8d synthetic inputs)."""
    init_log, final_log, trans_log, mapping = graph
    K = len(mapping)
    with np.errstate(under='ignore'):
        init = np.exp(init_log.astype(np.float64))
        A = np.exp(trans_log.astype(np.float64))
    init /= init.sum()
    A /= A.sum(axis=1, keepdims=True)
    cum = np.cumsum(A, axis=1)
    utts = []
    for _ in range(n_utts):
        T = n_frames if np.isscalar(n_frames) else int(rng.choice(n_frames))
        s = np.empty(T, dtype=np.int64)
        s[0] = rng.choice(K, p=init)
        u = rng.random(T)
        for t in range(1, T):
            s[t] = min(np.searchsorted(cum[s[t - 1]], u[t]), K - 1)
        pdf = np.asarray(mapping)[s]
        utts.append((means[pdf] + noise * rng.standard_normal((T, means.shape[1])))
                    .astype(np.float32))
    return utts


# --------------------------------------------------------------------------
# fbank front-end (beer/features.py) -- "next" row of SURVEY section 8(f)
# --------------------------------------------------------------------------

def hz2mel(freq_hz):
    """features.py:9-11."""
    return 1127 * np.log(1 + freq_hz / 700.0)


def mel2hz(mel):
    """features.py:14-16."""
    return 700.0 * (np.exp(mel / 1127.0) - 1)


def _triangle(center, start, end, freqs):
    """Triangular filter sampled on the FFT bins (features.py:30-43): both slopes are linspaces
    between the first and the last bin inside [start, center] resp. [center, end]."""
    out = np.zeros(len(freqs))
    up = (freqs >= start) & (freqs <= center)
    if up.any():
        f = freqs[up]
        out[up] = np.linspace((f[0] - start) / (center - start), (f[-1] - start) / (center - start), len(f))
    down = (freqs >= center) & (freqs <= end)
    if down.any():
        f = freqs[down]
        out[down] = np.linspace((end - f[0]) / (end - center), (end - f[-1]) / (end - center), len(f))
    return out


def create_fbank(nfilters, fft_len=512, srate=16000, lowfreq=0, highfreq=None):
    """Mel filterbank with the filter centres aligned to FFT bins (features.py:47-79)."""
    highfreq = highfreq or srate / 2
    centers = np.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilters + 2)
    centers = np.floor(fft_len * mel2hz(centers) / srate)
    bins = np.arange(0, fft_len // 2)
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.stack([_triangle(centers[i], centers[i - 1], centers[i + 1], bins)
                         for i in range(1, nfilters + 1)])


def fbank(signal, flen=0.025, frate=0.01, hifreq=8000, lowfreq=20, nfilters=26, preemph=0.97, srate=16000):
    """log(1 + mel filterbank energies of |rFFT| of the pre-emphasised, Hamming-windowed frames)
    (features.py:145-204; pre-emphasis on the whole signal in fp32, first sample against itself)."""
    frate_samp, flen_samp = int(srate * frate), int(srate * flen)
    nframes = (len(signal) - flen_samp) // frate_samp + 1
    s_t = np.array(signal, dtype=np.float32)
    s_t -= np.float32(preemph) * np.r_[s_t[0], s_t[:-1]]
    idx = np.arange(nframes)[:, None] * frate_samp + np.arange(flen_samp)[None, :]
    frames = s_t[idx] * np.hamming(flen_samp)[None, :]
    fft_len = int(2 ** np.floor(np.log2(flen_samp) + 1))
    magspec = np.abs(np.fft.rfft(frames, n=fft_len, axis=-1)[:, :-1])
    filters = create_fbank(nfilters, fft_len, srate=srate, lowfreq=lowfreq, highfreq=hifreq)
    return np.log(magspec @ filters.T + 1)


def short_term_mspec(signal, flen=0.025, frate=0.01, preemph=0.97, srate=16000):
    """Magnitude spectrum of `beer features extract` (features.py:102-143): mean removed, pre-emphasis inside each
    frame (first sample against itself), Hamming window, |rFFT| without the Nyquist bin.  Returns (mspec, fft_len)."""
    signal = np.asarray(signal) - np.asarray(signal).mean()
    frate_samp, flen_samp = int(srate * frate), int(srate * flen)
    nframes = (len(signal) - flen_samp) // frate_samp + 1
    idx = np.arange(nframes)[:, None] * frate_samp + np.arange(flen_samp)[None, :]
    frames = signal[idx].copy()
    frames -= preemph * np.c_[frames[:, 0], frames[:, :-1]]
    frames = frames * np.hamming(flen_samp)[None, :]
    fft_len = int(2 ** np.floor(np.log2(flen_samp) + 1))
    return np.abs(np.fft.rfft(frames, n=fft_len, axis=-1)[:, :-1]), fft_len


def log_mel_spectrum(signal, nfilters=40, lowfreq=20, hifreq=8000, **kw):
    """extract.py:107-127 for an fbank configuration: log(1e-6 + mspec @ filters.T)."""
    mspec, fft_len = short_term_mspec(signal, **kw)
    return np.log(1e-6 + mspec @ create_fbank(nfilters, fft_len, lowfreq=lowfreq, highfreq=hifreq).T)


def add_deltas(fea, winlens=(2, 2)):
    """Append delta / delta-delta features: regression filter over +-wlen frames with the edges
    replicated (features.py:82-100)."""
    out = [fea]
    for wlen in winlens:
        k = np.arange(-wlen, wlen + 1)
        coef = k / (2.0 * (k @ k))
        pad = np.r_[fea[[0]].repeat(wlen, 0), fea, fea[[-1]].repeat(wlen, 0)]
        fea = sum(coef[i] * pad[i:i + len(fea)] for i in range(2 * wlen + 1))
        out.append(fea)
    return np.hstack(out)
