"""Model classes of the VB-EM hot path with the call surface of beer/models/ (the `Model`
interface that `evidence_lower_bound` drives, beer/models/basemodel.py:9-178):

    NormalSet, Categorical, CategoricalSet, Mixture, MixtureSet, JointModelSet,
    DynamicallyOrderedModelSet, HMM

Python here is host-side bookkeeping (which buffers, which order); every per-frame computation
runs in the sm_100a kernels of libbeer_b200.so:

    KA  emission llh + mixture log-sum-exp      ops.emission_llh / emission_llh_tc
    KB  forward-backward over a CompiledGraph    ops.hmm_forward_backward
    KV  Viterbi                                  ops.hmm_viterbi
    KC  posterior-weighted statistics            ops.accumulate_stats
    KM  expected statistics / KL / M-step        ops.normalgamma_* / ops.dirichlet_*

There is no CPU or PyTorch fallback: models must live on a CUDA device.
"""
import abc

import numpy as np
import torch

from . import ops
from .dists import Dirichlet, Gamma, NormalGamma, frames_of, frames_with_grad
from .engine import Utterances
from .parameters import ConjugateBayesianParameter

__all__ = ['Model', 'DiscreteLatentModel', 'ModelSet', 'NormalSet', 'Categorical', 'SBCategorical', 'SBCategoricalHyperPrior', 'CategoricalSet', 'Mixture',
           'MixtureSet', 'JointModelSet', 'DynamicallyOrderedModelSet', 'HMM', 'PhoneLoop', 'BigramPhoneLoop', 'UnknownCovarianceType']

f32, f64, i32, i64 = torch.float32, torch.float64, torch.int32, torch.int64


class UnknownCovarianceType(Exception):
    """beer/models/normal.py:15."""


# ---------------------------------------------------------------------------------------------
# base classes (beer/models/basemodel.py, modelset.py:9-38)
# ---------------------------------------------------------------------------------------------

class Model(torch.nn.Module, metaclass=abc.ABCMeta):
    def __init__(self):
        super().__init__()
        self._cache = {}

    @property
    def cache(self):
        """Intermediate results kept between expected_log_likelihood and accumulate."""
        return self._cache

    def clear_cache(self):
        self._cache = {}
        for module in self.modules():
            if module is not self and isinstance(module, Model):
                module.clear_cache()

    def bayesian_parameters(self, paramtype=None, paramfilter=None, keepgroups=False):
        def _select(group):
            for param in group:
                if paramtype is None or type(param) == paramtype:
                    if paramfilter is None or paramfilter(param):
                        yield param

        for group in self.mean_field_factorization():
            if not keepgroups:
                yield from _select(group)
            else:
                group = list(_select(group))
                if group:
                    yield group

    def conjugate_bayesian_parameters(self, keepgroups=False):
        return self.bayesian_parameters(paramtype=ConjugateBayesianParameter, keepgroups=keepgroups)

    def kl_div_posterior_prior(self):
        """Sum over all parameters of KL(posterior || prior): fp64 device scalar (basemodel.py:64-77)."""
        total = None
        for param in self.bayesian_parameters():
            total = param.posterior._kl(param.prior, out=total)
        if total is None:
            return torch.zeros((), dtype=f64)
        return total[0]

    @abc.abstractmethod
    def accumulate(self, s_stats, parent_msg=None):
        pass

    @abc.abstractmethod
    def expected_log_likelihood(self, s_stats, **kwargs):
        pass

    @abc.abstractmethod
    def mean_field_factorization(self):
        pass

    @abc.abstractmethod
    def sufficient_statistics(self, data):
        pass


class DiscreteLatentModel(Model, metaclass=abc.ABCMeta):
    def __init__(self, modelset):
        super().__init__()
        self.modelset = modelset

    @abc.abstractmethod
    def posteriors(self, data, **kwargs):
        pass


class ModelSet(Model, metaclass=abc.ABCMeta):
    @abc.abstractmethod
    def __getitem__(self, key):
        pass

    @abc.abstractmethod
    def __len__(self):
        pass


def _merge_groups(l1, l2):
    """Zip two mean-field factorisations, padding the shorter one (mixture.py:52-60)."""
    l1, l2 = [list(g) for g in l1], [list(g) for g in l2]
    n = max(len(l1), len(l2))
    l1 += [[] for _ in range(n - len(l1))]
    l2 += [[] for _ in range(n - len(l2))]
    return [u + v for u, v in zip(l1, l2)]


def _stats_tensor(data):
    """sufficient_statistics of frames or of a ragged batch of utterances."""
    from .dists import NormalDiagonalLikelihood
    if isinstance(data, Utterances):
        stats = NormalDiagonalLikelihood.sufficient_statistics(data.X)
        stats._beer_utts = data
        return stats
    return NormalDiagonalLikelihood.sufficient_statistics(data)


def _offsets_of(stats, n_frames, device):
    utts = getattr(stats, '_beer_utts', None)
    if utts is not None:
        return utts.offsets, utts
    return torch.tensor([0, n_frames], dtype=i64, device=device), None


# ---------------------------------------------------------------------------------------------
# emission bundle: the flat device view of a tree of model sets
# ---------------------------------------------------------------------------------------------

class _Leaf:
    __slots__ = ('normal', 'weights', 'n_pdfs', 'n_comp', 'g0', 'k0')

    def __init__(self, normal, weights, n_pdfs, n_comp):
        self.normal, self.weights, self.n_pdfs, self.n_comp = normal, weights, n_pdfs, n_comp


def _leaves(modelset):
    """[(Normal-Gamma parameter, Dirichlet parameter or None, #pdfs, Gaussians per pdf)] in pdf
    order: what JointModelSet / MixtureSet / NormalSet concatenate (modelset.py:71-75,
    mixtureset.py:85-98)."""
    if isinstance(modelset, DynamicallyOrderedModelSet):
        return _leaves(modelset.original_modelset)
    if isinstance(modelset, JointModelSet):
        return [leaf for ms in modelset.modelsets for leaf in _leaves(ms)]
    if isinstance(modelset, MixtureSet):
        if not isinstance(modelset.modelset, NormalSet):
            raise NotImplementedError('MixtureSet components must be a NormalSet')
        return [_Leaf(modelset.modelset.means_precisions, modelset.categoricalset.weights, len(modelset),
                      modelset.n_comp_per_mixture)]
    if isinstance(modelset, NormalSet):
        return [_Leaf(modelset.means_precisions, None, len(modelset), 1)]
    raise NotImplementedError(f'no B200 emission kernel for model sets of type {type(modelset).__name__}')


class _Emission:
    """Flat view of the Gaussians of a model-set tree + the kernels that evaluate them."""

    def __init__(self, leaves, single_pdf=False):
        self.leaves = leaves
        g0 = k0 = 0
        counts = []
        for leaf in leaves:
            leaf.g0, leaf.k0 = g0, k0
            g0 += leaf.n_pdfs * leaf.n_comp
            k0 += leaf.n_pdfs
            counts += [leaf.n_comp] * leaf.n_pdfs
        self.M = g0
        self.Kp = 1 if single_pdf else k0
        if single_pdf:
            counts = [self.M]
        self.comp_off_host = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        post0 = leaves[0].normal.posterior.params
        self.D = post0.mean.shape[1]
        self.device = post0.mean.device
        self.has_mixtures = self.Kp != self.M
        self.uniform_C = int(counts[0]) if len(set(counts)) == 1 else 0
        self.comp_off = (torch.as_tensor(self.comp_off_host, dtype=i32, device=self.device)
                         if self.has_mixtures else None)

    def _cat(self, which):
        parts = [getattr(leaf.normal, which).params.as_tuple() for leaf in self.leaves]
        if len(parts) == 1:
            return parts[0]
        return tuple(torch.cat([p[i] for p in parts]).contiguous() for i in range(4))

    def log_weights(self, extra=None):
        """E[ln pi] of every Gaussian ([M] or None when no set has mixture weights)."""
        if extra is None and all(leaf.weights is None for leaf in self.leaves):
            return None
        logw = torch.zeros(self.M, device=self.device, dtype=f32)
        for leaf in self.leaves:
            if leaf.weights is not None:
                n = leaf.n_pdfs * leaf.n_comp
                logw[leaf.g0:leaf.g0 + n] = leaf.weights.posterior.expected_log_weights().reshape(-1)
        if extra is not None:
            logw += extra
        return logw

    def llh(self, X, extra_logw=None, want_comp=None):
        """-> (pdf_llh [N,Kp], comp_llh [N,M] or None, frame_ref [N]), offset form."""
        post = self._cat('posterior')
        W, bias, ref = ops.emission_prepare(*post, logw=self.log_weights(extra_logw))
        want_comp = self.has_mixtures if want_comp is None else want_comp
        C = self.uniform_C
        if self.has_mixtures and np.diff(self.comp_off_host).max() > 128:
            # a pdf wider than one emission tile: per-Gaussian llhs, then a segmented log-sum-exp
            comp, _, fref = ops.emission_llh(X, W, bias, ref, comp_off=None, Kp=self.M)
            pdf = ops.segment_logsumexp(comp, comp_off=self.comp_off)
            return pdf, comp, fref
        if C and ops.emission_tc_supported(self.M, self.D, C) and X.data_ptr() % 16 == 0:
            img = ops.emission_tc_pack(W, bias, C)
            return ops.emission_llh_tc(X, img, ref, self.M, C, want_comp=want_comp)
        return ops.emission_llh(X, W, bias, ref, comp_off=self.comp_off, Kp=self.Kp, want_comp=want_comp)

    def accumulate(self, X, pdf_post, pdf_llh, comp_llh):
        """{parameter: accumulated statistics} of every leaf (fp64)."""
        acc = torch.zeros(self.M, 2 * self.D + 2, device=self.device, dtype=f64)
        if comp_llh is None:
            ops.accumulate_stats(X, acc, pdf_post=pdf_post, Kp=self.M)
        else:
            ops.accumulate_stats(X, acc, pdf_post=pdf_post, pdf_llh=pdf_llh, comp_llh=comp_llh,
                                 comp_off=self.comp_off, Kp=self.Kp)
        out = {}
        wst = None
        for leaf in self.leaves:
            n = leaf.n_pdfs * leaf.n_comp
            out[leaf.normal] = acc[leaf.g0:leaf.g0 + n]
            if leaf.weights is not None:
                if wst is None:
                    wst = ops.mixture_weight_stats(acc, self.D, comp_off=self.comp_off, Kp=self.Kp)
                shape = leaf.weights.posterior.params.concentrations.shape
                out[leaf.weights] = wst[leaf.g0:leaf.g0 + n].reshape(shape)
        return out


# ---------------------------------------------------------------------------------------------
# NormalSet (beer/models/normalset.py)
# ---------------------------------------------------------------------------------------------

class NormalSet(ModelSet):
    """Set of Normal densities with Normal-Gamma priors (diagonal covariance)."""

    @classmethod
    def create(cls, mean, cov, size, prior_strength=1, noise_std=1., cov_type='full', shared_cov=False):
        """normalset.py:84-102.  Only `cov_type='diagonal'` has B200 kernels (every BASELINE config
        is diagonal); 'full' / 'isotropic' raise NotImplementedError, anything else
        UnknownCovarianceType."""
        if shared_cov:
            import warnings
            warnings.warn('The "NormalSet" with shared covariance is not supported anymore. The argument '
                          'will be ignored.', DeprecationWarning, stacklevel=2)
        if cov_type not in ('full', 'diagonal', 'isotropic'):
            raise UnknownCovarianceType(f'Unknown covariance type: "{cov_type}"')
        if cov_type != 'diagonal':
            raise NotImplementedError(f'cov_type="{cov_type}" is outside the B200 hot path (diagonal only)')
        mean = mean.detach().to(f32)
        cov = cov.detach().to(f32)
        if cov.dim() == 0:
            cov = cov * torch.ones(len(mean), dtype=f32, device=mean.device)
        var = cov.diag() if cov.dim() == 2 else cov
        means = mean.repeat(size, 1)
        noise = torch.randn(size, len(mean), dtype=f32, device=mean.device) * noise_std * var.sqrt()[None, :]
        scale = torch.full((size, 1), float(prior_strength), dtype=f32, device=mean.device)
        shape = torch.full((size, 1), float(prior_strength), dtype=f32, device=mean.device)
        rates = (prior_strength * var).repeat(size, 1)
        prior = NormalGamma.from_std_parameters(means, scale, shape, rates)
        posterior = NormalGamma.from_std_parameters(means + noise, scale.clone(), shape.clone(), rates.clone())
        return cls(ConjugateBayesianParameter(prior, posterior))

    def __init__(self, means_precisions):
        super().__init__()
        self.means_precisions = means_precisions

    @property
    def dim(self):
        return self.means_precisions.posterior.params.mean.shape[1]

    def sufficient_statistics(self, data):
        return _stats_tensor(data)

    def mean_field_factorization(self):
        return [[self.means_precisions]]

    def expected_log_likelihood(self, stats):
        """Per-Gaussian expected log-likelihood [N, M] (normalset.py:117-119)."""
        X = frames_of(stats, self.dim)
        pdf, _, fref = _Emission(_leaves(self)).llh(X)
        return pdf + fref[:, None]

    def accumulate(self, stats, resps):
        """resps.T @ stats (normalset.py:121-123) -> {means_precisions: [M, 2D+2]}."""
        X = frames_of(stats, self.dim)
        resps = resps.detach().to(f32).contiguous()
        return _Emission(_leaves(self)).accumulate(X, resps, None, None)

    def __len__(self):
        return len(self.means_precisions)

    def __getitem__(self, key):
        raise NotImplementedError('slicing a NormalSet is not part of the B200 hot path')


# ---------------------------------------------------------------------------------------------
# Categorical / CategoricalSet (beer/models/categorical.py:39-79, categoricalset.py:15-66)
# ---------------------------------------------------------------------------------------------

def _dirichlet_param(weights, prior_strength):
    conc = weights.detach().to(f32) * prior_strength
    return ConjugateBayesianParameter(Dirichlet.from_std_parameters(conc.clone()),
                                      Dirichlet.from_std_parameters(conc.clone()))


class Categorical(Model):
    @classmethod
    def create(cls, weights, prior_strength=1.):
        return cls(_dirichlet_param(weights, prior_strength))

    def __init__(self, weights):
        super().__init__()
        self.weights = weights

    @property
    def mean(self):
        return self.weights.value()

    def sufficient_statistics(self, data):
        return self.weights.likelihood_fn.sufficient_statistics(data)

    def mean_field_factorization(self):
        return [[self.weights]]

    def expected_log_likelihood(self, stats):
        return self.weights.likelihood_fn(self.weights.natural_form(), stats)

    def accumulate(self, stats, parent_msg=None):
        return {self.weights: stats.sum(dim=0).to(f64)}

    def expected_log_weights(self):
        """E[ln pi]: what evaluating the model on eye(C) yields (phoneloop.py:53-59)."""
        return self.weights.posterior.expected_log_weights()


class SBCategorical(Model):
    """Categorical with a truncated stick-breaking prior (categorical.py:82-165): one Beta (a two-category Dirichlet
    row) per stick.  The counts a model accumulates for it are re-ordered by decreasing size and turned into the
    Beta statistics [n_k, sum_{j>k} n_j] by a callback right before the update (categorical.py:107-116)."""

    @classmethod
    def create(cls, truncation, prior_strength=1., device=None):
        params = torch.ones(truncation, 2, dtype=f32, device='cuda' if device is None else device)
        params[:, 1] = prior_strength
        return cls(ConjugateBayesianParameter(Dirichlet.from_std_parameters(params),
                                              Dirichlet.from_std_parameters(params.clone())))

    def __init__(self, stickbreaking):
        super().__init__()
        self.stickbreaking = stickbreaking
        conc = self.stickbreaking.posterior.params.concentrations
        self.ordering = torch.arange(conc.shape[0], device=conc.device)
        self.stickbreaking.register_callback(self._transform_stats, notify_before_update=True)

    def _transform_stats(self):
        stats = self.stickbreaking.stats
        self.ordering = stats.sort(descending=True, stable=True)[1]      # (ties as the CPU sort of the reference leaves them)
        stats = stats[self.ordering]
        s2 = torch.zeros_like(stats)
        s2[:-1] = stats[1:]
        s2 = torch.flip(torch.flip(s2, dims=(0,)).cumsum(dim=0), dims=(0,))
        new_stats = torch.cat([stats[:, None], s2[:, None]], dim=-1)
        new_stats[:, -1] += new_stats[:, :-1].sum(dim=-1)
        self.stickbreaking.stats = new_stats[self.reverse_ordering, :]

    def _log_v(self):
        c = self.stickbreaking.posterior.params.concentrations[self.ordering].double()
        s_dig = torch.digamma(c.sum(dim=-1))
        return torch.digamma(c[:, 0]) - s_dig, torch.digamma(c[:, 1]) - s_dig

    def _log_prob(self):
        log_v, log_1_v = self._log_v()
        log_prob = log_v
        log_prob[1:] += log_1_v[:-1].cumsum(dim=0)
        return log_prob, log_1_v

    @property
    def reverse_ordering(self):
        return torch.argsort(self.ordering)

    @property
    def mean(self):
        c = self.stickbreaking.posterior.params.concentrations[self.ordering].double()
        norm = c.sum(dim=-1) + torch.finfo(torch.float64).eps
        weights = c[:, 0] / norm
        residual = (c[:, 1] / norm).cumprod(dim=0)
        weights[1:] *= residual[:-1]
        return weights[self.reverse_ordering].to(f32)

    def sufficient_statistics(self, data):
        return data          # one-hot encodings

    def mean_field_factorization(self):
        return [[self.stickbreaking]]

    def expected_log_weights(self):
        log_prob, _ = self._log_prob()
        return log_prob[self.reverse_ordering].to(f32)

    def expected_log_likelihood(self, stats):
        return stats @ self.expected_log_weights().to(stats.dtype)

    def accumulate(self, stats, parent_msg=None):
        return {self.stickbreaking: stats.sum(dim=0).to(f64)}


class SBCategoricalHyperPrior(SBCategorical):
    """Stick-breaking weights with a Gamma hyper-prior over the concentration of the process
    (categorical.py:168-209; `gamma_dirichlet_process`, the CLI's default unit-weight prior): after every update of
    the sticks the concentration's posterior is re-estimated from E[ln(1 - v_k)], and its mean becomes the second
    concentration of every stick's prior."""

    @classmethod
    def create(cls, truncation, prior_strength=1., hyper_prior_strength=1., device=None):
        dev = 'cuda' if device is None else device
        mean = torch.ones(1, dtype=f32, device=dev) * prior_strength
        shape = torch.ones_like(mean) * hyper_prior_strength
        rate = hyper_prior_strength / mean
        concentration = ConjugateBayesianParameter(Gamma.from_std_parameters(shape, rate),
                                                   Gamma.from_std_parameters(shape.clone(), rate.clone()))
        params = torch.ones(truncation, 2, dtype=f32, device=dev)
        params[:, 1] = prior_strength
        sb = ConjugateBayesianParameter(Dirichlet.from_std_parameters(params),
                                        Dirichlet.from_std_parameters(params.clone()))
        return cls(sb, concentration)

    def __init__(self, stickbreaking, concentration):
        super().__init__(stickbreaking)
        self.concentration = concentration
        self.stickbreaking.register_callback(self._on_stickbreaking_update)
        self.concentration.register_callback(self._on_concentration_update)
        self._on_concentration_update()

    def _on_concentration_update(self):
        self.stickbreaking.prior.params.concentrations[:, 1] = self.concentration.value().to(f32)

    def _on_stickbreaking_update(self):
        _, log_1_v = self._log_prob()
        pad = torch.ones_like(log_1_v)
        sb_stats = torch.cat([log_1_v[:, None], pad[:, None]], dim=-1)
        self.concentration.stats = sb_stats.sum(dim=0)
        self.concentration.natural_grad_update(lrate=1.)


class CategoricalSet(ModelSet):
    @classmethod
    def create(cls, weights, prior_strength=1.):
        return cls(_dirichlet_param(weights, prior_strength))

    def __init__(self, weights):
        super().__init__()
        self.weights = weights

    @property
    def mean(self):
        return self.weights.value()

    def sufficient_statistics(self, data):
        return self.weights.likelihood_fn.sufficient_statistics(data)

    def mean_field_factorization(self):
        return [[self.weights]]

    def expected_log_likelihood(self, stats):
        return self.weights.likelihood_fn(self.weights.natural_form(), stats)

    def accumulate(self, stats, resps):
        return {self.weights: (resps.t() @ stats).to(f64)}

    def accumulate_from_jointresps(self, jointresps_stats):
        return {self.weights: jointresps_stats.sum(dim=0).to(f64)}

    def __len__(self):
        return len(self.weights)

    def __getitem__(self, key):
        raise NotImplementedError('slicing a CategoricalSet is not part of the B200 hot path')


# ---------------------------------------------------------------------------------------------
# Mixture (beer/models/mixture.py)
# ---------------------------------------------------------------------------------------------

class Mixture(DiscreteLatentModel):
    """Bayesian mixture of the components of a NormalSet."""

    @classmethod
    def create(cls, modelset, categorical=None, prior_strength=1.):
        if categorical is None:
            ref = modelset.mean_field_factorization()[0][0].posterior.params.mean
            weights = torch.ones(len(modelset), dtype=f32, device=ref.device) / len(modelset)
            categorical = Categorical.create(weights, prior_strength)
        return cls(categorical, modelset)

    def __init__(self, categorical, modelset):
        super().__init__(modelset)
        self.categorical = categorical

    def _emission(self):
        if not isinstance(self.modelset, NormalSet):
            raise NotImplementedError('Mixture components must be a NormalSet')
        leaf = _Leaf(self.modelset.means_precisions, self.categorical.weights, 1, len(self.modelset))
        return _Emission([leaf], single_pdf=True)

    def mean_field_factorization(self):
        return _merge_groups(self.modelset.mean_field_factorization(), self.categorical.mean_field_factorization())

    def sufficient_statistics(self, data):
        return self.modelset.sufficient_statistics(data)

    def expected_log_likelihood(self, stats, labels=None, **kwargs):
        """Per-frame E[ln p(x, z)] - KL(q(z) || p(z)) = logsumexp_c(llh_c + E ln pi_c); with
        `labels` the responsibilities are one-hot and there is no KL term (mixture.py:70-93)."""
        em = self._emission()
        X = frames_of(stats, em.D).detach()
        pdf, comp, fref = em.llh(X, want_comp=True)
        self.cache.update(X=X, pdf_llh=pdf, comp_llh=comp, emission=em, labels=None)
        Xg = frames_with_grad(stats)
        if labels is None:
            frame = pdf[:, 0] + fref
            if Xg is not None:      # sum_c r_tc grad llh_c(x_t): detached responsibilities (mixture.py:79-93)
                ones = torch.ones(X.shape[0], 1, device=X.device, dtype=f32)
                frame = _attach_frame_grad(Xg, frame, ones, em, comp, pdf)
            return frame
        labels = torch.as_tensor(labels, device=X.device).to(i32).contiguous()
        logw = em.log_weights()
        resps, frame = ops.path_posteriors(labels, em.M, pdf_llh=comp, frame_ref=fref)
        self.cache.update(labels=labels, resps=resps)
        frame = frame - logw[labels.long()]
        if Xg is not None:          # one-hot responsibilities: the gradient of the labelled component's llh
            frame = _attach_frame_grad(Xg, frame, resps, em, None, None)
        return frame

    def accumulate(self, stats, parent_msg=None):
        c = self.cache
        em = c['emission']
        if c['labels'] is None:
            return em.accumulate(c['X'], None, c['pdf_llh'], c['comp_llh'])
        acc = torch.zeros(em.M, 2 * em.D + 2, device=em.device, dtype=f64)
        ops.accumulate_stats(c['X'], acc, pdf_post=c['resps'], Kp=em.M)
        wst = ops.mixture_weight_stats(acc, em.D, comp_off=em.comp_off, Kp=1)
        return {self.modelset.means_precisions: acc,
                self.categorical.weights: wst.reshape(self.categorical.weights.posterior.params.concentrations.shape)}

    def posteriors(self, data):
        """Component responsibilities [N, M] (mixture.py:109-115; that method is broken in the
        reference -- `self.weights` does not exist -- this is what it was meant to return)."""
        em = self._emission()
        X = frames_of(self.sufficient_statistics(data), em.D)
        pdf, comp, _ = em.llh(X, want_comp=True)
        return torch.exp(comp - pdf)


# ---------------------------------------------------------------------------------------------
# MixtureSet / JointModelSet / DynamicallyOrderedModelSet
# ---------------------------------------------------------------------------------------------

class MixtureSet(ModelSet):
    """K mixtures with the same number of components over one NormalSet of K*C Gaussians
    (beer/models/mixtureset.py)."""

    @classmethod
    def create(cls, size, modelset, prior_strength=1.):
        ref = modelset.mean_field_factorization()[0][0].posterior.params.mean
        ncomp = len(modelset) // size
        weights = torch.ones(size, ncomp, dtype=f32, device=ref.device) / ncomp
        return cls(CategoricalSet.create(weights, prior_strength), modelset)

    def __init__(self, categoricalset, modelset):
        super().__init__()
        self.categoricalset = categoricalset
        self.modelset = modelset

    @property
    def n_comp_per_mixture(self):
        return len(self.modelset) // len(self)

    def mean_field_factorization(self):
        return _merge_groups(self.modelset.mean_field_factorization(),
                             self.categoricalset.mean_field_factorization())

    def sufficient_statistics(self, data):
        return self.modelset.sufficient_statistics(data)

    def expected_log_likelihood(self, stats):
        """Per-mixture log-normaliser [N, K]; the component llhs stay cached for accumulate
        (mixtureset.py:85-98)."""
        em = _Emission(_leaves(self))
        X = frames_of(stats, em.D).detach()
        pdf, comp, fref = em.llh(X)
        self.cache.update(X=X, pdf_llh=pdf, comp_llh=comp, emission=em)
        return pdf + fref[:, None]

    def accumulate(self, stats, resps):
        c = self.cache
        resps = resps.detach().to(f32).contiguous()
        return c['emission'].accumulate(c['X'], resps, c['pdf_llh'], c['comp_llh'])

    def __len__(self):
        return len(self.categoricalset)

    def __getitem__(self, key):
        raise NotImplementedError('slicing a MixtureSet is not part of the B200 hot path')


class JointModelSet(ModelSet):
    """Concatenation of model sets sharing one kind of statistics (modelset.py:43-109)."""

    def __init__(self, modelsets):
        super().__init__()
        self.modelsets = torch.nn.ModuleList(modelsets)

    def mean_field_factorization(self):
        groups = []
        for modelset in self.modelsets:
            m_groups = modelset.mean_field_factorization()
            if len(m_groups) > 1:
                raise ValueError('Invalid model set: more than 1 mean field group')
            groups += m_groups[0]
        return [groups]

    def sufficient_statistics(self, data):
        return self.modelsets[0].sufficient_statistics(data)

    def expected_log_likelihood(self, stats):
        em = _Emission(_leaves(self))
        X = frames_of(stats, em.D).detach()
        pdf, comp, fref = em.llh(X)
        self.cache.update(X=X, pdf_llh=pdf, comp_llh=comp, emission=em)
        return pdf + fref[:, None]

    def accumulate(self, stats, resps):
        c = self.cache
        resps = resps.detach().to(f32).contiguous()
        return c['emission'].accumulate(c['X'], resps, c['pdf_llh'], c['comp_llh'])

    def __getitem__(self, key):
        if key < 0:
            raise ValueError('Unsupported negative index')
        total = 0
        for modelset in self.modelsets:
            if key < total + len(modelset):
                return modelset[key - total]
            total += len(modelset)
        raise IndexError('index out of range')

    def __len__(self):
        return sum(len(m) for m in self.modelsets)


class DynamicallyOrderedModelSet(ModelSet):
    """Model set evaluated through an ordering with possibly repeated indices
    (modelset.py:112-164).  Inside an HMM the gather / scatter-add through `pdf_id_mapping` is
    folded into the forward-backward kernel; the two methods below are the standalone form."""

    def __init__(self, original_modelset):
        super().__init__()
        self.original_modelset = original_modelset

    def mean_field_factorization(self):
        return self.original_modelset.mean_field_factorization()

    def sufficient_statistics(self, data):
        return self.original_modelset.sufficient_statistics(data)

    def expected_log_likelihood(self, stats, order=None):
        if order is None:
            order = list(range(len(self.original_modelset)))
        pc_exp_llh = self.original_modelset.expected_log_likelihood(stats)
        self.cache['order'] = order
        return pc_exp_llh[:, torch.as_tensor(order, device=pc_exp_llh.device)]

    def accumulate(self, stats, resps):
        order = torch.as_tensor(self.cache['order'], device=resps.device)
        new_resps = torch.zeros((len(stats), len(self.original_modelset)), dtype=resps.dtype, device=resps.device)
        new_resps.index_add_(1, order, resps)
        return self.original_modelset.accumulate(stats, new_resps)

    def __getitem__(self, key):
        return self.original_modelset[key]

    def __len__(self):
        return len(self.original_modelset)


class _FrameLlhGrad(torch.autograd.Function):
    """Gradient of the per-frame expected log-likelihood w.r.t. the frames with the posteriors held fixed
    (hmm.py:79-87: the inference runs on `pc_llhs.detach()`, the returned value is `(pc_llhs * resps).sum(-1)`;
    mixture.py:76-93: detached responsibilities times attached per-component llhs):
        d/dx_t sum_j w_tj llh_j(x_t) = sum_j w_tj (E[lambda_j mu_j] - x_t E[lambda_j]),
    w = scale * pdf posteriors (x responsibilities inside the pdf for mixtures).  Forward hands back the values the
    kernels computed; backward is ONE tcgen05 kernel (csrc/emission_bwd.cu: w formed on chip as the tensor-memory
    operand of [N, M] x [M, 2D]; w is never stored)."""

    @staticmethod
    def forward(ctx, X, frame, pdf_post, ets, comp_llh, pdf_llh, pdf_of, scale):
        ctx.save_for_backward(X, pdf_post, ets, comp_llh, pdf_llh, pdf_of)
        ctx.scale = scale
        return frame.clone()

    @staticmethod
    def backward(ctx, grad_out):
        X, post, ets, comp, pdf, pdf_of = ctx.saved_tensors
        g = ops.emission_llh_bwd(X.detach(), ets, post, grad_out=grad_out.to(f32).contiguous(), comp_llh=comp,
                                 pdf_llh=pdf, pdf_of=pdf_of, scale=ctx.scale)
        return g, None, None, None, None, None, None, None


def _attach_frame_grad(Xg, frame, post, em, comp, pdf, scale=1.0):
    """`frame` (values of the kernels) with the autograd edge to the frames `Xg` described in _FrameLlhGrad."""
    if not ops.emission_bwd_supported(em.M, em.D):
        raise NotImplementedError(f'no gradient kernel w.r.t. the frames for D = {em.D} (2 D <= 128)')
    ets = ops.normalgamma_expected_stats(*em._cat('posterior'))
    pdf_of = None
    if comp is not None:
        pdf_of = torch.as_tensor(np.repeat(np.arange(em.Kp, dtype=np.int32), np.diff(em.comp_off_host)), device=em.device)
    return _FrameLlhGrad.apply(Xg, frame, post.contiguous(), ets, comp, pdf if comp is not None else None, pdf_of,
                               float(max(scale, 1.0)))


# ---------------------------------------------------------------------------------------------
# HMM (beer/models/hmm.py)
# ---------------------------------------------------------------------------------------------

class HMM(DiscreteLatentModel):
    """Hidden Markov Model with fixed transition probabilities."""

    @classmethod
    def create(cls, graph, modelset):
        return cls(graph, modelset)

    def __init__(self, graph, modelset):
        super().__init__(DynamicallyOrderedModelSet(modelset))
        self.graph = graph

    def mean_field_factorization(self):
        return self.modelset.mean_field_factorization()

    def sufficient_statistics(self, data):
        return self.modelset.sufficient_statistics(data)

    def _emission(self):
        return _Emission(_leaves(self.modelset))

    def expected_log_likelihood(self, stats, inference_graph=None, viterbi=False, state_path=None, scale=1.,
                                _unit_counts=None, _trans=None):
        """Per-frame sum_k p_tk gamma_tk with p = scale * llh[:, pdf_id_mapping] (hmm.py:73-92).
        Forward-backward by default; `viterbi=True` or a `state_path` give one-hot posteriors.
        `stats` may come from an `Utterances` batch: every utterance is then its own sequence.
        (`_unit_counts`: PhoneLoop's reduction of the transition posteriors, see below; `_trans` =
        (rows, cols) state ids: cache the block xi[:, rows, cols] of the transition posteriors the
        reference computes whenever no inference graph is given (hmm.py:76), as `trans_resps`.)"""
        graph = self.graph if inference_graph is None else inference_graph
        em = self._emission()
        X = frames_of(stats, em.D).detach()
        off, utts = _offsets_of(stats, X.shape[0], X.device)
        pdf, comp, fref = em.llh(X)
        if isinstance(graph, (list, tuple)):
            # one alignment graph per utterance of the batch (the loop of accumulate.py:47-57 as one launch)
            if viterbi or state_path is not None or _trans is not None:
                raise NotImplementedError('a list of inference graphs runs forward-backward only')
            if len(graph) != off.numel() - 1:
                raise ValueError('need one inference graph per utterance of the batch')
            r = ops.hmm_forward_backward_chains(ops.ChainBatch(graph, X.device), pdf, fref, off, scale=scale,
                                                want_frame_llh=True)
            self.cache.update(X=X, pdf_post=r['pdf_post'], pdf_llh=pdf, comp_llh=comp, emission=em, scale=scale,
                              utts=utts, utt_exp_llh=r['utt_exp_llh'])
            return r['frame_exp_llh']
        plan = graph.plan(n_pdfs=em.Kp)
        if viterbi or state_path is not None:
            if state_path is None:
                path = ops.hmm_viterbi(plan, pdf, off, scale=scale)
            else:
                path = torch.as_tensor(state_path, device=X.device).to(i32).contiguous()
            post, frame = ops.path_posteriors(path, em.Kp, pdf_map=graph.pdf_map_device(X.device), scale=scale,
                                              pdf_llh=pdf, frame_ref=fref)
            utt_ell = None
            self.cache['path'] = path
            if _trans is not None:
                self.cache['trans_resps'] = _onehot_transitions(path, off, *_trans)
        else:
            r = ops.hmm_forward_backward(plan, pdf, fref, off, scale=scale, want_frame_llh=True,
                                         unit_counts=_unit_counts, want_state_post=_trans is not None)
            post, frame, utt_ell = r['pdf_post'], r['frame_exp_llh'], r['utt_exp_llh']
            if _trans is not None:
                dev = X.device
                rows, cols = (torch.as_tensor(v, dtype=i32, device=dev) for v in _trans)
                self.cache['trans_resps'] = ops.hmm_transition_posteriors(
                    pdf, r['state_post'], off, graph.init_log_probs.detach().to(device=dev, dtype=f32).contiguous(),
                    graph.trans_log_probs.detach().to(device=dev, dtype=f32).contiguous(),
                    pdf_map=graph.pdf_map_device(dev), scale=scale, rows=rows, cols=cols)
                self.cache['first_state_post'] = r['state_post'][off[:-1][off[1:] > off[:-1]]]   # gamma_0 of every (non-empty) utterance
        self.cache.update(X=X, pdf_post=post, pdf_llh=pdf, comp_llh=comp, emission=em, scale=scale,
                          utts=utts, utt_exp_llh=utt_ell)
        Xg = frames_with_grad(stats)
        if Xg is not None:
            # (mixtures: mixtureset.py:92-98 returns a DETACHED log-normaliser, so the reference has no gradient here;
            # this is the gradient of sum_k gamma_tk sum_c r_tkc llh_kc(x_t), the same construction one level down)
            frame = _attach_frame_grad(Xg, frame, post, em, comp if em.has_mixtures else None, pdf, scale)
        return frame

    def accumulate(self, stats, parent_msg=None):
        """Statistics of every emission parameter from scale * posteriors, scatter-added onto pdf
        ids by the scan kernel (hmm.py:94-100, modelset.py:148-154)."""
        c = self.cache
        return c['emission'].accumulate(c['X'], c['pdf_post'], c['pdf_llh'], c['comp_llh'])

    def decode(self, data, inference_graph=None, scale=1.):
        """Best path as pdf ids, CPU LongTensor (hmm.py:105-114)."""
        graph = self.graph if inference_graph is None else inference_graph
        em = self._emission()
        stats = self.sufficient_statistics(data)
        X = frames_of(stats, em.D).detach()
        off, _ = _offsets_of(stats, X.shape[0], X.device)
        pdf, _, _ = em.llh(X, want_comp=False)
        path = ops.hmm_viterbi(graph.plan(n_pdfs=em.Kp), pdf, off, scale=scale).cpu().long()
        mapping = torch.as_tensor(np.asarray(graph.pdf_id_mapping), dtype=torch.int64)
        return mapping[path]

    def posteriors(self, data, inference_graph=None, scale=1.0):
        """State posteriors [N, K] (hmm.py:116-121; scaling the statistics instead of the llhs only
        moves a per-frame constant, which the posteriors do not see)."""
        graph = self.graph if inference_graph is None else inference_graph
        em = self._emission()
        stats = self.sufficient_statistics(data)
        X = frames_of(stats, em.D).detach()
        off, _ = _offsets_of(stats, X.shape[0], X.device)
        pdf, _, fref = em.llh(X, want_comp=False)
        r = ops.hmm_forward_backward(graph.plan(n_pdfs=em.Kp), pdf, fref, off, scale=scale, want_state_post=True,
                                     want_pdf_post=False)
        return r['state_post']


def _onehot_transitions(path, off, rows, cols):
    """One-hot transition posteriors of a state path (hmm.py:49-54), only the block [rows, cols]:
    [N - n_utts, R, C] with no entry across an utterance boundary."""
    dev = path.device
    rows = torch.as_tensor(rows, device=dev, dtype=path.dtype)
    cols = torch.as_tensor(cols, device=dev, dtype=path.dtype)
    keep = torch.ones(path.numel(), dtype=torch.bool, device=dev)
    keep[off[1:] - 1] = False                      # the last frame of an utterance starts no transition
    src = path[keep]
    dst = path[torch.roll(keep, 1)]                # frame t + 1 of every kept t
    return ((src[:, None] == rows[None, :])[:, :, None] & (dst[:, None] == cols[None, :])[:, None, :]).to(f32)


# ---------------------------------------------------------------------------------------------
# PhoneLoop (beer/models/phoneloop.py:12-101)
# ---------------------------------------------------------------------------------------------

class PhoneLoop(HMM):
    """Phone-loop HMM whose unit weights are learned: the end -> start transitions of the decoding
    graph are rewritten from E[ln w] after every update (phoneloop.py:53-65) and the unit counts
    come from the transition posteriors (phoneloop.py:83-101).  The (T-1, K, K) tensor of the
    reference is never formed: the forward-backward kernel reduces it to one count per unit."""

    @classmethod
    def create(cls, graph, start_pdf, end_pdf, modelset, categorical=None, prior_strength=1.0):
        if categorical is None:
            ref = modelset.mean_field_factorization()[0][0].posterior.params.mean
            weights = torch.ones(len(start_pdf), dtype=f32, device=ref.device) / len(start_pdf)
            categorical = Categorical.create(weights, prior_strength)
        return cls(graph, modelset, start_pdf, end_pdf, categorical)

    def __init__(self, graph, modelset, start_pdf, end_pdf, categorical):
        super().__init__(graph, modelset)
        self.start_pdf = start_pdf
        self.end_pdf = end_pdf
        self.categorical = categorical
        param = self.categorical.mean_field_factorization()[0][0]
        param.register_callback(self._on_weights_update)
        self._on_weights_update()

    def _on_weights_update(self):
        """ln A[end, starts] = ln(1 - A[end, end]) + E[ln w], in place (host side: P x P numbers per
        update; the device plan of the graph is rebuilt from the new values on its next use)."""
        log_weights = self.categorical.expected_log_weights()
        trans = self.graph.trans_log_probs
        log_weights = log_weights.to(device=trans.device, dtype=trans.dtype)
        start_idxs = [value for value in self.start_pdf.values()]
        for end_idx in self.end_pdf.values():
            loop_prob = trans[end_idx, end_idx].exp()
            trans[end_idx, start_idxs] = (1 - loop_prob).log() + log_weights

    def mean_field_factorization(self):
        return _merge_groups(self.modelset.mean_field_factorization(), self.categorical.mean_field_factorization())

    def expected_log_likelihood(self, stats, inference_graph=None, viterbi=False, state_path=None, scale=1.):
        counts = trans = None
        if inference_graph is None and not viterbi and state_path is None:
            # the reference switches the transition posteriors on when no inference graph is given
            # (hmm.py:76); here: one count per unit, reduced inside the backward sweep
            dev = self.categorical.mean_field_factorization()[0][0].posterior.params.concentrations.device
            plan = self.graph.plan(n_pdfs=self._emission().Kp)
            if plan.n_units == 0:
                # units of different lengths (or any other loop the fused reduction has no kernel for): the
                # ends x starts block of the transition posteriors, as the reference reads it
                trans = (list(self.end_pdf.values()), list(self.start_pdf.values()))
            else:
                counts = torch.zeros(plan.n_units, dtype=f64, device=dev)
        retval = super().expected_log_likelihood(stats, inference_graph=inference_graph, viterbi=viterbi,
                                                 state_path=state_path, scale=scale, _unit_counts=counts,
                                                 _trans=trans)
        if counts is not None:
            self.cache['unit_counts'] = counts
        elif inference_graph is None and 'path' in self.cache:
            self.cache['unit_path'] = self.cache['path']
        return retval

    def accumulate(self, stats, parent_msg=None):
        retval = super().accumulate(stats, parent_msg)
        weights = self.categorical.mean_field_factorization()[0][0]
        start_idxs = [value for value in self.start_pdf.values()]
        n_units = len(start_idxs)
        if 'unit_counts' in self.cache:
            counts = self.cache['unit_counts']
            su = self.graph.n_states // counts.numel()
            order = torch.as_tensor([s // su for s in start_idxs], device=counts.device)
            phone_resps = counts[order]
        elif 'trans_resps' in self.cache:
            block = self.cache['trans_resps'].to(f64)            # [N - n_utts, ends, starts]
            first = self.cache['first_state_post'].to(f64)
            phone_resps = block.sum(dim=(0, 1)) + first[:, torch.as_tensor(start_idxs, device=first.device)].sum(dim=0)
        elif 'unit_path' in self.cache:
            # one-hot transition posteriors of a Viterbi / given path (hmm.py:49-54)
            # every utterance of a batch is its own sequence: no transition across a boundary, one first frame each
            path = self.cache['unit_path'].long()
            utts = self.cache.get('utts')
            off = utts.offsets.to(path.device) if utts is not None else \
                torch.tensor([0, path.numel()], dtype=i64, device=path.device)
            off = off[:-1][off[1:] > off[:-1]]                      # first frame of every non-empty utterance
            ends = torch.zeros(self.graph.n_states, dtype=torch.bool, device=path.device)
            ends[torch.as_tensor(list(self.end_pdf.values()), device=path.device)] = True
            starts = torch.as_tensor(start_idxs, device=path.device)
            inner = torch.ones(path.numel(), dtype=torch.bool, device=path.device)
            inner[off] = False                                      # frame t is inner if t - 1 is in the same utterance
            hit = (ends[path[:-1]] & inner[1:])[:, None] & (path[1:, None] == starts[None, :])
            phone_resps = hit.sum(dim=0).to(f64) + (path[off][:, None] == starts[None, :]).sum(dim=0).to(f64)
        else:
            phone_resps = torch.zeros(n_units, dtype=f64, device=weights.posterior.params.concentrations.device)
        if isinstance(self.categorical, SBCategorical):
            retval[weights] = phone_resps.to(f64)        # plain counts: its callback makes the Beta statistics
            return retval
        stats_w = phone_resps.clone()
        stats_w[-1] = phone_resps.sum()
        retval[weights] = stats_w
        return retval


# ---------------------------------------------------------------------------------------------
# BigramPhoneLoop (beer/models/phoneloop.py:105-191)
# ---------------------------------------------------------------------------------------------

class BigramPhoneLoop(HMM):
    """Phone loop with a bigram model over the units: one Dirichlet per unit end over the unit starts.
    Its statistics are the ends x starts block of the transition posteriors (phoneloop.py:175-186), which
    `csrc/transitions.cu` produces directly ([T-1, P, P] instead of the reference's [T-1, K, K])."""

    @classmethod
    def create(cls, graph, start_pdf, end_pdf, modelset, categoricalset=None, prior_strength=1.0):
        if categoricalset is None:
            ref = modelset.mean_field_factorization()[0][0].posterior.params.mean
            n = len(start_pdf)
            weights = torch.ones(n, n, dtype=f32, device=ref.device) / n
            categoricalset = CategoricalSet.create(weights, prior_strength)
        return cls(graph, modelset, start_pdf, end_pdf, categoricalset)

    def __init__(self, graph, modelset, start_pdf, end_pdf, categoricalset):
        super().__init__(graph, modelset)
        self.start_pdf = start_pdf
        self.end_pdf = end_pdf
        self.categoricalset = categoricalset
        param = self.categoricalset.mean_field_factorization()[0][0]
        param.register_callback(self._on_weights_update)
        self._on_weights_update()

    def _on_weights_update(self):
        """phoneloop.py:145-157.  The reference evaluates the set on eye(P), which is indexed
        [class, model], and writes row i onto the arcs out of unit i's end state: arc (end_i -> start_m)
        gets E[ln pi_m(i)].  Kept as it is (results identical to the reference)."""
        logw = self.categoricalset.weights.posterior.expected_log_weights()       # [model, class]
        trans = self.graph.trans_log_probs
        logw = logw.to(device=trans.device, dtype=trans.dtype)
        start_idxs = [value for value in self.start_pdf.values()]
        for i, end_idx in enumerate(self.end_pdf.values()):
            loop_prob = trans[end_idx, end_idx].exp()
            trans[end_idx, start_idxs] = (1 - loop_prob).log() + logw[:, i]

    def mean_field_factorization(self):
        return _merge_groups(self.modelset.mean_field_factorization(), self.categoricalset.mean_field_factorization())

    def expected_log_likelihood(self, stats, inference_graph=None, viterbi=False, state_path=None, scale=1.):
        trans = None
        if inference_graph is None:
            trans = (list(self.end_pdf.values()), list(self.start_pdf.values()))
        return super().expected_log_likelihood(stats, inference_graph=inference_graph, viterbi=viterbi,
                                               state_path=state_path, scale=scale, _trans=trans)

    def accumulate(self, stats, parent_msg=None):
        retval = super().accumulate(stats, parent_msg)
        weights = self.categoricalset.weights
        conc = weights.posterior.params.concentrations
        if 'trans_resps' in self.cache:
            w_stats = self.cache['trans_resps'].to(f64).sum(dim=0)     # [ends, starts]
            w_stats[:, -1] = w_stats.sum(dim=-1)                       # dirichlet.py:18-21
        else:
            # forced alignments: the transitions are not trained (phoneloop.py:187-190)
            w_stats = torch.zeros(conc.shape, dtype=f64, device=conc.device)
        retval[weights] = w_stats
        return retval
