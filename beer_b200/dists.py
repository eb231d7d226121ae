"""Conjugate exponential-family pairs of the hot path, same call surface as beer/dists/
(NormalGamma <-> NormalDiagonalLikelihood, Dirichlet <-> CategoricalLikelihood, kl_div).

The standard parameters are fp32 CUDA buffers of `torch.nn.Module`s (so `.to()`, `state_dict`
and pickling behave as in the reference, beer/dists/normalgamma.py:69-74); every formula is
evaluated by the sm_100a kernels of libbeer_b200.so (csrc/dists.cu, fp64 arithmetic from the
fp32 parameters, rounded once).  There is no CPU path: the tensors must live on the GPU.
"""
import math

import torch

from . import ops

__all__ = ['Gamma', 'GammaStdParams', 'GammaLikelihood', 'ConjugateLikelihood', 'ExponentialFamily', 'NormalDiagonalLikelihood', 'NormalGamma',
           'NormalGammaStdParams', 'CategoricalLikelihood', 'Dirichlet', 'DirichletStdParams', 'kl_div',
           'DistributionTypeMismatch', 'SupportDimensionMismatch']


class DistributionTypeMismatch(Exception):
    """KL divergence between distributions of different families (basedist.py:22-24)."""


class SupportDimensionMismatch(Exception):
    """KL divergence between distributions with different supports (basedist.py:26-28)."""


class ConjugateLikelihood:
    """Descriptor of the likelihood conjugate to an ExponentialFamily prior (basedist.py:197-240)."""


class ExponentialFamily(torch.nn.Module):
    """A set of distributions of one exponential family (basedist.py:59-194)."""

    def __init__(self, params):
        super().__init__()
        self.params = params

    @classmethod
    def from_std_parameters(cls, *args, **kwargs):
        return cls(cls._std_params_cls(*args, **kwargs))

    def update_from_natural_parameters(self, natural_params):
        self.params = self.params.from_natural_parameters(natural_params)


def _f32(t):
    return torch.as_tensor(t).detach().to(torch.float32).contiguous()


# ---------------------------------------------------------------------------------------------
# Normal-Gamma  <->  Normal with diagonal covariance
# ---------------------------------------------------------------------------------------------

class NormalDiagonalLikelihood(ConjugateLikelihood):
    """beer/dists/normalgamma.py:11-59."""

    def __init__(self, dim):
        self.dim = dim

    def __eq__(self, other):
        return isinstance(other, NormalDiagonalLikelihood) and other.dim == self.dim

    def __hash__(self):
        return hash(('NormalDiagonalLikelihood', self.dim))

    def sufficient_statistics_dim(self, zero_stats=True):
        return 2 * self.dim + (2 if zero_stats else 0)

    @staticmethod
    def sufficient_statistics(data):
        """T(x) = [x, -x^2/2, -1/2, 1/2] (normalgamma.py:19-27).  The source frames stay attached
        to the result so that the kernels downstream read them instead of slicing T(x)."""
        if torch.is_tensor(data) and data.requires_grad:
            # differentiable inputs (an encoder in front of the model, vae.py): the statistics are formed by tensor
            # arithmetic so that the graph reaches `data`; the kernels downstream read the frames themselves and the
            # model's expected log-likelihood re-attaches its gradient to them (models._FrameLlhGrad)
            x = data.to(torch.float32)
            ones = torch.ones(len(x), 1, dtype=x.dtype, device=x.device)
            stats = torch.cat([x, -0.5 * x * x, -0.5 * ones, 0.5 * ones], dim=-1)
            stats._beer_frames = x
            return stats
        X = _f32(data)
        stats = ops.normal_sufficient_statistics(X)
        stats._beer_frames = X
        return stats

    def __call__(self, pdfvecs, stats):
        """stats @ pdfvecs.T - D/2 ln(2 pi) (normalgamma.py:55-59).  Dense product of two
        caller-supplied matrices: the one place where a library GEMM is the right tool; the
        engine's own path (NormalSet.expected_log_likelihood) uses the fused emission kernel."""
        if pdfvecs.dim() == 1:
            pdfvecs = pdfvecs.view(1, -1)
        return stats @ pdfvecs.t() - 0.5 * self.dim * math.log(2 * math.pi)


def frames_of(stats, dim):
    """Frames [N, D] behind a statistics tensor produced by `sufficient_statistics`."""
    X = getattr(stats, '_beer_frames', None)
    if X is None:
        X = stats[:, :dim].detach().to(torch.float32).contiguous()
    return X


def frames_with_grad(stats):
    """The frames behind `stats` when they are part of an autograd graph, else None."""
    X = getattr(stats, '_beer_frames', None)
    return X if (X is not None and X.requires_grad) else None


class NormalGammaStdParams(torch.nn.Module):
    """mean [M,D], scale [M,1], shape [M,1], rates [M,D] (normalgamma.py:62-94)."""

    def __init__(self, mean, scale, shape, rates):
        super().__init__()
        mean, rates = _f32(mean), _f32(rates)
        if mean.dim() == 1:
            mean, rates = mean.view(1, -1), rates.view(1, -1)
        M = mean.shape[0]
        self.register_buffer('mean', mean)
        self.register_buffer('scale', _f32(scale).reshape(M, 1).clone())
        self.register_buffer('shape', _f32(shape).reshape(M, 1).clone())
        self.register_buffer('rates', rates)

    def as_tuple(self):
        return self.mean, self.scale, self.shape, self.rates

    @classmethod
    def from_natural_parameters(cls, natural_params):
        return cls(*ops.normalgamma_from_natural(_f32(natural_params)))


class NormalGamma(ExponentialFamily):
    """beer/dists/normalgamma.py:97-183."""
    _std_params_cls = NormalGammaStdParams

    def __len__(self):
        return self.params.mean.shape[0]

    @property
    def dim(self):
        return (*self.params.mean.shape, self.params.rates.shape[-1])

    def conjugate(self):
        return NormalDiagonalLikelihood(self.params.mean.shape[-1])

    def expected_sufficient_statistics(self):
        """[a/b m, a/b, D/k + sum a/b m^2, sum psi(a) - ln b] (normalgamma.py:118-146)."""
        return ops.normalgamma_expected_stats(*self.params.as_tuple())

    def expected_value(self):
        return self.params.mean, self.params.shape / self.params.rates

    def log_norm(self):
        return ops.normalgamma_log_norm(*self.params.as_tuple())

    def natural_parameters(self):
        return ops.normalgamma_natural_params(*self.params.as_tuple())

    # engine hooks -------------------------------------------------------------------------
    def _kl(self, prior, out=None):
        return ops.normalgamma_kl(prior.params.as_tuple(), self.params.as_tuple(), out=out)

    def _natural_grad_update(self, prior, stats, lrate, stats_scale=1.0):
        """In-place eta <- eta + lrate (eta0 + s * stats - eta) (parameters.py:134-141)."""
        ops.normalgamma_update(prior.params.as_tuple(), self.params.as_tuple(), stats, stats_scale, lrate)


# ---------------------------------------------------------------------------------------------
# Dirichlet  <->  Categorical
# ---------------------------------------------------------------------------------------------

class CategoricalLikelihood(ConjugateLikelihood):
    """beer/dists/dirichlet.py:10-63."""

    def __init__(self, dim):
        self.dim = dim

    def __eq__(self, other):
        return isinstance(other, CategoricalLikelihood) and other.dim == self.dim

    def __hash__(self):
        return hash(('CategoricalLikelihood', self.dim))

    def sufficient_statistics_dim(self, zero_stats=True):
        return self.dim - 1 + (1 if zero_stats else 0)

    def sufficient_statistics(self, data):
        """Last column replaced by the row sum (dirichlet.py:18-21); pure re-indexing."""
        retval = data.clone().reshape(-1, data.shape[-1])
        retval[:, -1] = retval.sum(dim=-1)
        return retval.reshape(*data.shape)

    def __call__(self, pdfvecs, stats):
        return stats @ pdfvecs.t() if pdfvecs.dim() > 1 else stats @ pdfvecs


class DirichletStdParams(torch.nn.Module):
    """concentrations [K, C] or [C] (dirichlet.py:66-81)."""

    def __init__(self, concentrations):
        super().__init__()
        self.register_buffer('concentrations', _f32(concentrations).clone())

    @classmethod
    def from_natural_parameters(cls, natural_params):
        return cls(ops.dirichlet_from_natural(_f32(natural_params)))


class Dirichlet(ExponentialFamily):
    """beer/dists/dirichlet.py:84-162."""
    _std_params_cls = DirichletStdParams

    def __len__(self):
        shape = self.params.concentrations.shape
        return 1 if len(shape) <= 1 else shape[0]

    @property
    def dim(self):
        c = self.params.concentrations
        return len(c) if c.dim() <= 1 else tuple(c.shape)

    def conjugate(self):
        return CategoricalLikelihood(self.params.concentrations.shape[-1])

    def expected_sufficient_statistics(self):
        return ops.dirichlet_expected_stats(self.params.concentrations)

    def expected_log_weights(self):
        """E[ln pi] = psi(a) - psi(sum a): what the eye(C) evaluation of
        Mixture/MixtureSet._log_weights yields (mixtureset.py:64-67)."""
        return ops.dirichlet_expected_logw(self.params.concentrations)

    def expected_value(self):
        c = self.params.concentrations
        return c / c.sum(dim=-1, keepdim=True)

    def log_norm(self):
        return ops.dirichlet_log_norm(self.params.concentrations)

    def natural_parameters(self):
        return ops.dirichlet_natural_params(self.params.concentrations)

    # engine hooks -------------------------------------------------------------------------
    def _kl(self, prior, out=None):
        return ops.dirichlet_kl(prior.params.concentrations, self.params.concentrations, out=out)

    def _natural_grad_update(self, prior, stats, lrate, stats_scale=1.0):
        ops.dirichlet_update(prior.params.concentrations, self.params.concentrations, stats, stats_scale, lrate)


# ---------------------------------------------------------------------------------------------
# Gamma (beer/dists/gamma.py): the hyper-prior over the concentration of a stick-breaking process.  One or a few
# scalars per model, updated once per VB iteration from a callback: plain tensor arithmetic in fp64, no kernel.
# ---------------------------------------------------------------------------------------------

class GammaLikelihood(ConjugateLikelihood):
    """beer/dists/gamma.py:14-56."""

    def __init__(self, dim):
        self.dim = dim

    def sufficient_statistics_dim(self, zero_stats=True):
        return 2 * self.dim + (1 if zero_stats else 0)

    @staticmethod
    def sufficient_statistics(data):
        return torch.cat([-data, data.log(), torch.ones(len(data), 1, dtype=data.dtype, device=data.device)], dim=-1)

    def __call__(self, pdfvecs, stats):
        if pdfvecs.dim() == 1:
            pdfvecs = pdfvecs.view(1, -1)
        return stats @ pdfvecs.t()


class GammaStdParams(torch.nn.Module):
    """shape, rate (gamma.py:59-79)."""

    def __init__(self, shape, rate):
        super().__init__()
        self.register_buffer('shape', _f32(shape).clone())
        self.register_buffer('rate', _f32(rate).clone())

    @classmethod
    def from_natural_parameters(cls, natural_params):
        nat = natural_params.reshape(-1, natural_params.shape[-1])
        dim = nat.shape[-1] // 2
        shape, rate = nat[:, dim:] + 1, -nat[:, :dim]
        if natural_params.dim() == 1:
            return cls(shape.reshape(-1), rate.reshape(-1))
        return cls(shape, rate)


class Gamma(ExponentialFamily):
    """beer/dists/gamma.py:82-153."""
    _std_params_cls = GammaStdParams

    @property
    def dim(self):
        shape = self.params.shape
        return len(shape) if shape.dim() <= 1 else tuple(shape.shape)

    def conjugate(self):
        return GammaLikelihood(self.params.shape.shape[-1])

    def expected_sufficient_statistics(self):
        shape, rate = self.params.shape.double(), self.params.rate.double()
        return torch.cat([shape / rate, torch.digamma(shape) - torch.log(rate)], dim=-1)

    def expected_value(self):
        return self.params.shape / self.params.rate

    def log_norm(self):
        shape, rate = self.params.shape.double(), self.params.rate.double()
        return (torch.lgamma(shape) - shape * torch.log(rate)).sum(dim=-1)

    def natural_parameters(self):
        return torch.cat([-self.params.rate.double(), self.params.shape.double() - 1], dim=-1)

    def _kl(self, prior, out=None):
        """basedist.py:243-263 with this distribution as pdf1."""
        kl = prior.log_norm() - self.log_norm() - torch.sum(
            self.expected_sufficient_statistics() * (prior.natural_parameters() - self.natural_parameters()), dim=-1)
        kl = kl.reshape(-1).sum().reshape(1)
        if out is None:
            return kl
        out += kl
        return out

    def _natural_grad_update(self, prior, stats, lrate, stats_scale=1.0):
        eta_p, eta_q = prior.natural_parameters(), self.natural_parameters()
        new = eta_q + lrate * (eta_p + stats_scale * stats.to(eta_q) - eta_q)
        upd = GammaStdParams.from_natural_parameters(new)
        self.params.shape.copy_(upd.shape.reshape(self.params.shape.shape))
        self.params.rate.copy_(upd.rate.reshape(self.params.rate.shape))


def kl_div(model1, model2):
    """KL(model1 || model2) = A(eta2) - A(eta1) - <E_1[T], eta2 - eta1> (basedist.py:243-263).
    Returns the divergence SUMMED over the distributions of the set as a 1-element fp64 device
    tensor (every caller in the reference sums the per-distribution vector right away,
    basemodel.py:76, objectives.py:183)."""
    if type(model1) is not type(model2):
        raise DistributionTypeMismatch('Cannot compute KL divergence between distributions of different '
                                       f'types: {type(model1).__name__} and {type(model2).__name__}')
    if model1.dim != model2.dim:
        raise SupportDimensionMismatch('Cannot compute KL divergence between distributions with different '
                                       f'supports: {model1.dim} and {model2.dim}')
    return model1._kl(model2)
