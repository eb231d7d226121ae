"""Bayesian parameters (prior + posterior + accumulated statistics), same call surface as
beer/models/parameters.py.  The M-step `natural_grad_update` runs in the conjugate-update
kernels of libbeer_b200.so; statistics are accumulated in fp64 on the device."""
import uuid

import torch

from .dists import kl_div

__all__ = ['BayesianParameter', 'ConjugateBayesianParameter']


class BayesianParameter(torch.nn.Module):
    """A parameter with a prior and a posterior distribution (parameters.py:11-80).  Identity is
    a uuid, so that parameters survive pickling as dictionary keys (EvidenceLowerBoundInstance
    keeps `{parameter: statistics}`, objectives.py:10-34)."""

    def __init__(self, prior, posterior=None):
        super().__init__()
        self.prior = prior
        self.posterior = posterior
        self.uuid = uuid.uuid4()
        self._callbacks = set()

    def __len__(self):
        return len(self.prior)

    def __repr__(self):
        post = self.posterior.__class__.__qualname__ if self.posterior is not None else '<unspecified>'
        return f'{self.__class__.__qualname__}(prior={self.prior.__class__.__qualname__}, posterior={post})'

    def __hash__(self):
        return hash(self.uuid)

    def __eq__(self, other):
        return isinstance(other, BayesianParameter) and hash(self) == hash(other)

    def dispatch(self, before_update=False):
        """Notify the observers that the parameter is about to change / has changed."""
        for callback, notify_before_update in self._callbacks:
            if notify_before_update == before_update:
                callback()

    def register_callback(self, callback, notify_before_update=False):
        self._callbacks.add((callback, notify_before_update))

    def value(self):
        return self.posterior.expected_value()

    def kl_div_posterior_prior(self):
        return kl_div(self.posterior, self.prior)


class ConjugateBayesianParameter(BayesianParameter):
    """Parameter whose likelihood is conjugate to its prior (parameters.py:83-141)."""

    def __init__(self, prior, posterior, init_stats=None, likelihood_fn=None):
        super().__init__(prior, posterior)
        if init_stats is None:
            nat = prior.natural_parameters()
            init_stats = torch.zeros(nat.shape, dtype=torch.float64, device=nat.device)
        self.register_buffer('stats', init_stats.clone().detach())
        self.likelihood_fn = prior.conjugate() if likelihood_fn is None else likelihood_fn

    def __len__(self):
        return 1 if self.stats.dim() <= 1 else self.stats.shape[0]

    def zero_stats(self):
        self.stats.zero_()

    def store_stats(self, acc_stats):
        """Keep the (already scaled) accumulated statistics for the next update; never part of
        an autograd graph (parameters.py:115-129)."""
        self.stats = acc_stats.detach() if acc_stats.requires_grad else acc_stats

    def natural_form(self):
        return self.posterior.expected_sufficient_statistics()

    def natural_grad_update(self, lrate):
        """eta <- eta + lrate (eta_prior + stats - eta), posterior rewritten in place; observers
        are called before and after (parameters.py:134-141)."""
        self.dispatch(before_update=True)
        stats = self.stats.to(torch.float64)
        self.posterior._natural_grad_update(self.prior, stats.contiguous(), float(lrate))
        self.dispatch(before_update=False)
