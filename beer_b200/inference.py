"""Variational-Bayes objective and optimizers, same call surface as beer/inference/
(objectives.py: evidence_lower_bound, EvidenceLowerBoundInstance; optimizers.py:
VBConjugateOptimizer, VBOptimizer).  The legacy spellings of beer/vbi.py are aliased in
beer_b200/vbi.py.

The ELBO value is an fp64 scalar ON THE DEVICE (no host synchronisation until `float(elbo)`),
the accumulated statistics are fp64 device tensors keyed by parameter.
"""
import torch

from .engine import Utterances

__all__ = ['evidence_lower_bound', 'EvidenceLowerBoundInstance', 'VBConjugateOptimizer', 'VBOptimizer',
           'add_acc_stats', 'scale_acc_stats']


def add_acc_stats(acc_stats1, acc_stats2):
    """Union of two {parameter: statistics} dictionaries, summing shared keys (objectives.py:10-34)."""
    new_stats = dict(acc_stats1)
    for key, val in acc_stats2.items():
        new_stats[key] = new_stats[key] + val if key in new_stats else val
    return new_stats


def scale_acc_stats(acc_stats, scale):
    return {key: scale * val for key, val in acc_stats.items()}


class EvidenceLowerBoundInstance:
    """ELBO of (a part of) a data set: value + the statistics needed for the update
    (objectives.py:54-116).  Instances add up; `backward()` hands the statistics, scaled by
    datasize / (frames seen), to the parameters."""

    def __init__(self, value, acc_stats, model_parameters, minibatchsize, datasize):
        self.value = value
        self._acc_stats = acc_stats
        self._model_parameters = set(model_parameters)
        self._minibatchsize = minibatchsize
        self._datasize = datasize

    def __repr__(self):
        return f'EvidenceLowerBoundInstance(value={float(self):.6f})'

    def __float__(self):
        return float(self.value.detach() if isinstance(self.value, torch.Tensor) else self.value)

    def __add__(self, other):
        if not isinstance(other, EvidenceLowerBoundInstance):
            raise ValueError('EvidenceLowerBoundInstance')
        if self._datasize != other._datasize:
            raise ValueError('Cannot add ELBOs evaluated on different data set')
        return EvidenceLowerBoundInstance(self.value + other.value,
                                          add_acc_stats(self._acc_stats, other._acc_stats),
                                          self._model_parameters.union(other._model_parameters),
                                          self._minibatchsize + other._minibatchsize, self._datasize)

    def backward(self, std_params=True):
        if std_params and isinstance(self.value, torch.Tensor) and self.value.requires_grad:
            (-self.value).backward()
        scale = self._datasize / self._minibatchsize
        for parameter in self._model_parameters:
            try:
                parameter.store_stats(scale * self._acc_stats[parameter])
            except KeyError:
                pass

    def sync(self, model):
        """Re-attach the parameters of `model` after the instance went through pickle
        (objectives.py:109-116)."""
        self._model_parameters = set(model.bayesian_parameters())


def evidence_lower_bound(model=None, minibatch_data=None, datasize=-1, **kwargs):
    """ELBO of `minibatch_data` under `model` (objectives.py:119-190):

        (datasize / T) * sum_t E[ln p(x_t)] - KL(q || p)

    Called with only `datasize`, returns an empty accumulator.  `minibatch_data` is one sequence
    of frames [T, D] (CUDA tensor) or an `Utterances` ragged batch; the batch form returns what
    summing the instances of its utterances returns (the loop of `beer hmm accumulate`,
    beer/cli/subcommands/hmm/accumulate.py:39-59), from one launch per kernel."""
    if model is None and minibatch_data is None and datasize > 0:
        return EvidenceLowerBoundInstance(0., {}, [], 0, datasize)
    if model is None or minibatch_data is None:
        raise ValueError('if datasize is not provided, need at least "model" and "minibatch_data"')

    mb_datasize = len(minibatch_data)
    if datasize <= 0:
        datasize = mb_datasize
    stats = model.sufficient_statistics(minibatch_data)
    exp_llh = model.expected_log_likelihood(stats, **kwargs)
    kl_div = model.kl_div_posterior_prior().sum()
    if isinstance(minibatch_data, Utterances) and minibatch_data.n_utts != 1:
        utts = minibatch_data
        lens = torch.as_tensor(utts.lengths, dtype=torch.float64, device=exp_llh.device)
        per_utt = model.cache.get('utt_exp_llh') if hasattr(model, 'cache') else None
        if per_utt is None:
            csum = torch.cat([exp_llh.new_zeros(1, dtype=torch.float64), exp_llh.double().cumsum(0)])
            per_utt = csum[utts.offsets[1:]] - csum[utts.offsets[:-1]]
        nonempty = lens > 0
        elbo_value = (float(datasize) * (per_utt[nonempty] / lens[nonempty])).sum() - utts.n_utts * kl_div
    else:
        scale = datasize / float(mb_datasize)
        elbo_value = float(scale) * exp_llh.double().sum() - kl_div
    acc_stats = model.accumulate(stats)
    model.clear_cache()
    return EvidenceLowerBoundInstance(elbo_value, acc_stats, model.bayesian_parameters(), mb_datasize, datasize)


class VBConjugateOptimizer:
    """Coordinate-ascent natural-gradient optimizer over mean-field groups (optimizers.py:5-31)."""

    def __init__(self, groups, lrate=1.):
        self.groups = [[param for param in group] for group in groups]
        self.lrate = lrate
        self.update_count = 0

    def state_dict(self):
        return {'lrate': self.lrate, 'update_count': self.update_count}

    def load_state_dict(self, state_dict):
        self.lrate = state_dict['lrate']
        self.update_count = state_dict['update_count']

    def init_step(self):
        for group in self.groups:
            for param in group:
                param.zero_stats()

    def step(self):
        if len(self.groups) > 0:
            for parameter in self.groups[self.update_count % len(self.groups)]:
                parameter.natural_grad_update(self.lrate)
        self.update_count += 1


class VBOptimizer:
    """Conjugate + standard (torch) optimizer pair (optimizers.py:34-67)."""

    def __init__(self, cjg_optim=None, std_optim=None):
        self.cjg_optim = cjg_optim
        self.std_optim = std_optim

    def state_dict(self):
        state = {}
        if self.cjg_optim is not None:
            state['cjg_optim'] = self.cjg_optim.state_dict()
        if self.std_optim is not None:
            state['std_optim'] = self.std_optim.state_dict()
        return state

    def load_state_dict(self, state_dict):
        if self.cjg_optim is not None:
            self.cjg_optim.load_state_dict(state_dict['cjg_optim'])
        if self.std_optim is not None:
            self.std_optim.load_state_dict(state_dict['std_optim'])

    def init_step(self):
        if self.cjg_optim is not None:
            self.cjg_optim.init_step()
        if self.std_optim is not None:
            self.std_optim.zero_grad()

    def step(self):
        if self.std_optim is not None:
            self.std_optim.step()
        if self.cjg_optim is not None:
            self.cjg_optim.step()
