"""Variational auto-encoder with an arbitrary beer model as the prior over the latent space (beer/models/vae.py).

The encoder / decoder networks, the reparameterised sample and the decoder likelihood are ordinary torch autograd
(they are the user's networks).  The prior's E-step on the latent samples -- `prior.expected_log_likelihood` on an HMM,
PhoneLoop or Mixture -- runs on the kernels of the VB hot path, and its gradient w.r.t. the samples (what trains the
encoder) comes from ONE tcgen05 kernel, csrc/emission_bwd.cu, with the posteriors held fixed as hmm.py:79-87 /
mixture.py:76-93 prescribe.
"""
import math

import torch

from .models import Model

__all__ = ['VAE', 'NormalDiagonalCovariance', 'MeanLogDiagCov']


class MeanLogDiagCov(torch.nn.Module):
    """Normal parameterised by the mean and the log of the diagonal covariance (vae.py:14-24)."""

    def __init__(self, mean, log_diag_cov):
        super().__init__()
        self.mean = mean
        self.log_diag_cov = log_diag_cov

    @property
    def diag_cov(self):
        return 1e-5 + self.log_diag_cov.exp()         # never exactly zero (vae.py:22-24)


class NormalDiagonalCovariance(torch.nn.Module):
    """One Normal with diagonal covariance per row of its parameters (beer/dists/normaldiag.py:69-200), differentiable
    in the parameters: the variational posteriors q(z | x) and the decoder densities p(x | z) of the VAE."""

    def __init__(self, params):
        super().__init__()
        self.params = params

    @property
    def dim(self):
        return self.params.mean.shape[-1]

    def natural_parameters(self):
        prec = 1. / self.params.diag_cov
        return torch.cat([prec * self.params.mean, prec], dim=-1)

    @staticmethod
    def sufficient_statistics(data):
        return torch.cat([data, -.5 * (data ** 2)], dim=-1)

    def forward(self, stats, pdfwise=False):
        """ln N(x; m, s) of every row from its statistics (normaldiag.py:87-106)."""
        mean, diag_cov = self.params.mean, self.params.diag_cov
        nparams = self.natural_parameters()
        lnorm = .5 * (diag_cov.log().sum(dim=-1) + ((1. / diag_cov) * mean ** 2).sum(dim=-1))
        base = -.5 * self.dim * math.log(2 * math.pi)
        if pdfwise:
            return torch.sum(nparams * stats, dim=-1) - lnorm + base
        return nparams @ stats.t() - lnorm[:, None] + base

    def sample(self, nsamples):
        """[N, nsamples, D] reparameterised samples (normaldiag.py:152-164)."""
        mean, diag_cov = self.params.mean, self.params.diag_cov
        noise = torch.randn(mean.shape[0], nsamples, mean.shape[-1], dtype=mean.dtype, device=mean.device)
        return mean[:, None, :] + diag_cov.sqrt()[:, None, :] * noise


class VAE(Model):
    """vae.py:27-89: `prior` is any beer model over the latent space, `encoder` / `decoder` are torch modules with
    `dim_in` / `dim_out` attributes."""

    def __init__(self, prior, encoder, decoder):
        super().__init__()
        self.prior = prior
        self.encoder = encoder
        self.decoder = decoder
        self.enc_mean_layer = torch.nn.Linear(encoder.dim_out, decoder.dim_in)
        self.enc_var_layer = torch.nn.Linear(encoder.dim_out, decoder.dim_in)
        self.dec_mean_layer = torch.nn.Linear(decoder.dim_out, encoder.dim_in)
        self.dec_var_layer = torch.nn.Linear(decoder.dim_out, encoder.dim_in)

    def posteriors(self, X):
        """Variational posteriors q(z | x) of the encoder (vae.py:39-44)."""
        H = self.encoder(X)
        return NormalDiagonalCovariance(MeanLogDiagCov(self.enc_mean_layer(H), self.enc_var_layer(H)))

    def pdfs(self, Z):
        """Decoder densities p(x | z) (vae.py:46-51)."""
        Z1 = self.decoder(Z)
        return NormalDiagonalCovariance(MeanLogDiagCov(self.dec_mean_layer(Z1), self.dec_var_layer(Z1)))

    # -- Model interface ------------------------------------------------------
    def mean_field_factorization(self):
        return self.prior.mean_field_factorization()

    def sufficient_statistics(self, data):
        return data

    def expected_log_likelihood(self, data, nsamples=1, llh_weight=1., kl_weight=1., **kwargs):
        """llh_weight * E_q[ln p(x | z)] - kl_weight * (E_q[-ln p(z)] - H[q]) by sampling (vae.py:63-86).  The value has
        the reference's shape: the [N, 1] likelihood minus the [N] divergence broadcasts to [N, N] there (vae.py:86),
        and `evidence_lower_bound` sums that matrix -- reproduced as is so that ELBO values agree."""
        if nsamples != 1:
            # the reference averages the prior's STATISTICS over the samples (vae.py:73-75); the kernels of the prior
            # take frames, not free-form statistics
            raise NotImplementedError('the latent prior runs on one sample per frame (nsamples = 1, the default)')
        # a ragged batch of utterances (`Utterances`): the networks see all frames at once, the prior runs every
        # utterance as its own sequence (one launch per kernel); the value is then the per-frame vector [N]
        from .engine import Utterances
        utts = data if isinstance(data, Utterances) else None
        if utts is not None:
            data = utts.X
        posts = self.posteriors(data)
        samples = posts.sample(nsamples)
        s_samples = posts.sufficient_statistics(samples).mean(dim=1)
        ent = -posts(s_samples, pdfwise=True)
        z = samples.view(-1, samples.shape[-1])
        # (carries z: the prior's kernels read the frames)
        prior_stats = self.prior.sufficient_statistics(z if utts is None else Utterances(z, utts.lengths))
        self.cache['prior_stats'] = prior_stats
        xent = -self.prior.expected_log_likelihood(prior_stats, **kwargs).to(ent.dtype)
        local_kl_div = xent - ent
        pdfs = self.pdfs(z)
        r_data = data[:, None, :].repeat(1, nsamples, 1).view(-1, data.shape[-1])
        llh = pdfs(pdfs.sufficient_statistics(r_data), pdfwise=True)
        llh = llh.reshape(len(data), nsamples, -1).mean(dim=1)
        if utts is not None:
            return llh_weight * llh.reshape(-1) - kl_weight * local_kl_div
        return llh_weight * llh - kl_weight * local_kl_div

    def accumulate(self, stats, parent_msg=None):
        return self.prior.accumulate(self.cache['prior_stats'])

    def clear_cache(self):
        super().clear_cache()
