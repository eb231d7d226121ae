"""Batched VB-EM engine: one VB iteration of an HMM-GMM over a ragged batch of utterances
resident in HBM, the way `beer hmm accumulate` + `beer hmm update` compose it
(beer/cli/subcommands/hmm/accumulate.py:37-63, update.py:37-62):

    for every utterance u:  elbo += evidence_lower_bound(model, X_u, datasize=N, ...)
    elbo.backward(); optim.step()

but as a handful of kernel launches over all utterances at once:

    KM  emission weights + KL        (beer_emission_prepare, beer_*_kl)
    KA  per-frame per-pdf llh        (beer_emission_llh)
    KB  forward-backward             (beer_hmm_forward_backward)
    KC  sufficient statistics        (beer_accumulate_stats)
    --  one all-reduce of the flat fp64 statistics buffer (torch.distributed, NCCL)
    KM  natural-gradient M-step      (beer_normalgamma_update, beer_dirichlet_update)

Utterances shard over ranks (one process per GPU); the only exchange is the all-reduce.
"""
import math
import os

import numpy as np
import torch

from . import ops

f32, f64, i32, i64 = torch.float32, torch.float64, torch.int32, torch.int64


class Utterances:
    """Ragged batch: frames of all utterances concatenated in one [N, D] fp32 tensor."""

    def __init__(self, X, lengths, device=None):
        lengths = np.asarray(lengths, dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(lengths)])
        if isinstance(X, (list, tuple)):
            X = torch.cat([torch.as_tensor(x, dtype=f32) for x in X]) if len(X) else torch.zeros(0, 1)
        X = torch.as_tensor(X, dtype=f32)
        if X.shape[0] != off[-1]:
            raise ValueError('sum(lengths) must equal the number of frames')
        self.X = X.to(device).contiguous() if device is not None else X.contiguous()
        self.lengths = lengths
        self.offsets_host = off
        self.offsets = torch.as_tensor(off, dtype=i64, device=self.X.device)

    @classmethod
    def from_list(cls, utts, device=None):
        return cls(list(utts), [len(u) for u in utts], device=device)

    def __len__(self):           # number of frames, like len(minibatch_data) in the reference
        return int(self.offsets_host[-1])

    @property
    def n_utts(self):
        return len(self.lengths)


class WeightGroup:
    """Dirichlet prior/posterior over the mixture weights of a run of pdfs with the same
    number of components (one MixtureSet of a JointModelSet)."""

    def __init__(self, pdf_start, n_pdfs, n_comp, prior_conc, post_conc):
        self.pdf_start, self.n_pdfs, self.n_comp = pdf_start, n_pdfs, n_comp
        self.prior, self.post = prior_conc, post_conc      # [n_pdfs, n_comp] fp32 on device


class EmissionParams:
    """Flat device view of the emission model: M diagonal Gaussians (Normal-Gamma prior and
    posterior), grouped into Kp pdfs by `comp_off`, with optional mixture-weight groups."""

    def __init__(self, prior, post, comp_off=None, weight_groups=()):
        self.prior = tuple(prior)      # (mean [M,D], scale [M], shape [M], rates [M,D])
        self.post = tuple(post)
        self.M, self.D = self.post[0].shape
        self.device = self.post[0].device
        self.weight_groups = list(weight_groups)
        if comp_off is None:
            self.comp_off_host = np.arange(self.M + 1, dtype=np.int32)
            self.comp_off = None
            self.Kp = self.M
        else:
            self.comp_off_host = np.asarray(comp_off, dtype=np.int32)
            self.Kp = len(self.comp_off_host) - 1
            self.comp_off = torch.as_tensor(self.comp_off_host, dtype=i32, device=self.device)
        self.has_mixtures = self.Kp != self.M
        self.logw = torch.zeros(self.M, device=self.device, dtype=f32) if self.weight_groups else None
        # uniform number of Gaussians per pdf -> tensor-core (tcgen05) emission kernel
        counts = np.diff(self.comp_off_host)
        self.uniform_C = int(counts[0]) if len(counts) and np.all(counts == counts[0]) else 0
        self.use_tc = bool(self.uniform_C) and ops.emission_tc_supported(self.M, self.D, self.uniform_C)
        # mixtures of 4 / 8 / 16 Gaussians per pdf: the fp16-split kernels that keep the per-Gaussian llhs on chip
        # (single-Gaussian pdfs, C = 1, take the same kernels: statistics / posteriors as tensor-memory operands)
        self.use16 = (bool(self.uniform_C) and os.environ.get('BEER_B200_NO_MIX16') is None
                      and ops.mix16_supported(self.M, self.D, self.uniform_C))
        # a plain GMM (one pdf of M components): its components as pseudo-pdfs of 8 / 16 / 4 for the same kernels
        self.gmm_C = 0
        if self.Kp == 1 and self.M > 1 and os.environ.get('BEER_B200_NO_MIX16') is None:
            self.gmm_C = next((c for c in (8, 16, 4) if self.M % c == 0 and ops.mix16_supported(self.M, self.D, c)), 0)
        self._image = None

    def refresh(self, pack_tc=True):
        """(W, bias, ref) of the current posteriors."""
        for g in self.weight_groups:
            j0 = int(self.comp_off_host[g.pdf_start])
            self.logw[j0:j0 + g.n_pdfs * g.n_comp] = ops.dirichlet_expected_logw(g.post).reshape(-1)
        W, bias, ref = ops.emission_prepare(*self.post, logw=self.logw)
        if self.use_tc and pack_tc:
            self._image = ops.emission_tc_pack(W, bias, self.uniform_C, out=self._image)
        return W, bias, ref

    def llh(self, X, W, bias, ref, out, out_comp, out_ref):
        """KA over a run of frames: tcgen05 path when the shape has one, SIMT kernel otherwise."""
        if self.use_tc:
            return ops.emission_llh_tc(X, self._image, ref, self.M, self.uniform_C, out=out, out_comp=out_comp,
                                       out_ref=out_ref)
        return ops.emission_llh(X, W, bias, ref, comp_off=self.comp_off, Kp=self.Kp, out=out, out_comp=out_comp,
                                out_ref=out_ref)

    def kl(self, out=None):
        out = ops.normalgamma_kl(self.prior, self.post, out=out)
        for g in self.weight_groups:
            ops.dirichlet_kl(g.prior, g.post, out=out)
        return out

    def update(self, acc, stats_scale, lrate):
        """Natural-gradient step of every parameter from the accumulated statistics."""
        if self.weight_groups:
            wst = ops.mixture_weight_stats(acc, self.D, comp_off=self.comp_off, Kp=self.Kp)
            for g in self.weight_groups:
                j0 = int(self.comp_off_host[g.pdf_start])
                gacc = wst[j0:j0 + g.n_pdfs * g.n_comp]
                ops.dirichlet_update(g.prior, g.post, gacc, stats_scale, lrate)
        ops.normalgamma_update(self.prior, self.post, acc, stats_scale, lrate)


class UnitWeights:
    """Learned unit weights of a phone loop (PhoneLoop.categorical, beer/models/phoneloop.py:12-101):
    Dirichlet prior / posterior concentrations [P] on the device, the compiled decoding graph whose
    end -> start transitions are rewritten from E[ln w] after every update, and the start / end state of
    every unit (in the order of the weights)."""

    def __init__(self, prior_conc, post_conc, graph, start_idxs, end_idxs):
        self.prior, self.post = prior_conc, post_conc
        self.graph = graph
        self.start_idxs = [int(i) for i in start_idxs]
        self.end_idxs = [int(i) for i in end_idxs]

    def expected_log_weights(self):
        return ops.dirichlet_expected_logw(self.post)

    def kl(self, out):
        """out (fp64 [1]) += KL(posterior || prior) of the weights."""
        ops.dirichlet_kl(self.prior, self.post, out=out)

    def update(self, counts, stats_scale, lrate):
        """Natural-gradient step from the unit counts (in the order of the weights): Dirichlet statistics with the
        last entry replaced by the total (dirichlet.py:18-21)."""
        stats = counts.clone()
        stats[-1] = stats.sum()
        ops.dirichlet_update(self.prior, self.post, stats, stats_scale, lrate)

    def rewrite_graph(self):
        """ln A[end, starts] = ln(1 - A[end, end]) + E[ln w] (phoneloop.py:53-65), on the host copy of the
        graph; its device plan is rebuilt on the next use."""
        logw = self.expected_log_weights()
        trans = self.graph.trans_log_probs
        logw = logw.to(device=trans.device, dtype=trans.dtype)
        for e in self.end_idxs:
            trans[e, self.start_idxs] = (1 - trans[e, e].exp()).log() + logw


class CategoricalUnitWeights(UnitWeights):
    """Unit weights held by a model of `beer_b200.models`: `Categorical` (Dirichlet), `SBCategorical` or
    `SBCategoricalHyperPrior` (the stick-breaking priors of categorical.py:82-209; `gamma_dirichlet_process` is the
    default of `beer hmm mkphoneloop`).  The engine hands it the unit counts of the iteration; the model's own
    parameter objects do the statistics transform, the update and the callbacks (a handful of numbers per iteration)."""

    def __init__(self, categorical, graph, start_idxs, end_idxs):
        self.categorical = categorical
        self.graph = graph
        self.start_idxs = [int(i) for i in start_idxs]
        self.end_idxs = [int(i) for i in end_idxs]

    def expected_log_weights(self):
        return self.categorical.expected_log_weights()

    def kl(self, out):
        out += self.categorical.kl_div_posterior_prior().sum().to(out.dtype)

    def update(self, counts, stats_scale, lrate):
        from .models import SBCategorical
        stats = counts.clone().to(f64)
        if not isinstance(self.categorical, SBCategorical):
            stats[-1] = stats.sum()            # (stick-breaking weights take the plain counts: categorical.py:107-116)
        param = self.categorical.mean_field_factorization()[0][0]
        param.store_stats(stats * stats_scale)
        param.natural_grad_update(lrate)


class BigramUnitWeights(UnitWeights):
    """Bigram unit weights of a `BigramPhoneLoop` (beer/models/phoneloop.py:105-191): a `CategoricalSet` of one Dirichlet
    per unit over the unit starts.  The engine hands it the ends x starts block of the transition posteriors summed over
    the frames of the iteration (phoneloop.py:175-186)."""

    needs_block = True

    def __init__(self, categoricalset, graph, start_idxs, end_idxs):
        self.categoricalset = categoricalset
        self.graph = graph
        self.start_idxs = [int(i) for i in start_idxs]
        self.end_idxs = [int(i) for i in end_idxs]

    def expected_log_weights(self):
        return self.categoricalset.weights.posterior.expected_log_weights()       # [model, class]

    def kl(self, out):
        out += self.categoricalset.kl_div_posterior_prior().sum().to(out.dtype)

    def update(self, counts, stats_scale, lrate):
        """counts: [ends, starts] block; every row is the statistics of one Dirichlet, last entry = the row total
        (dirichlet.py:18-21)."""
        stats = counts.clone().to(f64)
        stats[:, -1] = counts.sum(dim=-1)
        param = self.categoricalset.weights
        param.store_stats(stats * stats_scale)
        param.natural_grad_update(lrate)

    def rewrite_graph(self):
        """Arc (end_i -> start_m) = ln(1 - A[end_i, end_i]) + E[ln pi]_{m i}: the reference evaluates the set on eye(P),
        indexed [class, model], and writes row i onto the arcs out of unit i (phoneloop.py:145-157); kept as it is."""
        logw = self.expected_log_weights()
        trans = self.graph.trans_log_probs
        logw = logw.to(device=trans.device, dtype=trans.dtype)
        for i, e in enumerate(self.end_idxs):
            trans[e, self.start_idxs] = (1 - trans[e, e].exp()).log() + logw[:, i]


class _StageTimer:
    """CUDA events around one kernel stage on the current stream (only when profiling)."""

    def __init__(self, profile, name):
        self.profile, self.name = profile, name

    def __enter__(self):
        if self.profile is not None:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()

    def __exit__(self, *exc):
        if self.profile is not None:
            self.ev[1].record()
            self.profile.setdefault(self.name, []).append(self.ev)
        return False


def bind_to_gpu_cpus(device_index):
    """Pin this process to the CPUs NVML reports as local to GPU `device_index` (same NUMA node / PCIe root), so
    that the pinned host buffers it allocates afterwards sit next to the GPU they feed.  Matters for the
    host-streaming mode at several ranks per box; returns the CPU set or None when NVML / affinity is unavailable."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1 and 64 * w + b < n}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:       # no NVML, restricted container, non-Linux: placement stays as the launcher left it
        return None


def shard_utterances(lengths, world_size):
    """Static utterance -> rank assignment balancing the frames per rank (longest-processing-time
    greedy): the `split` step of the reference's job arrays (recipes/zrc2019/utils/parallel/split.sh)
    with balanced shards.  Returns a list of index arrays, one per rank."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind='stable')
    load = np.zeros(world_size, dtype=np.int64)
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))
        shards[r].append(int(i))
        load[r] += lengths[i]
    return [np.asarray(sorted(s), dtype=np.int64) for s in shards]


def elbo_from_flat(extras, kl, datasize):
    """Summed ELBO of `beer hmm accumulate` + `update` from the reduced bookkeeping terms
    extras = [sum_u ell_u / T_u, sum_u T_u, n_utts, sum_u ell_u]: every utterance contributes
    (datasize / T_u) * ell_u - KL (objectives.py:180-184 added up as in accumulate.py:39-59)."""
    return datasize * extras[0] - extras[2] * kl


def stats_scale_from_flat(total_frames, datasize):
    """`backward()` hands datasize / (frames seen) * statistics to the parameters (objectives.py:98-106)."""
    return datasize / total_frames


class VBEngine:
    """One process per GPU; `utts` is this rank's shard.  Its frames stay resident in HBM, or --
    when `utts.X` is a pinned HOST tensor -- are streamed through two device staging buffers, the
    copy of chunk i+1 overlapping the kernels of chunk i; with `prefetch` (default) the first chunk of the NEXT VB
    iteration is copied under the last chunk of this one, so that in steady state no copy is exposed (the host tensor
    must then not be modified between `step()` calls).  `sparse_stats` (default: on unless BEER_B200_DENSE_STATS is set):
    mixtures on the fp16-split kernels let the forward-backward mark the (frame tile, Gaussian tile) pairs with posterior
    mass and run the statistics kernel over those only (the others are exact zeros in its operands)."""

    def __init__(self, emission, plan, utts, datasize=None, scale=1.0, lrate=1.0, chunk_frames=None,
                 process_group=None, distributed=None, use_graph=False, unit_weights=None, viterbi=False,
                 prefetch=True, sparse_stats=None):
        self.em, self.plan, self.utts = emission, plan, utts
        self.prefetch = bool(prefetch)
        # Viterbi training (hmm.py:42-58 with viterbi=True, what recipes/zrc2019 trains with): one-hot posteriors of
        # the best path instead of forward-backward
        self.viterbi = bool(viterbi)
        self.scale, self.lrate = float(scale), float(lrate)
        self.dev = emission.device
        self.pg = process_group
        if distributed is None:
            distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.distributed = distributed
        M, D = emission.M, emission.D
        self.Q = 2 * D + 2
        # flat reduction buffer: [acc (M*Q) | unit counts (P) | sum_u (N/T_u) ell_u | sum_u T_u | n_utts | sum_u ell_u]
        self.units = unit_weights
        # per-utterance alignment chains (ops.ChainBatch: chain i belongs to utterance i of this shard) or one graph plan
        # plan None: a GMM without an HMM on top (Mixture, beer/models/mixture.py:70-102): no forward-backward at all
        self.gmm = plan is None
        if self.gmm and (emission.gmm_C == 0 or unit_weights is not None or viterbi):
            raise ValueError('the batched GMM mode needs one pdf whose number of components is a multiple of 32 '
                             '(D = 20 or 40); other mixtures go through the model API (Mixture)')
        self.chains = isinstance(plan, ops.ChainBatch)
        # one GraphPlan per utterance (alignment graphs that are not plain chains, e.g. units with skip arcs: the
        # forward-backward runs utterance by utterance, emission and statistics kernels stay batched)
        self.per_utt = isinstance(plan, (list, tuple))
        if self.per_utt and (len(plan) != utts.n_utts or self.viterbi):
            raise ValueError('a list of graph plans needs one plan per utterance of the shard and runs forward-backward')
        aligned = self.chains or self.per_utt
        if self.chains and (plan.n_utts != utts.n_utts or self.viterbi):
            raise ValueError('a ChainBatch needs one chain per utterance of the shard and runs forward-backward')
        # aligned training gives the unit weights ZERO statistics (phoneloop.py:98-100): with `unit_weights` their KL
        # still enters the ELBO and their posterior takes the natural-gradient step towards the prior, as in the reference
        P = plan.n_units if (unit_weights is not None and not aligned) else 0
        # phone loops the fused unit counting does not take (units of uneven length, shared pdfs, a bigram over the
        # units): the ends x starts block of the transition posteriors (csrc/transitions.cu) summed over the frames,
        # plus the start-state posteriors of every utterance's first frame -- what PhoneLoop / BigramPhoneLoop.accumulate
        # reduce (phoneloop.py:83-101, 175-186)
        self._xi_block = unit_weights is not None and not aligned and (P == 0 or getattr(unit_weights, 'needs_block', False))
        if self._xi_block:
            self._nE, self._nS = len(unit_weights.end_idxs), len(unit_weights.start_idxs)
            P = self._nE * self._nS + self._nS
        self.flat = torch.zeros(M * self.Q + P + 4, device=self.dev, dtype=f64)
        self.acc = self.flat[:M * self.Q].view(M, self.Q)
        self.unit_counts = self.flat[M * self.Q:M * self.Q + P] if P else None
        self.extras = self.flat[M * self.Q + P:]
        self.kl = torch.zeros(1, device=self.dev, dtype=f64)
        self.local_frames = len(utts)
        self.datasize = float(datasize) if datasize is not None else None
        lens = torch.as_tensor(utts.lengths, dtype=f64, device=self.dev)
        self.utt_len = lens
        self.inv_len = torch.where(lens > 0, 1.0 / lens.clamp(min=1.0), torch.zeros_like(lens))
        self.profile = None      # optional {stage: [(start_event, end_event), ...]} (bench.py)
        self.host_mode = not utts.X.is_cuda
        if self.host_mode and chunk_frames is None:
            # the whole shard as ONE chunk while two staging buffers of it are affordable (4.2 M frames of 40 dimensions:
            # 2 x 670 MB): the copy of the next VB iteration then runs under all the kernels of this one, and the
            # forward-backward keeps its full rounds of resident utterances (cfg3 with two chunks of 625 utterances:
            # 9.1e7 frames/s end to end, one chunk: 1.0e8); larger shards in ~4 M-frame chunks
            chunk_frames = max(1, min(len(utts), 4_200_000))
        self._chunks = self._make_chunks(chunk_frames)
        nmax = max((c[3] for c in self._chunks), default=0)
        if self.host_mode:
            if not utts.X.is_pinned():
                raise ValueError('host-resident features must be in pinned memory')
            self._stage_buf = [torch.empty(nmax, D, device=self.dev, dtype=f32) for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=self.dev)
            self._ready = [torch.cuda.Event() for _ in range(2)]
            self._free = [torch.cuda.Event() for _ in range(2)]
            self._free_valid = [False, False]
            self._copy_seq, self._pending = 0, None      # copies issued so far; staging buffer of the chunk in flight
        Kp = emission.Kp if not self.gmm else emission.M // emission.gmm_C
        self.pdf_llh = torch.empty(nmax, Kp, device=self.dev, dtype=f32)
        self._nonident = not self.gmm and (aligned or not (plan.info['map_identity'] and plan.n_states == Kp))
        self.pdf_post = (torch.zeros if self._nonident else torch.empty)(nmax, Kp, device=self.dev, dtype=f32)
        # mixtures through the fp16-split kernels (forward-backward over one graph plan): no per-Gaussian llhs at all,
        # `pdf_llh` holds log2 values; the feature images of resident chunks are built once, here
        self.mix16 = None
        # (Viterbi training of mixtures takes the fp16 emission kernel only; its statistics are sparse along the path)
        if self.gmm or (emission.use16 and not aligned and not (self.viterbi and emission.uniform_C == 1)):
            self.mix16 = ops.Mix16(M, D, emission.gmm_C if self.gmm else emission.uniform_C, self.dev)
            self._images = [None] * len(self._chunks)
            if not self.host_mode:
                for ci, (u0, u1, f0, nf, rel) in enumerate(self._chunks):
                    if nf > 0:
                        self._images[ci] = self.mix16.build_images(utts.X[f0:f0 + nf])
            self._stage_images = None
        # skip (frame tile, Gaussian tile) pairs without posterior mass in the mixture statistics kernel (they are exact
        # zeros: see beer_mix16_accumulate_blocks); BEER_B200_DENSE_STATS=1 runs every pair
        self.sparse_stats = (os.environ.get('BEER_B200_DENSE_STATS') is None) if sparse_stats is None else bool(sparse_stats)
        self._blocks = None
        self.active_fraction = None
        self.tensor_kind = 'f16' if self.mix16 is not None else 'tf32'
        self.comp_llh = (torch.empty(nmax, M, device=self.dev, dtype=f32)
                         if emission.has_mixtures and self.mix16 is None else None)
        if self.per_utt:
            ws_bytes = max([p.workspace_bytes(int(t)) for p, t in zip(plan, utts.lengths)] + [4])
            self._one_off = [torch.tensor([0, int(t)], dtype=i64, device=self.dev) for t in utts.lengths]
        else:
            ws_bytes = plan.workspace_bytes(nmax) if not self.gmm else 4
        if self.viterbi:
            ws_bytes = max(ws_bytes, nmax * plan.n_states * 2 + 4)      # uint16 back-pointers
            self._pdf_map = torch.as_tensor(np.asarray(plan.pdf_map), dtype=i32, device=self.dev)
            self._frame_llh = torch.empty(nmax, device=self.dev, dtype=f32)
            # single-Gaussian pdfs in one statistics tile: the one-hot posteriors are never written, KC reads pdf ids
            self._path_kc = (not emission.has_mixtures) and ops.accumulate_path_supported(M, D)
            # mixtures: only the Gaussians of the frame's pdf carry weight -- sparse statistics along the path
            self._path_mix = self.mix16 is not None and emission.uniform_C > 1 and D <= 64
            self._pdf_ids = (torch.zeros(nmax + 4, device=self.dev, dtype=i32)
                             if (self._path_kc or self._path_mix) else None)
        self.ws = torch.empty((ws_bytes + 3) // 4, device=self.dev, dtype=f32)
        self.frame_ref = torch.empty(nmax, device=self.dev, dtype=f32)
        self.utt_ell = torch.zeros(utts.n_utts, device=self.dev, dtype=f64)
        self.gpu_launches = 0
        # the SIMT emission kernel reads the component offsets back to size its grid: not capturable
        self.use_graph = (bool(use_graph) and (emission.use_tc or not emission.has_mixtures or self.mix16 is not None)
                          and unit_weights is None)      # the graph rewrite syncs with the host
        self._shard_counts = torch.tensor([float(self.local_frames), float(utts.n_utts)], device=self.dev, dtype=f64)
        self._graph, self._graph_elbo, self._graph_launches, self._eager_steps = None, None, 0, 0

    def _make_chunks(self, chunk_frames):
        """Split the shard into runs of whole utterances of at most `chunk_frames` frames:
        (utt_begin, utt_end, frame_begin, n_frames, offsets tensor rebased to the chunk)."""
        off = self.utts.offsets_host
        n = self.utts.n_utts
        if chunk_frames is None:
            chunk_frames = 1 << 62
        chunks, u0 = [], 0
        while u0 < n:
            u1 = u0 + 1
            while u1 < n and off[u1 + 1] - off[u0] <= chunk_frames:
                u1 += 1
            rel = torch.as_tensor(off[u0:u1 + 1] - off[u0], dtype=i64, device=self.dev)
            chunks.append((u0, u1, int(off[u0]), int(off[u1] - off[u0]), rel))
            u0 = u1
        return chunks

    def _stage(self, name):
        return _StageTimer(self.profile, name)

    # -- one VB iteration -----------------------------------------------------
    def e_step(self):
        """Accumulate statistics and ELBO terms of the local shard into `self.flat`."""
        em, plan = self.em, self.plan
        self.flat.zero_()
        self.kl.zero_()
        # single-Gaussian pdfs keep the 3xTF32 emission kernel (one weight chunk per frame tile: the fp16 kernel's
        # per-tile statistics gather is not hidden there, measured 1.26 vs 0.60 ms on cfg2) and take only the statistics
        # kernel from the fp16 path
        ka16 = self.mix16 is not None and self.mix16.C > 1
        W, bias, ref = em.refresh(pack_tc=not ka16)
        em.kl(out=self.kl)
        if self.units is not None:
            self.units.kl(self.kl)
        self.gpu_launches += 2 + 1 + 2 * len(em.weight_groups) + int(em.use_tc)
        nonident = self._nonident
        chunks = [c for c in self._chunks if c[3] > 0]
        chunk_ids = [i for i, c in enumerate(self._chunks) if c[3] > 0]
        if self.host_mode and chunks and self._pending is None:
            self._pending = self._next_copy(chunks[0])
        for ci, (u0, u1, f0, nf, rel) in enumerate(chunks):
            if self.host_mode:
                b = self._pending
                nxt = chunks[ci + 1] if ci + 1 < len(chunks) else (chunks[0] if self.prefetch else None)
                self._pending = self._next_copy(nxt) if nxt is not None else None
                torch.cuda.current_stream().wait_event(self._ready[b])
                X = self._stage_buf[b][:nf]
            else:
                X = self.utts.X[f0:f0 + nf]
            pdf_llh = self.pdf_llh[:nf]
            pdf_post = self.pdf_post[:nf]
            comp = self.comp_llh[:nf] if self.comp_llh is not None else None
            images = None
            direct = False
            blocks = None
            nmax_tiles = (max(c[3] for c in self._chunks) + 63) // 64
            if self.mix16 is not None:
                if self.host_mode:       # streamed features: the images of the chunk are rebuilt behind its copy
                    with self._stage('KI_feature_images'):
                        images = self._stage_images = self.mix16.build_images(X, out=self._stage_images)
                    self.gpu_launches += 4
                else:
                    images = self._images[chunk_ids[ci]]
            with self._stage('KA_emission_llh'):
                if ka16:
                    self.mix16.pack(W, bias, images['alpha'])
                    fref = self.mix16.frame_ref(X, ref, out=self.frame_ref[:nf])
                    self.mix16.emission(images, out=pdf_llh)
                    self.gpu_launches += 2
                else:
                    _, _, fref = em.llh(X, W, bias, ref, pdf_llh, comp, self.frame_ref[:nf])
            if nonident:
                pdf_post.zero_()
            with self._stage('KB_forward_backward'):
                if self.gmm:
                    # no graph: the softmax over the pseudo-pdfs of a frame is the whole "inference"
                    self.utt_ell[u0:u1].zero_()
                    self.mix16.gmm_posteriors(pdf_llh, fref, rel, scale=self.scale, out=pdf_post,
                                              out_utt_exp_llh=self.utt_ell[u0:u1])
                    self.gpu_launches += 1
                elif self.viterbi and self._path_mix:
                    # the emission kernel wrote log2 llhs: the scale of the Viterbi recursion carries ln 2
                    path = ops.hmm_viterbi(plan, pdf_llh, rel, scale=self.scale * math.log(2.0), workspace=self.ws)
                    self._path_unit_counts(path, rel)
                    if plan.info['map_identity']:
                        self._pdf_ids[:nf].copy_(path)
                    else:
                        torch.index_select(self._pdf_map, 0, path, out=self._pdf_ids[:nf])
                elif self.viterbi:
                    path = ops.hmm_viterbi(plan, pdf_llh, rel, scale=self.scale, workspace=self.ws)
                    self._path_unit_counts(path, rel)
                    _, frame = ops.path_posteriors(path, em.Kp, pdf_map=self._pdf_map, scale=self.scale,
                                                   pdf_llh=pdf_llh, frame_ref=fref, want_post=not self._path_kc,
                                                   out_post=None if self._path_kc else pdf_post,
                                                   out_frame=self._frame_llh[:nf])
                    if self._path_kc:
                        torch.index_select(self._pdf_map, 0, path, out=self._pdf_ids[:nf])
                    # per-utterance sums of the per-frame expected llh (fp64 prefix sums, differences at the offsets)
                    cs = torch.cat([torch.zeros(1, dtype=f64, device=self.dev), frame.double().cumsum(0)])
                    self.utt_ell[u0:u1] = cs[rel[1:]] - cs[rel[:-1]]
                elif self.per_utt:
                    off_h = self.utts.offsets_host
                    for i in range(u0, u1):
                        a, b = int(off_h[i] - off_h[u0]), int(off_h[i + 1] - off_h[u0])
                        if b > a:
                            ops.hmm_forward_backward(plan[i], pdf_llh[a:b], fref[a:b], self._one_off[i], scale=self.scale,
                                                     workspace=self.ws, out_pdf_post=pdf_post[a:b],
                                                     out_utt_exp_llh=self.utt_ell[i:i + 1])
                    self.gpu_launches += u1 - u0
                elif self.chains:
                    ops.hmm_forward_backward_chains(plan, pdf_llh, fref, rel, scale=self.scale, first_utt=u0,
                                                    workspace=self.ws, out_pdf_post=pdf_post,
                                                    out_utt_exp_llh=self.utt_ell[u0:u1])
                else:
                    direct = (images is not None and not nonident and plan.writes_log2_posteriors
                              and self.mix16.C > 1)         # single-Gaussian pdfs: the statistics kernel takes pdf_post
                    # activity map: (tile of 64 frames, pdfs of one Gaussian tile) pairs with posterior mass; the statistics
                    # kernel skips the others (their weights are exactly zero in its fp16 operands)
                    blocks = None
                    if direct and self.sparse_stats and plan.marks_active_blocks(self.unit_counts is not None):
                        nb = (em.Kp + self.mix16.pdfs_per_block - 1) // self.mix16.pdfs_per_block
                        if self._blocks is None or self._blocks.shape[1] != nb:
                            self._blocks = torch.empty((nmax_tiles, nb), device=self.dev, dtype=torch.uint8)
                        blocks = self._blocks[:(nf + 63) // 64]
                        blocks.zero_()
                        self.gpu_launches += 1
                    r = ops.hmm_forward_backward(plan, pdf_llh, fref, rel, scale=self.scale, workspace=self.ws,
                                                 want_pdf_post=not direct, out_pdf_post=None if direct else pdf_post,
                                             out_pdf_lpost=pdf_post if direct else None,
                                             out_utt_exp_llh=self.utt_ell[u0:u1],
                                             unit_counts=None if self._xi_block else self.unit_counts,
                                             llh_log2=ka16, lpost_relative=direct, block_active=blocks,
                                             pdfs_per_block=self.mix16.pdfs_per_block if blocks is not None else 0,
                                             want_state_post=self._xi_block)
                    if self._xi_block:
                        self._block_counts(r['state_post'], pdf_llh, rel, u0, u1, self.scale * (math.log(2.0) if ka16 else 1.0))
                    if images is not None and not direct and self.mix16.C > 1:      # graphs without a loop kernel: log2 of pdf_post
                        self.mix16.log2_posteriors(pdf_post, out=pdf_post)
                        self.gpu_launches += 1
            with self._stage('KC_accumulate'):
                if self.viterbi and self._path_mix:
                    frame = ops.path_accumulate_mix(X, self._pdf_ids[:nf], W, bias, self.mix16.C, self.acc, frame_ref=fref,
                                                    scale=self.scale, out_frame=self._frame_llh[:nf])
                    cs = torch.cat([torch.zeros(1, dtype=f64, device=self.dev), frame.double().cumsum(0)])
                    self.utt_ell[u0:u1] = cs[rel[1:]] - cs[rel[:-1]]
                elif images is not None:
                    # `direct`: the scan wrote log2 posterior - log2 llh, the one array the statistics kernel adds to z
                    self.mix16.accumulate(images, pdf_post, None if direct else pdf_llh, self.acc, scale=self.scale,
                                          relative=direct, block_active=blocks)
                elif self.viterbi and self._path_kc:
                    ops.accumulate_stats_path(X, self.acc, self._pdf_ids[:nf], scale=self.scale)
                else:
                    ops.accumulate_stats(X, self.acc, pdf_post=pdf_post,
                                         pdf_llh=pdf_llh if comp is not None else None, comp_llh=comp,
                                         comp_off=em.comp_off if (comp is not None and not em.uniform_C) else None,
                                         Kp=em.Kp)
            self.gpu_launches += 3
            if blocks is not None and self.profile is not None:      # (outside the stage timers)
                self.active_fraction = blocks.float().mean()
            if self.host_mode:
                self._free[b].record()
                self._free_valid[b] = True
        # ELBO bookkeeping of the shard (objectives.py:176-190 summed as in accumulate.py:39-59)
        self.extras[1:3].copy_(self._shard_counts)       # device -> device: capturable in a CUDA graph
        self.extras[3] = self.utt_ell.sum()
        # sum_u ell_u / T_u; multiplied by the global datasize after the reduction
        self.extras[0] = (self.utt_ell * self.inv_len).sum()

    def _block_counts(self, state_post, pdf_llh, rel, u0, u1, scale, max_floats=1 << 26):
        """unit_counts[:E * S] += sum_t xi_t[ends, starts], unit_counts[E * S:] += sum_u gamma_0[starts] over the
        utterances of a chunk.  The per-step block [frames, E, S] is formed for groups of utterances of at most
        `max_floats` elements and summed in float64."""
        u = self.units
        nE, nS = self._nE, self._nS
        g = u.graph
        if not hasattr(self, '_xi_rows'):
            self._xi_rows = torch.as_tensor(u.end_idxs, dtype=i32, device=self.dev)
            self._xi_cols = torch.as_tensor(u.start_idxs, dtype=i32, device=self.dev)
            self._xi_map = torch.as_tensor(np.asarray(g.pdf_id_mapping), dtype=i32, device=self.dev)
        init = g.init_log_probs.detach().to(device=self.dev, dtype=f32).contiguous()
        trans = g.trans_log_probs.detach().to(device=self.dev, dtype=f32).contiguous()     # rewritten every iteration
        block = self.unit_counts[:nE * nS].view(nE, nS)
        first = self.unit_counts[nE * nS:]
        oh = self.utts.offsets_host
        off_h = [int(oh[k] - oh[u0]) for k in range(u0, u1 + 1)]
        n_utts = len(off_h) - 1
        per_frame = max(nE * nS, 1)
        i = 0
        while i < n_utts:
            j = i + 1
            while j < n_utts and (off_h[j + 1] - off_h[i]) * per_frame <= max_floats:
                j += 1
            a, b = off_h[i], off_h[j]
            if b > a:
                xi = ops.hmm_transition_posteriors(pdf_llh[a:b], state_post[a:b], (rel[i:j + 1] - a).contiguous(), init, trans,
                                                   pdf_map=self._xi_map, scale=scale, rows=self._xi_rows, cols=self._xi_cols)
                block += xi.sum(dim=0, dtype=f64)
                self.gpu_launches += 2
            i = j
        starts = rel[:-1][rel[1:] > rel[:-1]]               # first frame of every non-empty utterance
        first += state_post[starts][:, self._xi_cols.long()].sum(dim=0, dtype=f64)

    def _path_unit_counts(self, path, rel):
        """Unit counts of a state path: its one-hot transition posteriors summed over the ends x starts block, plus the
        first frame of every utterance (hmm.py:49-54, phoneloop.py:83-101).  Index arithmetic on [N] integers."""
        if self.unit_counts is None:
            return
        p = path.long()
        n = p.numel()
        first = torch.zeros(n + 1, dtype=torch.bool, device=self.dev)
        first[rel[:-1]] = True            # an empty utterance marks the next one's first frame (or the slot past the end)
        if self._xi_block:
            # block counting: (unit whose end state frame t - 1 is in, unit whose start state frame t is in) pairs
            nE, nS = self._nE, self._nS
            if not hasattr(self, '_end_of'):
                K = self.plan.n_states
                self._end_of = torch.full((K,), -1, dtype=i64, device=self.dev)
                self._end_of[torch.as_tensor(self.units.end_idxs, device=self.dev)] = torch.arange(nE, device=self.dev)
                self._start_of = torch.full((K,), -1, dtype=i64, device=self.dev)
                self._start_of[torch.as_tensor(self.units.start_idxs, device=self.dev)] = torch.arange(nS, device=self.dev)
            s_unit = self._start_of[p]
            e_unit = torch.full_like(s_unit, -1)
            e_unit[1:] = self._end_of[p[:-1]]
            hit = (s_unit >= 0) & (e_unit >= 0) & ~first[:n]
            self.unit_counts[:nE * nS].index_add_(0, (e_unit * nS + s_unit).clamp(min=0), hit.to(f64))
            self.unit_counts[nE * nS:].index_add_(0, s_unit.clamp(min=0), ((s_unit >= 0) & first[:n]).to(f64))
            return
        su = self.plan.n_states // self.unit_counts.numel()
        is_start = (p % su) == 0
        after_end = torch.zeros(n, dtype=torch.bool, device=self.dev)
        after_end[1:] = (p[:-1] % su) == su - 1
        hit = is_start & (after_end | first[:n])
        self.unit_counts.index_add_(0, p // su, hit.to(f64))

    def _next_copy(self, chunk):
        """Issue the copy of `chunk` into the staging buffer whose turn it is; returns the buffer."""
        b = self._copy_seq & 1
        self._copy_seq += 1
        self._issue_copy(chunk, b)
        return b

    def _issue_copy(self, chunk, b):
        """H2D copy of one chunk into staging buffer b on the copy stream."""
        _, _, f0, nf, _ = chunk
        with torch.cuda.stream(self._copy_stream):
            if self._free_valid[b]:
                self._copy_stream.wait_event(self._free[b])   # the kernels that last read b are done
            self._stage_buf[b][:nf].copy_(self.utts.X[f0:f0 + nf], non_blocking=True)
            self._ready[b].record()

    def reduce(self):
        """The one exchange step: sum the flat statistics buffer over ranks (NCCL)."""
        if self.distributed:
            torch.distributed.all_reduce(self.flat, group=self.pg)

    def m_step(self):
        """Replicated on every rank from the reduced statistics (no broadcast needed).
        Returns the summed ELBO as a device scalar (no host sync)."""
        total_frames = self.extras[1]
        n_utts = self.extras[2]
        # python float for the kernel arguments would force a sync; stats_scale is known on the
        # host because datasize and the frame counts are host-side constants of the run.
        datasize = self.datasize if self.datasize is not None else self._global_frames()
        stats_scale = stats_scale_from_flat(self._global_frames(), datasize)
        elbo = elbo_from_flat(self.extras, self.kl[0], datasize)
        self.em.update(self.acc, stats_scale, self.lrate)
        self.gpu_launches += 1 + (1 + len(self.em.weight_groups) if self.em.weight_groups else 0)
        if self.units is not None:
            # unit counts in the order of the weights -> Dirichlet statistics (last entry = total,
            # dirichlet.py:18-21) -> natural-gradient step -> rewrite the graph -> new device plan
            u = self.units
            if self.chains or self.per_utt:
                # aligned training: zero statistics (phoneloop.py:98-100, 187-190)
                shape = (len(u.end_idxs), len(u.start_idxs)) if getattr(u, 'needs_block', False) else (len(u.start_idxs),)
                counts = torch.zeros(shape, device=self.dev, dtype=f64)
            elif self._xi_block:
                block = self.unit_counts[:self._nE * self._nS].view(self._nE, self._nS)
                counts = block if getattr(u, 'needs_block', False) else block.sum(dim=0) + self.unit_counts[self._nE * self._nS:]
            else:
                su = self.plan.n_states // self.unit_counts.numel()
                order = torch.as_tensor([s // su for s in u.start_idxs], device=self.dev)
                counts = self.unit_counts[order]
            u.update(counts, stats_scale, self.lrate)
            u.rewrite_graph()
            if not (self.chains or self.per_utt):
                self.plan = u.graph.plan(n_pdfs=self.em.Kp)
            self.gpu_launches += 2
        return elbo

    def _global_frames(self):
        if not hasattr(self, '_gframes'):
            n = torch.tensor([float(self.local_frames)], device=self.dev, dtype=f64)
            if self.distributed:
                torch.distributed.all_reduce(n, group=self.pg)
            self._gframes = float(n.item())
        return self._gframes

    def step(self):
        """E-step + all-reduce + M-step; returns the summed ELBO (device fp64 scalar).

        With `use_graph` (resident features, no stage profiling) the ~16 launches of an iteration
        are captured once into a CUDA graph and replayed: the kernels are unchanged, only the
        launch gaps between the small parameter kernels go away.  The returned tensor is then a
        static buffer that the next step overwrites."""
        self._global_frames()
        if self.use_graph and not self.host_mode and self.profile is None:
            if self._graph is None and self._eager_steps >= 1:
                self._capture()
            if self._graph is not None:
                self._graph.replay()
                self.gpu_launches += self._graph_launches
                return self._graph_elbo
        self._eager_steps += 1
        return self._step_eager()

    def _step_eager(self):
        self.e_step()
        self.reduce()
        return self.m_step()

    def _capture(self):
        """Capture one iteration (side stream, as CUDA requires) after an eager one has run."""
        before = self.gpu_launches
        try:
            torch.cuda.synchronize(self.dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                elbo = self._step_eager()
            self._graph, self._graph_elbo = graph, elbo
            self._graph_launches = self.gpu_launches - before
        except Exception as exc:       # capture is an optimisation: fall back to plain launches
            import warnings
            warnings.warn(f'CUDA graph capture of the VB iteration failed ({exc}); launching eagerly')
            self.use_graph = False
        self.gpu_launches = before

    def elbo_per_frame(self, elbo):
        """The figure `beer hmm update` logs (update.py:72): ELBO / (n_utts * datasize)."""
        datasize = self.datasize if self.datasize is not None else self._global_frames()
        return elbo / (self.extras[2] * datasize)
