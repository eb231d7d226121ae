"""`beer hmm accumulate` + `beer hmm update` as ONE multi-GPU command over the reference's own files.

    python -m beer_b200.hmm_train [-a ALIS.npz] [-s SCALE] [-l LRATE] [-e EPOCHS] [-u UTTIDS] MODEL DATASET OUT_MODEL
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m beer_b200.hmm_train ...

What the recipes do per epoch with a job array and pickles on a shared file system (`beer hmm accumulate` per split:
beer/cli/subcommands/hmm/accumulate.py:22-63, then `beer hmm update`: update.py:22-72) is one `VBEngine.step()` here:
the utterances are dealt to the ranks, every rank keeps its shard resident in HBM, the statistics meet in ONE
all-reduce and every rank takes the same natural-gradient step.  MODEL is a model pickled by the reference (`beer hmm
mkphoneloop`, or an earlier `update`), DATASET the pickle of `beer dataset create`, ALIS the archive of `beer hmm
mkaligraph`; OUT_MODEL is written by rank 0 as a pickle the reference loads (`beer hmm decode`, the next `accumulate`).
Neither side needs the other installed (beer_b200/refpickle.py).

Covered: HMM / PhoneLoop over NormalSet (diagonal) / MixtureSet / JointModelSet emissions; unit weights with a
Dirichlet, stick-breaking or Gamma-stick-breaking prior (the CLI default); forward-backward over the decoding graph or
over per-utterance alignment graphs (chains in one launch; any other alignment graph utterance by utterance).  Not
covered (use the model API): BigramPhoneLoop, full covariances.
"""
import argparse
import os
import sys

import numpy as np
import torch

from . import refpickle
from .dataset import Alignments, Dataset

f32, f64 = torch.float32, torch.float64


def load_dataset(path):
    """The pickle of `beer dataset create` (create.py:44-60: a dataclass with feapath / mean / var / size).  The stored
    path is absolute on the machine that created it: when it does not exist here, the archive of the same name next
    to the pickle is taken."""
    obj = refpickle.load(path)
    d = obj.__dict__
    feapath = d['feapath']
    if not os.path.exists(feapath):
        local = os.path.join(os.path.dirname(os.path.abspath(path)), os.path.basename(feapath))
        if not os.path.exists(local):
            raise FileNotFoundError(f'features archive {feapath} (nor {local})')
        feapath = local
    return Dataset(feapath, d['mean'], d['var'], int(d['size']))


def _categorical_from_reference(cat, device):
    """The unit-weight model of a pickled PhoneLoop as a model of `beer_b200.models` (same parameters)."""
    from . import models
    from .dists import Dirichlet, Gamma
    from .parameters import ConjugateBayesianParameter
    V = refpickle.ModelView
    kind = type(cat).__qualname__

    def stats_of(p):       # (given explicitly: sizing them from the natural parameters would launch a kernel)
        return p._buffers['stats'].detach().to(device, f64).clone()

    def dirichlet_param(p):
        return ConjugateBayesianParameter(
            Dirichlet.from_std_parameters(V.concentrations(p, 'prior').to(device, f32).clone()),
            Dirichlet.from_std_parameters(V.concentrations(p, 'posterior').to(device, f32).clone()),
            init_stats=stats_of(p))

    if kind == 'Categorical':
        return models.Categorical(dirichlet_param(cat._modules['weights']))
    if kind == 'CategoricalSet':          # the bigram weights of `beer hmm mkphoneloopbigram --weights-prior dirichlet2`
        return models.CategoricalSet(dirichlet_param(cat._modules['weights']))
    if kind in ('SBCategorical', 'SBCategoricalHyperPrior'):
        sb = dirichlet_param(cat._modules['stickbreaking'])
        if kind == 'SBCategorical':
            out = models.SBCategorical(sb)
        else:
            c = cat._modules['concentration']
            pri, pos = c._modules['prior'].params._buffers, c._modules['posterior'].params._buffers
            conc = ConjugateBayesianParameter(
                Gamma.from_std_parameters(pri['shape'].to(device, f32).clone(), pri['rate'].to(device, f32).clone()),
                Gamma.from_std_parameters(pos['shape'].to(device, f32).clone(), pos['rate'].to(device, f32).clone()),
                init_stats=stats_of(c))
            out = models.SBCategoricalHyperPrior(sb, conc)
        out.ordering = cat.__dict__['ordering'].to(device)
        return out
    raise NotImplementedError(f'unit weights of type {kind}: use the model API')


def _categorical_to_reference(model_cat, cat):
    """Write the parameters of the trained unit-weight model back into the pickled tree."""
    V = refpickle.ModelView
    kind = type(cat).__qualname__

    def put(dist_ref, name, value):
        old = dist_ref.params._buffers[name]
        dist_ref.params._buffers[name] = value.detach().to(device=old.device, dtype=old.dtype).reshape(old.shape).clone()

    def put_stats(param_ref, param):
        old = param_ref._buffers['stats']
        param_ref._buffers['stats'] = param.stats.detach().to(device=old.device, dtype=old.dtype).clone()

    if kind in ('Categorical', 'CategoricalSet'):
        V.set_concentrations(cat._modules['weights'], model_cat.weights.posterior.params.concentrations)
        put_stats(cat._modules['weights'], model_cat.weights)
        return
    sb_ref, sb = cat._modules['stickbreaking'], model_cat.stickbreaking
    put(sb_ref._modules['prior'], 'concentrations', sb.prior.params.concentrations)       # (hyper-prior: second column)
    put(sb_ref._modules['posterior'], 'concentrations', sb.posterior.params.concentrations)
    put_stats(sb_ref, sb)
    cat.__dict__['ordering'] = model_cat.ordering.detach().cpu().to(cat.__dict__['ordering'].dtype)
    if kind == 'SBCategoricalHyperPrior':
        c_ref, c = cat._modules['concentration'], model_cat.concentration
        for name in ('shape', 'rate'):
            put(c_ref._modules['posterior'], name, getattr(c.posterior.params, name))
        put_stats(c_ref, c)


class ReferenceModel:
    """A pickled reference HMM-GMM model opened for training on the engine: flat device tensors of its parameters
    (`emission`, `unit_weights`, `graph`) and `save()` that writes them back under the reference's class names."""

    def __init__(self, path, device):
        from .engine import BigramUnitWeights, CategoricalUnitWeights, EmissionParams, WeightGroup
        from .graph import CompiledGraph
        self.device = device
        self.tree = refpickle.load(path)
        self.view = v = refpickle.ModelView(self.tree)
        init, final, trans, pdf_map = v.graph_arrays()
        self.graph = CompiledGraph(init.detach().float().cpu().clone(), final.detach().float().cpu().clone(),
                                   trans.detach().float().cpu().clone(), pdf_map)

        def cat(which, i):
            return torch.cat([v.normal_gamma(g['normal'], which)[i].to(device, f32) for g in v.groups]).contiguous()

        prior = tuple(cat('prior', i) for i in range(4))
        post = tuple(cat('posterior', i) for i in range(4))
        comp_off, groups, k0 = [0], [], 0
        for g in v.groups:
            for _ in range(g['n_pdfs']):
                comp_off.append(comp_off[-1] + g['n_comp'])
            if g['weights'] is not None:
                groups.append(WeightGroup(k0, g['n_pdfs'], g['n_comp'],
                                          v.concentrations(g['weights'], 'prior').to(device, f32).clone(),
                                          v.concentrations(g['weights'], 'posterior').to(device, f32).clone()))
            k0 += g['n_pdfs']
        mixtures = any(g['n_comp'] > 1 or g['weights'] is not None for g in v.groups)
        self.emission = EmissionParams(prior, post, comp_off=np.asarray(comp_off) if mixtures else None,
                                       weight_groups=groups)
        self.n_pdfs = k0
        self.categorical = self.unit_weights = None
        if v.kind == 'PhoneLoop':
            self.categorical = _categorical_from_reference(v.categorical, device)
            self.unit_weights = CategoricalUnitWeights(self.categorical, self.graph, list(v.start_pdf.values()),
                                                       list(v.end_pdf.values()))
        elif v.kind == 'BigramPhoneLoop':
            # one Dirichlet per unit over the unit starts (phoneloop.py:105-191); the engine sums the ends x starts
            # block of the transition posteriors for it
            self.categorical = _categorical_from_reference(v.categorical, device)
            self.unit_weights = BigramUnitWeights(self.categorical, self.graph, list(v.start_pdf.values()),
                                                  list(v.end_pdf.values()))

    def save(self, path, acc=None, stats_scale=1.0):
        """Posteriors (and, as `update` leaves them, the stored statistics) back into the tree, then the pickle."""
        from . import ops
        v, em = self.view, self.emission
        j0 = 0
        wg = {g.pdf_start: g for g in em.weight_groups}
        k0 = 0
        wst = None
        if acc is not None and em.weight_groups:      # Dirichlet statistics of the mixture weights (mixtureset.py:100-112)
            wst = ops.mixture_weight_stats(acc, em.D, comp_off=em.comp_off, Kp=em.Kp)
        for g in v.groups:
            m = g['n_pdfs'] * g['n_comp']
            v.set_normal_gamma(g['normal'], *(t[j0:j0 + m] for t in em.post))
            if acc is not None:
                old = g['normal']._buffers['stats']
                g['normal']._buffers['stats'] = (acc[j0:j0 + m] * stats_scale).to(device=old.device, dtype=old.dtype)
            if g['weights'] is not None:
                v.set_concentrations(g['weights'], wg[k0].post)
                if wst is not None:
                    old = g['weights']._buffers['stats']
                    g['weights']._buffers['stats'] = (wst[j0:j0 + m] * stats_scale).reshape(old.shape).to(
                        device=old.device, dtype=old.dtype)
            j0 += m
            k0 += g['n_pdfs']
        if self.categorical is not None:
            _categorical_to_reference(self.categorical, v.categorical)
            b = v.graph._buffers                           # the end -> start arcs rewritten from E[ln w]
            b['trans_log_probs'] = self.graph.trans_log_probs.detach().to(
                device=b['trans_log_probs'].device, dtype=b['trans_log_probs'].dtype).clone()
        refpickle.dump(self.tree, path)


def _read_uttids(args, dataset):
    if args.uttids:
        with open(args.uttids) as f:
            lines = f.readlines()
    elif not sys.stdin.isatty() and int(os.environ.get('WORLD_SIZE', '1')) == 1:
        try:
            lines = sys.stdin.readlines()      # as `beer hmm accumulate`: one utterance id per line
        except OSError:                        # (no usable stdin, e.g. under a test runner: the whole data set)
            lines = []
    else:
        lines = []
    ids = [line.strip().split()[0] for line in lines if line.strip()]
    return ids if ids else sorted(dataset.fea_dict.keys())


def main(argv=None):
    ap = argparse.ArgumentParser(prog='python -m beer_b200.hmm_train', description=__doc__.split('\n\n')[0])
    ap.add_argument('-a', '--alis', help='alignment graphs in a "npz" archive (beer hmm mkaligraph)')
    ap.add_argument('-s', '--acoustic-scale', default=1., type=float, help='scaling factor of the acoustic model')
    ap.add_argument('-l', '--learning-rate', default=1., type=float, help='learning rate of the update')
    ap.add_argument('-e', '--epochs', default=1, type=int, help='accumulate + update rounds')
    ap.add_argument('-u', '--uttids', help='file of utterance ids (default: stdin when piped, else the whole data set)')
    ap.add_argument('model', help='hmm based model pickled by the reference')
    ap.add_argument('dataset', help='data set pickled by `beer dataset create`')
    ap.add_argument('out_model', help='updated model (pickle the reference loads)')
    args = ap.parse_args(argv)

    from .engine import Utterances, VBEngine, bind_to_gpu_cpus, shard_utterances
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('beer_b200.hmm_train needs a CUDA device (there is no CPU path)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        bind_to_gpu_cpus(local)
        torch.distributed.init_process_group('nccl', device_id=dev)

    dataset = load_dataset(args.dataset)
    model = ReferenceModel(args.model, dev)
    alis = Alignments(args.alis) if args.alis else None
    ids = _read_uttids(args, dataset)
    known = set(dataset.fea_dict.keys())
    for u in ids:
        if u not in known and rank == 0:
            print(f'warning: no utterance {u} in {args.dataset}', file=sys.stderr)
    ids = [u for u in ids if u in known]
    if alis is not None:
        for u in ids:
            if u not in alis and rank == 0:
                print(f'warning: no alignment graph for utterance "{u}": skipped', file=sys.stderr)
        ids = [u for u in ids if u in alis]
    lens = [len(dataset.fea_dict[u]) for u in ids]
    mine = shard_utterances(lens, world)[rank]
    feats = [np.asarray(dataset.fea_dict[ids[i]], dtype=np.float32) for i in mine]
    X = torch.from_numpy(np.concatenate(feats) if feats else np.zeros((0, model.emission.D), np.float32))
    utts = Utterances(X, [lens[i] for i in mine], device=dev)
    if alis is not None:
        try:
            plan = alis.chain_batch([ids[i] for i in mine], dev)       # all chains in one launch
        except ValueError:
            # alignment graphs that are not plain left-to-right chains (units with skip arcs, e.g. the silence model of
            # recipes/timit_v2/conf_61phns/hmm_gmm/hmm.yml): one graph plan and one forward-backward launch per utterance
            plan = [alis[ids[i]].plan(n_pdfs=model.n_pdfs) for i in mine]
    else:
        plan = model.graph.plan(n_pdfs=model.n_pdfs)
    engine = VBEngine(model.emission, plan, utts, datasize=float(dataset.size), scale=args.acoustic_scale,
                      lrate=args.learning_rate, unit_weights=model.unit_weights, distributed=world > 1)
    for epoch in range(args.epochs):
        elbo = engine.step()
        if rank == 0:
            # the figure `beer hmm update` logs (update.py:72)
            print(f'epoch {epoch + 1}: accumulated ELBO={float(engine.elbo_per_frame(elbo)):.3f}', flush=True)
    if rank == 0:
        from .engine import stats_scale_from_flat
        scale = stats_scale_from_flat(engine._global_frames(), float(dataset.size))
        model.save(args.out_model, acc=engine.acc, stats_scale=scale)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
