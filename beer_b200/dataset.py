"""Data-set format of the reference CLI (beer/cli/dataset.py:38-80, subcommands/dataset/create.py:14-38):
an `.npz` archive of per-utterance float arrays [T_u, D] plus the global mean / variance / frame count.
Host-side I/O only; `Dataset.shard(rank, world)` returns this rank's utterances as the ragged batch the
engine keeps resident in HBM (the `split` step of the reference's job arrays)."""
import contextlib
import sys

import numpy as np
import torch

from .engine import Utterances, shard_utterances

__all__ = ['Dataset', 'Utterance', 'Alignments']


class Utterance:
    """An utterance id with its features (dataset.py:10-15)."""

    def __init__(self, id, features):
        self.id, self.features = id, features


class Dataset:
    """Utterances of a features archive with the statistics `beer dataset create` stores."""

    def __init__(self, feapath, mean=None, var=None, size=None):
        self.feapath = feapath
        self._fea = None
        if mean is None or var is None or size is None:
            mean, var, size = self.accumulate(feapath)
        self.mean, self.var, self.size = mean, var, size

    @staticmethod
    def accumulate(feapath):
        """Global mean, variance and number of frames (create.py:14-38)."""
        feats = np.load(feapath)
        keys = list(feats.keys())
        dim = feats[keys[0]].shape[1]
        tot, tot2, n = np.zeros(dim), np.zeros(dim), 0
        for k in keys:
            x = feats[k]
            tot += x.sum(axis=0)
            tot2 += (x ** 2).sum(axis=0)
            n += len(x)
        mean = tot / n
        return torch.from_numpy(mean).float(), torch.from_numpy(tot2 / n - mean ** 2).float(), int(n)

    @property
    def fea_dict(self):
        if self._fea is None:
            self._fea = np.load(self.feapath)
        return self._fea

    def __getstate__(self):
        state = self.__dict__.copy()
        state['_fea'] = None
        return state

    def __len__(self):
        return len(self.fea_dict.files)

    def __getitem__(self, key):
        return Utterance(key, torch.from_numpy(self.fea_dict[key]).float())

    def utterances(self, random_order=False):
        ids = sorted(self.fea_dict.keys())
        if random_order:
            import random
            random.shuffle(ids)
        for uttid in ids:
            yield self[uttid]

    def shard(self, rank=0, world_size=1, device='cuda', pinned_host=False):
        """(utterance ids, Utterances) of this rank: utterances dealt to ranks so that the frames per rank
        are balanced.  `pinned_host`: keep the frames in pinned host memory (the engine then streams them)."""
        ids = sorted(self.fea_dict.keys())
        lens = [len(self.fea_dict[i]) for i in ids]
        mine = shard_utterances(lens, world_size)[rank]
        feats = [np.asarray(self.fea_dict[ids[i]], dtype=np.float32) for i in mine]
        dim = self.fea_dict[ids[0]].shape[1]
        X = torch.from_numpy(np.concatenate(feats) if feats else np.zeros((0, dim), np.float32))
        if pinned_host:
            return [ids[i] for i in mine], Utterances(X.pin_memory(), [lens[i] for i in mine])
        return [ids[i] for i in mine], Utterances(X, [lens[i] for i in mine], device=device)


@contextlib.contextmanager
def _reference_module_names():
    """While an archive written by the reference is unpickled, `beer.graph.CompiledGraph` resolves to this
    package's class (same buffers, same attribute names), so the reference need not be installed."""
    from . import graph as _graph
    import beer_b200 as _pkg
    saved = {k: sys.modules.get(k) for k in ('beer', 'beer.graph')}
    if saved['beer'] is None:
        sys.modules['beer'], sys.modules['beer.graph'] = _pkg, _graph
    try:
        yield
    finally:
        if saved['beer'] is None:
            for k in ('beer', 'beer.graph'):
                sys.modules.pop(k, None)


class Alignments:
    """Alignment graphs of `beer hmm mkaligraph` as the recipes archive them for `beer hmm accumulate --alis`
    (mkaligraph.py:40-63, accumulate.py:33-51): an npz with one `np.array([CompiledGraph])` per utterance id.
    `alis[uttid]` is the utterance's CompiledGraph; `chain_batch(uttids, device)` flattens the graphs of a shard
    into the per-utterance chains the batched engine runs in one launch (ops.ChainBatch)."""

    def __init__(self, path):
        self.path = path
        with _reference_module_names():
            archive = np.load(path, allow_pickle=True)
            self._graphs = {k: archive[k][0] for k in archive.files}

    def __contains__(self, uttid):
        return uttid in self._graphs

    def __len__(self):
        return len(self._graphs)

    def keys(self):
        return self._graphs.keys()

    def __getitem__(self, uttid):
        return self._graphs[uttid]

    def chain_batch(self, uttids, device):
        from . import ops
        return ops.ChainBatch([self._graphs[u] for u in uttids], device)


def create_dataset(feapath, out):
    """`beer dataset create` (beer/cli/subcommands/dataset/create.py:44-60): global mean / variance / frame count of a
    features archive, pickled as the reference's `beer.cli.dataset.Dataset` (a dataclass: feapath, mean, var, size) so
    that `beer hmm accumulate`, `beer hmm decode` and `python -m beer_b200.hmm_train` all read it."""
    import os
    from . import refpickle
    mean, var, size = Dataset.accumulate(feapath)
    obj = refpickle.new_object('beer.cli.dataset', 'Dataset', feapath=os.path.abspath(feapath), mean=mean, var=var,
                               size=size, _fea_dict=None)
    refpickle.dump(obj, out)
    return obj


if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser(prog='python -m beer_b200.dataset', description='compile a data set with the given features')
    ap.add_argument('features', help='features archive (npz format)')
    ap.add_argument('out', help='output compiled dataset')
    a = ap.parse_args()
    d = create_dataset(a.features, a.out)
    print(f'created dataset (total frame count: {d.size})')
