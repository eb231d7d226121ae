"""Model / data-set pickles of the reference, read and written WITHOUT the reference installed.

The recipes pass models between `beer hmm mkphoneloop`, `beer hmm accumulate`, `beer hmm update` and `beer hmm decode`
as `pickle.dump(model)` files (beer/cli/subcommands/hmm/mkphoneloop.py:82-83, accumulate.py:23-25, update.py:64-66).
A reference model is a tree of `torch.nn.Module`s whose pickled state is the instance `__dict__` (`_buffers`,
`_modules`, plain attributes); this module unpickles such a file into stand-in objects that keep exactly that state
under the reference's class names (`load`), exposes the standard parameters of the tree as tensors (`ModelView`), and
pickles the tree back under the same names (`dump`), so that the reference loads the result with `pickle.load`.
Nothing of the reference is imported: class names are resolved to stand-ins that are created on the fly.

Host-side file I/O only; the arithmetic stays in the engine.
"""
import io
import pickle
import sys
import types

import torch

__all__ = ['load', 'dump', 'dumps', 'loads', 'RefObject', 'ModelView', 'new_object']

_PREFIX = 'beer'
_classes = {}


class RefObject:
    """Stand-in for an instance of a reference class: holds the pickled state as its `__dict__`."""

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):      # (dict state, slots state)
            state = {**(state[0] or {}), **state[1]}
        self.__dict__.update(state)

    def __getattr__(self, name):
        # nn.Module-style access to buffers / sub-modules; bound methods pickled by the reference (parameter
        # callbacks: models/parameters.py:54-66 keeps `(self._on_weights_update, flag)` pairs) come back as
        # placeholders that pickle to the same `getattr(obj, name)` again
        if name.startswith('__'):
            raise AttributeError(name)
        d = self.__dict__
        for store in ('_buffers', '_modules', '_parameters'):
            if store in d and name in d[store]:
                return d[store][name]
        return _RefMethod(self, name)

    def ref_class(self):
        return f'{type(self).__module__}.{type(self).__qualname__}'

    def __repr__(self):
        return f'<reference {self.ref_class()}>'


class _RefMethod:
    """A bound method of a reference object (only ever stored and pickled back, never called)."""

    def __init__(self, obj, name):
        self.obj, self.name = obj, name

    def __reduce__(self):
        return getattr, (self.obj, self.name)

    def __hash__(self):
        return hash((id(self.obj), self.name))

    def __eq__(self, other):
        return isinstance(other, _RefMethod) and other.obj is self.obj and other.name == self.name

    def __call__(self, *args, **kwargs):
        raise TypeError(f'{self.obj!r}.{self.name} is a placeholder for a method of the reference')


def _stand_in(module, name):
    key = (module, name)
    if key not in _classes:
        _classes[key] = type(name.rsplit('.', 1)[-1], (RefObject,), {'__module__': module, '__qualname__': name})
    return _classes[key]


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == _PREFIX or module.startswith(_PREFIX + '.'):
            return _stand_in(module, name)
        return super().find_class(module, name)


class _FakeModules:
    """While a tree is pickled, `beer.*` module names resolve to modules holding the stand-in classes (pickle stores
    a class as module + qualified name and checks that the name resolves to the very same object)."""

    def __enter__(self):
        self.saved = {}
        names = set()
        for module, _ in _classes:
            parts = module.split('.')
            for i in range(1, len(parts) + 1):
                names.add('.'.join(parts[:i]))
        for n in sorted(names):
            self.saved[n] = sys.modules.get(n)
            m = types.ModuleType(n)
            m.__path__ = []
            sys.modules[n] = m
        for (module, name), cls in _classes.items():
            setattr(sys.modules[module], name, cls)
        for n in names:
            if '.' in n:
                parent, child = n.rsplit('.', 1)
                setattr(sys.modules[parent], child, sys.modules[n])
        return self

    def __exit__(self, *exc):
        for n, m in self.saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m


def new_object(module, name, **state):
    """A stand-in instance of the reference class `module.name` with the given attributes: what `dump` pickles so that
    the reference unpickles a real instance (classes whose pickled state is their `__dict__`: dataclasses, plain
    objects, modules of `torch.nn`)."""
    obj = _stand_in(module, name)()
    obj.__dict__.update(state)
    return obj


def loads(data):
    return _Unpickler(io.BytesIO(data)).load()


def load(path):
    """Unpickle a file written by the reference (a model, `(hmms, emissions)`, a Dataset, ...)."""
    with open(path, 'rb') as f:
        return _Unpickler(f).load()


def dumps(obj, protocol=4):
    with _FakeModules():
        return pickle.dumps(obj, protocol=protocol)


def dump(obj, path, protocol=4):
    """Pickle a tree of stand-ins under the reference's class names (the reference's `pickle.load` reads it)."""
    data = dumps(obj, protocol)
    with open(path, 'wb') as f:
        f.write(data)


# ---------------------------------------------------------------------------------------------
# view of the parameters of a pickled HMM-GMM model
# ---------------------------------------------------------------------------------------------

def _cls(obj):
    return type(obj).__qualname__ if isinstance(obj, RefObject) else None


def _std(dist):
    """Standard-parameter buffers of a distribution (dists/normalgamma.py:62-74, dirichlet.py:60-68, gamma.py)."""
    return dist.params._buffers


class ModelView:
    """The parts of a reference HMM-GMM model the VB iteration touches, as references INTO the unpickled tree (so
    writing the tensors back updates what `dump` pickles):

      kind         'PhoneLoop' | 'HMM' | 'BigramPhoneLoop'
      graph        stand-in of the CompiledGraph (buffers init/final/trans_log_probs, `pdf_id_mapping`)
      groups       one entry per emission group of the JointModelSet (or the single set), in pdf order:
                   dict(n_pdfs, n_comp, normal=<parameter means_precisions>, weights=<parameter> or None)
      categorical  stand-in of the unit-weight model (PhoneLoop / BigramPhoneLoop) or None
      start_pdf, end_pdf   the unit -> state dictionaries of the phone loop

    Model stacks of the CLI: PhoneLoop(DynamicallyOrderedModelSet(JointModelSet([MixtureSet(NormalSet), ...])))
    (mkphones.py:24-61, mkphoneloop.py:62-79); HMM over a NormalSet / MixtureSet directly (examples/HMM.ipynb)."""

    def __init__(self, model):
        self.model = model
        self.kind = _cls(model)
        if self.kind not in ('PhoneLoop', 'HMM', 'BigramPhoneLoop'):
            raise TypeError(f'not an HMM-based model of the reference: {model!r}')
        self.graph = model._modules['graph']
        ms = model._modules['modelset']
        while _cls(ms) == 'DynamicallyOrderedModelSet':      # (mkphoneloopbigram wraps the unigram loop's wrapped set again)
            ms = ms._modules['original_modelset']
        sets = list(ms._modules['modelsets']._modules.values()) if _cls(ms) == 'JointModelSet' else [ms]
        self.groups = [self._group(s) for s in sets]
        self.categorical = model._modules.get('categorical', model._modules.get('categoricalset'))
        self.start_pdf = model.__dict__.get('start_pdf')
        self.end_pdf = model.__dict__.get('end_pdf')

    @staticmethod
    def _group(s):
        if _cls(s) == 'MixtureSet':
            normal = s._modules['modelset']._modules['means_precisions']
            weights = s._modules['categoricalset']._modules['weights']
            conc = _std(weights._modules['prior'])['concentrations']
            return dict(n_pdfs=conc.shape[0], n_comp=conc.shape[1], normal=normal, weights=weights)
        if _cls(s) is not None and _cls(s).startswith('NormalSet'):
            normal = s._modules['means_precisions']
            if 'NormalGamma' != _cls(normal._modules['prior']):
                raise NotImplementedError(f'{_cls(normal._modules["prior"])} emissions: only diagonal covariances '
                                          '(NormalGamma) are on the engine path')
            return dict(n_pdfs=_std(normal._modules['prior'])['mean'].shape[0], n_comp=1, normal=normal, weights=None)
        raise NotImplementedError(f'emission set {s!r}')

    @staticmethod
    def normal_gamma(param, which):
        """(mean [M, D], scale [M], shape [M], rates [M, D]) of the prior / posterior of a means_precisions parameter."""
        b = _std(param._modules[which])
        return b['mean'], b['scale'].reshape(-1), b['shape'].reshape(-1), b['rates']

    @staticmethod
    def set_normal_gamma(param, mean, scale, shape, rates):
        b = _std(param._modules['posterior'])
        for name, new in (('mean', mean), ('scale', scale), ('shape', shape), ('rates', rates)):
            old = b[name]
            b[name] = new.detach().to(device=old.device, dtype=old.dtype).reshape(old.shape).clone()

    @staticmethod
    def concentrations(param, which):
        return _std(param._modules[which])['concentrations']

    @staticmethod
    def set_concentrations(param, conc):
        b = _std(param._modules['posterior'])
        old = b['concentrations']
        b['concentrations'] = conc.detach().to(device=old.device, dtype=old.dtype).reshape(old.shape).clone()

    def graph_arrays(self):
        b = self.graph._buffers
        return b['init_log_probs'], b['final_log_probs'], b['trans_log_probs'], list(self.graph.__dict__['pdf_id_mapping'])
