"""Model / data-set pickles of the reference read and written without the reference (beer_b200/refpickle.py,
beer_b200/hmm_train.py): SURVEY 8(f) row 4.  The fixtures under tests/golden/cli/ were written by the LIVE reference
(tests/golden/make_goldens.py gold_cli_files: `pickle.dump(model)` as `beer hmm mkphoneloop` / `update` do)."""
import os
import pickletools

import numpy as np
import pytest
import torch

CLI = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'cli')


def _tensors(obj, prefix='', seen=None, out=None):
    """Every tensor of an unpickled tree with its path."""
    from beer_b200.refpickle import RefObject
    seen = set() if seen is None else seen
    out = {} if out is None else out
    if id(obj) in seen:
        return out
    seen.add(id(obj))
    if isinstance(obj, torch.Tensor):
        out[prefix] = obj
    elif isinstance(obj, (RefObject, torch.nn.Module)):      # (a JointModelSet keeps a real torch ModuleList)
        _tensors(obj.__dict__, prefix, seen, out)
    elif isinstance(obj, dict):
        for k, v in obj.items():
            _tensors(v, f'{prefix}/{k}', seen, out)
    elif isinstance(obj, (list, tuple)):
        for i, v in enumerate(obj):
            _tensors(v, f'{prefix}/{i}', seen, out)
    return out


@pytest.mark.parametrize('name', ['ploop_0.mdl', 'ploop_2.mdl', 'ploop_sbhp_0.mdl', 'ploop_sbhp_ali_1.mdl'])
def test_round_trip_keeps_classes_and_tensors(name):
    from beer_b200 import refpickle
    path = os.path.join(CLI, name)
    tree = refpickle.load(path)
    assert tree.ref_class() == 'beer.models.phoneloop.PhoneLoop'
    data = refpickle.dumps(tree)

    def globals_of(blob):
        names, strings = set(), []
        for op, arg, _ in pickletools.genops(blob):
            if op.name in ('SHORT_BINUNICODE', 'BINUNICODE'):
                strings.append(arg)
            elif op.name == 'STACK_GLOBAL':
                names.add((strings[-2], strings[-1]))
        return names

    with open(path, 'rb') as f:
        original = f.read()
    assert globals_of(data) == globals_of(original)        # the same classes under the same names
    again = refpickle.loads(data)
    a, b = _tensors(tree), _tensors(again)
    assert a.keys() == b.keys() and len(a) > 20
    for k in a:
        assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k
    import sys
    assert not any(m == 'beer' or m.startswith('beer.') for m in sys.modules)      # nothing left behind


def test_model_view_and_engine_parameters():
    from beer_b200 import refpickle
    from beer_b200.hmm_train import ReferenceModel, load_dataset
    v = refpickle.ModelView(refpickle.load(os.path.join(CLI, 'ploop_0.mdl')))
    assert v.kind == 'PhoneLoop' and [(g['n_pdfs'], g['n_comp']) for g in v.groups] == [(6, 4), (6, 2)]
    assert list(v.start_pdf.values()) == [0, 3, 6, 9] and list(v.end_pdf.values()) == [2, 5, 8, 11]
    init, final, trans, pmap = v.graph_arrays()
    assert trans.shape == (12, 12) and pmap == list(range(12))
    m = ReferenceModel(os.path.join(CLI, 'ploop_0.mdl'), 'cpu')
    em = m.emission
    assert (em.M, em.D, em.Kp) == (36, 4, 12) and list(em.comp_off_host) == [0, 4, 8, 12, 16, 20, 24, 26, 28, 30, 32, 34, 36]
    assert [(g.pdf_start, g.n_pdfs, g.n_comp) for g in em.weight_groups] == [(0, 6, 4), (6, 6, 2)]
    g0 = v.normal_gamma(v.groups[0]['normal'], 'posterior')
    assert torch.equal(em.post[0][:24], g0[0]) and torch.equal(em.post[3][24:], v.normal_gamma(v.groups[1]['normal'], 'posterior')[3])
    ds = load_dataset(os.path.join(CLI, 'dataset.pkl'))
    assert ds.size == sum(len(ds.fea_dict[k]) for k in ds.fea_dict.keys()) and len(ds) == 5
    np.testing.assert_allclose(ds.mean.numpy(), np.concatenate([ds.fea_dict[k] for k in ds.fea_dict.keys()]).mean(0),
                               rtol=1e-5, atol=1e-6)


def test_save_without_training_is_the_identity(tmp_path):
    from beer_b200 import refpickle
    from beer_b200.hmm_train import ReferenceModel
    src = os.path.join(CLI, 'ploop_2.mdl')
    m = ReferenceModel(src, 'cpu')
    out = str(tmp_path / 'same.mdl')
    m.save(out)
    a, b = _tensors(refpickle.load(src)), _tensors(refpickle.load(out))
    assert a.keys() == b.keys()
    for k in a:
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
        if k.endswith('/stats'):
            continue                  # (the unit weights' statistics buffer is fp64 on this side before the first update)
        assert torch.equal(a[k], b[k]), k


def test_dataset_pickle_written_for_the_reference(tmp_path):
    """`python -m beer_b200.dataset` = `beer dataset create` (create.py:44-60): same class name, same fields, same
    statistics as the pickle the live reference wrote for the same archive."""
    from beer_b200 import refpickle
    from beer_b200.dataset import create_dataset
    from beer_b200.hmm_train import load_dataset
    out = str(tmp_path / 'dataset.pkl')
    create_dataset(os.path.join(CLI, 'feats.npz'), out)
    got, want = refpickle.load(out), refpickle.load(os.path.join(CLI, 'dataset.pkl'))
    assert got.ref_class() == want.ref_class() == 'beer.cli.dataset.Dataset'
    assert set(got.__dict__) == set(want.__dict__)
    assert got.__dict__['size'] == want.__dict__['size']
    for k in ('mean', 'var'):
        assert got.__dict__[k].dtype == want.__dict__[k].dtype
        np.testing.assert_allclose(got.__dict__[k].numpy(), want.__dict__[k].numpy(), rtol=1e-5, atol=1e-6)
    assert len(load_dataset(out)) == 5


def test_bigram_phone_loop_pickle(tmp_path):
    """A BigramPhoneLoop written by `beer hmm mkphoneloopbigram --weights-prior dirichlet2` (mkphoneloopbigram.py:35-55):
    the unigram loop's already wrapped emission set is wrapped once more, the unit weights are a CategoricalSet of one
    Dirichlet per unit; opened for the engine and saved untrained, every tensor comes back as it was."""
    from beer_b200 import refpickle
    from beer_b200.engine import BigramUnitWeights
    from beer_b200.hmm_train import ReferenceModel
    src = os.path.join(CLI, 'ploop_bigram_0.mdl')
    tree = refpickle.load(src)
    assert tree.ref_class() == 'beer.models.phoneloop.BigramPhoneLoop'
    v = refpickle.ModelView(tree)
    assert v.kind == 'BigramPhoneLoop' and [(g['n_pdfs'], g['n_comp']) for g in v.groups] == [(6, 4), (6, 2)]
    assert refpickle.ModelView.concentrations(v.categorical._modules['weights'], 'posterior').shape == (4, 4)
    m = ReferenceModel(src, 'cpu')
    assert isinstance(m.unit_weights, BigramUnitWeights)
    assert m.unit_weights.start_idxs == [0, 3, 6, 9] and m.unit_weights.end_idxs == [2, 5, 8, 11]
    assert tuple(m.categorical.weights.posterior.params.concentrations.shape) == (4, 4)
    out = str(tmp_path / 'same.mdl')
    m.save(out)
    a, b = _tensors(refpickle.load(src)), _tensors(refpickle.load(out))
    # (an untrained model's posterior shares tensors with its prior; the walker lists a shared tensor once, the saved
    # posterior holds its own copies: more paths on that side, the same values)
    assert set(a) <= set(b)
    for k in a:
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
        if not k.endswith('/stats'):
            assert torch.equal(a[k], b[k]), k
    for k in set(b) - set(a):
        assert '/posterior/' in k and torch.equal(b[k], b[k.replace('/posterior/', '/prior/')]), k
