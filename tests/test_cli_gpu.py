"""`python -m beer_b200.hmm_train` (= `beer hmm accumulate` + `beer hmm update`, accumulate.py:22-63, update.py:22-72) on
files written by the LIVE reference, against the models the reference's own accumulate + update produced from them
(tests/golden/make_goldens.py gold_cli_files; the reference ran in float32 as its CLI does)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CLI = os.path.join(HERE, 'golden', 'cli')


def _compare(got_path, want_path, rtol, skip=()):
    from beer_b200 import refpickle
    from test_refpickle import _tensors
    got, want = _tensors(refpickle.load(got_path)), _tensors(refpickle.load(want_path))
    assert got.keys() == want.keys()
    worst = 0.0
    for k in want:
        if any(s in k for s in skip):
            continue
        g, w = got[k].double(), want[k].double()
        assert g.shape == w.shape and got[k].dtype == want[k].dtype, k
        if w.dtype.is_floating_point and w.numel():
            fin = torch.isfinite(w)
            assert torch.equal(fin, torch.isfinite(g)), k
            err = ((g - w)[fin].abs().max() / w[fin].abs().max().clamp(min=1e-3)).item() if fin.any() else 0.0
            worst = max(worst, err)
            assert err <= rtol, (k, err)
    return worst


def test_unsupervised_two_epochs(tmp_path, capsys):
    from beer_b200 import hmm_train
    ids = tmp_path / 'utts'
    ids.write_text(''.join(f'{u} extra columns are ignored\n' for u in ('utt_a', 'utt_b', 'utt_c', 'utt_d', 'utt_e')))
    out = str(tmp_path / 'ploop_2.mdl')
    assert hmm_train.main(['-e', '2', '-u', str(ids), os.path.join(CLI, 'ploop_0.mdl'),
                           os.path.join(CLI, 'dataset.pkl'), out]) == 0
    want = np.load(os.path.join(CLI, 'expected.npz'))
    logged = [float(line.split('=')[1]) for line in capsys.readouterr().out.splitlines() if 'ELBO=' in line]
    np.testing.assert_allclose(logged, [want['unsup_elbo_1'], want['unsup_elbo_2']], atol=2e-3)
    # every tensor of the pickle: emission posteriors + their statistics, mixture weights, unit weights, the decoding
    # graph rewritten from the new unit weights (fp32 reference: 1e-4)
    _compare(out, os.path.join(CLI, 'ploop_2.mdl'), rtol=5e-4)


def test_aligned_training_with_the_cli_default_unit_prior(tmp_path, capsys):
    from beer_b200 import hmm_train
    ids = tmp_path / 'utts'
    ids.write_text('utt_a\nutt_b\nutt_c\nutt_missing\n')
    out = str(tmp_path / 'ploop_ali.mdl')
    assert hmm_train.main(['-a', os.path.join(HERE, 'golden', 'alis.npz'), '-s', '0.8', '-l', '0.5', '-u', str(ids),
                           os.path.join(CLI, 'ploop_sbhp_0.mdl'), os.path.join(CLI, 'dataset.pkl'), out]) == 0
    want = np.load(os.path.join(CLI, 'expected.npz'))
    cap = capsys.readouterr()
    logged = [float(line.split('=')[1]) for line in cap.out.splitlines() if 'ELBO=' in line]
    np.testing.assert_allclose(logged, [want['ali_elbo_1']], atol=2e-3)
    assert 'utt_missing' in cap.err
    _compare(out, os.path.join(CLI, 'ploop_sbhp_ali_1.mdl'), rtol=5e-4)


def test_decode_matches_the_reference_cli(tmp_path):
    """`python -m beer_b200.hmm_decode` == `beer hmm decode` (decode.py:41-87) of the same pickled model: the unit
    sequences, and with --per-frame every frame's unit (alignments bit-exact), also under an acoustic scale."""
    import io
    from beer_b200 import hmm_decode
    model, data = os.path.join(CLI, 'ploop_2.mdl'), os.path.join(CLI, 'dataset.pkl')
    buf = io.StringIO()
    assert hmm_decode.main([model, data], out=buf) == 0
    with open(os.path.join(CLI, 'decode_ploop_2.txt')) as f:
        assert buf.getvalue() == f.read()
    buf = io.StringIO()
    assert hmm_decode.main(['--per-frame', '-s', '0.5', model, data], out=buf) == 0
    with open(os.path.join(CLI, 'decode_ploop_2_per_frame_scale.txt')) as f:
        assert buf.getvalue() == f.read()
    # on the alignment graphs the units are the aligned sequences themselves (make_goldens.py gold_alignment_archive)
    ids = tmp_path / 'utts'
    ids.write_text('utt_c\nutt_a\nutt_b\n')
    buf = io.StringIO()
    assert hmm_decode.main(['-a', os.path.join(HERE, 'golden', 'alis.npz'), '-u', str(ids), model, data], out=buf) == 0
    assert buf.getvalue() == 'utt_c u3 u3 u0 u2\nutt_a u2 u0\nutt_b u1\n'


def test_training_command_on_the_tensor_core_kernels(tmp_path, capsys):
    """The same file-level check at a shape the tcgen05 kernels take (40-d features, 48 states x 8 Gaussians: fp16-split
    emission, forward-backward with fused unit counts and relative log-posteriors, statistics with the responsibilities
    recomputed on chip): two epochs against accumulate + update of the live reference replayed in float64 on the same
    pickled files (make_goldens.py gold_cli_files_tc)."""
    from beer_b200 import hmm_train, refpickle
    tc = os.path.join(HERE, 'golden', 'cli_tc')
    out = str(tmp_path / 'ploop_2.mdl')
    assert hmm_train.main(['-e', '2', os.path.join(tc, 'ploop_0.mdl'), os.path.join(tc, 'dataset.pkl'), out]) == 0
    want = np.load(os.path.join(tc, 'expected.npz'))
    logged = [float(line.split('=')[1]) for line in capsys.readouterr().out.splitlines() if 'ELBO=' in line]
    np.testing.assert_allclose(logged, [want['elbo_1'], want['elbo_2']], rtol=1e-5, atol=6e-4)      # (logged with 3 decimals)
    v = refpickle.ModelView(refpickle.load(out))
    assert [(g['n_pdfs'], g['n_comp']) for g in v.groups] == [(48, 8)]
    mean, scale, shape, rates = (t.double().numpy() for t in v.normal_gamma(v.groups[0]['normal'], 'posterior'))
    for got, key in ((mean, 'post_mean'), (scale, 'post_scale'), (shape, 'post_shape'), (rates, 'post_rates')):
        w = want[key].reshape(got.shape)
        assert np.abs(got - w).max() <= 2e-4 * np.abs(w).max(), (key, np.abs(got - w).max(), np.abs(w).max())
    mix = v.concentrations(v.groups[0]['weights'], 'posterior').double().numpy()
    assert np.abs(mix - want['mix_conc']).max() <= 2e-4 * np.abs(want['mix_conc']).max()
    unit = v.concentrations(v.categorical._modules['weights'], 'posterior').double().numpy()
    assert np.abs(unit - want['unit_conc']).max() <= 2e-4 * np.abs(want['unit_conc']).max()
    trans = v.graph_arrays()[2].double().numpy()
    fin = np.isfinite(want['trans'])
    assert (np.isfinite(trans) == fin).all() and np.abs(trans[fin] - want['trans'][fin]).max() <= 2e-4


def test_aligned_training_on_graphs_that_are_not_chains(tmp_path, capsys):
    """Alignment graphs whose units have a skip arc (pickled by the live reference, gold_cli_files): no ChainBatch, the
    command falls back to one graph plan per utterance; against the reference's own accumulate --alis + update."""
    from beer_b200 import hmm_train
    ids = tmp_path / 'utts'
    ids.write_text('utt_a\nutt_b\nutt_c\n')
    out = str(tmp_path / 'ploop_skipali.mdl')
    assert hmm_train.main(['-a', os.path.join(CLI, 'alis_skip.npz'), '-u', str(ids), os.path.join(CLI, 'ploop_0.mdl'),
                           os.path.join(CLI, 'dataset.pkl'), out]) == 0
    want = np.load(os.path.join(CLI, 'expected.npz'))
    logged = [float(line.split('=')[1]) for line in capsys.readouterr().out.splitlines() if 'ELBO=' in line]
    np.testing.assert_allclose(logged, [want['skipali_elbo_1']], atol=2e-3)
    _compare(out, os.path.join(CLI, 'ploop_skipali_1.mdl'), rtol=5e-4)


def test_bigram_phone_loop(tmp_path, capsys):
    """A BigramPhoneLoop pickled by `beer hmm mkphoneloopbigram --weights-prior dirichlet2` (mkphoneloopbigram.py:35-55):
    one epoch of the batched engine (ends x starts block of the transition posteriors, phoneloop.py:175-186) against the
    reference's own accumulate + update of the same file -- logged ELBO and every tensor of the pickle."""
    from beer_b200 import hmm_train
    out = str(tmp_path / 'ploop_bigram_1.mdl')
    assert hmm_train.main([os.path.join(CLI, 'ploop_bigram_0.mdl'), os.path.join(CLI, 'dataset.pkl'), out]) == 0
    want = np.load(os.path.join(CLI, 'expected_bigram.npz'))
    logged = [float(line.split('=')[1]) for line in capsys.readouterr().out.splitlines() if 'ELBO=' in line]
    np.testing.assert_allclose(logged, [want['bigram_elbo_1']], atol=2e-3)
    _compare(out, os.path.join(CLI, 'ploop_bigram_1.mdl'), rtol=5e-4)
