"""CPU tests of the host-side logic: the C-ABI library loads and exports every symbol the header
declares (no compute calls), utterance sharding, the world_size-2 reduction of the flat statistics
buffer over `gloo` against the single-process oracle iteration, the ELBO value object, the graph
builder against the reference's compiled graphs."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from conftest import load_golden  # noqa: E402
from oracle import beer_oracle as O  # noqa: E402


def test_library_exports_every_declared_symbol():
    from beer_b200 import _lib, build
    build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'beer_b200.h')).read()
    declared = set(re.findall(r'BEER_API\s+[\w\s\*]+?\b(beer_\w+)\s*\(', header))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/beer_b200.h but not exported'
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.beer_b200_version() >= 100
    assert lib.beer_emission_tc_supported(100, 40, 1) == 1 and lib.beer_emission_tc_supported(100, 13, 1) == 0
    assert lib.beer_accumulate_tc_supported(100, 40) == 1


def test_package_imports_without_gpu_and_fails_loudly():
    import beer_b200 as beer
    assert beer.evidence_lower_bound and beer.HMM and beer.vbi.BayesianModelOptimizer is beer.VBOptimizer
    if not torch.cuda.is_available():
        with pytest.raises(beer._lib.BeerB200Error):
            beer.NormalSet.create(torch.zeros(3), torch.ones(3), size=4, cov_type='diagonal')


def test_shard_utterances_balances_frames():
    from beer_b200.engine import shard_utterances
    rng = np.random.default_rng(0)
    lens = rng.integers(50, 3000, size=1000)
    for world in (1, 2, 4, 8):
        shards = shard_utterances(lens, world)
        allidx = np.sort(np.concatenate(shards))
        np.testing.assert_array_equal(allidx, np.arange(len(lens)))      # a partition
        loads = np.array([lens[s].sum() for s in shards])
        assert loads.max() - loads.min() <= lens.max()                    # LPT bound
    assert [len(s) for s in shard_utterances([5, 5], 4)] == [1, 1, 0, 0]


def test_graph_builder_matches_reference_compile():
    """beer_b200.graph.Graph.compile against the reference's compiled graphs (graph.py:185-240)."""
    from beer_b200.graph import Graph
    from beer_b200.synthetic import unit_graph
    g = load_golden('graph_compile')
    pl = Graph()
    pl.start_state, pl.end_state = pl.add_state(), pl.add_state()
    pivot = pl.add_state()
    us = [pl.add_state() for _ in range(5)]
    pl.add_arc(pl.start_state, pivot)
    pl.add_arc(pivot, pl.end_state)
    for s in us:
        pl.add_arc(pivot, s)
        pl.add_arc(s, pivot)
    pl.normalize()
    for i, s in enumerate(us):
        pl.replace_state(s, unit_graph(3, i * 3))
    pl.normalize()
    cg = pl.compile()
    with np.errstate(invalid='ignore'):
        for name, t in (('init', cg.init_log_probs), ('final', cg.final_log_probs), ('trans', cg.trans_log_probs)):
            want = g['pl_' + name]
            got = t.numpy()
            assert np.array_equal(np.isinf(got), np.isinf(want))
            np.testing.assert_allclose(got[~np.isinf(got)], want[~np.isinf(want)], rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(np.asarray(cg.pdf_id_mapping), g['pl_map'])


def test_elbo_instance_semantics():
    """EvidenceLowerBoundInstance.__add__ / backward / sync (objectives.py:78-116) on plain tensors."""
    from beer_b200.inference import EvidenceLowerBoundInstance, evidence_lower_bound

    class P:
        def __init__(self):
            self.stored = None

        def store_stats(self, s):
            self.stored = s

    p1, p2 = P(), P()
    a = EvidenceLowerBoundInstance(torch.tensor(-3.), {p1: torch.ones(2)}, [p1], 10, 100)
    b = EvidenceLowerBoundInstance(torch.tensor(-4.), {p1: torch.ones(2), p2: 2 * torch.ones(3)}, [p1, p2], 30, 100)
    c = evidence_lower_bound(datasize=100) + a + b
    assert float(c) == -7. and c._minibatchsize == 40
    c.backward()
    np.testing.assert_allclose(p1.stored.numpy(), 100 / 40 * 2 * np.ones(2))
    np.testing.assert_allclose(p2.stored.numpy(), 100 / 40 * 2 * np.ones(3))
    with pytest.raises(ValueError):
        a + EvidenceLowerBoundInstance(0., {}, [], 1, 99)
    with pytest.raises(ValueError):
        evidence_lower_bound(model=object())


# ---------------------------------------------------------------------------------------------
# world_size 2 over gloo: shards + one all-reduce of the flat buffer == single process
# ---------------------------------------------------------------------------------------------

def _make_case():
    rng = np.random.default_rng(5)
    P, S, D = 4, 3, 6
    graph, _, _ = O.phone_loop_graph(P, S)
    K = P * S
    means = 2.0 * rng.standard_normal((K, D))
    lens = [40, 25, 61, 33, 18, 50]
    utts = [O.sample_utterances(rng, graph, means, 1, n)[0] for n in lens]
    prior = (np.zeros((K, D)), np.ones((K, 1)), np.ones((K, 1)), np.ones((K, D)))
    post = (rng.standard_normal((K, D)), np.ones((K, 1)), np.ones((K, 1)), np.ones((K, D)))
    return graph, utts, prior, post


def _rank_main(rank, world, port, out):
    import torch.distributed as dist
    from beer_b200.engine import elbo_from_flat, shard_utterances, stats_scale_from_flat
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    graph, utts, prior, post = _make_case()
    mine = shard_utterances([len(u) for u in utts], world)[rank]
    K, D = post[0].shape
    Q = 2 * D + 2
    # the engine's flat buffer layout: [acc (M*Q) | sum_u ell_u/T_u | sum_u T_u | n_utts | sum_u ell_u]
    flat = torch.zeros(K * Q + 4, dtype=torch.float64)
    for i in mine:
        r = O.hmm_estep(utts[i], post, None, graph)      # the E-step itself is the GPU kernels' job
        flat[:K * Q] += torch.from_numpy(r['acc_normal'].reshape(-1))
        ell = float(r['exp_llh'].sum())
        flat[K * Q + 0] += ell / len(utts[i])
        flat[K * Q + 1] += len(utts[i])
        flat[K * Q + 2] += 1
        flat[K * Q + 3] += ell
    dist.all_reduce(flat)                                 # the one exchange step of a VB iteration
    datasize = float(sum(len(u) for u in utts))
    kl = float(O.normalgamma_kl(post, prior).sum())
    elbo = float(elbo_from_flat(flat[K * Q:], kl, datasize))
    scale = stats_scale_from_flat(float(flat[K * Q + 1]), datasize)
    if rank == 0:
        np.savez(out, flat=flat.numpy(), elbo=elbo, scale=scale)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduction_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / 'rank0.npz')
    port = 29500 + os.getpid() % 2000
    mp.spawn(_rank_main, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    graph, utts, prior, post = _make_case()
    want_elbo, _, _, info = O.vb_iteration_hmm(utts, prior, post, None, None, graph)
    K, D = post[0].shape
    np.testing.assert_allclose(got['flat'][:K * (2 * D + 2)].reshape(K, -1), info['acc_normal'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(float(got['elbo']), want_elbo, rtol=1e-12)
    np.testing.assert_allclose(float(got['scale']), 1.0)
    assert got['flat'][-3] == sum(len(u) for u in utts) and got['flat'][-2] == len(utts)


def test_filterbank_matrix_matches_reference():
    from beer_b200 import features as F
    g = load_golden('fbank')
    for nf in (40, 26):
        np.testing.assert_allclose(F.create_fbank(nf, 512, lowfreq=20, highfreq=8000), g[f'filters{nf}'], atol=1e-15)


def test_dataset_archive_roundtrip(tmp_path):
    """The reference's features archive (npz of per-utterance arrays) -> statistics and balanced shards."""
    from beer_b200.dataset import Dataset
    rng = np.random.default_rng(3)
    utts = {f'utt{i:02d}': rng.standard_normal((int(n), 5)).astype(np.float32) + 2
            for i, n in enumerate(rng.integers(5, 60, size=11))}
    path = str(tmp_path / 'feats.npz')
    np.savez(path, **utts)
    ds = Dataset(path)
    allx = np.concatenate(list(utts.values()))
    assert ds.size == len(allx) and len(ds) == 11
    np.testing.assert_allclose(ds.mean.numpy(), allx.mean(0), rtol=1e-5)
    np.testing.assert_allclose(ds.var.numpy(), allx.var(0), rtol=1e-4)
    seen = []
    for rank in range(3):
        ids, batch = ds.shard(rank, 3, device='cpu')
        seen += ids
        assert batch.n_utts == len(ids) and len(batch) == sum(len(utts[i]) for i in ids)
        off = batch.offsets_host
        for j, i in enumerate(ids):
            np.testing.assert_array_equal(batch.X[off[j]:off[j + 1]].numpy(), utts[i])
    assert sorted(seen) == sorted(utts)
    import pickle
    ds2 = pickle.loads(pickle.dumps(ds))
    assert ds2.size == ds.size and len(ds2['utt03'].features) == len(utts['utt03'])


def test_chain_batch_host_logic():
    """Alignment graphs -> flat chain arrays (mkaligraph.py:18-39 after Graph.compile): lengths, pdf ids and
    weights of every chain, rejection of graphs that are not chains, and the chains of sampled state paths."""
    from beer_b200 import ops, synthetic
    gr, _, _ = O.phone_loop_graph(3, 2)
    with pytest.raises(ValueError):
        ops.ChainBatch([gr], 'cpu')
    # two hand-made chains as (init, final, trans, map)
    def chain(L, pdfs, loop):
        trans = np.full((L, L), -np.inf, dtype=np.float32)
        trans[np.arange(L), np.arange(L)] = np.log(loop)
        trans[np.arange(L - 1), np.arange(1, L)] = np.log(1 - loop)
        init = np.full(L, -np.inf, dtype=np.float32)
        init[0] = 0.
        final = np.full(L, -np.inf, dtype=np.float32)
        final[-1] = np.log(1 - loop)
        return init, final, trans, pdfs
    cb = ops.ChainBatch([chain(3, [4, 5, 4], 0.75), chain(1, [2], 0.5)], 'cpu')
    assert cb.n_utts == 2 and cb.max_len == 3 and cb.n_pdfs == 6 and cb.row_stride == 128
    np.testing.assert_array_equal(cb.chain_off.numpy(), [0, 3, 4])
    np.testing.assert_array_equal(cb.pdf.numpy(), [4, 5, 4, 2])
    np.testing.assert_allclose(cb.log_self.numpy(), np.log([.75, .75, .75, .5]), rtol=1e-6)
    np.testing.assert_allclose(cb.log_next.numpy(), np.log([.25, .25, .25, .5]), rtol=1e-6)
    assert cb.workspace_bytes(10) == 10 * (128 + 32) * 4         # alpha rows + the 32 lane offsets of every row
    # chains of sampled paths: unit instances in visiting order, re-entry of the same unit is a new instance
    paths = np.array([[0, 0, 1, 1, 0, 1, 2, 3, 3, 2]])           # 2-state units: u0, u0 again, u1, u1 again
    off, pdf, ls, ln, li = synthetic.alignment_chains(paths, 2)
    np.testing.assert_array_equal(off, [0, 8])
    np.testing.assert_array_equal(pdf, [0, 1, 0, 1, 2, 3, 2, 3])
    cb2 = ops.ChainBatch.from_arrays(off, pdf, ls, ln, li, 'cpu')
    assert cb2.max_len == 8 and cb2.n_pdfs == 4


def test_onehot_transitions_of_a_path():
    """hmm.py:49-54 restricted to rows x columns, with no transition across an utterance boundary."""
    from beer_b200.models import _onehot_transitions
    path = torch.tensor([0, 1, 1, 2, 0, 2, 2], dtype=torch.int32)
    off = torch.tensor([0, 4, 7])
    rows, cols = [1, 2], [0, 2]
    got = _onehot_transitions(path, off, rows, cols).numpy()
    want = []
    for a, b in ((0, 4), (4, 7)):
        for t in range(a, b - 1):
            want.append([[float(path[t] == r and path[t + 1] == c) for c in cols] for r in rows])
    np.testing.assert_array_equal(got, np.asarray(want, dtype=np.float32))


def test_alignment_archive_written_by_the_reference():
    """tests/golden/alis.npz was pickled by the live reference (mkaligraph.py:40-63 layout): it loads without the
    reference installed, gives the same graphs as the plain-array dump, and flattens into per-utterance chains."""
    from beer_b200 import Alignments, CompiledGraph
    here = os.path.join(ROOT, 'tests', 'golden')
    assert 'beer' not in sys.modules or sys.modules['beer'].__name__ != 'beer'
    alis = Alignments(os.path.join(here, 'alis.npz'))
    want = np.load(os.path.join(here, 'alis_expected.npz'))
    assert sorted(alis.keys()) == ['utt_a', 'utt_b', 'utt_c'] and 'utt_b' in alis and len(alis) == 3
    for u in alis.keys():
        g = alis[u]
        assert isinstance(g, CompiledGraph)
        np.testing.assert_array_equal(g.trans_log_probs.numpy(), want[u + '_trans'])
        np.testing.assert_array_equal(g.init_log_probs.numpy(), want[u + '_init'])
        np.testing.assert_array_equal(g.final_log_probs.numpy(), want[u + '_final'])
        assert list(g.pdf_id_mapping) == list(want[u + '_map'])
    cb = alis.chain_batch(['utt_c', 'utt_b'], 'cpu')
    assert list(cb.lengths) == [12, 3] and cb.max_len == 12
    np.testing.assert_array_equal(cb.pdf.numpy()[:12], want['utt_c_map'])
    np.testing.assert_allclose(cb.log_self.numpy()[12:], np.diagonal(want['utt_b_trans']), rtol=1e-6)
    assert 'beer' not in sys.modules or sys.modules['beer'].__name__ != 'beer'


def test_gamma_distribution_math():
    """beer/dists/gamma.py restated: natural parameters, expected statistics, log-normaliser, KL
    (basedist.py:243-263) and the natural-gradient step with lrate 1 (posterior = prior + statistics), against scipy."""
    from scipy.special import digamma, gammaln
    from beer_b200.dists import Gamma, kl_div
    a1, b1, a0, b0 = 5.0, 3.5, 1.0, 0.5
    q = Gamma.from_std_parameters(torch.tensor([a1]), torch.tensor([b1]))
    p = Gamma.from_std_parameters(torch.tensor([a0]), torch.tensor([b0]))
    np.testing.assert_allclose(q.natural_parameters().numpy(), [-b1, a1 - 1])
    np.testing.assert_allclose(q.expected_sufficient_statistics().numpy(), [a1 / b1, digamma(a1) - np.log(b1)], rtol=1e-12)
    np.testing.assert_allclose(float(q.log_norm()), gammaln(a1) - a1 * np.log(b1), rtol=1e-12)
    np.testing.assert_allclose(float(q.expected_value()), a1 / b1, rtol=1e-6)
    # closed form KL(Gamma(a1, b1) || Gamma(a0, b0))
    want = (a1 - a0) * digamma(a1) - gammaln(a1) + gammaln(a0) + a0 * (np.log(b1) - np.log(b0)) + a1 * (b0 - b1) / b1
    np.testing.assert_allclose(float(kl_div(q, p)), want, rtol=1e-10)
    q._natural_grad_update(p, torch.tensor([-2.0, 4.0], dtype=torch.float64), 1.0)
    np.testing.assert_allclose([float(q.params.shape), float(q.params.rate)], [a0 + 4.0, b0 + 2.0], rtol=1e-6)


def test_utils_onehot_logsumexp():
    """beer_b200.utils mirrors beer/utils.py:84-123: `onehot`, and a `logsumexp` that hands back +-inf when the
    maximum is +-inf instead of NaN."""
    import torch
    from beer_b200 import utils
    labels = [2, 0, 3, 3]
    got = utils.onehot(labels, 4, torch.float64, 'cpu').numpy()
    np.testing.assert_array_equal(got, O.onehot(np.asarray(labels), 4))
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 7)) * 30
    x[1, :] = -np.inf
    x[3, 2] = -np.inf
    for dim in (0, 1):
        got = utils.logsumexp(torch.from_numpy(x), dim=dim).numpy()
        with np.errstate(all='ignore'):
            want = O.logsumexp(x, axis=dim)
        np.testing.assert_allclose(got, want, rtol=1e-12)
    assert utils.logsumexp(torch.tensor([float('inf'), 1.0]), dim=0).item() == float('inf')
    assert utils.logsumexp(torch.full((3,), float('-inf')), dim=0).item() == float('-inf')
