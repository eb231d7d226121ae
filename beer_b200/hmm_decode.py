"""`beer hmm decode` over the reference's own files: the most likely unit sequence of every utterance.

    python -m beer_b200.hmm_decode [-a ALIS.npz] [--per-frame] [-s SCALE] [-u UTTS|-] MODEL DATASET > transcriptions

Same arguments and output as beer/cli/subcommands/hmm/decode.py:13-87 (`<utterance id> <unit> <unit> ...` per line; with
`--per-frame` one unit per frame).  MODEL is a phone-loop model pickled by the reference (read without it:
beer_b200/refpickle.py).  All utterances are decoded as ONE ragged batch: emission kernel -> `beer_hmm_viterbi` (max-plus
scan + on-device backtrack, first-maximum ties as `torch.argmax` in graph.py:329-344) -> pdf ids; the unit names are
looked up on the host as decode.py:24-38 does.  With `--alis` every utterance is decoded on its own alignment graph.
"""
import argparse
import sys

import numpy as np
import torch

from . import ops
from .dataset import Alignments
from .hmm_train import ReferenceModel, load_dataset

f32 = torch.float32


def state2phone(path, start_pdf, per_frame):
    """decode.py:24-38: a new unit starts whenever the path ENTERS a unit-start pdf from another pdf."""
    starts = set(start_pdf.values())
    state2sym = {value: key for key, value in start_pdf.items()}
    previous = path[0]
    last = state2sym[previous]
    phones = [last]
    for state in path[1:]:
        if state != previous and state in starts:
            last = state2sym[state]
            phones.append(last)
        elif per_frame:
            phones.append(last)
        previous = state
    return phones


def decode_batch(model, feats, scale=1.0, graphs=None):
    """Best paths (pdf ids, numpy int64) of a list of [T, D] feature arrays: one emission launch + one Viterbi launch
    over the decoding graph; `graphs` (one CompiledGraph per utterance) decodes every utterance on its own graph."""
    dev = model.device
    em = model.emission
    lens = [len(x) for x in feats]
    off_host = np.concatenate([[0], np.cumsum(lens)])
    X = torch.from_numpy(np.concatenate(feats).astype(np.float32)).to(dev)
    W, bias, ref = em.refresh()
    pdf_llh = torch.empty(len(X), em.Kp, device=dev, dtype=f32)
    comp = torch.empty(len(X), em.M, device=dev, dtype=f32) if em.has_mixtures else None
    em.llh(X, W, bias, ref, pdf_llh, comp, torch.empty(len(X), device=dev, dtype=f32))
    out = []
    if graphs is None:
        plan = model.graph.plan(n_pdfs=em.Kp)
        off = torch.as_tensor(off_host, dtype=torch.int64, device=dev)
        path = ops.hmm_viterbi(plan, pdf_llh, off, scale=scale).cpu().numpy()
        mapping = np.asarray(model.graph.pdf_id_mapping, dtype=np.int64)
        for u in range(len(lens)):
            out.append(mapping[path[off_host[u]:off_host[u + 1]]])
        return out
    for u, g in enumerate(graphs):
        a, b = int(off_host[u]), int(off_host[u + 1])
        off = torch.tensor([0, b - a], dtype=torch.int64, device=dev)
        path = ops.hmm_viterbi(g.plan(n_pdfs=em.Kp), pdf_llh[a:b], off, scale=scale).cpu().numpy()
        out.append(np.asarray(g.pdf_id_mapping, dtype=np.int64)[path])
    return out


def main(argv=None, out=None):
    ap = argparse.ArgumentParser(prog='python -m beer_b200.hmm_decode', description=__doc__.split('\n\n')[0])
    ap.add_argument('-a', '--alis', help='alignment graphs in a "npz" archive')
    ap.add_argument('--per-frame', action='store_true', help='output the per-frame transcription')
    ap.add_argument('-s', '--acoustic-scale', default=1., type=float, help='scaling factor of the acoustic model')
    ap.add_argument('-u', '--utts', help='decode the given utterances ("-") for stdin')
    ap.add_argument('model', help='hmm based model pickled by the reference')
    ap.add_argument('dataset', help='data set pickled by `beer dataset create`')
    args = ap.parse_args(argv)
    out = sys.stdout if out is None else out
    if not torch.cuda.is_available():
        raise RuntimeError('beer_b200.hmm_decode needs a CUDA device (there is no CPU path)')
    dev = torch.device('cuda', torch.cuda.current_device())
    model = ReferenceModel(args.model, dev)
    if model.view.start_pdf is None:
        raise TypeError('decoding to unit names needs a phone-loop model (start_pdf)')
    dataset = load_dataset(args.dataset)
    alis = Alignments(args.alis) if args.alis else None
    if args.utts:
        lines = sys.stdin.readlines() if args.utts == '-' else open(args.utts).readlines()
        utts = [line.strip().split()[0] for line in lines if line.strip()]
    else:
        utts = sorted(dataset.fea_dict.keys())
    known = set(dataset.fea_dict.keys())
    for u in utts:
        if u not in known:
            print(f'warning: no utterance {u} in {args.dataset}', file=sys.stderr)
    utts = [u for u in utts if u in known]
    feats = [np.asarray(dataset.fea_dict[u], dtype=np.float32) for u in utts]
    count = 0
    if alis is None:
        paths = decode_batch(model, feats, scale=args.acoustic_scale) if utts else []
    else:
        # utterances without an alignment graph fall back to the decoding graph (decode.py:70-76)
        paths = [None] * len(utts)
        with_g = [i for i, u in enumerate(utts) if u in alis]
        without = [i for i, u in enumerate(utts) if u not in alis]
        for i in without:
            print(f'warning: no alignment graph for utterance "{utts[i]}"', file=sys.stderr)
        if with_g:
            for i, p in zip(with_g, decode_batch(model, [feats[i] for i in with_g], scale=args.acoustic_scale,
                                                 graphs=[alis[utts[i]] for i in with_g])):
                paths[i] = p
        if without:
            for i, p in zip(without, decode_batch(model, [feats[i] for i in without], scale=args.acoustic_scale)):
                paths[i] = p
    for u, path in zip(utts, paths):
        print(u, ' '.join(state2phone([int(s) for s in path], model.view.start_pdf, args.per_frame)), file=out)
        count += 1
    print(f'successfully decoded {count} utterances.', file=sys.stderr)
    return 0


if __name__ == '__main__':
    sys.exit(main())
