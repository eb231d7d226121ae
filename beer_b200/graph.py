"""Acoustic graphs: the mutable `Graph` builder and the `CompiledGraph` the kernels run on.

Same call surface as beer/graph.py (Graph.add_state / add_arc / normalize / replace_state /
compile, CompiledGraph.posteriors / best_path / n_states; reference lines cited per method).
The builder is host-side bookkeeping; inference (forward-backward, Viterbi) runs in the
sm_100a scan kernels through a device-resident sparse plan that is rebuilt whenever the
dense log-matrices change (PhoneLoop rewrites rows in place, beer/models/phoneloop.py:53-65).
"""
import numpy as np
import torch

from . import ops

__all__ = ['Graph', 'CompiledGraph']


class Graph:
    """Weighted directed graph with emitting (pdf_id set) and non-emitting states.

    Arcs are unique per (start, end): adding an existing pair keeps the first weight, as the
    reference's set of hash-by-endpoints arcs does (beer/graph.py:19-29, 110-113)."""

    def __init__(self):
        self._next_id = 0
        self._pdf = {}            # state id -> pdf id or None, insertion ordered
        self._succ = {}           # state id -> {end: weight}
        self._pred = {}           # state id -> {start: weight}
        self.symbols = {}
        self.start_state = None
        self.end_state = None

    # -- inspection (graph.py:72-101) ---------------------------------------------------
    def states(self):
        return self._pdf.keys()

    def state_from_id(self, state_id):
        return _StateView(state_id, self._pdf[state_id])

    def arcs(self, state_id=None, incoming=False):
        if state_id is None:
            for s, outs in self._succ.items():
                for e, w in outs.items():
                    yield _ArcView(s, e, w)
        elif not incoming:
            for e, w in self._succ.get(state_id, {}).items():
                yield _ArcView(state_id, e, w)
        else:
            for s, w in self._pred.get(state_id, {}).items():
                yield _ArcView(s, state_id, w)

    # -- construction (graph.py:103-156) ------------------------------------------------
    def add_state(self, pdf_id=None):
        sid = self._next_id
        self._next_id += 1
        self._pdf[sid] = pdf_id
        self._succ[sid] = {}
        self._pred[sid] = {}
        return sid

    def add_arc(self, start, end, weight=1.0):
        if end not in self._succ.setdefault(start, {}):
            self._succ[start][end] = weight
            self._pred.setdefault(end, {})[start] = weight
        return _ArcView(start, end, self._succ[start][end])

    def _set_weight(self, start, end, weight):
        self._succ[start][end] = weight
        self._pred[end][start] = weight

    def _remove_arc(self, start, end):
        self._succ.get(start, {}).pop(end, None)
        self._pred.get(end, {}).pop(start, None)

    def normalize(self):
        """Outgoing weights of every state sum to one (graph.py:115-121)."""
        for s in self._pdf:
            outs = self._succ.get(s, {})
            total = 0.
            for w in outs.values():
                total += w
            for e in list(outs):
                self._set_weight(s, e, outs[e] / total)

    def replace_state(self, old_state_id, graph):
        """Splice `graph` in place of a state (graph.py:123-156): arcs into the old state go
        to the sub-graph's start, arcs out of it leave from the sub-graph's end."""
        remap = {s: self.add_state(pdf_id=graph._pdf[s]) for s in graph.states()}
        for arc in graph.arcs():
            self.add_arc(remap[arc.start], remap[arc.end], arc.weight)
        outgoing = list(self._succ.get(old_state_id, {}).items())
        incoming = list(self._pred.get(old_state_id, {}).items())
        for end, w in outgoing:
            self.add_arc(remap[graph.end_state], end, w)
        for start, w in incoming:
            self.add_arc(start, remap[graph.start_state], w)
        for end, _ in outgoing:
            self._remove_arc(old_state_id, end)
        for start, _ in incoming:
            self._remove_arc(start, old_state_id)
        del self._pdf[old_state_id]
        self._succ.pop(old_state_id, None)
        self._pred.pop(old_state_id, None)

    # -- epsilon closure (graph.py:158-184) ---------------------------------------------
    def _closure(self, origin, weight, forward):
        nbrs = self._succ if forward else self._pred
        stack = [(n, w, weight) for n, w in nbrs.get(origin, {}).items()]
        seen = {origin}
        while stack:
            node, arc_w, acc_w = stack.pop()
            if self._pdf[node] is not None:
                yield node, acc_w * arc_w
            elif node not in seen:
                stack.extend((n, w, arc_w * acc_w) for n, w in nbrs.get(node, {}).items())
                seen.add(node)

    def find_next_pdf_ids(self, start_state, init_weight=1.0):
        return self._closure(start_state, init_weight, True)

    def find_previous_pdf_ids(self, start_state, init_weight=1.0):
        return self._closure(start_state, init_weight, False)

    def compile(self):
        """Dense log-probabilities over the emitting states (graph.py:185-240): non-emitting
        states are removed by following their arcs; rows are renormalised keeping the
        self-loop probability.  fp32, as the reference builds them."""
        index, mapping = {}, []
        for s, pdf in self._pdf.items():
            if pdf is not None:
                index[s] = len(mapping)
                mapping.append(pdf)
        K = len(mapping)
        f = np.float32
        init, final, trans = np.zeros(K, f), np.zeros(K, f), np.zeros((K, K), f)
        for s, w in self.find_next_pdf_ids(self.start_state, 1.0):
            init[index[s]] += f(w)
        init /= init.sum()
        for s, w in self.find_previous_pdf_ids(self.end_state, 1.0):
            final[index[s]] += f(w)
        final /= final.sum()
        for arc in self.arcs():
            if self._pdf[arc.start] is None:
                continue                      # reached through the closure of its predecessors
            row = index[arc.start]
            if self._pdf[arc.end] is None:
                for s, w in self.find_next_pdf_ids(arc.end, arc.weight):
                    trans[row, index[s]] += f(w)
            else:
                trans[row, index[arc.end]] += f(arc.weight)
        for k in range(K):
            diag = trans[k, k]
            off = trans[k].sum() - diag
            if diag > 0. and off > 0:
                trans[k] /= off / (1 - diag)
                trans[k, k] = diag
        with np.errstate(divide='ignore'):
            return CompiledGraph(torch.from_numpy(np.log(init)), torch.from_numpy(np.log(final)),
                                 torch.from_numpy(np.log(trans)), mapping)


class _StateView:
    __slots__ = ('id', 'pdf_id')

    def __init__(self, sid, pdf_id):
        self.id, self.pdf_id = sid, pdf_id


class _ArcView:
    __slots__ = ('start', 'end', 'weight')

    def __init__(self, start, end, weight):
        self.start, self.end, self.weight = start, end, weight


class CompiledGraph(torch.nn.Module):
    """Inference graph of an HMM (beer/graph.py:243-344).  The three log-probability tensors
    are module buffers (so `.to()` / pickling work as in the reference); the device plan is
    derived from them lazily and refreshed when they are modified in place."""

    def __init__(self, init_log_probs, final_log_probs, trans_log_probs, pdf_id_mapping=None):
        super().__init__()
        self.register_buffer('init_log_probs', init_log_probs)
        self.register_buffer('final_log_probs', final_log_probs)
        self.register_buffer('trans_log_probs', trans_log_probs)
        if pdf_id_mapping is None:
            pdf_id_mapping = list(range(len(init_log_probs)))
        self.pdf_id_mapping = pdf_id_mapping
        self._plans = {}

    def __repr__(self):
        return '<CompiledGraph>'

    def __getstate__(self):
        state = self.__dict__.copy()
        state['_plans'] = {}           # device handles are not picklable; rebuilt on demand
        return state

    def __setstate__(self, state):
        # also reached by graphs pickled by the reference (no `_plans` there)
        self.__dict__.update(state)
        self.__dict__.setdefault('_plans', {})

    @property
    def n_states(self):
        return len(self.trans_log_probs)

    def _key(self):
        bufs = (self.init_log_probs, self.final_log_probs, self.trans_log_probs)
        return tuple((b.data_ptr(), b._version) for b in bufs)

    def plan(self, n_pdfs=None, state_level=False, factorize=True):
        """Device plan.  `state_level`: the llh columns are graph states (identity map, what
        CompiledGraph.posteriors / best_path receive); otherwise columns are pdf ids."""
        tag = ('state', factorize) if state_level else ('pdf', n_pdfs, factorize)
        cached = self._plans.get(tag)
        key = self._key()
        if cached is not None and cached[0] == key:
            return cached[1]
        K = self.n_states
        pmap = np.arange(K) if state_level else np.asarray(self.pdf_id_mapping)
        if n_pdfs is None or state_level:
            n_pdfs = int(pmap.max()) + 1 if K else 0
        plan = ops.GraphPlan(self.init_log_probs.detach().float().cpu().numpy(),
                             self.final_log_probs.detach().float().cpu().numpy(),
                             self.trans_log_probs.detach().float().cpu().numpy(), pmap, n_pdfs=n_pdfs,
                             factorize=factorize)
        self._plans[tag] = (key, plan)
        return plan

    def pdf_map_device(self, device):
        """pdf_id_mapping as an int32 tensor on `device` (cached)."""
        cached = self._plans.get(('map', str(device)))
        if cached is None:
            cached = torch.as_tensor(np.asarray(self.pdf_id_mapping), dtype=torch.int32, device=device)
            self._plans[('map', str(device))] = cached
        return cached

    @staticmethod
    def _as_device_llhs(llhs):
        if not llhs.is_cuda:
            raise ops._lib.BeerB200Error('CompiledGraph inference runs on the GPU: pass CUDA tensors '
                                         '(there is no CPU fallback)')
        return llhs.detach().to(torch.float32).contiguous()

    def posteriors(self, llhs, trans_posteriors=False):
        """State posteriors of one sequence (graph.py:289-326).  Returns
        (posteriors [T,K], log-evidence); with `trans_posteriors` the first element is the
        pair (state posteriors, transition posteriors [T-1,K,K])."""
        x = self._as_device_llhs(llhs)
        T = x.shape[0]
        off = torch.tensor([0, T], dtype=torch.int64, device=x.device)
        r = ops.hmm_forward_backward(self.plan(state_level=True), x, None, off, want_state_post=True,
                                     want_pdf_post=False, want_logz=True)
        gamma = r['state_post'].to(llhs.dtype)
        logz = r['utt_logz'][0].to(llhs.dtype)
        if trans_posteriors:
            return (gamma, self._transition_posteriors(x, gamma)), logz
        return gamma, logz

    def _transition_posteriors(self, llhs, gamma):
        """Dense (T-1, K, K) tensor of graph.py:308-323, for API parity only (O(T K^2) memory,
        not on the training path, which only needs reductions of it).  Built from the scan kernel's
        posteriors and one more forward recursion:
        xi_t[i,j] = gamma_{t+1}[j] alpha_t[i] A_ij / sum_i' alpha_t[i'] A_i'j  (csrc/transitions.cu)."""
        T = llhs.shape[0]
        off = torch.tensor([0, T], dtype=torch.int64, device=llhs.device)
        dev = llhs.device
        return ops.hmm_transition_posteriors(
            llhs, gamma.to(torch.float32).contiguous(), off,
            self.init_log_probs.detach().to(device=dev, dtype=torch.float32).contiguous(),
            self.trans_log_probs.detach().to(device=dev, dtype=torch.float32).contiguous()).to(gamma.dtype)

    def best_path(self, llhs):
        """Viterbi path, first-max tie-breaking (graph.py:329-344); CPU LongTensor like the
        reference's torch.LongTensor(path)."""
        x = self._as_device_llhs(llhs)
        T = x.shape[0]
        off = torch.tensor([0, T], dtype=torch.int64, device=x.device)
        path = ops.hmm_viterbi(self.plan(state_level=True), x, off)
        return path.to(torch.int64).cpu()
