"""Functional wrappers: torch CUDA tensors -> C-ABI calls of libbeer_b200.so.

PyTorch is used for device memory and streams only; every computation of the VB
E-step / M-step below runs in the hand-written sm_100a kernels.  Each wrapper
launches on the current torch CUDA stream and never synchronises.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

f32, f64, i32, i64 = torch.float32, torch.float64, torch.int32, torch.int64


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t, dtype=None, allow_none=False):
    """Device pointer of a contiguous CUDA tensor."""
    if t is None:
        if allow_none:
            return None
        raise ValueError('missing tensor argument')
    if not t.is_cuda:
        raise _lib.BeerB200Error('beer_b200 kernels need CUDA tensors (there is no CPU fallback)')
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f'expected {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise ValueError('tensor must be contiguous')
    return C.c_void_p(t.data_ptr())


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.BeerB200Error('beer_b200 needs a CUDA device (B200, sm_100a); '
                                 'there is no CPU fallback')
    return _lib.load()


# ---------------------------------------------------------------------------
# parameter math
# ---------------------------------------------------------------------------

def normalgamma_expected_stats(mean, scale, shape, rates):
    lib = require_cuda()
    M, D = mean.shape
    ets = torch.empty(M, 2 * D + 2, device=mean.device, dtype=f32)
    _lib.check(lib.beer_normalgamma_expected_stats(_p(mean, f32), _p(scale, f32), _p(shape, f32),
                                                   _p(rates, f32), M, D, _p(ets), _stream()),
               'beer_normalgamma_expected_stats')
    return ets


def dirichlet_expected_logw(conc):
    lib = require_cuda()
    c2 = conc.reshape(-1, conc.shape[-1])
    K, Cc = c2.shape
    out = torch.empty_like(c2)
    _lib.check(lib.beer_dirichlet_expected_logw(_p(c2, f32), K, Cc, _p(out), _stream()),
               'beer_dirichlet_expected_logw')
    return out.reshape(conc.shape)


def emission_prepare(mean, scale, shape, rates, logw=None):
    """-> (W [M,2D], bias [M], ref [D+1])."""
    lib = require_cuda()
    M, D = mean.shape
    W = torch.empty(M, 2 * D, device=mean.device, dtype=f32)
    bias = torch.empty(M, device=mean.device, dtype=f32)
    ref = torch.empty(D + 1, device=mean.device, dtype=f32)
    _lib.check(lib.beer_emission_prepare(_p(mean, f32), _p(scale, f32), _p(shape, f32), _p(rates, f32),
                                         _p(logw, f32, True), M, D, _p(W), _p(bias), _p(ref), _stream()),
               'beer_emission_prepare')
    return W, bias, ref


def emission_llh(X, W, bias, ref, comp_off=None, Kp=None, want_comp=False, out=None, out_comp=None,
                 out_ref=None):
    """-> (pdf_llh [N,Kp], comp_llh [N,M] or None, frame_ref [N]); offset form."""
    lib = require_cuda()
    N, D = X.shape
    M = W.shape[0]
    if Kp is None:
        Kp = M if comp_off is None else comp_off.numel() - 1
    pdf_llh = out if out is not None else torch.empty(N, Kp, device=X.device, dtype=f32)
    comp = out_comp
    if comp is None and want_comp:
        comp = torch.empty(N, M, device=X.device, dtype=f32)
    fref = out_ref if out_ref is not None else torch.empty(N, device=X.device, dtype=f32)
    _lib.check(lib.beer_emission_llh(_p(X, f32), N, D, _p(W, f32), _p(bias, f32), _p(ref, f32), M,
                                     _p(comp_off, i32, True), Kp, _p(pdf_llh, f32), pdf_llh.stride(0),
                                     _p(comp, f32, True), _p(fref), _stream()),
               'beer_emission_llh')
    return pdf_llh, comp, fref


def emission_tc_supported(M, D, C):
    return bool(_lib.load().beer_emission_tc_supported(int(M), int(D), int(C)))


def emission_tc_pack(W, bias, C, out=None):
    """Pack (W, bias) into the tensor-core weight image (hi/lo core-matrix layout)."""
    lib = require_cuda()
    M, D = W.shape[0], W.shape[1] // 2
    n = int(lib.beer_emission_tc_image_floats(M, D, int(C)))
    if n < 0:
        raise _lib.BeerB200Error('no tensor-core emission path for this shape')
    img = out if out is not None else torch.empty(n, device=W.device, dtype=f32)
    _lib.check(lib.beer_emission_tc_pack(_p(W, f32), _p(bias, f32), M, D, int(C), _p(img, f32), _stream()),
               'beer_emission_tc_pack')
    return img


def emission_llh_tc(X, image, ref, M, C, want_comp=False, out=None, out_comp=None, out_ref=None):
    """tcgen05 version of emission_llh (uniform C components per pdf)."""
    lib = require_cuda()
    N, D = X.shape
    Kp = M // C
    pdf_llh = out if out is not None else torch.empty(N, Kp, device=X.device, dtype=f32)
    comp = out_comp
    if comp is None and want_comp:
        comp = torch.empty(N, M, device=X.device, dtype=f32)
    fref = out_ref if out_ref is not None else torch.empty(N, device=X.device, dtype=f32)
    _lib.check(lib.beer_emission_llh_tc(_p(X, f32), N, D, _p(image, f32), _p(ref, f32), M, int(C),
                                        _p(pdf_llh, f32), pdf_llh.stride(0), _p(comp, f32, True), _p(fref),
                                        _stream()), 'beer_emission_llh_tc')
    return pdf_llh, comp, fref


def path_accumulate_mix(X, pdf_ids, W, bias, C, acc, frame_ref=None, scale=1.0, out_frame=None, want_frame=True):
    """Sparse statistics of a mixture model along a state path (beer_path_accumulate_mix): acc [M, 2D+2] fp64 +=,
    returns the per-frame expected llh [N] (scale * (log-sum-exp over the components of the frame's pdf + frame_ref))."""
    lib = require_cuda()
    N, D = X.shape
    M = W.shape[0]
    frame = out_frame if out_frame is not None else (torch.empty(N, device=X.device, dtype=f32) if want_frame else None)
    _lib.check(lib.beer_path_accumulate_mix(_p(X, f32), N, D, _p(pdf_ids, i32), _p(W, f32), _p(bias, f32), M, int(C),
                                            _p(frame_ref, f32, True), float(scale), _p(acc, f64), _p(frame, f32, True),
                                            _stream()), 'beer_path_accumulate_mix')
    return frame


def emission_bwd_supported(M, D):
    return bool(require_cuda().beer_emission_bwd_supported(int(M), int(D)))


def emission_llh_bwd(X, exp_stats, pdf_post, grad_out=None, comp_llh=None, pdf_llh=None, pdf_of=None, scale=1.0):
    """KA backward (beer_emission_llh_bwd): d/dX of sum_t grad_out[t] sum_k pdf_post[t,k] llh_k(x_t) with the posteriors
    held fixed (hmm.py:79-87, vae.py:63-89).  `exp_stats` [M, >= 2D] = E[T(theta)] (normalgamma_expected_stats);
    mixtures pass the per-Gaussian and per-pdf llhs of the forward call and `pdf_of` [M] (int32)."""
    lib = require_cuda()
    N, D = X.shape
    M = exp_stats.shape[0]
    dev = X.device
    image = torch.empty(lib.beer_emission_bwd_image_bytes(M, D), dtype=torch.uint8, device=dev)
    inv_scale = torch.empty(2 * D, dtype=f32, device=dev)
    scratch = torch.empty(2 * D, dtype=i32, device=dev)
    _lib.check(lib.beer_emission_bwd_pack(_p(exp_stats, f32), M, D, exp_stats.stride(0), _p(image), _p(inv_scale),
                                          _p(scratch), _stream()), 'beer_emission_bwd_pack')
    grad = torch.empty(N, D, dtype=f32, device=dev)
    mixt = comp_llh is not None
    _lib.check(lib.beer_emission_llh_bwd(
        _p(X, f32), N, D, _p(image), _p(inv_scale), M, _p(pdf_post, f32), pdf_post.stride(0),
        _p(comp_llh, f32, True), comp_llh.stride(0) if mixt else 0, _p(pdf_llh, f32, True),
        pdf_llh.stride(0) if mixt else 0, _p(pdf_of, i32, True), _p(grad_out, f32, True), float(scale), _p(grad),
        _stream()), 'beer_emission_llh_bwd')
    return grad


# ---------------------------------------------------------------------------
# graph plan
# ---------------------------------------------------------------------------

class GraphPlan:
    """Device-resident sparse form of a compiled graph (beer_graph_plan)."""

    def __init__(self, init_log, final_log, trans_log, pdf_map, n_pdfs=None, factorize=True):
        lib = require_cuda()
        init = np.ascontiguousarray(np.asarray(init_log, dtype=np.float32))
        final = np.ascontiguousarray(np.asarray(final_log, dtype=np.float32))
        trans = np.ascontiguousarray(np.asarray(trans_log, dtype=np.float32))
        pmap = np.ascontiguousarray(np.asarray(pdf_map, dtype=np.int32))
        K = init.shape[0]
        if trans.shape != (K, K) or final.shape != (K,) or pmap.shape != (K,):
            raise ValueError('inconsistent graph shapes')
        self.n_states = K
        self.n_pdfs = int(n_pdfs) if n_pdfs is not None else int(pmap.max()) + 1
        self.pdf_map = pmap
        handle = C.c_void_p()
        _lib.check(lib.beer_graph_plan_create(init.ctypes.data, final.ctypes.data, trans.ctypes.data,
                                              pmap.ctypes.data, K, self.n_pdfs, int(bool(factorize)),
                                              C.byref(handle)), 'beer_graph_plan_create')
        self._h = handle
        self._lib = lib
        info = np.zeros(8, dtype=np.int32)
        _lib.check(lib.beer_graph_plan_info(self._h, info.ctypes.data), 'beer_graph_plan_info')
        self.info = dict(K=int(info[0]), junctions=int(info[1]), direct_arcs=int(info[2]),
                         junction_in=int(info[3]), junction_out=int(info[4]), states_per_lane=int(info[5]),
                         map_identity=bool(info[6]), dense_nnz=int(info[7]))

    @property
    def n_units(self):
        """Units of an aligned left-to-right loop (0 if the graph is not one)."""
        return int(self._lib.beer_hmm_unit_count_size(self._h))

    def workspace_bytes(self, n_frames):
        return int(self._lib.beer_hmm_workspace_bytes(self._h, int(n_frames)))

    @property
    def writes_log2_posteriors(self):
        """The forward-backward kernel of this graph can hand out log2 posteriors (`out_pdf_lpost`)."""
        return bool(self._lib.beer_hmm_lpost_supported(self._h))

    def marks_active_blocks(self, with_unit_counts=False):
        """The forward-backward kernel of this graph can fill the activity map (`block_active`) the statistics kernel of
        mixtures skips by."""
        return bool(self._lib.beer_hmm_block_activity_supported(self._h, int(bool(with_unit_counts))))

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            self._lib.beer_graph_plan_destroy(h)


def hmm_forward_backward(plan, pdf_llh, frame_ref, utt_off, scale=1.0, want_state_post=False,
                         want_pdf_post=True, want_frame_llh=False, want_logz=False, workspace=None,
                         out_pdf_post=None, out_utt_exp_llh=None, unit_counts=None, llh_log2=False, out_pdf_lpost=None,
                         lpost_relative=False, block_active=None, pdfs_per_block=0):
    """Forward-backward for a ragged batch.  -> dict(state_post, pdf_post, frame_exp_llh,
    utt_exp_llh (fp64), utt_logz (fp64)).  `out_pdf_post` must be zero-filled by the caller
    when the graph's pdf map is not the identity (the kernel then scatter-adds).  `lpost_relative`: `out_pdf_lpost` =
    log2(scale posterior) - log2 llh, the form `Mix16.accumulate(..., relative=True)` reads (BEER_FB_LPOST_RELATIVE).
    `block_active` (uint8 [ceil(N / 64), >= ceil(Kp / pdfs_per_block)], zeroed by the caller): the activity map of
    beer_hmm_forward_backward_blocks, for `Mix16.accumulate(..., block_active=...)`."""
    lib = require_cuda()
    N = pdf_llh.shape[0]
    n_utts = utt_off.numel() - 1
    dev = pdf_llh.device
    K, Kp = plan.n_states, plan.n_pdfs
    if pdf_llh.shape[1] < Kp:
        raise ValueError('pdf_llh has fewer columns than the graph has pdfs')
    nbytes = plan.workspace_bytes(N)
    if workspace is None or workspace.numel() * workspace.element_size() < nbytes:
        workspace = torch.empty((nbytes + 3) // 4, device=dev, dtype=f32)
    state_post = torch.empty(N, K, device=dev, dtype=f32) if want_state_post else None
    pdf_post = out_pdf_post
    if pdf_post is None and want_pdf_post:
        pdf_post = (torch.empty if plan.info['map_identity'] and Kp == K else torch.zeros)(
            N, Kp, device=dev, dtype=f32)
    frame = torch.empty(N, device=dev, dtype=f32) if want_frame_llh else None
    utt_ell = out_utt_exp_llh if out_utt_exp_llh is not None else torch.empty(n_utts, device=dev, dtype=f64)
    utt_logz = torch.empty(n_utts, device=dev, dtype=f64) if want_logz else None
    if block_active is not None and (block_active.dtype != torch.uint8 or block_active.shape[0] < (N + 63) // 64):
        raise ValueError('block_active: uint8 [ceil(N / 64), blocks]')
    _lib.check(lib.beer_hmm_forward_backward_blocks(
        plan._h, _p(pdf_llh, f32), pdf_llh.stride(0), _p(frame_ref, f32, True), _p(utt_off, i64), n_utts,
        float(scale), _p(state_post, f32, True), _p(pdf_post, f32, True),
        pdf_post.stride(0) if pdf_post is not None else 0, _p(frame, f32, True),
        _p(utt_ell, f64), _p(utt_logz, f64, True), _p(unit_counts, f64, True),
        (1 if llh_log2 else 0) | (2 if lpost_relative else 0),
        _p(out_pdf_lpost, f32, True), out_pdf_lpost.stride(0) if out_pdf_lpost is not None else 0,
        _p(block_active, None, True), block_active.stride(0) if block_active is not None else 0, int(pdfs_per_block),
        _p(workspace), _stream()), 'beer_hmm_forward_backward')
    return dict(state_post=state_post, pdf_post=pdf_post, frame_exp_llh=frame, utt_exp_llh=utt_ell,
                utt_logz=utt_logz, workspace=workspace)


class ChainBatch:
    """Per-utterance alignment graphs (mkaligraph.py:18-39) of a ragged batch, as left-to-right chains on the
    device: `chain_off` [n+1] and per state the pdf id, ln a(j,j) and ln a(j,j+1) (last state: final weight)."""

    def __init__(self, graphs, device):
        offs, pdf, lself, lnext, linit = [0], [], [], [], []
        for g in graphs:
            init, final, trans, pmap = _graph_arrays(g)
            K = len(init)
            fin = np.isfinite(trans)
            band = np.eye(K, dtype=bool) | np.eye(K, k=1, dtype=bool)
            if K == 0 or (fin & ~band).any() or np.isfinite(init[1:]).any() or np.isfinite(final[:-1]).any() \
                    or not np.isfinite(init[0]) or not np.isfinite(final[-1]):
                raise ValueError('not a left-to-right chain (self loop + one arc to the next state, start in the '
                                 'first state, end in the last): use a GraphPlan for this graph')
            offs.append(offs[-1] + K)
            pdf.append(np.asarray(pmap, dtype=np.int32))
            lself.append(np.diagonal(trans).astype(np.float32))
            lnext.append(np.concatenate([np.diagonal(trans, 1), final[-1:]]).astype(np.float32))
            linit.append(init[0])
        self._set(np.asarray(offs, dtype=np.int64), np.concatenate(pdf) if pdf else np.zeros(0, np.int32),
                  np.concatenate(lself) if pdf else np.zeros(0, np.float32),
                  np.concatenate(lnext) if pdf else np.zeros(0, np.float32), np.asarray(linit, dtype=np.float32), device)

    @classmethod
    def from_arrays(cls, chain_off, pdf, log_self, log_next, log_init, device):
        """The flat form directly (host arrays): no dense K x K matrix per utterance."""
        self = cls.__new__(cls)
        self._set(np.asarray(chain_off, dtype=np.int64), np.asarray(pdf, dtype=np.int32),
                  np.asarray(log_self, dtype=np.float32), np.asarray(log_next, dtype=np.float32),
                  np.asarray(log_init, dtype=np.float32), device)
        return self

    def _set(self, offs, pdf, lself, lnext, linit, device):
        self.n_utts = len(offs) - 1
        self.lengths = np.diff(offs)
        self.max_len = int(self.lengths.max()) if self.n_utts else 0
        self.n_pdfs = int(pdf.max()) + 1 if len(pdf) else 0
        self.chain_off = torch.as_tensor(offs, device=device)
        self.pdf = torch.as_tensor(pdf, device=device)
        self.log_self = torch.as_tensor(lself, device=device)
        self.log_next = torch.as_tensor(lnext, device=device)
        self.log_init = torch.as_tensor(linit, device=device)
        self.row_stride = int(_lib.load().beer_hmm_chain_row_stride(max(self.max_len, 1)))
        if self.row_stride < 0:
            raise _lib.BeerB200Error('alignment chains longer than 1024 states are not supported')

    def workspace_bytes(self, N):
        return int(_lib.load().beer_hmm_chain_workspace_bytes(max(self.max_len, 1), int(N)))


def _graph_arrays(g):
    """(init, final, trans, pdf map) of a CompiledGraph-like object or a 4-tuple, as numpy."""
    if isinstance(g, (tuple, list)):
        init, final, trans, pmap = g
    else:
        init, final, trans, pmap = g.init_log_probs, g.final_log_probs, g.trans_log_probs, g.pdf_id_mapping
    def to_np(x):
        return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)
    return to_np(init).astype(np.float32), to_np(final).astype(np.float32), to_np(trans).astype(np.float32), \
        np.asarray(pmap)


def hmm_forward_backward_chains(chains, pdf_llh, frame_ref, utt_off, scale=1.0, first_utt=0, want_state_post=False,
                                want_frame_llh=False, want_logz=False, workspace=None, out_pdf_post=None,
                                out_utt_exp_llh=None):
    """Forward-backward of a ragged batch in which utterance i runs over chain `first_utt + i` of `chains`
    (hmm.py:73-92 with a per-utterance inference_graph).  `out_pdf_post` must be zero-filled by the caller.
    -> dict like hmm_forward_backward (state_post has row stride chains.row_stride)."""
    lib = require_cuda()
    N = pdf_llh.shape[0]
    n_utts = utt_off.numel() - 1
    dev = pdf_llh.device
    if first_utt + n_utts > chains.n_utts:
        raise ValueError('more utterances than alignment chains')
    if pdf_llh.shape[1] < chains.n_pdfs:
        raise ValueError('pdf_llh has fewer columns than the chains have pdfs')
    nbytes = chains.workspace_bytes(N)
    if workspace is None or workspace.numel() * workspace.element_size() < nbytes:
        workspace = torch.empty((nbytes + 3) // 4, device=dev, dtype=f32)
    state_post = torch.zeros(N, chains.row_stride, device=dev, dtype=f32) if want_state_post else None
    pdf_post = out_pdf_post if out_pdf_post is not None else torch.zeros(N, pdf_llh.shape[1], device=dev, dtype=f32)
    frame = torch.empty(N, device=dev, dtype=f32) if want_frame_llh else None
    utt_ell = out_utt_exp_llh if out_utt_exp_llh is not None else torch.empty(n_utts, device=dev, dtype=f64)
    utt_logz = torch.empty(n_utts, device=dev, dtype=f64) if want_logz else None
    chain_off = chains.chain_off[first_utt:first_utt + n_utts + 1]
    _lib.check(lib.beer_hmm_forward_backward_chains(
        _p(pdf_llh, f32), pdf_llh.stride(0), _p(frame_ref, f32, True), _p(utt_off, i64), n_utts, _p(chain_off, i64),
        _p(chains.pdf, i32), _p(chains.log_self, f32), _p(chains.log_next, f32),
        _p(chains.log_init[first_utt:first_utt + n_utts], f32), max(chains.max_len, 1), float(scale),
        _p(state_post, f32, True), _p(pdf_post, f32), pdf_post.stride(0), _p(frame, f32, True), _p(utt_ell, f64),
        _p(utt_logz, f64, True), _p(workspace), _stream()), 'beer_hmm_forward_backward_chains')
    return dict(state_post=state_post, pdf_post=pdf_post, frame_exp_llh=frame, utt_exp_llh=utt_ell,
                utt_logz=utt_logz, workspace=workspace)


def hmm_transition_posteriors(pdf_llh, state_post, utt_off, log_init, log_trans, pdf_map=None, scale=1.0, rows=None,
                              cols=None):
    """Per-step normalised transition posteriors [N - n_utts, R, C] of a ragged batch (graph.py:308-323);
    `rows` / `cols` (int32 state ids) keep a sub-block, e.g. unit ends x unit starts."""
    lib = require_cuda()
    N, K = state_post.shape
    # the kernel places utterance u's rows at t0 - u: every utterance must hold a frame, empty ones are dropped here
    nonempty = utt_off[1:] > utt_off[:-1]
    if not bool(nonempty.all()):
        utt_off = torch.cat([utt_off[:1], utt_off[1:][nonempty]]).contiguous()
    n_utts = utt_off.numel() - 1
    R = K if rows is None else rows.numel()
    Cn = K if cols is None else cols.numel()
    xi = torch.empty(max(N - n_utts, 0), R, Cn, device=pdf_llh.device, dtype=f32)
    _lib.check(lib.beer_hmm_transition_posteriors(
        _p(pdf_llh, f32), pdf_llh.stride(0), _p(pdf_map, i32, True), float(scale), _p(log_init, f32),
        _p(log_trans, f32), K, _p(state_post, f32), _p(utt_off, i64), n_utts, _p(rows, i32, True),
        0 if rows is None else R, _p(cols, i32, True), 0 if cols is None else Cn, _p(xi, f32), _stream()),
        'beer_hmm_transition_posteriors')
    return xi


def hmm_viterbi(plan, pdf_llh, utt_off, scale=1.0, workspace=None):
    """Best state path (int32 [N]) of every utterance of a ragged batch."""
    lib = require_cuda()
    N = pdf_llh.shape[0]
    n_utts = utt_off.numel() - 1
    nbytes = N * plan.n_states * 2
    if workspace is None or workspace.numel() * workspace.element_size() < nbytes:
        workspace = torch.empty((nbytes + 3) // 4 + 1, device=pdf_llh.device, dtype=f32)
    path = torch.empty(N, device=pdf_llh.device, dtype=i32)
    _lib.check(lib.beer_hmm_viterbi(plan._h, _p(pdf_llh, f32), pdf_llh.stride(0), _p(utt_off, i64), n_utts,
                                    float(scale), _p(path), _p(workspace), _stream()), 'beer_hmm_viterbi')
    return path


def accumulate_tc_supported(M, D):
    return bool(_lib.load().beer_accumulate_tc_supported(int(M), int(D)))


def accumulate_stats(X, acc_normal, pdf_post=None, pdf_llh=None, comp_llh=None, comp_off=None, Kp=None,
                     tensor_cores=None):
    """acc_normal [M, 2D+2] (fp64) += posterior-weighted statistics.  `tensor_cores`: None = use
    the tcgen05 kernel when the shape has one, False = SIMT kernel, True = require tcgen05."""
    lib = require_cuda()
    N, D = X.shape
    M = acc_normal.shape[0]
    if Kp is None:
        Kp = M if comp_llh is None else (comp_off.numel() - 1 if comp_off is not None else pdf_llh.shape[1])
    if tensor_cores is None:
        tensor_cores = accumulate_tc_supported(M, D) and X.data_ptr() % 16 == 0
    fn = lib.beer_accumulate_stats_tc if tensor_cores else lib.beer_accumulate_stats
    _lib.check(fn(
        _p(X, f32), N, D, _p(pdf_post, f32, True), pdf_post.stride(0) if pdf_post is not None else 0,
        _p(pdf_llh, f32, True), pdf_llh.stride(0) if pdf_llh is not None else 0, _p(comp_llh, f32, True),
        _p(comp_off, i32, True), Kp, M, _p(acc_normal, f64), _stream()), 'beer_accumulate_stats')
    return acc_normal


def accumulate_path_supported(M, D):
    return M <= 128 and D in (20, 40)


def accumulate_stats_path(X, acc_normal, pdf_ids, scale=1.0):
    """acc_normal [M, 2D+2] (fp64) += statistics of one-hot posteriors scale * onehot(pdf_ids[t]) (Viterbi training).
    `pdf_ids`: int32 with at least N rounded up to 4 entries allocated (a view of a padded buffer)."""
    lib = require_cuda()
    N, D = X.shape
    M = acc_normal.shape[0]
    if pdf_ids.untyped_storage().nbytes() - pdf_ids.storage_offset() * 4 < ((N + 3) // 4) * 16:
        raise ValueError('pdf_ids must be a view of a buffer padded to a multiple of 4 entries')
    _lib.check(lib.beer_accumulate_stats_path(_p(X, f32), N, D, _p(pdf_ids, i32), float(scale), M, _p(acc_normal, f64),
                                              _stream()), 'beer_accumulate_stats_path')
    return acc_normal


def mixture_weight_stats(acc_normal, D, comp_off=None, Kp=None):
    lib = require_cuda()
    M = acc_normal.shape[0]
    if Kp is None:
        Kp = comp_off.numel() - 1
    out = torch.empty(M, device=acc_normal.device, dtype=f64)
    _lib.check(lib.beer_mixture_weight_stats(_p(acc_normal, f64), M, D, _p(comp_off, i32, True), Kp, _p(out),
                                             _stream()), 'beer_mixture_weight_stats')
    return out


def normalgamma_update(prior, post, acc, stats_scale=1.0, lrate=1.0):
    """In-place natural-gradient step; prior/post = (mean, scale, shape, rates)."""
    lib = require_cuda()
    M, D = post[0].shape
    _lib.check(lib.beer_normalgamma_update(*[_p(t, f32) for t in prior], *[_p(t, f32) for t in post],
                                           _p(acc, f64), float(stats_scale), float(lrate), M, D, _stream()),
               'beer_normalgamma_update')


def normalgamma_kl(prior, post, out=None):
    lib = require_cuda()
    M, D = post[0].shape
    if out is None:
        out = torch.zeros(1, device=post[0].device, dtype=f64)
    _lib.check(lib.beer_normalgamma_kl(*[_p(t, f32) for t in prior], *[_p(t, f32) for t in post], M, D,
                                       _p(out, f64), _stream()), 'beer_normalgamma_kl')
    return out


def dirichlet_update(prior, post, acc, stats_scale=1.0, lrate=1.0):
    lib = require_cuda()
    p2, q2 = prior.reshape(-1, prior.shape[-1]), post.reshape(-1, post.shape[-1])
    K, Cc = q2.shape
    _lib.check(lib.beer_dirichlet_update(_p(p2, f32), _p(q2, f32), _p(acc, f64), float(stats_scale),
                                         float(lrate), K, Cc, _stream()), 'beer_dirichlet_update')


def dirichlet_kl(prior, post, out=None):
    lib = require_cuda()
    p2, q2 = prior.reshape(-1, prior.shape[-1]), post.reshape(-1, post.shape[-1])
    K, Cc = q2.shape
    if out is None:
        out = torch.zeros(1, device=post.device, dtype=f64)
    _lib.check(lib.beer_dirichlet_kl(_p(p2, f32), _p(q2, f32), K, Cc, _p(out, f64), _stream()),
               'beer_dirichlet_kl')
    return out


# ---------------------------------------------------------------------------
# beer.dists accessors
# ---------------------------------------------------------------------------

def normal_sufficient_statistics(X):
    """T(x) = [x, -x^2/2, -1/2, 1/2] -> [N, 2D+2]."""
    lib = require_cuda()
    N, D = X.shape
    out = torch.empty(N, 2 * D + 2, device=X.device, dtype=f32)
    _lib.check(lib.beer_normal_sufficient_statistics(_p(X, f32), N, D, _p(out), _stream()),
               'beer_normal_sufficient_statistics')
    return out


def normalgamma_natural_params(mean, scale, shape, rates):
    lib = require_cuda()
    M, D = mean.shape
    out = torch.empty(M, 2 * D + 2, device=mean.device, dtype=f32)
    _lib.check(lib.beer_normalgamma_natural_params(_p(mean, f32), _p(scale, f32), _p(shape, f32), _p(rates, f32),
                                                   M, D, _p(out), _stream()), 'beer_normalgamma_natural_params')
    return out


def normalgamma_from_natural(nat):
    """-> (mean [M,D], scale [M,1], shape [M,1], rates [M,D])."""
    lib = require_cuda()
    nat2 = nat.reshape(-1, nat.shape[-1]).contiguous()
    M, Q = nat2.shape
    D = (Q - 2) // 2
    dev = nat.device
    mean, rates = torch.empty(M, D, device=dev, dtype=f32), torch.empty(M, D, device=dev, dtype=f32)
    scale, shape = torch.empty(M, 1, device=dev, dtype=f32), torch.empty(M, 1, device=dev, dtype=f32)
    _lib.check(lib.beer_normalgamma_from_natural(_p(nat2, f32), M, D, _p(mean), _p(scale), _p(shape), _p(rates),
                                                 _stream()), 'beer_normalgamma_from_natural')
    return mean, scale, shape, rates


def normalgamma_log_norm(mean, scale, shape, rates):
    lib = require_cuda()
    M, D = rates.shape
    out = torch.empty(M, device=rates.device, dtype=f64)
    _lib.check(lib.beer_normalgamma_log_norm(_p(scale, f32), _p(shape, f32), _p(rates, f32), M, D, _p(out),
                                             _stream()), 'beer_normalgamma_log_norm')
    return out


def _dirichlet_rows(fn_name, conc, out_dtype=f32, per_row=False):
    lib = require_cuda()
    c2 = conc.reshape(-1, conc.shape[-1]).contiguous()
    K, Cc = c2.shape
    out = torch.empty(K if per_row else (K, Cc), device=conc.device, dtype=out_dtype)
    _lib.check(getattr(lib, fn_name)(_p(c2, f32), K, Cc, _p(out), _stream()), fn_name)
    if per_row:
        return out if conc.dim() > 1 else out[0]
    return out.reshape(conc.shape)


def dirichlet_natural_params(conc):
    return _dirichlet_rows('beer_dirichlet_natural_params', conc)


def dirichlet_expected_stats(conc):
    return _dirichlet_rows('beer_dirichlet_expected_stats', conc)


def dirichlet_from_natural(nat):
    return _dirichlet_rows('beer_dirichlet_from_natural', nat)


def dirichlet_log_norm(conc):
    return _dirichlet_rows('beer_dirichlet_log_norm', conc, out_dtype=f64, per_row=True)


def segment_logsumexp(comp_llh, comp_off=None, Kp=1, out=None):
    """pdf_llh[t, k] = logsumexp of comp_llh[t, comp_off[k]:comp_off[k+1]]."""
    lib = require_cuda()
    N, M = comp_llh.shape
    if comp_off is not None:
        Kp = comp_off.numel() - 1
    pdf_llh = out if out is not None else torch.empty(N, Kp, device=comp_llh.device, dtype=f32)
    _lib.check(lib.beer_segment_logsumexp(_p(comp_llh, f32), N, M, _p(comp_off, i32, True), Kp, _p(pdf_llh, f32),
                                          pdf_llh.stride(0), _stream()), 'beer_segment_logsumexp')
    return pdf_llh


def path_posteriors(path, n_pdfs, pdf_map=None, scale=1.0, pdf_llh=None, frame_ref=None, want_post=True,
                    want_frame_llh=True, out_post=None, out_frame=None):
    """Posteriors / per-frame expected llh of a given state path (int32 [N])."""
    lib = require_cuda()
    N = path.numel()
    dev = path.device
    post = out_post if out_post is not None else (torch.empty(N, n_pdfs, device=dev, dtype=f32) if want_post else None)
    frame = out_frame if out_frame is not None else (torch.empty(N, device=dev, dtype=f32) if want_frame_llh else None)
    _lib.check(lib.beer_path_posteriors(_p(path, i32), N, _p(pdf_map, i32, True), float(scale),
                                        _p(pdf_llh, f32, True), pdf_llh.stride(0) if pdf_llh is not None else 0,
                                        _p(frame_ref, f32, True), _p(post, f32, True), n_pdfs, n_pdfs,
                                        _p(frame, f32, True), _stream()), 'beer_path_posteriors')
    return post, frame


# ---------------------------------------------------------------------------
# fbank front-end
# ---------------------------------------------------------------------------

def fbank(signal, window, filters_t, frame_shift, preemph, fft_len):
    """signal [L] fp32 -> log(1 + mel energies) [n_frames, n_filters]."""
    lib = require_cuda()
    L, flen = signal.numel(), window.numel()
    n_filters = filters_t.shape[1]
    nframes = max(0, (L - flen) // frame_shift + 1) if L >= flen else 0
    out = torch.empty(nframes, n_filters, device=signal.device, dtype=f32)
    if nframes == 0:
        return out
    _lib.check(lib.beer_fbank(_p(signal, f32), L, flen, int(frame_shift), float(preemph), _p(window, f32),
                              _p(filters_t, f32), int(fft_len), n_filters, _p(out), _stream()), 'beer_fbank')
    return out


def short_term_mspec(signal, window, frame_shift, preemph, fft_len, dc_offset, filters_t=None, log_offset=1e-6):
    """signal [L] fp32 -> magnitude spectrum [n_frames, fft_len / 2] (filters_t None) or log(log_offset + mel energies)."""
    lib = require_cuda()
    L, flen = signal.numel(), window.numel()
    width = fft_len // 2 if filters_t is None else filters_t.shape[1]
    nframes = max(0, (L - flen) // frame_shift + 1) if L >= flen else 0
    out = torch.empty(nframes, width, device=signal.device, dtype=f32)
    if nframes == 0:
        return out
    _lib.check(lib.beer_short_term_mspec(_p(signal, f32), L, flen, int(frame_shift), float(preemph), float(dc_offset),
                                         _p(window, f32), _p(filters_t, f32, True), int(fft_len), int(width),
                                         float(log_offset), _p(out), _stream()), 'beer_short_term_mspec')
    return out


def add_deltas(fea, wlen):
    lib = require_cuda()
    T, F = fea.shape
    out = torch.empty_like(fea)
    _lib.check(lib.beer_add_deltas(_p(fea, f32), T, F, int(wlen), _p(out), _stream()), 'beer_add_deltas')
    return out


# ---------------------------------------------------------------------------
# mixture path without per-Gaussian llhs in HBM (csrc/mix16.cu)
# ---------------------------------------------------------------------------

def mix16_supported(M, D, C):
    return bool(_lib.load().beer_mix16_supported(int(M), int(D), int(C)))


class Mix16:
    """Buffers and calls of the fp16-split mixture kernels for M Gaussians in pdfs of C, dimension D."""

    def __init__(self, M, D, C, device):
        lib = require_cuda()
        if not lib.beer_mix16_supported(int(M), int(D), int(C)):
            raise _lib.BeerB200Error('no mix16 kernels for this shape')
        self.M, self.D, self.C, self.Kp, self.device = int(M), int(D), int(C), int(M) // int(C), device
        sz = self._geometry(0)
        self.NB, self.KP = int(sz[4]), int(sz[5])
        self.wimg = torch.zeros(int(sz[1]), device=device, dtype=torch.float16)
        self.wtm = torch.zeros(int(sz[2]), device=device, dtype=i32)
        self.k12 = torch.zeros(int(sz[3]), device=device, dtype=f32)
        self._absmax = torch.zeros(D, device=device, dtype=i32)

    def _geometry(self, N):
        sz = np.zeros(6, dtype=np.int64)
        _lib.check(_lib.load().beer_mix16_geometry(self.M, self.D, self.C, int(N), sz.ctypes.data), 'beer_mix16_geometry')
        return sz

    def image_halfs(self, N):
        return int(self._geometry(N)[0])

    def build_images(self, X, out=None):
        """Feature images of a run of frames: dict(alpha [2D], img1, img2, N).  `out` = a dict from an earlier call
        with at least as many frames (its buffers are reused)."""
        lib = _lib.load()
        N = X.shape[0]
        n = self.image_halfs(N)
        if out is None or out['img1'].numel() < n:
            out = dict(alpha=torch.empty(2 * self.D, device=X.device, dtype=f32),
                       img1=torch.empty(n, device=X.device, dtype=torch.float16),
                       img2=torch.empty(n, device=X.device, dtype=torch.float16))
        _lib.check(lib.beer_mix16_feature_images(_p(X, f32), N, self.D, _p(out['alpha']), _p(self._absmax),
                                                 _p(out['img1']), _p(out['img2']), _stream()), 'beer_mix16_feature_images')
        out['N'] = N
        return out

    def pack(self, W, bias, alpha):
        _lib.check(_lib.load().beer_mix16_pack(_p(W, f32), _p(bias, f32), _p(alpha, f32), self.M, self.D, self.C,
                                               _p(self.wimg), _p(self.wtm), _p(self.k12), _stream()),
                   'beer_mix16_pack')

    def frame_ref(self, X, ref, out=None):
        N = X.shape[0]
        out = out if out is not None else torch.empty(N, device=X.device, dtype=f32)
        _lib.check(_lib.load().beer_mix16_frame_ref(_p(X, f32), N, self.D, _p(ref, f32), _p(out), _stream()),
                   'beer_mix16_frame_ref')
        return out

    def emission(self, images, out=None):
        """llh2 [N, Kp]: log2-domain pdf llhs in offset form (add frame_ref / ln 2 for absolute values)."""
        N = images['N']
        llh2 = out if out is not None else torch.empty(N, self.Kp, device=self.device, dtype=f32)
        _lib.check(_lib.load().beer_mix16_emission(_p(images['img1']), N, self.D, _p(self.wimg), _p(self.k12), self.M,
                                                   self.C, _p(llh2, f32), llh2.stride(0), _stream()),
                   'beer_mix16_emission')
        return llh2

    @property
    def pdfs_per_block(self):
        """pdfs of one tile of 128 Gaussians of the statistics kernel: the block width of the activity map."""
        return 128 // self.C

    def accumulate(self, images, pdf_lpost, llh2, acc_normal, scale=1.0, relative=False, block_active=None):
        """`pdf_lpost` [N, Kp]: log2 of the (scaled) pdf posteriors, see `log2_posteriors`; for single-Gaussian pdfs
        (C = 1) the posteriors themselves, `llh2` is then not read.  `relative`: `pdf_lpost` already holds
        log2 posterior - llh2 (`hmm_forward_backward(..., lpost_relative=True)`), `llh2` must be None.  `block_active`
        (uint8 [ceil(N / 64), >= M / 128], from `hmm_forward_backward(..., block_active=..., pdfs_per_block=
        self.pdfs_per_block)`): work only on the marked (frame tile, Gaussian tile) pairs; the others are exact zeros."""
        N = images['N']
        if self.C > 1 and relative != (llh2 is None):
            raise ValueError('relative=True takes llh2=None (and only then)')
        ld_llh = llh2.stride(0) if llh2 is not None else pdf_lpost.stride(0)
        _lib.check(_lib.load().beer_mix16_accumulate_blocks(
            _p(images['img1']), _p(images['img2']), N, self.D, _p(self.wtm), _p(self.k12),
            _p(images['alpha']), self.M, self.C, _p(pdf_lpost, f32), pdf_lpost.stride(0), _p(llh2, f32, True), ld_llh,
            float(scale), _p(block_active, None, True), block_active.stride(0) if block_active is not None else 0,
            _p(acc_normal, f64), _stream()), 'beer_mix16_accumulate')
        return acc_normal

    def gmm_posteriors(self, llh2, frame_ref, utt_off, scale=1.0, out=None, out_utt_exp_llh=None, want_frame_llh=False):
        """GMM without an HMM: softmax over the pseudo-pdfs of every frame -> (lpost [N, Kp], frame llh or None);
        `out_utt_exp_llh` (fp64, zeroed by the caller) += the per-utterance sums of the frame llhs."""
        N = llh2.shape[0]
        lpost = out if out is not None else torch.empty(N, self.Kp, device=llh2.device, dtype=f32)
        frame = torch.empty(N, device=llh2.device, dtype=f32) if want_frame_llh else None
        n_utts = utt_off.numel() - 1 if utt_off is not None else 0
        _lib.check(_lib.load().beer_mix16_gmm_posteriors(
            _p(llh2, f32), N, self.Kp, llh2.stride(0), _p(frame_ref, f32, True), _p(utt_off, i64, True), n_utts,
            float(scale), _p(lpost, f32), lpost.stride(0), _p(frame, f32, True), _p(out_utt_exp_llh, f64, True),
            _stream()), 'beer_mix16_gmm_posteriors')
        return lpost, frame

    def log2_posteriors(self, pdf_post, out=None):
        N = pdf_post.shape[0]
        out = out if out is not None else torch.empty(N, self.Kp, device=pdf_post.device, dtype=f32)
        _lib.check(_lib.load().beer_mix16_log2_posteriors(_p(pdf_post, f32), N, self.Kp, pdf_post.stride(0), _p(out, f32),
                                                          out.stride(0), _stream()), 'beer_mix16_log2_posteriors')
        return out


# ---------------------------------------------------------------------------
# roofline probes (measurement only)
# ---------------------------------------------------------------------------

def _timed(fn, reps):
    """Best CUDA-event time in seconds of `fn()` launched on the current stream."""
    fn()
    torch.cuda.synchronize()
    best = float('inf')
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def probe_mma_tflops(kind, n_mma=20000, reps=5):
    """Measured dispatch-limited tcgen05 peak (TFLOP/s, dense) of MMA kind 'tf32' or 'f16' on this GPU."""
    lib = require_cuda()
    flops = C.c_double(0.0)
    k = {'tf32': 0, 'f16': 1}[kind]

    def run():
        _lib.check(lib.beer_probe_mma(k, int(n_mma), C.byref(flops), _stream()), 'beer_probe_mma')
    t = _timed(run, reps)
    return flops.value / t / 1e12


def probe_dram_gbs(mode, nbytes=4 << 30, reps=5):
    """Measured DRAM stream rate in GB/s: mode 'fill_st' (float4 stores), 'fill_bulk' (bulk copies shared -> global)
    or 'read'."""
    lib = require_cuda()
    buf = torch.empty(nbytes // 4, device='cuda', dtype=f32)
    sink = torch.zeros(1, device='cuda', dtype=f32)
    if mode == 'read':
        buf.zero_()

        def run():
            _lib.check(lib.beer_probe_read(_p(buf), nbytes, _p(sink), _stream()), 'beer_probe_read')
    else:
        m = {'fill_st': 0, 'fill_bulk': 1}[mode]

        def run():
            _lib.check(lib.beer_probe_fill(_p(buf), nbytes, m, _stream()), 'beer_probe_fill')
    t = _timed(run, reps)
    return nbytes / t / 1e9


def probe_tma_gbs(mode, chunk_bytes, stages, src_mib=64, copies=2000, reps=3, issuers=1, shared_walk=False):
    """Aggregate global -> shared copy-engine rate (GB/s over 148 SMs) from an L2-sized buffer."""
    lib = require_cuda()
    buf = torch.zeros(src_mib << 18, device='cuda', dtype=f32)
    m = {'bulk': 0, 'tensor': 1}[mode]

    def run():
        _lib.check(lib.beer_probe_tma(_p(buf), buf.numel() * 4, m, int(chunk_bytes), int(stages), int(copies), int(issuers),
                                      int(bool(shared_walk)), _stream()),
                   'beer_probe_tma')
    t = _timed(run, reps)
    return 148 * copies * chunk_bytes / t / 1e9


def probe_mma_shape_cycles(N, n_acc, n_buf, a_tmem, elect=True, n_mma=20000, reps=3, mhz=1965.0):
    """Average cycles per tcgen05.mma (kind::f16, M = 128, width N, one k-step) at `mhz`."""
    lib = require_cuda()
    flops = C.c_double(0.0)

    def run():
        _lib.check(lib.beer_probe_mma_shape(int(n_mma), int(N), int(n_acc), int(n_buf), int(bool(a_tmem)), int(bool(elect)),
                                            C.byref(flops), _stream()), 'beer_probe_mma_shape')
    t = _timed(run, reps)
    return t / n_mma * mhz * 1e6
