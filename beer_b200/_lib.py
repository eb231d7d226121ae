"""ctypes binding of libbeer_b200.so (the C ABI declared in include/beer_b200.h).

There is no CPU fallback: if the library is missing or a call fails, an exception
is raised.  Loading the library itself needs no GPU (the symbol table can be
inspected on a CPU-only machine); calling any kernel entry point does.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(_HERE, 'lib', 'libbeer_b200.so')

c_f32p = C.c_void_p   # device pointers travel as integers
c_ptr = C.c_void_p

# name -> (restype, argtypes); mirrors include/beer_b200.h one to one
SIGNATURES = {
    'beer_b200_version': (C.c_int, []),
    'beer_normalgamma_expected_stats': (C.c_int, [c_ptr] * 4 + [C.c_int, C.c_int, c_ptr, c_ptr]),
    'beer_dirichlet_expected_logw': (C.c_int, [c_ptr, C.c_int, C.c_int, c_ptr, c_ptr]),
    'beer_emission_prepare': (C.c_int, [c_ptr] * 5 + [C.c_int, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    'beer_emission_llh': (C.c_int, [c_ptr, C.c_int64, C.c_int, c_ptr, c_ptr, c_ptr, C.c_int, c_ptr,
                                    C.c_int, c_ptr, C.c_int64, c_ptr, c_ptr, c_ptr]),
    'beer_emission_tc_supported': (C.c_int, [C.c_int, C.c_int, C.c_int]),
    'beer_emission_tc_image_floats': (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    'beer_emission_tc_pack': (C.c_int, [c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr]),
    'beer_emission_llh_tc': (C.c_int, [c_ptr, C.c_int64, C.c_int, c_ptr, c_ptr, C.c_int, C.c_int, c_ptr,
                                       C.c_int64, c_ptr, c_ptr, c_ptr]),
    'beer_graph_plan_create': (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_void_p)]),
    'beer_graph_plan_destroy': (None, [c_ptr]),
    'beer_graph_plan_info': (C.c_int, [c_ptr, c_ptr]),
    'beer_hmm_workspace_bytes': (C.c_int64, [c_ptr, C.c_int64]),
    'beer_hmm_forward_backward': (C.c_int, [c_ptr, c_ptr, C.c_int64, c_ptr, c_ptr, C.c_int, C.c_float,
                                            c_ptr, c_ptr, C.c_int64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    'beer_hmm_unit_count_size': (C.c_int, [c_ptr]),
    'beer_hmm_forward_backward_units': (C.c_int, [c_ptr, c_ptr, C.c_int64, c_ptr, c_ptr, C.c_int, C.c_float,
                                                  c_ptr, c_ptr, C.c_int64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    'beer_hmm_viterbi': (C.c_int, [c_ptr, c_ptr, C.c_int64, c_ptr, C.c_int, C.c_float, c_ptr, c_ptr,
                                   c_ptr]),
    'beer_accumulate_stats': (C.c_int, [c_ptr, C.c_int64, C.c_int, c_ptr, C.c_int64, c_ptr, C.c_int64,
                                        c_ptr, c_ptr, C.c_int, C.c_int, c_ptr, c_ptr]),
    'beer_accumulate_tc_supported': (C.c_int, [C.c_int, C.c_int]),
    'beer_accumulate_stats_tc': (C.c_int, [c_ptr, C.c_int64, C.c_int, c_ptr, C.c_int64, c_ptr, C.c_int64,
                                           c_ptr, c_ptr, C.c_int, C.c_int, c_ptr, c_ptr]),
    'beer_accumulate_stats_path': (C.c_int, [c_ptr, C.c_int64, C.c_int, c_ptr, C.c_float, C.c_int, c_ptr, c_ptr]),
    'beer_mixture_weight_stats': (C.c_int, [c_ptr, C.c_int, C.c_int, c_ptr, C.c_int, c_ptr, c_ptr]),
    'beer_normalgamma_update': (C.c_int, [c_ptr] * 9 + [C.c_double, C.c_double, C.c_int, C.c_int, c_ptr]),
    'beer_normalgamma_kl': (C.c_int, [c_ptr] * 8 + [C.c_int, C.c_int, c_ptr, c_ptr]),
    'beer_dirichlet_update': (C.c_int, [c_ptr, c_ptr, c_ptr, C.c_double, C.c_double, C.c_int, C.c_int,
                                        c_ptr]),
    'beer_dirichlet_kl': (C.c_int, [c_ptr, c_ptr, C.c_int, C.c_int, c_ptr, c_ptr]),
    'beer_normal_sufficient_statistics': (C.c_int, [c_ptr, C.c_int64, C.c_int, c_ptr, c_ptr]),
    'beer_normalgamma_natural_params': (C.c_int, [c_ptr] * 4 + [C.c_int, C.c_int, c_ptr, c_ptr]),
    'beer_normalgamma_from_natural': (C.c_int, [c_ptr, C.c_int, C.c_int] + [c_ptr] * 5),
    'beer_normalgamma_log_norm': (C.c_int, [c_ptr] * 3 + [C.c_int, C.c_int, c_ptr, c_ptr]),
    'beer_dirichlet_natural_params': (C.c_int, [c_ptr, C.c_int, C.c_int, c_ptr, c_ptr]),
    'beer_dirichlet_expected_stats': (C.c_int, [c_ptr, C.c_int, C.c_int, c_ptr, c_ptr]),
    'beer_dirichlet_log_norm': (C.c_int, [c_ptr, C.c_int, C.c_int, c_ptr, c_ptr]),
    'beer_dirichlet_from_natural': (C.c_int, [c_ptr, C.c_int, C.c_int, c_ptr, c_ptr]),
    'beer_segment_logsumexp': (C.c_int, [c_ptr, C.c_int64, C.c_int, c_ptr, C.c_int, c_ptr, C.c_int64, c_ptr]),
    'beer_fbank': (C.c_int, [c_ptr, C.c_int64, C.c_int, C.c_int, C.c_float, c_ptr, c_ptr, C.c_int, C.c_int, c_ptr, c_ptr]),
    'beer_short_term_mspec': (C.c_int, [c_ptr, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_float, c_ptr, c_ptr, C.c_int,
                                        C.c_int, C.c_float, c_ptr, c_ptr]),
    'beer_add_deltas': (C.c_int, [c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr]),
    'beer_hmm_chain_row_stride': (C.c_int, [C.c_int]),
    'beer_hmm_chain_workspace_bytes': (C.c_int64, [C.c_int, C.c_int64]),
    'beer_hmm_forward_backward_chains': (C.c_int, [c_ptr, C.c_int64, c_ptr, c_ptr, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                                                   c_ptr, C.c_int, C.c_float, c_ptr, c_ptr, C.c_int64, c_ptr, c_ptr,
                                                   c_ptr, c_ptr, c_ptr]),
    'beer_hmm_transition_posteriors': (C.c_int, [c_ptr, C.c_int64, c_ptr, C.c_float, c_ptr, c_ptr, C.c_int, c_ptr,
                                                 c_ptr, C.c_int, c_ptr, C.c_int, c_ptr, C.c_int, c_ptr, c_ptr]),
    'beer_hmm_forward_backward_ex': (C.c_int, [c_ptr, c_ptr, C.c_int64, c_ptr, c_ptr, C.c_int, C.c_float,
                                               c_ptr, c_ptr, C.c_int64, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, c_ptr, C.c_int64, c_ptr, c_ptr]),
    'beer_hmm_lpost_supported': (C.c_int, [c_ptr]),
    'beer_hmm_block_activity_supported': (C.c_int, [c_ptr, C.c_int]),
    'beer_hmm_forward_backward_blocks': (C.c_int, [c_ptr, c_ptr, C.c_int64, c_ptr, c_ptr, C.c_int, C.c_float,
                                                   c_ptr, c_ptr, C.c_int64, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, c_ptr,
                                                   C.c_int64, c_ptr, C.c_int64, C.c_int, c_ptr, c_ptr]),
    'beer_mix16_weight_exponent': (C.c_int, [C.c_float]),
    'beer_mix16_accumulate_blocks': (C.c_int, [c_ptr, c_ptr, C.c_int64, C.c_int, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int,
                                               c_ptr, C.c_int64, c_ptr, C.c_int64, C.c_float, c_ptr, C.c_int64, c_ptr,
                                               c_ptr]),
    'beer_mix16_gmm_posteriors': (C.c_int, [c_ptr, C.c_int64, C.c_int, C.c_int64, c_ptr, c_ptr, C.c_int, C.c_float, c_ptr,
                                            C.c_int64, c_ptr, c_ptr, c_ptr]),
    'beer_mix16_log2_posteriors': (C.c_int, [c_ptr, C.c_int64, C.c_int, C.c_int64, c_ptr, C.c_int64, c_ptr]),
    'beer_mix16_set_trace': (None, [c_ptr]),
    'beer_path_accumulate_mix': (C.c_int, [c_ptr, C.c_int64, C.c_int, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, c_ptr, C.c_float,
                                           c_ptr, c_ptr, c_ptr]),
    'beer_emission_bwd_supported': (C.c_int, [C.c_int, C.c_int]),
    'beer_emission_bwd_image_bytes': (C.c_int64, [C.c_int, C.c_int]),
    'beer_emission_bwd_pack': (C.c_int, [c_ptr, C.c_int, C.c_int, C.c_int64, c_ptr, c_ptr, c_ptr, c_ptr]),
    'beer_emission_llh_bwd': (C.c_int, [c_ptr, C.c_int64, C.c_int, c_ptr, c_ptr, C.c_int, c_ptr, C.c_int64, c_ptr, C.c_int64,
                                        c_ptr, C.c_int64, c_ptr, c_ptr, C.c_float, c_ptr, c_ptr]),
    'beer_mix16_supported': (C.c_int, [C.c_int, C.c_int, C.c_int]),
    'beer_mix16_geometry': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, c_ptr]),
    'beer_mix16_feature_images': (C.c_int, [c_ptr, C.c_int64, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    'beer_mix16_pack': (C.c_int, [c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    'beer_mix16_frame_ref': (C.c_int, [c_ptr, C.c_int64, C.c_int, c_ptr, c_ptr, c_ptr]),
    'beer_mix16_emission': (C.c_int, [c_ptr, C.c_int64, C.c_int, c_ptr, c_ptr, C.c_int, C.c_int, c_ptr,
                                      C.c_int64, c_ptr]),
    'beer_mix16_accumulate': (C.c_int, [c_ptr, c_ptr, C.c_int64, C.c_int, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int,
                                        c_ptr, C.c_int64, c_ptr, C.c_int64, C.c_float, c_ptr, c_ptr]),
    'beer_probe_mma': (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double), c_ptr]),
    'beer_probe_mma_shape': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), c_ptr]),
    'beer_probe_tma': (C.c_int, [c_ptr, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_ptr]),
    'beer_probe_fill': (C.c_int, [c_ptr, C.c_int64, C.c_int, c_ptr]),
    'beer_probe_read': (C.c_int, [c_ptr, C.c_int64, c_ptr, c_ptr]),
    'beer_path_posteriors': (C.c_int, [c_ptr, C.c_int64, c_ptr, C.c_float, c_ptr, C.c_int64, c_ptr, c_ptr,
                                       C.c_int64, C.c_int, c_ptr, c_ptr]),
}

_lib = None


class BeerB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise BeerB200Error(
            f'{LIBPATH} is missing: build it with `python -m beer_b200.build` '
            '(there is no CPU or PyTorch fallback for the VB E-step)')
    lib = C.CDLL(LIBPATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the ABI and the binding diverge
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


_ENGINE_ERRORS = {-1: 'invalid argument', -2: 'shape not supported by the sm_100a kernels',
                  -3: 'allocation failed'}


def check(code, what):
    if code == 0:
        return
    if code < 0:
        raise BeerB200Error(f'{what}: {_ENGINE_ERRORS.get(code, code)}')
    raise BeerB200Error(f'{what}: CUDA error {code}')
