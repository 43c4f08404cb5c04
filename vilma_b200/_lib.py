"""ctypes binding of libvilma_b200.so (the C ABI in include/vilma_b200.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('VILMA_B200_LIB', os.path.join(_HERE, 'libvilma_b200.so'))
_lib = None

c_dp = C.POINTER(C.c_double)
c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)

# name -> (restype, argtypes); must list every symbol include/vilma_b200.h declares
SIGNATURES = {
    'vb_abi_version': (C.c_int, []),
    'vb_source_hash': (C.c_char_p, []),
    'vb_last_error': (C.c_char_p, []),
    'vb_set_option': (C.c_int, [C.c_char_p, C.c_int64]),
    'vb_ld_sym_nmax': (C.c_int64, []),
    'vb_ld_fac_nmax': (C.c_int64, []),
    'vb_debug_tile_plan': (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int),
                                     C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    'vb_ctx_create': (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    'vb_ctx_destroy': (C.c_int, [C.c_void_p]),
    'vb_ctx_sync': (C.c_int, [C.c_void_p]),
    'vb_ctx_launch_count': (C.c_int64, [C.c_void_p]),
    'vb_ctx_profile': (C.c_int, [C.c_void_p, C.c_int]),
    'vb_ctx_profile_read': (C.c_int, [C.c_void_p, c_dp, c_i64p]),
    'vb_ld_create': (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_i64p, c_i64p, C.POINTER(C.c_void_p)]),
    'vb_ld_destroy': (C.c_int, [C.c_void_p]),
    'vb_ld_set_dense': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int]),
    'vb_ld_set_factor': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]),
    'vb_ld_finalize': (C.c_int, [C.c_void_p, c_i64p, C.c_int64]),
    'vb_ld_dot': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'vb_ld_bytes': (C.c_int64, [C.c_void_p]),
    'vb_setup_nmax': (C.c_int64, []),
    'vb_setup_dense': (C.c_int, [C.c_void_p, C.c_int64, c_i64p] + [C.c_void_p] * 10),
    'vb_fit_create': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    'vb_fit_destroy': (C.c_int, [C.c_void_p]),
    'vb_fit_set_fusion': (C.c_int, [C.c_void_p, C.c_int]),
    'vb_fit_set_snp_data': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'vb_fit_set_mixture': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'vb_fit_set_hyper': (C.c_int, [C.c_void_p, C.c_void_p]),
    'vb_fit_set_delta_grad': (C.c_int, [C.c_void_p, C.c_void_p]),
    'vb_fit_set_tau': (C.c_int, [C.c_void_p, C.c_void_p]),
    'vb_fit_set_params': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'vb_fit_get_params': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'vb_fit_set_params_dev': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'vb_fit_get_params_dev': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'vb_fit_set_shard': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    'vb_fit_set_params_shard': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'vb_fit_get_params_shard': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'vb_host_accessible': (C.c_int, [C.c_void_p]),
    'vb_host_register': (C.c_int, [C.c_void_p, C.c_int64]),
    'vb_host_unregister': (C.c_int, [C.c_void_p]),
    'vb_fit_eval': (C.c_int, [C.c_void_p, C.c_void_p]),
    'vb_fit_beta_trial': (C.c_int, [C.c_void_p, C.c_double, C.c_void_p]),
    'vb_fit_refresh_delta': (C.c_int, [C.c_void_p, C.c_void_p]),
    'vb_fit_accept': (C.c_int, [C.c_void_p]),
    'vb_fit_sum_annotations': (C.c_int, [C.c_void_p, C.c_void_p]),
    'vb_fit_posterior': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'vb_fit_pm_diff': (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_void_p]),
    'vb_fit_pm_mark': (C.c_int, [C.c_void_p, C.c_int]),
    'vb_fit_vi_sigma': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'vb_fit_init_delta': (C.c_int, [C.c_void_p, C.c_void_p]),
    'vb_fit_init_mu': (C.c_int, [C.c_void_p]),
}


class StepIO(C.Structure):
    """vb_step_io of include/vilma_b200.h"""
    _fields_ = [('L', C.c_double * 5), ('line_search_rate', C.c_double),
                ('running_elbo_delta', C.c_double), ('obj', C.c_double),
                ('elbo_delta', C.c_double), ('atol', C.c_double), ('rtol', C.c_double),
                ('diff', C.c_double * 10), ('has_running', C.c_int32), ('trials', C.c_int32),
                ('evals', C.c_int32), ('do_diff', C.c_int32), ('speculate', C.c_int32),
                ('rejects', C.c_int32)]


SIGNATURES.update({
    'vb_nccl_unique_id': (C.c_int, [C.c_char_p]),
    'vb_comm_init': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    'vb_xr_create': (C.c_int, [C.c_void_p, C.c_char_p]),
    'vb_xr_open': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    'vb_fit_set_constants': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    'vb_fit_timing': (C.c_int, [C.c_void_p, c_dp]),
    'vb_fit_iteration': (C.c_int, [C.c_void_p, C.POINTER(StepIO), C.c_void_p, C.c_void_p, C.c_void_p]),
})


class VilmaB200Error(RuntimeError):
    pass


def load():
    """Load the CUDA library.  Raises if it is missing: there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VilmaB200Error(
            'libvilma_b200.so not found at %s. Build it with '
            '`python -c "import __graft_entry__ as g; g.build()"` (needs nvcc). '
            'vilma_b200 has no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.vb_abi_version() != 1:
        raise VilmaB200Error('libvilma_b200.so ABI version mismatch')
    _lib = lib
    # VILMA_B200_OPTIONS="name=value,name=value": process-wide vb_set_option calls at load (kernel
    # selection for experiments and for the worker processes of the multi-GPU tests)
    for item in filter(None, os.environ.get('VILMA_B200_OPTIONS', '').split(',')):
        name, _, value = item.partition('=')
        if lib.vb_set_option(name.strip().encode(), int(value)) != 0:
            raise VilmaB200Error(lib.vb_last_error().decode())
    return lib


def check(rc):
    if rc != 0:
        raise VilmaB200Error(load().vb_last_error().decode())


def np_ptr(a):
    """Raw pointer of a C-contiguous numpy array."""
    assert a.flags['C_CONTIGUOUS']
    return C.c_void_p(a.ctypes.data)
