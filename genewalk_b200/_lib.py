"""ctypes binding of libgenewalk_b200.so (the C ABI declared in include/genewalk_b200.h).

There is NO CPU fallback: if the library is missing it is built with nvcc, and if that fails, or
no CUDA device is present when a compute entry point is called, the call raises.
"""
import ctypes

from . import _build

GW_OK = 0
GW_ERR_INVALID = -1
GW_ERR_CUDA = -2
GW_ERR_UNSUPPORTED = -3
GW_ERR_NO_DEVICE = -4

_vp = ctypes.c_void_p
_i32 = ctypes.c_int32
_i64 = ctypes.c_int64
_u64 = ctypes.c_uint64
_f64 = ctypes.c_double

# name -> argtypes; every entry returns int.  Kept in one table so tests can check that the .so
# exports exactly what include/genewalk_b200.h declares.
SIGNATURES = {
    'gw_version': [],
    'gw_device_count': [],
    'gw_csr_from_edges': [_i32, _i64, _vp, _vp, _vp, _vp, ctypes.POINTER(_i64), _vp],
    'gw_config_model': [_i32, _vp, _i64, _u64, _vp, _vp, _vp],
    'gw_walk_uniform': [_i32, _i64, _vp, _vp, _i32, _i32, _u64, _u64, _i64, _i64, _vp, _vp],
    'gw_count_tokens': [_vp, _i64, _i32, _vp, _vp],
    'gw_vocab_rank': [_i32, _vp, _vp, _vp, _vp, _vp],
    'gw_alias_build': [_i32, _vp, _f64, _vp, _vp],
    'gw_sgns_init': [_i32, _i32, _vp, _vp, _vp, _vp, _vp],
    'gw_walk_count': [_i32, _i64, _vp, _vp, _i32, _i32, _u64, _vp, _vp],
    'gw_sgns_plan': [_i32, _vp, _i32, _i64, _i64, _i32, _i64, _vp, _i64, _vp, _vp],
    'gw_sgns_train': [_i32, _i32, _vp, _vp, _vp, _i32, _u64, _i64, _i32, _i32, _i32, _i32, _i32, _i32,
                      _f64, _f64, _i64, _vp, _vp, _vp, _vp, _vp, _u64, _i64, _i32, _vp],
    'gw_sgns_auto_shape': [_i32, _i64, _i64, _i64, ctypes.POINTER(_i64), ctypes.POINTER(_i64),
                           ctypes.POINTER(_i32)],
    'gw_cosine_pairs': [_i32, _vp, _i64, _vp, _vp, _vp, _vp],
    'gw_null_similarities': [_i32, _vp, _vp, _i32, _vp, _vp, ctypes.POINTER(_i64), _vp],
    'gw_sort_f32': [_vp, _i64, _vp],
    'gw_psim': [_vp, _i64, _vp, _i64, _vp, _vp, _vp],
    'gw_bh_segmented': [_vp, _i64, _vp, _i64, _vp, _vp],
    'gw_log_stats': [_vp, _i64, _i32, _vp, _vp],
    'gw_mean_sem_f32': [_vp, _i64, _i32, _vp, _vp, _vp],
}

# functions that do not return a gw_status
INT64_RETURNS = {'gw_sgns_plan_capacity': [_i32, _i64, _i64, _i32]}
# plain int results that are not a gw_status
INT_RETURNS = {'gw_sgns_resident_trainings': [_i32, _i64]}
ABI_VERSION = 200   # GW_VERSION of include/genewalk_b200.h this binding was written against

_lib = None


class GeneWalkCudaError(RuntimeError):
    pass


def library_path():
    return _build.SO_PATH


def load():
    """Load (building first if needed) the CUDA library.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    # rebuild when the library is missing or older than its sources (a stale .so under the same
    # symbol names would marshal wrong arguments into the kernels)
    if _build.needs_build():
        _build.build()
    lib = ctypes.CDLL(_build.SO_PATH)
    lib.gw_version.restype = ctypes.c_int
    if lib.gw_version() != ABI_VERSION:
        raise GeneWalkCudaError('%s has ABI version %d, this package needs %d: rebuild it '
                                '(python -m genewalk_b200._build --force)'
                                % (_build.SO_PATH, lib.gw_version(), ABI_VERSION))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale: fail loudly
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    for name, argtypes in INT64_RETURNS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int64
    for name, argtypes in INT_RETURNS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    lib.gw_last_error.argtypes = []
    lib.gw_last_error.restype = ctypes.c_char_p
    lib.gw_launch_count.argtypes = []
    lib.gw_launch_count.restype = ctypes.c_longlong
    _lib = lib
    return lib


def launch_count():
    """Kernels launched by libgenewalk_b200 since load."""
    return int(load().gw_launch_count())


def last_error():
    return load().gw_last_error().decode('utf-8', 'replace')


def check(rc):
    if rc == GW_OK:
        return
    msg = last_error()
    if rc == GW_ERR_INVALID:
        raise ValueError(msg)
    if rc == GW_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise GeneWalkCudaError('libgenewalk_b200 error %d: %s' % (rc, msg))


def call(name, *args):
    check(getattr(load(), name)(*args))


def require_cuda():
    """The product has no CPU path: raise unless a CUDA device is usable."""
    import torch
    if not torch.cuda.is_available():
        raise GeneWalkCudaError(
            'genewalk_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    n = load().gw_device_count()
    if n <= 0:
        raise GeneWalkCudaError('libgenewalk_b200: no CUDA device (%s)' % last_error())


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)
