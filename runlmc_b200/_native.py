"""ctypes binding of liblmc_b200.so (include/lmc_b200.h).

The CUDA library is the product: there is no CPU fallback.  Importing this
module without the built library, or calling into it without a CUDA device,
raises."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_lib', 'liblmc_b200.so')


class NativeError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            'CUDA extension {} is missing: build it with '
            '`python -c "import __graft_entry__ as g; g.build()"` '
            '(there is no CPU fallback)'.format(LIB_PATH))
    return ctypes.CDLL(LIB_PATH)


lib = _load()

_p = ctypes.c_void_p
_i = ctypes.c_int
_l = ctypes.c_long
_d = ctypes.c_double

SIGNATURES = {
    'lmc_version': (_i, []),
    'lmc_last_error': (ctypes.c_char_p, []),
    'lmc_launch_count': (ctypes.c_ulonglong, []),
    'lmc_fp64_peak': (_i, [_p]),
    'lmc_profile_ncat': (_i, []),
    'lmc_profile_name': (ctypes.c_char_p, [_i]),
    'lmc_profile_begin': (_i, []),
    'lmc_profile_end': (_i, [_p, _p]),
    'lmc_op_create': (_i, [ctypes.POINTER(_p), _i, _i, _p, _p, _p, _p, _p]),
    'lmc_op_create_dev': (_i, [ctypes.POINTER(_p), _i, _i, _p, _p, _p, _p, _p, _p]),
    'lmc_op_destroy': (_i, [_p]),
    'lmc_op_set_kernels': (_i, [_p, _i, _p, _p, _p, _p]),
    'lmc_op_num_kernel_tops': (_i, [_p, _i]),
    'lmc_op_kernel_tops': (_i, [_p, _i, _p]),
    'lmc_grad_grams_kernels': (_i, [_p, _p, _p, _p, _l, _i, _p, _p, _p, _p, _p]),
    'lmc_op_set_params': (_i, [_p, _i, _p, _p, _p]),
    'lmc_op_set_coreg_factors': (_i, [_p, _p, _p, _p]),
    'lmc_op_n': (_l, [_p]),
    'lmc_op_grid_cells': (_l, [_p]),
    'lmc_op_embed_bins': (_l, [_p]),
    'lmc_op_max_tile_points': (_i, [_p]),
    'lmc_op_perm': (_i, [_p, _p]),
    'lmc_mvm': (_i, [_p, _p, _l, _i, _p, _p]),
    'lmc_mvm_host': (_i, [_p, _p, _l, _i, _p]),
    'lmc_mvm_sorted': (_i, [_p, _p, _l, _i, _p, _p]),
    'lmc_mvm_rows': (_i, [_p, _p, _l, _i, _p, _l, _p]),
    'lmc_mvm_rows_host': (_i, [_p, _p, _l, _i, _p, _l]),
    'lmc_to_grid': (_i, [_p, _p, _l, _i, _p, _p]),
    'lmc_grid_mvm': (_i, [_p, _p, _i, _p, _p]),
    'lmc_from_grid': (_i, [_p, _p, _i, _p, _l, _p]),
    'lmc_minres': (_i, [_p, _p, _l, _i, _p, _d, _i, _i, _p, _p, _p, _p]),
    'lmc_minres_host': (_i, [_p, _p, _l, _i, _p, _d, _i, _i, _p, _p, _p]),
    'lmc_minres_generic': (_i, [_p, _p, _l, _p, _p, _p, _l, _i, _p, _d, _i, _i, _p, _p, _p, _p]),
    'lmc_minres_pre': (_i, [_p, _p, _l, _i, _p, _d, _i, _i, _i, _p, _p, _p, _p]),
    'lmc_minres_lanczos': (_i, [_p, _p, _l, _i, _p, _d, _i, _i, _p, _p, _p, _i, _p, _p, _p]),
    'lmc_op_diagonal': (_i, [_p, _p]),
    'lmc_minres_generic_pre': (_i, [_p, _p, _p, _l, _p, _p, _p, _l, _i, _p, _d, _i, _i, _p, _p, _p, _p]),
    'lmc_cg': (_i, [_p, _p, _l, _i, _p, _d, _i, _i, _p, _p, _p, _p]),
    'lmc_cg_host': (_i, [_p, _p, _l, _i, _p, _d, _i, _i, _p, _p, _p]),
    'lmc_cg_generic': (_i, [_p, _p, _l, _p, _p, _p, _l, _i, _p, _d, _i, _i, _p, _p, _p, _p]),
    'lmc_block_dot': (_i, [_p, _l, _p, _l, _l, _i, _p, _p]),
    'lmc_grad_grams': (_i, [_p, _p, _p, _p, _l, _i, _i, _p, _p, _p, _p, _p, _p]),
    'lmc_bttb_create': (_i, [ctypes.POINTER(_p), _i, _p, _p]),
    'lmc_bttb_destroy': (_i, [_p]),
    'lmc_bttb_apply': (_i, [_p, _p, _i, _p, _p]),
    'lmc_dense_apply': (_i, [_p, _i, _i, _p, _i, _i, _p, _p]),
    'lmc_transpose': (_i, [_p, _i, _i, _i, _p, _p]),
    'lmc_csr_apply': (_i, [_i, _p, _p, _p, _p, _l, _i, _p, _l, _p]),
    'lmc_axpby': (_i, [_l, _d, _p, _d, _p, _p]),
    'lmc_diag_apply': (_i, [_p, _l, _p, _i, _p, _p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc):
    if rc != 0:
        msg = lib.lmc_last_error().decode('utf-8', 'replace')
        if rc == 1:
            raise ValueError(msg)
        raise NativeError(msg)


def host_ptr(a):
    """Pointer to a C-contiguous numpy array (kept alive by the caller)."""
    assert a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(ctypes.c_void_p)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise NativeError('no CUDA device visible: the runlmc_b200 hot path '
                          'runs on the GPU only (there is no CPU fallback)')
    return torch


def current_stream_ptr():
    torch = require_cuda()
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def profile_begin():
    check(lib.lmc_profile_begin())


def profile_end():
    """-> {family: (total_ms, launches)} since profile_begin()."""
    ncat = lib.lmc_profile_ncat()
    ms = np.zeros(ncat)
    cnt = np.zeros(ncat, dtype=np.int32)
    check(lib.lmc_profile_end(host_ptr(ms), host_ptr(cnt)))
    return {lib.lmc_profile_name(c).decode(): (float(ms[c]), int(cnt[c]))
            for c in range(ncat) if cnt[c]}
