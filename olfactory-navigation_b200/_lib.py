'''
ctypes binding of the C ABI declared in include/olfnav_b200.h.

There is NO fallback: if libolfnav_b200.so is missing the import fails loudly, and every compute
entry point returns an error without a CUDA device.
'''
from __future__ import annotations

import ctypes
import os

import torch  # noqa: F401  (device memory / streams; loads the CUDA runtime first)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libolfnav_b200.so')

if not os.path.exists(LIB_PATH):
    raise ImportError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                      f'or {_HERE}/csrc/build.sh (needs nvcc, sm_100a). There is no CPU fallback.')

lib = ctypes.CDLL(LIB_PATH)

_p = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_f = ctypes.c_float
_d = ctypes.c_double

# name -> (restype, argtypes); mirrors include/olfnav_b200.h one to one
SIGNATURES = {
    'olfnav_abi_version': (_i, []),
    'olfnav_last_error': (ctypes.c_char_p, []),
    'olfnav_padded_width': (_i, [_i]),
    'olfnav_padded_states': (_i64, [_i, _i]),
    'olfnav_padded_width_for': (_i, [_i, _i]),
    'olfnav_kernel_launches': (ctypes.c_uint64, []),
    'olfnav_model_create': (_i, [ctypes.POINTER(_p), _i, _i, _i, _i, _i, _p, _i, _p, _p, _p]),
    'olfnav_model_create_padded': (_i, [ctypes.POINTER(_p), _i, _i, _i, _i, _i, _p, _i, _p, _p, _p, _i]),
    'olfnav_model_destroy': (None, [_p]),
    'olfnav_model_dims': (_i, [_p, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    'olfnav_pack_rows': (_i, [_p, _p, _p, _p, _i64, _i64, _p]),
    'olfnav_unpack_rows': (_i, [_p, _p, _i64, _i64, _p, _p, _p, _p]),
    'olfnav_belief_init': (_i, [_p, _p, _i64, _i64, _p, _p]),
    'olfnav_model_create_dense': (_i, [ctypes.POINTER(_p), _i, _i, _i, _i, _p, _p, _p, _p]),
    'olfnav_belief_step': (_i, [_p, _p, _p, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p]),
    'olfnav_belief_step_evaluate': (_i, [_p, _p, _p, _i64, _p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    'olfnav_can_fuse_policy': (_i, [_p, _p]),
    'olfnav_fused_out_rows': (_i64, [_p, _i64]),
    'olfnav_copy_rows': (_i, [_p, _p, _i64, _i64, _p, _p, _p, _i64, _p]),
    'olfnav_alphas_create': (_i, [ctypes.POINTER(_p), _p, _p, _i64, _p, _i, _p]),
    'olfnav_alphas_destroy': (None, [_p]),
    'olfnav_evaluate': (_i, [_p, _p, _p, _i64, _i64, _p, _p, _p, _p, _i, _p]),
    'olfnav_infotaxis_choose': (_i, [_p, _p, _i64, _i64, _p, _p, _p, _p]),
    'olfnav_infotaxis_last_unsure': (_i, [_p, ctypes.POINTER(ctypes.c_int32), _p]),
    'olfnav_entropies': (_i, [_p, _p, _i64, _i64, _p, _p, _p]),
    'olfnav_backup_gamma': (_i, [_p, _p, _i64, _i, _f, _p, _i64, _p]),
    'olfnav_backup_select': (_i, [_p, _p, _i64, _i64, _p, _p, _i64, _i, _p, _p, _p, _i, _p]),
    'olfnav_backup_assemble': (_i, [_p, _p, _i64, _i, _p, _p, _i64, _p, _i64, _p]),
    'olfnav_backup_select_alphas': (_i, [_p, _p, _i64, _i64, _p, _p, _i64, _i, _f, _p, _p, _p, _p]),
    'olfnav_backup_assemble_alphas': (_i, [_p, _p, _i64, _i, _f, _p, _p, _i64, _p, _i64, _p]),
    'olfnav_vi_sweep': (_i, [_p, _p, _p, _i64, _d, _p, _p]),
    'olfnav_env_create': (_i, [ctypes.POINTER(_p), _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _d, _d, _d]),
    'olfnav_env_create_layered': (_i, [ctypes.POINTER(_p), _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _i, _d, _d, _d]),
    'olfnav_env_destroy': (None, [_p]),
    'olfnav_env_step': (_i, [_p, _i64, _p, _p, _p, _p, _i64, _p, _i, _p, _p, _p, _p]),
}



class SimStepArgs(ctypes.Structure):
    'Mirror of olfnav_sim_step_args (include/olfnav_b200.h).'
    _fields_ = [
        ('b_in', _p), ('b_out', _p), ('ld', _i64),
        ('scale_in', _p), ('scale_out', _p),
        ('pos_in', _p), ('pos_out', _p),
        ('ts_in', _p), ('ts_out', _p),
        ('n', _i64), ('step', _i64),
        ('moves', _p),
        ('thresholds', _p), ('n_thr', ctypes.c_int32),
        ('time_mode', ctypes.c_int32), ('time_loop', ctypes.c_int32), ('eval_path', ctypes.c_int32), ('have_act', ctypes.c_int32),
        ('act', _p), ('obs_id', _p), ('src_rows', _p), ('act_c', _p), ('obs_c', _p),
        ('ok', _p), ('at_end', _p), ('term_c', _p),
        ('hist', _p),
        ('eval_val', _p), ('eval_idx', _p),
        ('rec_pos', _p), ('rec_odor', _p), ('rec_reached', _p), ('rec_term', _p),
        ('counts', _p),
        ('ev_policy_end', _p), ('ev_update_begin', _p), ('ev_update_end', _p),
        ('act_next', _p),
        ('brow_in', _p), ('brow_out', _p), ('belief_rows', _i64), ('inplace', ctypes.c_int32),
        ('act_layer', _p), ('thr_per_layer', ctypes.c_int32), ('hist_layer', _p), ('ev_plan_done', _p),
    ]


SIGNATURES['olfnav_pair_reduce'] = (_i, [_i, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i64, ctypes.c_float, ctypes.c_float, _i64, _p, _p])
SIGNATURES['olfnav_sim_step'] = (_i, [_p, _p, _p, ctypes.POINTER(SimStepArgs), _p])

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)   # AttributeError here = the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM = range(5)
BOUNDARY = {'stop': 0, 'clip': 0, 'wrap': 1, 'wrap_vertical': 2, 'wrap_horizontal': 3}


class OlfnavError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != OK:
        msg = lib.olfnav_last_error().decode('utf-8', 'replace')
        raise OlfnavError(f'olfnav_b200 error {rc}: {msg}')


def ptr(t) -> ctypes.c_void_p | None:
    'Device (or host) pointer of a torch tensor / numpy array, None passes NULL.'
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        assert t.is_contiguous() or t.stride(-1) == 1, 'tensor must be contiguous in its last dimension'
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)   # numpy


def stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise OlfnavError('olfnav_b200 needs a CUDA device (sm_100a); there is no CPU fallback.')
