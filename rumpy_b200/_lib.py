"""ctypes binding of librumpy_b200.so (C ABI declared in include/rumpy_b200.h).

The library is the product: if it is missing, or no sm_100 device is present, every compute call raises.
There is deliberately no CPU / eager-PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'librumpy_b200.so')

_c = ctypes
_vp, _fp, _i, _u, _f = _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_uint, _c.c_float

# name -> argtypes (restype is int unless noted); must list every symbol of include/rumpy_b200.h
SIGNATURES = {
    'rumpy_version': [],
    'rumpy_last_error': [],
    'rumpy_device_check': [],
    'rumpy_pack_conv3x3': [_fp, _vp, _i, _i, _i, _i, _i, _vp],
    'rumpy_pack_bias': [_fp, _fp, _i, _i, _i, _vp],
    'rumpy_conv3x3': [_vp, _vp, _fp, _fp, _vp, _vp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _u, _f, _vp],
    'rumpy_conv3x3_wgrad_workspace': [_i, _i, _i, _i, _i],
    'rumpy_conv3x3_wgrad': [_vp, _vp, _fp, _vp, _i, _i, _i, _i, _i, _i, _f, _i, _vp],
    'rumpy_conv3x3_tail': [_vp, _vp, _fp, _fp, _i, _i, _i, _i, _i, _vp],
    'rumpy_head_conv': [_fp, _fp, _fp, _fp, _vp, _i, _i, _i, _i, _i, _vp],
    'rumpy_ca_apply': [_fp, _vp, _i, _fp, _fp, _fp, _fp, _fp, _fp, _vp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _vp],
    'rumpy_nchw_to_nhwc': [_fp, _fp, _vp, _i, _i, _i, _i, _vp],
    'rumpy_nhwc_to_nchw': [_vp, _i, _fp, _i, _i, _i, _i, _vp],
    'rumpy_pool_sum': [_fp, _fp, _i, _i, _i, _i, _vp],
    'rumpy_pool_rows': [_i, _i, _i],
    'rumpy_net_create': [_c.POINTER(_c.c_void_p), _i, _i, _i, _i, _i, _i, _f, _i, _i, _i],
    'rumpy_net_create_q': [_c.POINTER(_c.c_void_p), _i, _i, _i, _i, _i, _i, _f, _i, _i, _i, _i, _c.c_char_p, _i, _i],
    'rumpy_net_set_metadata': [_vp, _fp, _i, _i],
    'rumpy_psnr_y_workspace': [_i],
    'rumpy_psnr_y': [_fp, _fp, _fp, _vp, _i, _i, _i, _f, _vp],
    'rumpy_quantize_u8': [_fp, _vp, _i, _i, _i, _i, _vp],
    'rumpy_patch_batch': [_vp, _vp, _vp, _fp, _fp, _i, _i, _i, _vp],
    'rumpy_bicubic_workspace': [_i, _i, _i],
    'rumpy_bicubic_upsample': [_fp, _fp, _vp, _i, _i, _i, _i, _i, _vp],
    'rumpy_lanczos_workspace': [_i, _i, _i],
    'rumpy_lanczos_upsample': [_fp, _fp, _vp, _i, _i, _i, _i, _i, _vp],
    'rumpy_net_backward_chunks': [_vp, _vp, _i],
    'rumpy_net_set_backward_events': [_vp, _vp, _i],
    'rumpy_net_destroy': [_vp],
    'rumpy_net_num_params': [_vp],
    'rumpy_net_num_launches': [_vp],
    'rumpy_net_num_launches_backward': [_vp],
    'rumpy_net_trunk_mode': [_vp],
    'rumpy_net_set_option': [_vp, _c.c_char_p, _c.c_longlong],
    'rumpy_net_get_option': [_vp, _c.c_char_p],
    'rumpy_net_set_trunk_events': [_vp, _vp, _vp],
    'rumpy_net_set_timeline': [_vp, _vp, _i],
    'rumpy_net_packed_bytes': [_vp, _i],
    'rumpy_net_workspace_bytes': [_vp, _i, _i, _i, _i],
    'rumpy_net_pack': [_vp, _vp, _vp, _i, _vp],
    'rumpy_net_forward': [_vp, _vp, _vp, _fp, _fp, _vp, _i, _i, _i, _i, _vp],
    'rumpy_net_backward': [_vp, _vp, _vp, _fp, _fp, _vp, _vp, _i, _i, _i, _vp],
    'rumpy_l1_workspace_floats': [],
    'rumpy_l1_loss_grad': [_fp, _fp, _fp, _fp, _fp, _c.c_longlong, _f, _vp],
    'rumpy_grad_clip_coef': [_fp, _c.c_longlong, _f, _fp, _fp, _vp],
    'rumpy_adam_step': [_fp, _fp, _fp, _fp, _c.c_longlong, _f, _f, _f, _f, _i, _fp, _f, _vp],
}

_LONGLONG = {'rumpy_net_packed_bytes', 'rumpy_net_workspace_bytes', 'rumpy_conv3x3_wgrad_workspace',
             'rumpy_l1_workspace_floats', 'rumpy_psnr_y_workspace', 'rumpy_bicubic_workspace', 'rumpy_lanczos_workspace', 'rumpy_net_get_option'}

_lib = None

# developer switches, read when an engine is created and applied to ITS handle (rumpy_net_set_option): env var ->
# (option, value when the variable holds that string)
ENV_OPTIONS = {'RUMPY_B200_PDL': ('pdl', {'0': 0}),                'RUMPY_B200_TRUNK': ('trunk', {'0': 0}), 'RUMPY_B200_BAND': ('band', {'1': 1}),
               'RUMPY_B200_CLUSTER': ('cluster', {'0': 0}), 'RUMPY_B200_CLUSTER_GROUPS': ('cluster_groups', {'2': 2, '4': 4}),
               'RUMPY_B200_FUSED_CA': ('fused_ca', {'1': 1}), 'RUMPY_B200_CLUSTER_SPLIT': ('cluster_split', {'0': 0}),
               'RUMPY_B200_CONV_DBG': ('conv_dbg', {str(i): i for i in range(1, 128)})}


def env_options():
    out = {}
    for var, (name, table) in ENV_OPTIONS.items():
        v = os.environ.get(var)
        if v in table:
            out[name] = table[v]
    return out


class RumpyB200Error(RuntimeError):
    pass


def load():
    """Loads the shared library (once).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RumpyB200Error(
            f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(rumpy_b200 has no CPU fallback)')
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = (ctypes.c_char_p if name == 'rumpy_last_error' else
                      ctypes.c_longlong if name in _LONGLONG else ctypes.c_int)
    _lib = lib
    return lib


def check(code: int, what: str = ''):
    if code != 0:
        msg = load().rumpy_last_error()
        raise RumpyB200Error(f'{what} failed ({code}): {msg.decode() if msg else "?"}')


def call(name: str, *args):
    lib = load()
    check(getattr(lib, name)(*args), name)
