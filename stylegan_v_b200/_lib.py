"""ctypes binding of libsgv_b200.so (the C ABI declared in include/sgv_b200.h).

There is no CPU implementation behind this module: if the shared library is missing, or was built
without an entry point the header declares, importing a kernel raises immediately.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libsgv_b200.so')

c_int, c_i64, c_f32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

SGV_F32, SGV_F16, SGV_F64 = 0, 1, 2
ABI_VERSION = 2


class UpfirdnParams(ctypes.Structure):
    """struct sgv_upfirdn2d_params"""
    _fields_ = [
        ('x', c_vp), ('f', c_vp), ('y', c_vp), ('dtype', c_int),
        ('up_x', c_int), ('up_y', c_int), ('down_x', c_int), ('down_y', c_int),
        ('pad_x0', c_int), ('pad_x1', c_int), ('pad_y0', c_int), ('pad_y1', c_int),
        ('flip', c_int), ('gain', c_f32),
        ('in_w', c_int), ('in_h', c_int), ('in_c', c_int), ('in_n', c_int),
        ('in_stride_x', c_i64), ('in_stride_y', c_i64), ('in_stride_c', c_i64), ('in_stride_n', c_i64),
        ('f_w', c_int), ('f_h', c_int), ('f_stride_x', c_i64), ('f_stride_y', c_i64),
        ('out_w', c_int), ('out_h', c_int),
        ('out_stride_x', c_i64), ('out_stride_y', c_i64), ('out_stride_c', c_i64), ('out_stride_n', c_i64),
        ('epi_scale', c_vp), ('epi_bias', c_vp), ('epi_act', c_int),
        ('epi_alpha', c_f32), ('epi_gain', c_f32), ('epi_clamp', c_f32), ('epi_round_tf32', c_int),
        ('epi_noise', c_vp), ('epi_noise_stride_n', c_i64), ('epi_noise_stride_y', c_i64), ('epi_noise_stride_x', c_i64),
    ]


class BiasActParams(ctypes.Structure):
    """struct sgv_bias_act_params"""
    _fields_ = [
        ('x', c_vp), ('b', c_vp), ('xref', c_vp), ('yref', c_vp), ('dy', c_vp), ('y', c_vp),
        ('dtype', c_int), ('grad', c_int), ('act', c_int),
        ('alpha', c_f32), ('gain', c_f32), ('clamp', c_f32),
        ('size_x', c_int), ('size_b', c_int), ('step_b', c_int),
        ('db_accum', c_vp),
    ]


CONV_MAX_TAPS = 16


class ConvParams(ctypes.Structure):
    """struct sgv_conv_params (include/sgv_b200_conv.h)"""
    _fields_ = [
        ('x', c_vp), ('wp', c_vp), ('y', c_vp),
        ('n', c_int), ('h', c_int), ('w', c_int), ('cin', c_int), ('cout', c_int),
        ('out_h', c_int), ('out_w', c_int),
        ('out_stride_n', c_i64), ('out_stride_y', c_i64), ('out_stride_x', c_i64),
        ('in_stride', c_int), ('ntaps', c_int),
        ('tap_dy', c_int * CONV_MAX_TAPS), ('tap_dx', c_int * CONV_MAX_TAPS),
        ('a_scale', c_vp), ('o_scale', c_vp), ('bias', c_vp),
        ('act', c_int), ('alpha', c_f32), ('gain', c_f32), ('clamp', c_f32),
        ('in_stride_n', c_i64), ('in_stride_y', c_i64), ('in_stride_x', c_i64), ('accumulate', c_int),
        ('red_x', c_vp), ('red_out', c_vp), ('a_ready', c_int),
        ('wp_lo', c_vp), ('noise', c_vp), ('noise_stride_n', c_i64), ('noise_stride_y', c_i64), ('noise_stride_x', c_i64),
    ]


class ConvVariant(ctypes.Structure):
    """struct sgv_conv_variant"""
    _fields_ = [('kernel', c_int), ('bn', c_int), ('mh', c_int), ('cluster', c_int), ('cta_pair', c_int), ('x3', c_int)]


class WgradVariant(ctypes.Structure):
    """struct sgv_wgrad_variant"""
    _fields_ = [('kernel', c_int), ('nt', c_int), ('stages', c_int), ('ksplit', c_int), ('passes', c_int)]


class WgradParams(ctypes.Structure):
    """struct sgv_wgrad_params (include/sgv_b200_conv.h)"""
    _fields_ = [
        ('g', c_vp), ('x', c_vp), ('dw', c_vp),
        ('n', c_int), ('gh', c_int), ('gw', c_int), ('xh', c_int), ('xw', c_int), ('cin', c_int), ('cout', c_int),
        ('out_h', c_int), ('out_w', c_int), ('g_stride', c_int), ('x_stride', c_int), ('ntaps', c_int),
        ('g_dy', c_int * CONV_MAX_TAPS), ('g_dx', c_int * CONV_MAX_TAPS), ('x_dy', c_int * CONV_MAX_TAPS), ('x_dx', c_int * CONV_MAX_TAPS),
        ('g_scale', c_vp), ('x_scale', c_vp),
        ('x_stride_n', c_i64), ('x_stride_y', c_i64), ('x_stride_x', c_i64),
        ('use_dw_slot', c_int), ('dw_slot', c_int * CONV_MAX_TAPS),
        ('g_ready', c_int), ('x_ready', c_int), ('precision', c_int),
    ]


class AdamParams(ctypes.Structure):
    """struct sgv_adam_params (include/sgv_b200_aux.h)"""
    _fields_ = [
        ('param', c_vp), ('grad', c_vp), ('exp_avg', c_vp), ('exp_avg_sq', c_vp), ('param_ema', c_vp),
        ('numel', c_i64),
        ('lr', c_f32), ('beta1', c_f32), ('beta2', c_f32), ('eps', c_f32),
        ('ema_beta', c_f32), ('grad_scale', c_f32), ('grad_clamp', c_f32),
        ('step', c_int), ('step_count', c_vp), ('advance_step', c_int), ('zero_grad', c_int),
    ]


class DenseParams(ctypes.Structure):
    """struct sgv_dense_params (include/sgv_b200_aux.h)"""
    _fields_ = [
        ('a', c_vp), ('a_row_off', c_vp), ('lda', c_i64), ('w', c_vp), ('bias', c_vp), ('y', c_vp), ('ldy', c_i64),
        ('m', c_int), ('n', c_int), ('k', c_int), ('w_gain', c_f32), ('b_gain', c_f32), ('act', c_int), ('alpha', c_f32), ('gain', c_f32),
        ('groups', c_int), ('group_col', c_vp), ('group_off', c_vp),
        ('dy', c_vp), ('lddy', c_i64), ('da', c_vp), ('ldda', c_i64), ('dw', c_vp), ('db', c_vp), ('accumulate', c_int),
    ]


STRUCTS = {'sgv_dense_params': DenseParams, 'sgv_upfirdn2d_params': UpfirdnParams, 'sgv_bias_act_params': BiasActParams, 'sgv_conv_params': ConvParams,
           'sgv_wgrad_params': WgradParams, 'sgv_adam_params': AdamParams, 'sgv_conv_variant': ConvVariant, 'sgv_wgrad_variant': WgradVariant}

# every symbol include/sgv_b200*.h declares: (name, restype, argtypes)
SYMBOLS = [
    ('sgv_abi_version', c_int, []),
    ('sgv_last_error', ctypes.c_char_p, []),
    ('sgv_device_check', c_int, []),
    ('sgv_kernel_launch_count', c_i64, []),
    ('sgv_upfirdn2d_out_size', c_int, [c_int] * 6),
    ('sgv_upfirdn2d', c_int, [ctypes.POINTER(UpfirdnParams), c_vp]),
    ('sgv_bias_act', c_int, [ctypes.POINTER(BiasActParams), c_vp]),
    ('sgv_conv_prep_weights', c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_int,
                                      ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_vp, c_vp]),
    ('sgv_conv_prep_weights_pair', c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_vp,
                                           c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_vp, c_vp]),
    ('sgv_conv_prep_weights_ex', c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_int,
                                         ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_f32, c_vp, c_vp, c_vp]),
    ('sgv_conv_prep_weights_pair_x3', c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_vp, c_vp,
                                              c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_vp, c_vp, c_vp]),
    ('sgv_conv2d_tf32', c_int, [ctypes.POINTER(ConvParams), c_vp]),
    ('sgv_conv2d_tf32_variant', c_int, [ctypes.POINTER(ConvParams), ctypes.POINTER(ConvVariant)]),
    ('sgv_conv2d_wgrad_tf32', c_int, [ctypes.POINTER(WgradParams), c_vp]),
    ('sgv_conv2d_wgrad_tf32_variant', c_int, [ctypes.POINTER(WgradParams), ctypes.POINTER(WgradVariant)]),
    ('sgv_modconv_act_bwd', c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f32, c_f32, c_vp]),
    ('sgv_modconv_act_bwd_rgb', c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f32, c_f32, c_vp]),
    ('sgv_modconv_act_bwd_ex', c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f32, c_f32, c_vp]),
    ('sgv_modconv_scale_reduce', c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp]),
    ('sgv_torgb_fwd', c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp]),
    ('sgv_torgb_bwd', c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp]),
    ('sgv_torgb_wmod_fwd', c_int, [c_vp, c_vp, c_i64, c_vp, c_int, c_int, c_int, c_f32, c_vp]),
    ('sgv_torgb_wmod_bwd', c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_int, c_int, c_int, c_f32, c_vp]),
    ('sgv_time_encoder_fwd', c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_f32, c_vp]),
    ('sgv_time_encoder_bwd', c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_f32, c_vp]),
    ('sgv_adam_ema_step', c_int, [ctypes.POINTER(AdamParams), c_vp]),
    ('sgv_fromrgb_fwd', c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f32, c_int, c_f32, c_f32, c_vp]),
    ('sgv_fromrgb_bwd', c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f32, c_vp]),
    ('sgv_mbstd_fwd', c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    ('sgv_mbstd_bwd', c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    ('sgv_dense_f32_fwd', c_int, [ctypes.POINTER(DenseParams), c_vp]),
    ('sgv_dense_f32_dgrad', c_int, [ctypes.POINTER(DenseParams), c_vp]),
    ('sgv_dense_f32_wgrad', c_int, [ctypes.POINTER(DenseParams), c_vp]),
    ('sgv_demod_fwd', c_int, [c_vp, c_vp, c_i64, c_vp, c_int, c_int, c_int, c_int, c_f32, c_vp]),
    ('sgv_demod_bwd_styles', c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp]),
    ('sgv_demod_bwd_weight', c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
]

_lib = None
_lock = threading.Lock()


class SgvError(RuntimeError):
    pass


def lib():
    """Loads (once) and returns the shared library; raises if it is absent or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise SgvError(f'{LIB_PATH} not found: build it with `python -m stylegan_v_b200.build` '
                           '(there is no CPU or PyTorch fallback for the CUDA ops)')
        L = ctypes.CDLL(LIB_PATH)
        for name, restype, argtypes in SYMBOLS:
            try:
                fn = getattr(L, name)
            except AttributeError as e:
                raise SgvError(f'{LIB_PATH} does not export {name}; rebuild the library') from e
            fn.restype = restype
            fn.argtypes = argtypes
        if L.sgv_abi_version() != ABI_VERSION:
            raise SgvError(f'ABI mismatch: library {L.sgv_abi_version()} vs binding {ABI_VERSION}')
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().sgv_last_error()
        raise SgvError(f'{what} failed (status {rc}): {msg.decode() if msg else "?"}')


def launch_count() -> int:
    return int(lib().sgv_kernel_launch_count())
