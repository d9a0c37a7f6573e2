"""ctypes binding of libemsanet_b200.so (the C ABI declared in include/emsanet_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, this raises.
The library is loaded lazily (first op), never at import time, because the reference forks
DataLoader workers and wandb processes (SURVEY.md §8b "threading").
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('EB200_LIB') or os.path.join(HERE, 'lib', 'libemsanet_b200.so')   # EB200_LIB: A/B builds
MAX_TAPS = 9

BIAS, RELU, AUX_ADD, AUX_MASK, STATS, STATS_SUM_ONLY, BN_BWD = 1, 2, 4, 8, 16, 32, 64


class View(C.Structure):
    _fields_ = [('ptr', C.c_void_p), ('n', C.c_int), ('h', C.c_int), ('w', C.c_int), ('c', C.c_int),
                ('sn', C.c_longlong), ('sh', C.c_longlong), ('sw', C.c_longlong)]


class ConvDesc(C.Structure):
    _fields_ = [('inp', View * 2), ('n', C.c_int), ('h', C.c_int), ('w', C.c_int), ('cin', C.c_int),
                ('cout', C.c_int), ('taps', C.c_int), ('tap_view', C.c_int * MAX_TAPS),
                ('tap_dy', C.c_int * MAX_TAPS), ('tap_dx', C.c_int * MAX_TAPS), ('tap_w', C.c_int * MAX_TAPS),
                ('weight_taps', C.c_int), ('weight', C.c_void_p),
                ('cout_pad', C.c_int), ('cin_pad', C.c_int), ('out', C.c_void_p), ('out_sn', C.c_longlong),
                ('out_sh', C.c_longlong), ('out_sw', C.c_longlong), ('aux', C.c_void_p), ('aux_sn', C.c_longlong),
                ('aux_sh', C.c_longlong), ('aux_sw', C.c_longlong), ('bias', C.c_void_p), ('stats', C.c_void_p),
                ('flags', C.c_uint32), ('bn_scale', C.c_void_p), ('bn_shift', C.c_void_p)]


class WgradDesc(C.Structure):
    _fields_ = [('dy', View), ('x', View * 2), ('taps', C.c_int), ('tap_view', C.c_int * MAX_TAPS),
                ('tap_dy', C.c_int * MAX_TAPS), ('tap_dx', C.c_int * MAX_TAPS), ('dw', C.c_void_p),
                ('dw_sco', C.c_longlong), ('dw_sci', C.c_longlong), ('dw_st', C.c_longlong),
                ('ws', C.c_void_p), ('ws_floats', C.c_longlong)]


class PackEntry(C.Structure):
    _fields_ = [('w', C.c_void_p), ('fwd', C.c_void_p), ('bwd', C.c_void_p), ('cout', C.c_int), ('cin', C.c_int),
                ('taps', C.c_int), ('fwd_rows', C.c_int), ('fwd_cols', C.c_int), ('bwd_rows', C.c_int),
                ('bwd_cols', C.c_int), ('co_off', C.c_int), ('ci_off', C.c_int), ('pad_', C.c_int)]


class OptimEntry(C.Structure):
    _fields_ = [('p', C.c_void_p), ('g', C.c_void_p), ('m', C.c_void_p), ('v', C.c_void_p), ('numel', C.c_longlong),
                ('pack', C.c_int), ('pad_', C.c_int)]


class OptimHyper(C.Structure):
    _fields_ = [('kind', C.c_int), ('lr', C.c_float), ('momentum', C.c_float), ('beta2', C.c_float), ('eps', C.c_float),
                ('weight_decay', C.c_float), ('bias_correction1', C.c_float), ('bias_correction2_sqrt', C.c_float),
                ('nesterov', C.c_int), ('step', C.c_int), ('flags', C.c_int), ('one_minus_beta1', C.c_float),
                ('one_minus_beta2', C.c_float), ('step_size', C.c_float), ('decay', C.c_float), ('pad_', C.c_int)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> argtypes (all return int status); must list every symbol of include/emsanet_b200.h
SIGNATURES = {
    'eb200_conv2d': [C.POINTER(ConvDesc), _P],
    'eb200_conv2d_wgrad': [C.POINTER(WgradDesc), _P],
    'eb200_conv2d_pair': [C.POINTER(ConvDesc), C.POINTER(ConvDesc), _P],
    'eb200_conv2d_wgrad_pair': [C.POINTER(WgradDesc), C.POINTER(WgradDesc), _P],
    'eb200_pack_conv_weight': [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P],
    'eb200_pack_conv_weights_batched': [_P, _P, _P, _I, _P],
    'eb200_bn_finalize': [_P, _L, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _I, _P],
    'eb200_bn_apply': [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    'eb200_bn_bwd_reduce': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    'eb200_bn_bwd_apply': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    'eb200_bn_apply_train': [_P, _P, _P, _L, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I,
                             _I, _P],
    'eb200_bn_bwd_reduce_rep': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    'eb200_bn_bwd_fused': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    'eb200_bn_bwd_apply_raw': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    'eb200_bn_bwd_param': [_P, _P, _P, _I, _P],
    'eb200_colsum': [_P, _P, _L, _I, _I, _I, _P],
    'eb200_im2col_stem': [_P, _P, _I, _I, _I, _I, _I, _P],
    'eb200_maxpool_fwd': [_P, _P, _P, _I, _I, _I, _I, _P],
    'eb200_maxpool_bwd': [_P, _P, _P, _I, _I, _I, _I, _P],
    'eb200_gap': [_P, _P, _I, _I, _I, _P],
    'eb200_se_mlp_fwd': [_P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    'eb200_se_mlp_bwd': [_P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    'eb200_se_fuse_fwd': [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    'eb200_se_fuse_bwd_reduce': [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    'eb200_se_fuse_bwd_apply': [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    'eb200_adaptive_pool_fwd': [_P, _P, _I, _I, _I, _I, _I, _P],
    'eb200_adaptive_pool_bwd': [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    'eb200_bilinear_fwd': [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    'eb200_bilinear_bwd': [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    'eb200_upsample_dw_fwd': [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    'eb200_upsample_dw_bwd_input': [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    'eb200_upsample_dw_bwd_weight': [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    'eb200_upsample_dw_fwd_nchw': [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    'eb200_upsample_dw_bwd_nchw': [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    'eb200_nhwc_to_nchw': [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    'eb200_nchw_to_nhwc_grad': [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    'eb200_linear_fwd': [_P, _P, _P, _P, _I, _I, _I, _P],
    'eb200_linear_bwd': [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    'eb200_add_inplace': [_P, _P, _L, _P],
    'eb200_copy_channels': [_P, _P, _L, _I, _I, _I, _I, _I, _I, _P],
    # inference post-processing (csrc/postproc.cu)
    'eb200_pp_softmax_argmax': [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    'eb200_pp_instance_centers': [_P, _I, _I, _I, _F, _I, _I, _P, _P, _L, _P, _P, _P, _P, _P],
    'eb200_pp_instance_assign': [_P, _P, _P, _P, _I, _I, _I, _F, _F, _F, _P, _I, _P, _P, _P, _P],
    'eb200_pp_panoptic_merge': [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    'eb200_pp_nearest_resize': [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    'eb200_pp_instance_orientation': [_P, _P, _I, _P, _I, _I, _I, _I, _P, _P],
    # fused semantic cross-entropy (csrc/loss.cu)
    'eb200_ce_loss_fwd': [_P, _P, _I, _P, _F, _I, _I, _I, _I, _P, _P, _P],
    'eb200_ce_loss_bwd': [_P, _P, _I, _P, _F, _P, _I, _I, _I, _I, _P, _P],
    'eb200_masked_loss_fwd': [_I, _P, _P, _P, _I, _I, _L, _L, _L, _L, _F, _P, _P, _P],
    'eb200_masked_loss_bwd': [_I, _P, _P, _P, _I, _I, _L, _L, _L, _L, _F, _P, _P, _P],
    'eb200_memset_zero': [_P, _L, _P],
    # GPU-side input normalisation (csrc/preproc.cu)
    'eb200_normalize_rgb': [_P, _P, _I, _I, _I, C.POINTER(C.c_float), C.POINTER(C.c_float), _P],
    'eb200_normalize_depth': [_P, _I, _P, _I, _I, _I, _F, _F, _I, _F, _P],
    # fused optimizer step + weight re-layout (csrc/optim.cu)
    'eb200_optim_step': [_P, _P, _P, _P, _I, C.POINTER(OptimHyper), _P],
}
OTHER_SYMBOLS = ('eb200_last_error', 'eb200_version', 'eb200_launch_count', 'eb200_pp_centers_ws_bytes',
                 'eb200_optim_chunk')

_lib = None


class EB200Error(RuntimeError):
    pass


def load(path: str = LIB_PATH):
    """Load the shared library and bind every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise EB200Error(f'{path} not found: build it with `python -m emsanet_b200.build` '
                         '(there is no CPU / PyTorch fallback for the EMSANet hot path)')
    lib = C.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.eb200_last_error.restype = C.c_char_p
    lib.eb200_version.restype = C.c_int
    lib.eb200_launch_count.restype = C.c_longlong
    lib.eb200_pp_centers_ws_bytes.restype = C.c_longlong
    lib.eb200_pp_centers_ws_bytes.argtypes = [_I, _I, _I, _I]
    _lib = lib
    return lib


def call(name: str, *args):
    lib = _lib or load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise EB200Error(f'{name}: {lib.eb200_last_error().decode()}')


def launch_count() -> int:
    return int((_lib or load()).eb200_launch_count())
