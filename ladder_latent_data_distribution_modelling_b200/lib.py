"""ctypes binding of libladder_sm100.so (the C ABI declared in include/ladder_sm100.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libladder_sm100.so')

_lib = None

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)
ptr = C.c_void_p          # device pointers travel as integers
stream_t = C.c_void_p

# name -> (restype, argtypes); every symbol include/ladder_sm100.h declares.
SIGNATURES = {
    'ladder_version': (C.c_int, []),
    'ladder_last_error': (C.c_char_p, []),
    'ladder_launch_count': (C.c_ulonglong, []),
    'ladder_device_check': (C.c_int, [C.c_int]),
    'ladder_crc32c': (C.c_uint, [C.c_uint, C.c_void_p, C.c_size_t]),
    'ladder_mixture_table_stride': (C.c_int, [C.c_int, C.c_int]),
    'ladder_mixture_pack_full': (C.c_int, [c_double_p, c_double_p, c_double_p, C.c_int, C.c_int, c_float_p, c_float_p]),
    'ladder_mixture_pack_diag': (C.c_int, [c_double_p, c_double_p, c_double_p, C.c_int, C.c_int, C.c_int,
                                           c_float_p, c_float_p, c_float_p]),
    'ladder_mixture_workspace_bytes': (C.c_size_t, [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int]),
    'ladder_mixture_logprob': (C.c_int, [ptr, C.c_longlong, C.c_int, ptr, C.c_int, C.c_int, C.c_float, C.c_float,
                                         ptr, ptr, ptr, ptr, ptr, C.c_size_t, stream_t]),
    'ladder_mixture_pack_diag_device': (C.c_int, [ptr, ptr, C.c_int, C.c_int, ptr, ptr, stream_t]),
    'ladder_mixture_logprob_devref': (C.c_int, [ptr, C.c_longlong, C.c_int, ptr, C.c_int, C.c_int, ptr, ptr, ptr, ptr,
                                                C.c_size_t, stream_t]),
    'ladder_mixture_diag_param_grad': (C.c_int, [ptr, C.c_longlong, C.c_int, ptr, ptr, C.c_int, ptr, C.c_float, ptr, ptr,
                                                 stream_t]),
    # conv / dense
    'ladder_conv2d_workspace_bytes': (C.c_size_t, [C.c_int] * 7),
    'ladder_conv2d_fprop': (C.c_int, [ptr, ptr, ptr, ptr] + [C.c_int] * 14 + [ptr, C.c_size_t, stream_t]),
    'ladder_conv2d_dgrad': (C.c_int, [ptr, ptr, ptr, ptr] + [C.c_int] * 15 + [stream_t]),
    'ladder_conv2d_wgrad': (C.c_int, [ptr, ptr, ptr, ptr] + [C.c_int] * 12 + [ptr, C.c_size_t, stream_t]),
    'ladder_tap_sum': (C.c_int, [ptr, C.c_int, ptr, ptr] + [C.c_int] * 11 + [stream_t]),
    'ladder_tap_scatter': (C.c_int, [ptr, ptr] + [C.c_int] * 11 + [stream_t]),
    'ladder_colsum': (C.c_int, [ptr, C.c_longlong, C.c_int, ptr, stream_t]),
    'ladder_conv2d_tc_workspace_bytes': (C.c_size_t, [C.c_int] * 7),
    'ladder_conv2d_fprop_tc': (C.c_int, [ptr, ptr, ptr, ptr] + [C.c_int] * 14 + [ptr, C.c_size_t, stream_t]),
    'ladder_conv2d_dgrad_tc': (C.c_int, [ptr, ptr, ptr, ptr] + [C.c_int] * 15 + [ptr, C.c_size_t, stream_t]),
    'ladder_conv2d_wgrad_tc_supported': (C.c_int, [C.c_int, C.c_int]),
    'ladder_conv2d_wgrad_tc': (C.c_int, [ptr, ptr, ptr] + [C.c_int] * 12 + [ptr, C.c_size_t, stream_t]),
    'ladder_conv2d_tma_supported': (C.c_int, [C.c_int] * 11),
    'ladder_conv2d_tma_workspace_bytes': (C.c_size_t, [C.c_int] * 4),
    'ladder_conv2d_tma_bn': (C.c_int, [C.c_int] * 8),
    'ladder_conv2d_tma_pack_bytes': (C.c_size_t, [C.c_int] * 6),
    'ladder_conv2d_tma_pack': (C.c_int, [ptr, ptr, C.c_size_t] + [C.c_int] * 6 + [stream_t]),
    'ladder_pack_weights_multi': (C.c_int, [ptr, ptr, ptr, C.c_int, C.c_longlong, stream_t]),
    'ladder_conv2d_fprop_tma': (C.c_int, [ptr, ptr, ptr, ptr, C.c_int] + [C.c_int] * 14 + [ptr, C.c_size_t, ptr, C.c_int, stream_t]),
    'ladder_conv2d_tma_set_halo': (C.c_int, [C.c_int, C.c_int]),
    'ladder_conv2d_dgrad_tma': (C.c_int, [ptr, ptr, ptr, C.c_int, ptr, C.c_int] + [C.c_int] * 15 + [ptr, C.c_size_t, stream_t]),
    'ladder_conv2d_wgrad_tma': (C.c_int, [ptr, ptr, ptr] + [C.c_int] * 12 + [stream_t]),
    'ladder_tap_dgrad': (C.c_int, [ptr, ptr, ptr, C.c_int, ptr, C.c_int] + [C.c_int] * 14 + [stream_t]),
    'ladder_im2col64_bf16': (C.c_int, [ptr, ptr] + [C.c_int] * 11 + [stream_t]),
    'ladder_thin_k_supported': (C.c_int, [C.c_int] * 4),
    'ladder_thin_k_fprop': (C.c_int, [ptr, ptr, ptr, ptr, C.c_int] + [C.c_int] * 13 + [stream_t]),
    'ladder_thin_k_wgrad': (C.c_int, [ptr, ptr, ptr, ptr] + [C.c_int] * 12 + [stream_t]),
    'ladder_thin_n_dgrad': (C.c_int, [ptr, ptr, ptr, C.c_longlong, C.c_int, C.c_int, C.c_int, stream_t]),
    'ladder_thin_wgrad_1x1': (C.c_int, [ptr, C.c_int, ptr, ptr, C.c_longlong, C.c_int, C.c_int, stream_t]),
    'ladder_tap_scatter_bf16': (C.c_int, [ptr, ptr] + [C.c_int] * 10 + [stream_t]),
    'ladder_f32_to_bf16': (C.c_int, [ptr, ptr, C.c_longlong, stream_t]),
    'ladder_bf16_to_f32': (C.c_int, [ptr, ptr, C.c_longlong, stream_t]),
    'ladder_colsum_bf16': (C.c_int, [ptr, C.c_longlong, C.c_int, ptr, stream_t]),
    # layout / elementwise
    'ladder_sym_pad': (C.c_int, [ptr, ptr] + [C.c_int] * 5 + [stream_t]),
    'ladder_sym_pad_bwd': (C.c_int, [ptr, ptr] + [C.c_int] * 5 + [stream_t]),
    'ladder_depth_to_space': (C.c_int, [ptr, ptr] + [C.c_int] * 5 + [stream_t]),
    'ladder_space_to_depth_actgrad': (C.c_int, [ptr, ptr, ptr] + [C.c_int] * 6 + [stream_t]),
    'ladder_act_bwd': (C.c_int, [ptr, ptr, C.c_longlong, C.c_int, stream_t]),
    'ladder_axpy': (C.c_int, [ptr, ptr, C.c_float, C.c_longlong, stream_t]),
    # CelebA-only layers
    'ladder_bn_stats': (C.c_int, [ptr, C.c_longlong, C.c_int, ptr, stream_t]),
    'ladder_bn_apply': (C.c_int, [ptr, ptr, ptr, ptr, ptr, C.c_longlong, C.c_int, C.c_longlong, C.c_float, C.c_int, stream_t]),
    'ladder_bn_bwd_stats': (C.c_int, [ptr, ptr, ptr, ptr, C.c_longlong, C.c_int, C.c_longlong, C.c_float, C.c_int, ptr, stream_t]),
    'ladder_bn_bwd_apply': (C.c_int, [ptr] * 7 + [C.c_longlong, C.c_int, C.c_longlong, C.c_float, C.c_int, stream_t]),
    'ladder_instnorm_style_fwd': (C.c_int, [ptr, ptr, ptr, ptr, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, stream_t]),
    'ladder_instnorm_style_bwd': (C.c_int, [ptr] * 7 + [C.c_int, C.c_int, C.c_int, C.c_int, stream_t]),
    'ladder_resize_bilinear_fwd': (C.c_int, [ptr, ptr] + [C.c_int] * 6 + [stream_t]),
    'ladder_resize_bilinear_bwd': (C.c_int, [ptr, ptr] + [C.c_int] * 6 + [stream_t]),
    'ladder_resize_bilinear_fwd_ex': (C.c_int, [ptr, C.c_int, ptr, C.c_int] + [C.c_int] * 6 + [stream_t]),
    'ladder_resize_bilinear_bwd_ex': (C.c_int, [ptr, C.c_int, ptr, C.c_int, ptr, C.c_int, C.c_int] + [C.c_int] * 6 + [stream_t]),
    'ladder_norm_fused_supported': (C.c_int, [C.c_int]),
    'ladder_bn_apply_bf16': (C.c_int, [ptr, ptr, ptr, ptr, ptr, C.c_longlong, C.c_int, C.c_longlong, C.c_float, C.c_int, stream_t]),
    'ladder_bn_bwd_stats_bf16': (C.c_int, [ptr, C.c_int, ptr, ptr, C.c_longlong, C.c_int, C.c_longlong, C.c_float, ptr, stream_t]),
    'ladder_bn_bwd_apply_bf16': (C.c_int, [ptr, C.c_int, ptr, ptr, ptr, ptr, ptr, C.c_longlong, C.c_int, C.c_longlong, C.c_float,
                                           ptr, stream_t]),
    'ladder_in_sums_bf16': (C.c_int, [ptr, C.c_int, C.c_int, C.c_int, ptr, stream_t]),
    'ladder_in_style_resize_bf16': (C.c_int, [ptr, ptr, ptr, ptr] + [C.c_int] * 6 + [C.c_float, C.c_int, stream_t]),
    'ladder_in_style_bwd_bf16': (C.c_int, [ptr, C.c_int, ptr, ptr, ptr, ptr, ptr, ptr, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                           stream_t]),
    # ELBO pieces
    'ladder_gauss_head_fwd': (C.c_int, [ptr, ptr, ptr, ptr, C.c_longlong, C.c_float, ptr, stream_t]),
    'ladder_gauss_head_bwd': (C.c_int, [ptr] * 8 + [C.c_longlong, C.c_float, C.c_float, C.c_float, stream_t]),
    'ladder_mc_sample': (C.c_int, [ptr, ptr, ptr, ptr, C.c_int, C.c_longlong, stream_t]),
    'ladder_mc_reduce': (C.c_int, [ptr, ptr, C.c_int, C.c_longlong, C.c_float, ptr, ptr, stream_t]),
    'ladder_sum': (C.c_int, [ptr, C.c_longlong, ptr, stream_t]),
    'ladder_l1_recon_fwd': (C.c_int, [ptr, ptr, C.c_longlong, ptr, stream_t]),
    'ladder_l1_recon_bwd': (C.c_int, [ptr, ptr, ptr, ptr, C.c_longlong, C.c_int, stream_t]),
    'ladder_code_recon_fwd': (C.c_int, [ptr, ptr, ptr, C.c_int, C.c_longlong, ptr, stream_t]),
    'ladder_code_recon_bwd': (C.c_int, [ptr, ptr, ptr, C.c_int, ptr, C.c_float, ptr, ptr, C.c_int, C.c_longlong, stream_t]),
    'ladder_elbo_scalars': (C.c_int, [ptr, ptr, ptr] + [C.c_int] * 8 + [C.c_float, C.c_float, C.c_int, C.c_int, stream_t]),
    'ladder_clip_adam': (C.c_int, [ptr, ptr, ptr, ptr, C.c_longlong, ptr, ptr, C.c_float, C.c_float, C.c_float, stream_t]),
    'ladder_increment': (C.c_int, [ptr, stream_t]),
    'ladder_philox_normal': (C.c_int, [ptr, C.c_int, C.c_int, ptr, C.c_int, C.c_int, ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_ulonglong, ptr, stream_t]),
    'ladder_gmm_param_stride': (C.c_int, [C.c_int]),
    'ladder_gmm_moment_stride': (C.c_int, [C.c_int]),
    'ladder_gmm_em_step': (C.c_int, [ptr, C.c_longlong, C.c_int, ptr, C.c_int, C.c_int, ptr, ptr, stream_t]),
    'ladder_gmm_score': (C.c_int, [ptr, C.c_longlong, C.c_int, ptr, C.c_int, ptr, ptr, stream_t]),
    'ladder_pipe_peak_launch': (C.c_int, [C.c_int, C.c_int, C.c_int, ptr, stream_t]),
    'ladder_mixture_tc_image_bytes': (C.c_size_t, [C.c_int, C.c_int]),
    'ladder_mixture_tc_pack_iso': (C.c_int, [c_double_p, C.c_double, c_double_p, C.c_int, C.c_int, c_float_p, c_float_p, c_float_p]),
    'ladder_mixture_tc_workspace_bytes': (C.c_size_t, [C.c_longlong, C.c_int]),
    'ladder_mixture_logprob_tc': (C.c_int, [ptr, C.c_longlong, C.c_int, ptr, ptr, C.c_int, C.c_float, C.c_float, ptr, ptr, C.c_size_t, stream_t]),
    'ladder_mixture_logprob_packed': (C.c_int, [ptr, C.c_longlong, C.c_int, ptr, C.c_int, C.c_int, C.c_float, C.c_float, ptr, C.c_int,
                                                ptr, C.c_size_t, stream_t]),
    'ladder_mixture_tc_grad_image_bytes': (C.c_size_t, [C.c_int, C.c_int]),
    'ladder_mixture_tc_pack_iso_grad': (C.c_int, [c_double_p, C.c_double, c_double_p, C.c_int, C.c_int, c_float_p, c_float_p, c_float_p]),
    'ladder_mixture_tc_grad_workspace_bytes': (C.c_size_t, [C.c_longlong, C.c_int, C.c_int]),
    'ladder_mixture_logprob_grad_tc': (C.c_int, [ptr, C.c_longlong, C.c_int, ptr, ptr, C.c_int, C.c_float, C.c_float, ptr, ptr, ptr,
                                                 C.c_size_t, stream_t]),
    'ladder_mixture_logprob_tc_packed': (C.c_int, [ptr, C.c_longlong, C.c_int, ptr, ptr, C.c_int, C.c_float, C.c_float, ptr, C.c_int,
                                                   ptr, C.c_size_t, stream_t]),
    'ladder_mixture_bigd_table_stride': (C.c_size_t, [C.c_int]),
    'ladder_mixture_bigd_workspace_bytes': (C.c_size_t, [C.c_longlong, C.c_int]),
    'ladder_mixture_logprob_bigd': (C.c_int, [ptr, C.c_longlong, C.c_int, ptr, C.c_int, ptr, ptr, ptr, C.c_size_t, stream_t]),
    'ladder_mixture_diag_bigd': (C.c_int, [ptr, C.c_longlong, C.c_int, ptr, ptr, C.c_int, ptr, ptr, ptr, stream_t]),
    'ladder_mixture_diag_bigd_param_grad': (C.c_int, [ptr, C.c_longlong, C.c_int, ptr, ptr, C.c_int, ptr, C.c_float, ptr, ptr,
                                                      stream_t]),
    'ladder_mixture_combine_packed': (C.c_int, [ptr, C.c_int, C.c_longlong, C.c_int, C.c_int, ptr, ptr, stream_t]),
    'ladder_mixture_combine': (C.c_int, [ptr, ptr, ptr, C.c_int, C.c_longlong, C.c_int, ptr, ptr, stream_t]),
}


def load():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'libladder_sm100.so not found at %s -- build it with '
            '`python -m ladder_latent_data_distribution_modelling_b200.build` '
            '(there is no CPU or PyTorch fallback for the ELBO hot path)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, what=''):
    if code != 0:
        msg = load().ladder_last_error().decode(errors='replace')
        raise RuntimeError('libladder_sm100 %s failed (%d): %s' % (what, code, msg))
