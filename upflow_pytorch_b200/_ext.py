"""ctypes binding of libupflow_b200.so (include/upflow_b200.h).

The library is the product: there is NO fallback.  If it is missing, or a call
is attempted without a CUDA device, this module raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libupflow_b200.so")

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_F = _c.c_float
_LL = _c.c_longlong

# name -> (restype, argtypes); mirrors include/upflow_b200.h one to one
SIGNATURES = {
    "upf_abi_version": (_I, []),
    "upf_last_error": (_c.c_char_p, []),
    "upf_launch_count": (_LL, []),
    "upf_last_kernel": (_c.c_char_p, []),
    "upf_corr_lrelu_fwd": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _F, _I, _P]),
    "upf_corr_lrelu_fwd_planar": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _P]),
    "upf_corr_lrelu_fwd_lp": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _F, _P]),
    "upf_warp_fwd_lp": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P, _P]),
    "upf_corr_lrelu_bwd": (_I, [_P, _I, _P, _I, _P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _F, _P]),
    "upf_warp_fwd": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _F, _I, _P, _I, _P]),
    "upf_warp_bwd": (_I, [_P, _I, _P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _F, _P]),
    "upf_occ_check": (_I, [_P, _I, _P, _I, _I, _I, _I, _F, _F, _I, _I, _P]),
    "upf_featnorm_stats": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "upf_featnorm_apply": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    "upf_featnorm_combine": (_I, [_P, _P, _I, _P, _P, _I, _I, _LL, _I, _I, _P]),
    "upf_resize_bilinear": (_I, [_P, _I, _I, _I, _P, _I, _I, _I, _I, _I, _c.POINTER(_F), _P]),
    "upf_sgu_blend": (_I, [_P, _I, _P, _I, _I, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P]),
    "upf_conv2d_fwd": (_I, [_P, _I, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P]),
    "upf_conv_chain_fwd": (_I, [_P, _I, _I, _I, _I, _P]),
    "upf_debug_conv_chain": (_I, [_I, _I]),
    "upf_debug_conv_chain_probe": (_I, [_P]),
    "upf_conv3x3_tap_combine": (_I, [_P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _F, _I, _P]),
    "upf_conv_tc_packed_elems": (_LL, [_I, _I, _I]),
    "upf_conv_tc_pack_weights": (_I, [_P, _P, _I, _I, _I, _P]),
    "upf_debug_conv_halo": (_I, [_I, _I]),
    "upf_debug_probe": (_I, [_P]),
    "upf_debug_conv_win": (_I, [_I, _I, _I]),
    "upf_debug_corr_pipe": (_I, [_I]),
    "upf_debug_conv_tc": (_I, [_I]),
    "upf_debug_wgrad_taps": (_I, [_I]),
    "upf_conv2d_wgrad_workspace_elems": (_LL, [_I, _I, _I, _I, _I, _I, _I, _I]),
    "upf_conv2d_wgrad": (_I, [_P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "upf_conv2d_wgrad_tc_workspace_elems": (_LL, [_I, _I, _I, _I, _I, _I, _I]),
    "upf_conv2d_wgrad_tc": (_I, [_P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "upf_wgrad_tc_planar_pitch": (_LL, [_I, _I, _I, _I, _I]),
    "upf_wgrad_tc_planar_elems": (_LL, [_I, _I, _I, _I, _I, _I]),
    "upf_wgrad_tc_transpose_input": (_I, [_P, _I, _I, _P, _I, _I, _I, _I, _I, _P]),
    "upf_conv2d_wgrad_tc_planar": (_I, [_P, _I, _I, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "upf_pointwise": (_I, [_I, _P, _I, _P, _I, _P, _I, _LL, _I, _F, _P]),
    "upf_act_split": (_I, [_P, _I, _P, _I, _P, _I, _P, _I, _LL, _I, _F, _P]),
    "upf_blend_fwd": (_I, [_P, _I, _P, _I, _P, _I, _P, _I, _LL, _P]),
    "upf_blend_bwd": (_I, [_P, _I, _P, _I, _P, _I, _P, _I, _P, _I, _P, _I, _P, _I, _LL, _P]),
    "upf_featnorm_bwd_workspace_doubles": (_LL, [_I, _I]),
    "upf_featnorm_bwd": (_I, [_P, _I, _P, _P, _I, _P, _I, _P, _I, _I, _I, _I, _P]),
    "upf_resize_bilinear_bwd_workspace_elems": (_LL, [_I, _I, _I, _I]),
    "upf_resize_bilinear_bwd": (_I, [_P, _I, _I, _I, _P, _I, _I, _I, _I, _I, _c.POINTER(_F), _P, _P]),
    "upf_repack_conv_weight": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "upf_repack_conv_weight_tc": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "upf_loss_workspace_elems": (_LL, []),
    "upf_robust_loss_fwd": (_I, [_P, _I, _P, _I, _P, _I, _P, _P, _LL, _I, _I, _F, _P]),
    "upf_robust_loss_bwd": (_I, [_P, _I, _P, _I, _P, _I, _P, _P, _P, _I, _P, _I, _LL, _I, _I, _F, _P]),
    "upf_edge_smooth1_fwd": (_I, [_P, _I, _I, _P, _I, _I, _P, _P, _I, _I, _I, _P]),
    "upf_edge_smooth1_bwd": (_I, [_P, _I, _I, _P, _I, _I, _P, _P, _I, _I, _I, _I, _P]),
    "upf_census_loss_fwd": (_I, [_P, _I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "upf_census_loss_bwd": (_I, [_P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P]),
    "upf_boundary_warp_fwd": (_I, [_P, _I, _I, _I, _I, _P, _I, _P, _P, _I, _I, _I, _I, _P]),
    "upf_boundary_warp_bwd": (_I, [_P, _I, _I, _I, _I, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _P]),
    "upf_nchw_to_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "upf_nhwc_to_nchw": (_I, [_P, _I, _P, _I, _I, _I, _I, _P]),
    "upf_copy_channels": (_I, [_P, _I, _P, _I, _LL, _I, _I, _P]),
}



class ChainLayer(ctypes.Structure):
    """upf_chain_layer (include/upflow_b200.h)"""
    _fields_ = [("x", _P), ("ldx", _I), ("w_packed", _P), ("bias", _P), ("out", _P), ("ldo", _I), ("residual", _P), ("ldr", _I),
                ("out2", _P), ("ldo2", _I), ("Cin", _I), ("Cout", _I), ("ksize", _I), ("dilation", _I), ("slope", _F), ("flags", _I)]


CONV_FP32 = 0
CONV_TF32 = 1
CONV_ROUND_OUT = 0x100       # OR-ed into the precision: store the output rounded to the nearest TF32 value
FLAG_ROUND_TF32 = 1
DTYPE_F16, DTYPE_BF16 = 1, 2      # UPF_DTYPE_* (low-precision storage variants)
ABI_VERSION = 2
PW_LRELU_BWD, PW_SIGMOID, PW_SIGMOID_BWD = 0, 1, 2
LOSS_KINDS = {"abs_robust": 0, "charbonnier": 1, "L1": 2}

_lib = None


class UpflowLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (building is ``__graft_entry__.build()`` /
    ``python -m upflow_pytorch_b200.build``); raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UpflowLibraryError(
            "%s not found: build it with `python -m upflow_pytorch_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the header and the .so disagree
        fn.restype = res
        fn.argtypes = args
    if lib.upf_abi_version() != ABI_VERSION:
        raise UpflowLibraryError("ABI version mismatch")
    if os.environ.get("UPF_WGRAD_TAPS"):          # A/B runs: largest padded Cout served by the taps-along-N weight gradient
        lib.upf_debug_wgrad_taps(int(os.environ["UPF_WGRAD_TAPS"]))
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().upf_last_error()
        raise RuntimeError("upflow_b200.%s failed (%d): %s" % (what, code, msg.decode() if msg else ""))


def launch_count():
    return int(load().upf_launch_count())
