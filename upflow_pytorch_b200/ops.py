"""Tensor-level wrappers over the C ABI (include/upflow_b200.h).

Two layers:
  * ``k_*`` functions: thin launchers on PIXEL-MAJOR buffers (torch tensors of
    shape [N,H,W,ld], contiguous; a channel slice is (buffer, channel offset,
    channel count)).  The decoder engine (engine.py) uses only these.
  * NCHW-in / NCHW-out functions with the reference's operator semantics
    (``correlation``, ``warp``, ``normalize_features``, ``upsample2d_flow_as``,
    ``conv2d`` ...) used by the drop-in modules.  Inputs in torch
    ``channels_last`` memory format are used in place; NCHW-contiguous inputs
    are transposed by the library's own copy kernel.  Outputs are NCHW-shaped
    views of pixel-major storage (== channels_last tensors).

PyTorch is plumbing here: allocation, streams, autograd bookkeeping.  All
arithmetic runs in libupflow_b200.so; there is no fallback.
"""
import ctypes

import torch

from . import _ext

LRELU_SLOPE = 0.1
# diagnostic only (tests): replaces the reference's `mask >= 1.0` threshold of every module-level warp, like
# oracle/ref_port.MASK_THRESHOLD on the other side (see tests/test_gpu_engine.py for why)
DIAG_MASK_THRESHOLD = None


def _lib():
    return _ext.load()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t, offset_elems=0):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr() + 4 * offset_elems)


def _require_cuda(*ts, lp_ok=False):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("upflow_pytorch_b200 ops need CUDA tensors (no CPU path exists); got %s" % t.device)
        if lp_ok and t.dtype in (torch.float16, torch.bfloat16):
            continue                     # correlation / warp have fp16 / bf16 storage variants
        if t.dtype != torch.float32:
            raise RuntimeError("upflow_pytorch_b200 ops are fp32; got %s" % t.dtype)


class Slice:
    """Channels [c0, c0+C) of a pixel-major buffer [N,H,W,ld]."""
    __slots__ = ("buf", "c0", "C")

    def __init__(self, buf, c0=0, C=None):
        assert buf.dim() == 4 and buf.is_contiguous(), "pixel-major buffers are contiguous [N,H,W,ld]"
        self.buf = buf
        self.c0 = c0
        self.C = buf.shape[3] - c0 if C is None else C
        assert 0 <= c0 and c0 + self.C <= buf.shape[3]

    @property
    def ld(self):
        return self.buf.shape[3]

    @property
    def N(self):
        return self.buf.shape[0]

    @property
    def H(self):
        return self.buf.shape[1]

    @property
    def W(self):
        return self.buf.shape[2]

    def ptr(self):
        return _p(self.buf, self.c0)

    def nchw(self):
        """NCHW-shaped view (channels_last strides) of this slice."""
        return self.buf[..., self.c0:self.c0 + self.C].permute(0, 3, 1, 2)


def _as_slice(x):
    return x if isinstance(x, Slice) else Slice(x)


PLANAR_CORR = True              # NCHW inputs with 16-byte pitches go to the planar kernel (False: always pixel-major)
PLANAR_CORR_MIN_PIXELS = 16384  # smaller images: too few 120x4 tiles for 148 CTAs (measured: 47x156 29 vs 22 us pixel-major)

# ---------------------------------------------------------------- launchers
def k_corr(f1, f2, out, max_disp=4, stats1=None, stats2=None, f2_shift=0, slope=LRELU_SLOPE, round_tf32=False):
    f1, f2, out = _as_slice(f1), _as_slice(f2), _as_slice(out)
    assert f1.C == f2.C and out.C == (2 * max_disp + 1) ** 2
    _ext.check(_lib().upf_corr_lrelu_fwd(f1.ptr(), f1.ld, f2.ptr(), f2.ld, out.ptr(), out.ld, f1.N, f1.H, f1.W, f1.C,
                                         max_disp, _p(stats1), _p(stats2), f2_shift, slope,
                                         _ext.FLAG_ROUND_TF32 if round_tf32 else 0, _stream()), "corr_lrelu_fwd")


def last_kernel():
    """kernel family that served this thread's most recent library call (upf_last_kernel)"""
    return _lib().upf_last_kernel().decode()


def _planar_ok(t):
    """dense planar [N,C,H,W] whose row / plane / image pitches TMA can walk (multiples of 16 bytes)"""
    return (t.dim() == 4 and t.dtype == torch.float32 and t.stride(3) == 1 and all(st % 4 == 0 for st in t.stride()[:3])
            and t.stride(2) >= t.shape[3] and t.stride(1) >= t.stride(2) * t.shape[2] and t.stride(0) >= t.stride(1) * t.shape[1]
            and t.data_ptr() % 16 == 0)


def k_corr_planar(f1, f2, out, max_disp=4, f2_shift=0, slope=LRELU_SLOPE, round_tf32=False):
    """planar (NCHW) feature maps -> planar cost volume [N,(2d+1)^2,H,W], no layout conversion (corr_planar.cu)."""
    N, C, H, W = f1.shape
    assert f2.shape == f1.shape and tuple(out.shape) == (N, (2 * max_disp + 1) ** 2, H, W)
    assert _planar_ok(f1) and _planar_ok(f2) and _planar_ok(out) and max_disp <= 4
    pit = [(_ext.ctypes.c_longlong * 3)(t.stride(2), t.stride(1), t.stride(0)) for t in (f1, f2, out)]
    _ext.check(_lib().upf_corr_lrelu_fwd_planar(f1.data_ptr(), pit[0], f2.data_ptr(), pit[1], out.data_ptr(), pit[2],
                                                N, H, W, C, max_disp, f2_shift, slope,
                                                _ext.FLAG_ROUND_TF32 if round_tf32 else 0, _stream()), "corr_lrelu_fwd_planar")


_LP_DTYPES = {torch.float16: _ext.DTYPE_F16, torch.bfloat16: _ext.DTYPE_BF16}


def k_corr_lp(f1, f2, out, max_disp=4, stats1=None, stats2=None, f2_shift=0, slope=LRELU_SLOPE):
    """fp16 / bf16 storage variant: pixel-major [N,H,W,C] half tensors in, [N,H,W,(2d+1)^2] of the same dtype out, fp32
    arithmetic (upf_corr_lrelu_fwd_lp).  Pitches may exceed C (channel slices of a wider buffer)."""
    assert f1.dtype in _LP_DTYPES and f2.dtype == f1.dtype and out.dtype == f1.dtype
    N, H, W, C = f1.shape
    assert f1.stride(3) == 1 and f2.stride(3) == 1 and out.stride(3) == 1 and out.shape[3] == (2 * max_disp + 1) ** 2
    dp = lambda t: ctypes.c_void_p(t.data_ptr())
    _ext.check(_lib().upf_corr_lrelu_fwd_lp(dp(f1), f1.stride(2), dp(f2), f2.stride(2), dp(out), out.stride(2), _LP_DTYPES[f1.dtype],
                                            N, H, W, C, max_disp, _p(stats1), _p(stats2), f2_shift, slope, _stream()), "corr_lrelu_fwd_lp")


def k_warp_lp(x, flow, out, align_corners=False, use_mask=True, x_shift=0, stats=None):
    """fp16 / bf16 storage variant of k_warp: x, out pixel-major half tensors, flow fp32 [N,H,W,>=2]."""
    assert x.dtype in _LP_DTYPES and out.dtype == x.dtype and flow.dtype == torch.float32
    N, H, W, C = out.shape
    dp = lambda t: ctypes.c_void_p(t.data_ptr())
    _ext.check(_lib().upf_warp_fwd_lp(dp(x), x.stride(2), dp(flow), flow.stride(2), dp(out), out.stride(2), _LP_DTYPES[x.dtype],
                                      N, H, W, C, int(align_corners), _mask_thr(use_mask), x_shift, _p(stats), _stream()), "warp_fwd_lp")


def _pixel_major_any(x):
    """[N,C,H,W] of any dtype -> contiguous pixel-major [N,H,W,C] (a view when x is channels_last)"""
    return x.permute(0, 2, 3, 1).contiguous()


def k_corr_bwd(f1, f2, out, grad_out, grad_f1, grad_f2, max_disp=4, slope=1.0):
    f1, f2, grad_out = _as_slice(f1), _as_slice(f2), _as_slice(grad_out)
    out = _as_slice(out) if out is not None else None
    g1, g2 = _as_slice(grad_f1), _as_slice(grad_f2)
    _ext.check(_lib().upf_corr_lrelu_bwd(f1.ptr(), f1.ld, f2.ptr(), f2.ld, out.ptr() if out else None,
                                         out.ld if out else 0, grad_out.ptr(), grad_out.ld, g1.ptr(), g1.ld, g2.ptr(),
                                         g2.ld, f1.N, f1.H, f1.W, f1.C, max_disp, slope, _stream()), "corr_lrelu_bwd")


def _mask_thr(use_mask):
    """True -> 1.0 (the reference's `mask >= 1.0`), False -> no mask, float -> that threshold (diagnostic)."""
    if use_mask is True:
        return 1.0 if DIAG_MASK_THRESHOLD is None else float(DIAG_MASK_THRESHOLD)
    if use_mask is False or use_mask is None:
        return 0.0
    return float(use_mask)


def k_warp(x, flow, out, align_corners=False, use_mask=True, x_shift=0, stats=None, round_tf32=False):
    x, flow, out = _as_slice(x), _as_slice(flow), _as_slice(out)
    assert flow.C >= 2 and out.C == x.C
    _ext.check(_lib().upf_warp_fwd(x.ptr(), x.ld, flow.ptr(), flow.ld, out.ptr(), out.ld, out.N, out.H, out.W, x.C,
                                   int(align_corners), _mask_thr(use_mask), x_shift, _p(stats),
                                   _ext.FLAG_ROUND_TF32 if round_tf32 else 0, _stream()), "warp_fwd")


def k_warp_bwd(x, flow, grad_out, grad_x, grad_flow, align_corners=False, use_mask=True):
    x, flow, grad_out = _as_slice(x), _as_slice(flow), _as_slice(grad_out)
    gx = _as_slice(grad_x) if grad_x is not None else None
    gf = _as_slice(grad_flow) if grad_flow is not None else None
    _ext.check(_lib().upf_warp_bwd(x.ptr(), x.ld, flow.ptr(), flow.ld, grad_out.ptr(), grad_out.ld,
                                   gx.ptr() if gx else None, gx.ld if gx else 0, gf.ptr() if gf else None,
                                   gf.ld if gf else 0, x.N, x.H, x.W, x.C, int(align_corners), _mask_thr(use_mask), _stream()),
               "warp_bwd")


OCC_MODES = {"all": 0, "obj": 1, "out": 2}


def k_occ_check(flow, occ, alpha_1, alpha_2, mode="obj", align_corners=False):
    """flow [2B,H,W,2] = [forward flows; backward flows] -> occ [2B,H,W,1] (1 = visible)."""
    flow, occ = _as_slice(flow), _as_slice(occ)
    _ext.check(_lib().upf_occ_check(flow.ptr(), flow.ld, occ.ptr(), occ.ld, flow.N, flow.H, flow.W, float(alpha_1),
                                    float(alpha_2), OCC_MODES[mode], int(align_corners), _stream()), "occ_check")


def k_stats(x, stats):
    x = _as_slice(x)
    assert stats.dtype == torch.float64 and stats.numel() >= x.N * x.C * 2
    _ext.check(_lib().upf_featnorm_stats(x.ptr(), x.ld, x.N, x.H, x.W, x.C, ctypes.c_void_p(stats.data_ptr()), _stream()),
               "featnorm_stats")


def k_norm_combine(stats_a, stats_b, out_a, out_b, N, C, npix, shift=0, across_channels=False, across_images=False):
    """Rewrite the raw moments of the pairs (stats_a[n], stats_b[(n+shift)%N]) into equivalent per-channel moments of
    normalize_features' pooled modes (model/upflow.py:109-124)."""
    dp = lambda t: ctypes.c_void_p(t.data_ptr())
    for t in (stats_a, stats_b, out_a, out_b):
        assert t.dtype == torch.float64 and t.numel() >= N * C * 2
    _ext.check(_lib().upf_featnorm_combine(dp(stats_a), dp(stats_b), shift, dp(out_a), dp(out_b), N, C, int(npix),
                                           int(across_channels), int(across_images), _stream()), "featnorm_combine")


def k_norm_apply(x, stats, out):
    x, out = _as_slice(x), _as_slice(out)
    _ext.check(_lib().upf_featnorm_apply(x.ptr(), x.ld, ctypes.c_void_p(stats.data_ptr()), out.ptr(), out.ld, x.N, x.H,
                                         x.W, x.C, _stream()), "featnorm_apply")


def k_resize(src, out, scale=None):
    src, out = _as_slice(src), _as_slice(out)
    assert src.C == out.C <= 4
    sc = None
    if scale is not None:
        sc = (ctypes.c_float * 4)(*([float(s) for s in scale] + [1.0] * (4 - len(scale))))
    _ext.check(_lib().upf_resize_bilinear(src.ptr(), src.ld, src.H, src.W, out.ptr(), out.ld, out.H, out.W, src.N, src.C,
                                          sc, _stream()), "resize_bilinear")


def k_sgu_blend(flow_init, inter, out, align_corners=False, out_tc=None, round_tf32=False):
    """out_tc: a 4-channel slot of a convolution input buffer that receives (u, v, 0, 0), rounded to TF32 on request."""
    flow_init, inter, out = _as_slice(flow_init), _as_slice(inter), _as_slice(out)
    o2 = _as_slice(out_tc) if out_tc is not None else None
    assert o2 is None or o2.C == 4
    _ext.check(_lib().upf_sgu_blend(flow_init.ptr(), flow_init.ld, inter.ptr(), inter.ld, inter.H, inter.W, out.ptr(),
                                    out.ld, o2.ptr() if o2 else None, o2.ld if o2 else 0, out.N, out.H, out.W,
                                    int(align_corners), _ext.FLAG_ROUND_TF32 if round_tf32 else 0, _stream()), "sgu_blend")


def k_conv(x, weight, bias, out, ksize, stride=1, dilation=1, slope=LRELU_SLOPE, residual=None, precision=_ext.CONV_FP32):
    """weight: the layout matching ``precision`` (see pack_conv_weight)."""
    x, out = _as_slice(x), _as_slice(out)
    res = _as_slice(residual) if residual is not None else None
    _ext.check(_lib().upf_conv2d_fwd(x.ptr(), x.ld, _p(weight), _p(bias), out.ptr(), out.ld, res.ptr() if res else None,
                                     res.ld if res else 0, x.N, x.H, x.W, x.C, out.C, ksize, stride, dilation,
                                     float(slope), int(precision), _stream()), "conv2d_fwd")


def chain_layer(x, weight_tc, bias, out, ksize, dilation=1, slope=LRELU_SLOPE, residual=None, round_tf32=False, out2=None):
    """One layer of k_conv_chain: k_conv(..., precision=CONV_TF32) with stride 1; out2 = optional TF32-rounded second copy."""
    x, out = _as_slice(x), _as_slice(out)
    res = _as_slice(residual) if residual is not None else None
    o2 = _as_slice(out2) if out2 is not None else None
    L = _ext.ChainLayer()
    L.x, L.ldx, L.w_packed, L.bias = x.ptr(), x.ld, _p(weight_tc), _p(bias)
    L.out, L.ldo = out.ptr(), out.ld
    L.residual, L.ldr = (res.ptr(), res.ld) if res else (None, 0)
    L.out2, L.ldo2 = (o2.ptr(), o2.ld) if o2 else (None, 0)
    L.Cin, L.Cout, L.ksize, L.dilation = x.C, out.C, ksize, dilation
    L.slope, L.flags = float(slope), (_ext.FLAG_ROUND_TF32 if round_tf32 else 0)
    L._shape = (x.N, x.H, x.W)
    L._keep = (x.buf, weight_tc, bias, out.buf, res.buf if res else None, o2.buf if o2 else None)
    return L


def k_conv_chain(layers):
    """Dependent stride-1 tensor-core convolutions on one small feature map in ONE persistent launch (conv_chain.cu)."""
    N, H, W = layers[0]._shape
    assert all(L._shape == (N, H, W) for L in layers), "every layer of a chain works on the same [N,H,W] map"
    arr = (_ext.ChainLayer * len(layers))(*layers)
    _ext.check(_lib().upf_conv_chain_fwd(ctypes.cast(arr, ctypes.c_void_p), len(layers), N, H, W, _stream()), "conv_chain_fwd")


def k_wgrad_planar_input(x, ksize, dilation=1):
    """The planar, zero-padded transpose of slice x that the tensor-core weight gradient reads, made ONCE for several
    convolutions whose inputs are nested channel ranges of x (same ksize / dilation).  Returns (xt, rows): blocked planar
    [pitch / 32][rows = x.C][32]; a convolution reading channels [c0, c0 + Cin) passes planar=(xt, rows, c0)."""
    x = _as_slice(x)
    xt = torch.empty(int(_lib().upf_wgrad_tc_planar_elems(x.N, x.H, x.W, x.C, ksize, dilation)), dtype=torch.float32,
                     device=x.buf.device)
    _ext.check(_lib().upf_wgrad_tc_transpose_input(x.ptr(), x.ld, x.C, _p(xt), x.N, x.H, x.W, ksize, dilation, _stream()),
               "wgrad_tc_transpose_input")
    return xt, x.C


def k_conv_wgrad(x, grad_out, ksize, stride=1, dilation=1, want_bias=True, tensor_cores=False, planar=None):
    """Weight gradient [k*k, Cin, Cout] and bias gradient [Cout] of conv() from its input and the gradient wrt its
    PRE-activation output (both pixel-major).  tensor_cores: TF32 tcgen05 GEMM (stride 1), else fp32 SIMT.
    planar = (xt, rows, c0): k_wgrad_planar_input's transpose of a `rows`-channel buffer whose channels [c0, c0 + x.C) are x."""
    x, g = _as_slice(x), _as_slice(grad_out)
    lib = _lib()
    if tensor_cores and stride == 1:
        n = lib.upf_conv2d_wgrad_tc_workspace_elems(x.N, x.H, x.W, x.C, g.C, ksize, dilation)
        ws = torch.empty(n, dtype=torch.float32, device=x.buf.device)
        gw = torch.empty(ksize * ksize, x.C, g.C, dtype=torch.float32, device=x.buf.device)
        gb = torch.empty(g.C, dtype=torch.float32, device=x.buf.device) if want_bias else None
        if planar is not None:
            xt, rows, c0 = planar
            _ext.check(lib.upf_conv2d_wgrad_tc_planar(_p(xt), rows, c0, g.ptr(), g.ld, _p(gw), _p(gb), _p(ws), x.N, x.H, x.W,
                                                      x.C, g.C, ksize, dilation, _stream()), "conv2d_wgrad_tc_planar")
        else:
            _ext.check(lib.upf_conv2d_wgrad_tc(x.ptr(), x.ld, g.ptr(), g.ld, _p(gw), _p(gb), _p(ws), x.N, x.H, x.W, x.C, g.C,
                                               ksize, dilation, _stream()), "conv2d_wgrad_tc")
        return gw, gb
    n = lib.upf_conv2d_wgrad_workspace_elems(x.N, x.H, x.W, x.C, g.C, ksize, stride, dilation)
    ws = torch.empty(n, dtype=torch.float32, device=x.buf.device)
    gw = torch.empty(ksize * ksize, x.C, g.C, dtype=torch.float32, device=x.buf.device)
    gb = torch.empty(g.C, dtype=torch.float32, device=x.buf.device) if want_bias else None
    _ext.check(lib.upf_conv2d_wgrad(x.ptr(), x.ld, g.ptr(), g.ld, _p(gw), _p(gb), _p(ws), x.N, x.H, x.W, x.C, g.C, ksize,
                                    stride, dilation, _stream()), "conv2d_wgrad")
    return gw, gb


def k_pointwise(op, a, b, out, slope=LRELU_SLOPE):
    a, out = _as_slice(a), _as_slice(out)
    b = _as_slice(b) if b is not None else None
    _ext.check(_lib().upf_pointwise(op, a.ptr(), a.ld, b.ptr() if b else None, b.ld if b else 0, out.ptr(), out.ld,
                                    a.N * a.H * a.W, a.C, float(slope), _stream()), "pointwise")


def k_featnorm_bwd(x, stats, grad_out, grad_x):
    x, g, gx = _as_slice(x), _as_slice(grad_out), _as_slice(grad_x)
    lib = _lib()
    ws = torch.empty(lib.upf_featnorm_bwd_workspace_doubles(x.N, x.C), dtype=torch.float64, device=x.buf.device)
    _ext.check(lib.upf_featnorm_bwd(x.ptr(), x.ld, ctypes.c_void_p(stats.data_ptr()), g.ptr(), g.ld, gx.ptr(), gx.ld,
                                    ctypes.c_void_p(ws.data_ptr()), x.N, x.H, x.W, x.C, _stream()), "featnorm_bwd")


def k_resize_bwd(grad_out, grad_in, scale=None):
    g, gi = _as_slice(grad_out), _as_slice(grad_in)
    sc = None
    if scale is not None:
        sc = (ctypes.c_float * 4)(*([float(v) for v in scale] + [1.0] * (4 - len(scale))))
    lib = _lib()
    ws = torch.empty(lib.upf_resize_bilinear_bwd_workspace_elems(g.N, g.H, gi.W, g.C), dtype=torch.float32, device=g.buf.device)
    _ext.check(lib.upf_resize_bilinear_bwd(g.ptr(), g.ld, g.H, g.W, gi.ptr(), gi.ld, gi.H, gi.W, g.N, g.C, sc, _p(ws), _stream()),
               "resize_bilinear_bwd")


def k_tap_combine(y, bias, out, dilation=1, slope=LRELU_SLOPE, residual=None, round_tf32=False):
    """out = lrelu(bias + sum of the nine shifted tap slices of y) (+ residual); y holds 9*Cout channels."""
    y, out = _as_slice(y), _as_slice(out)
    res = _as_slice(residual) if residual is not None else None
    assert y.C == 9 * out.C
    _ext.check(_lib().upf_conv3x3_tap_combine(y.ptr(), y.ld, _p(bias), out.ptr(), out.ld, res.ptr() if res else None,
                                              res.ld if res else 0, out.N, out.H, out.W, out.C, dilation, float(slope),
                                              _ext.FLAG_ROUND_TF32 if round_tf32 else 0, _stream()), "conv3x3_tap_combine")


def expand_taps_weight(weight):
    """[Cout,Cin,3,3] -> [9*Cout,Cin,1,1] with row tap*Cout+co = weight[co,:,ky,kx], tap = ky*3+kx."""
    Cout, Cin, k, _ = weight.shape
    assert k == 3
    return weight.detach().permute(2, 3, 0, 1).reshape(9 * Cout, Cin, 1, 1).contiguous()


def k_act_split(t, out, out_lo=None, slope=1.0, residual=None):
    """out = lrelu(t) (+ residual); out_lo = out - trunc_tf32(out).  t=None: only split `out` into out_lo."""
    out = _as_slice(out)
    t = _as_slice(t) if t is not None else None
    lo = _as_slice(out_lo) if out_lo is not None else None
    res = _as_slice(residual) if residual is not None else None
    _ext.check(_lib().upf_act_split(t.ptr() if t else None, t.ld if t else 0, res.ptr() if res else None, res.ld if res else 0,
                                    out.ptr(), out.ld, lo.ptr() if lo else None, lo.ld if lo else 0, out.N * out.H * out.W,
                                    out.C, float(slope), _stream()), "act_split")


def k_copy(src, dst, round_tf32=False):
    """dst = src (src=None: zeros), optionally rounded to the nearest TF32 value."""
    dst = _as_slice(dst)
    src = _as_slice(src) if src is not None else None
    assert src is None or src.C == dst.C
    _ext.check(_lib().upf_copy_channels(src.ptr() if src else None, src.ld if src else 0, dst.ptr(), dst.ld,
                                        dst.N * dst.H * dst.W, dst.C, _ext.FLAG_ROUND_TF32 if round_tf32 else 0, _stream()),
               "copy_channels")


# ---------------------------------------------------------------- weights
def pack_conv_weight(weight, in_slots=None, cin_total=None, tc=False, flip_transpose=False, tc_only=False):
    """nn.Conv2d weight [Cout,Cin,k,k] -> library layouts.

    flip_transpose: the weights of the input-gradient convolution instead (taps flipped, Cin and Cout exchanged).
    tc_only (with tc): only the tensor-core layout, in one launch; the first element of the result is None.

    in_slots[i] = position, inside the input slice the kernel reads, of the
    reference's input channel i (identity when None); cin_total = width of that
    slice (unused slots get zero weights -- this is how the dense blocks'
    prepend-concatenation order (model/pwc_modules.py:280-284) is mapped onto
    append-only buffers).  Returns (w_simt [taps,cin_total,cout_pad4], w_tc or None)."""
    Cout, Cin, k, _ = weight.shape
    if flip_transpose or (in_slots is None and cin_total in (None, Cin) and weight.is_cuda):
        if in_slots is not None or cin_total not in (None, Cin):
            raise ValueError("flip_transpose packs the plain layout only")
        src = weight.detach().float().contiguous()
        _require_cuda(src)
        if flip_transpose:
            Cout, Cin = Cin, Cout
        if tc and tc_only:
            w_tc = torch.empty(_lib().upf_conv_tc_packed_elems(Cin, Cout, k), dtype=torch.float32, device=weight.device)
            _ext.check(_lib().upf_repack_conv_weight_tc(_p(src), _p(w_tc), weight.shape[0], weight.shape[1], k,
                                                        1 if flip_transpose else 0, _stream()), "repack_conv_weight_tc")
            return None, w_tc
        w = torch.empty(k * k, Cin, (Cout + 3) // 4 * 4, dtype=torch.float32, device=weight.device)
        _ext.check(_lib().upf_repack_conv_weight(_p(src), _p(w), weight.shape[0], weight.shape[1], k,
                                                 1 if flip_transpose else 0, _stream()), "repack_conv_weight")
        w_tc = None
        if tc:
            w_tc = torch.empty(_lib().upf_conv_tc_packed_elems(Cin, Cout, k), dtype=torch.float32, device=weight.device)
            _ext.check(_lib().upf_conv_tc_pack_weights(_p(w), _p(w_tc), Cin, Cout, k, _stream()), "conv_tc_pack_weights")
        return w, w_tc
    cin_total = cin_total or Cin
    cout_pad = (Cout + 3) // 4 * 4
    w = torch.zeros(k * k, cin_total, cout_pad, dtype=torch.float32, device=weight.device)
    src = weight.detach().float().permute(2, 3, 1, 0).reshape(k * k, Cin, Cout)
    if in_slots is None:
        w[:, :Cin, :Cout] = src
    else:
        idx = torch.as_tensor(in_slots, dtype=torch.long, device=weight.device)
        w[:, idx, :Cout] = src
    w = w.contiguous()
    w_tc = None
    if tc:
        n = _lib().upf_conv_tc_packed_elems(cin_total, Cout, k)
        w_tc = torch.empty(n, dtype=torch.float32, device=weight.device)
        _ext.check(_lib().upf_conv_tc_pack_weights(_p(w), _p(w_tc), cin_total, Cout, k, _stream()), "conv_tc_pack_weights")
    return w, w_tc


# ---------------------------------------------------------------- NCHW <-> pixel-major
def to_pixel_major(x, ld=None):
    """[N,C,H,W] tensor -> pixel-major buffer [N,H,W,ld>=C].  channels_last
    inputs with ld==C are returned as a view."""
    _require_cuda(x)
    N, C, H, W = x.shape
    v = x.permute(0, 2, 3, 1)
    if (ld is None or ld == C) and v.is_contiguous():
        return v
    ld = ld or C
    out = torch.empty(N, H, W, ld, dtype=torch.float32, device=x.device) if ld == C else \
        torch.zeros(N, H, W, ld, dtype=torch.float32, device=x.device)
    if x.is_contiguous():
        _ext.check(_lib().upf_nchw_to_nhwc(_p(x), _p(out), ld, N, C, H, W, _stream()), "nchw_to_nhwc")
    else:
        if not v.is_contiguous():
            xc = x.contiguous()
            _ext.check(_lib().upf_nchw_to_nhwc(_p(xc), _p(out), ld, N, C, H, W, _stream()), "nchw_to_nhwc")
        else:
            k_copy(Slice(v), Slice(out, 0, C))
    return out


def to_nchw_contiguous(buf, C=None):
    s = _as_slice(buf) if C is None else Slice(buf, 0, C)
    out = torch.empty(s.N, s.C, s.H, s.W, dtype=torch.float32, device=s.buf.device)
    _ext.check(_lib().upf_nhwc_to_nchw(s.ptr(), s.ld, _p(out), s.N, s.C, s.H, s.W, _stream()), "nhwc_to_nchw")
    return out


_ZERO_BIAS = {}


def _zero_bias(n, device):
    """A read-only all-zero bias vector (never written, kept alive: one fill per device instead of one per call)."""
    z = _ZERO_BIAS.get(device)
    if z is None or z.numel() < n:
        z = _ZERO_BIAS[device] = torch.zeros(max(1024, n), dtype=torch.float32, device=device)
    return z[:n]


def _new(N, H, W, C, like):
    return torch.empty(N, H, W, C, dtype=torch.float32, device=like.device)


# ---------------------------------------------------------------- reference-semantics operators (NCHW in/out)
class _CorrelationFn(torch.autograd.Function):
    """Correlation(pad=d, k=1, maxd=d, s1=s2=1) -- model/correlation_package/correlation.py:6-44."""

    @staticmethod
    def forward(ctx, in1, in2, max_disp, slope):
        a, b = to_pixel_major(in1), to_pixel_major(in2)
        N, H, W, C = a.shape
        out = _new(N, H, W, (2 * max_disp + 1) ** 2, in1)
        k_corr(a, b, out, max_disp, slope=slope)
        ctx.save_for_backward(a, b, out)
        ctx.max_disp, ctx.slope = max_disp, slope
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad):
        a, b, out = ctx.saved_tensors
        g = to_pixel_major(grad)
        g1, g2 = torch.empty_like(a), torch.empty_like(b)
        k_corr_bwd(a, b, out if ctx.slope != 1.0 else None, g, g1, g2, ctx.max_disp, ctx.slope)
        return g1.permute(0, 3, 1, 2), g2.permute(0, 3, 1, 2), None, None


def correlation(in1, in2, max_disp=4, leaky_slope=None):
    """Cost volume [B,(2d+1)^2,H,W]; ``leaky_slope`` fuses the LeakyReLU of model/upflow.py:563-564."""
    _require_cuda(in1, in2, lp_ok=True)
    if in1.shape != in2.shape:
        raise RuntimeError("correlation: shape mismatch %s vs %s" % (tuple(in1.shape), tuple(in2.shape)))
    if in1.dtype in _LP_DTYPES:
        # fp16 / bf16 storage (the reference's Half dispatch): inference only -- train in fp32
        if in2.dtype != in1.dtype:
            raise RuntimeError("correlation: dtype mismatch %s vs %s" % (in1.dtype, in2.dtype))
        if torch.is_grad_enabled() and (in1.requires_grad or in2.requires_grad):
            raise RuntimeError("correlation: the fp16 / bf16 storage variant has no backward; train in fp32")
        a, b = _pixel_major_any(in1), _pixel_major_any(in2)
        out = torch.empty(a.shape[0], a.shape[1], a.shape[2], (2 * max_disp + 1) ** 2, device=a.device, dtype=a.dtype)
        k_corr_lp(a, b, out, max_disp, slope=1.0 if leaky_slope is None else float(leaky_slope))
        return out.permute(0, 3, 1, 2)
    if (PLANAR_CORR and max_disp in (3, 4) and not (torch.is_grad_enabled() and (in1.requires_grad or in2.requires_grad))
            and _planar_ok(in1) and _planar_ok(in2) and in1.shape[2] * in1.shape[3] >= PLANAR_CORR_MIN_PIXELS):
        # the reference operator's own layout, inference: planar in, planar out, no transposes (corr_planar.cu)
        out = torch.empty(in1.shape[0], (2 * max_disp + 1) ** 2, in1.shape[2], in1.shape[3], device=in1.device, dtype=in1.dtype)
        k_corr_planar(in1, in2, out, max_disp, slope=1.0 if leaky_slope is None else float(leaky_slope))
        return out
    return _CorrelationFn.apply(in1, in2, max_disp, 1.0 if leaky_slope is None else float(leaky_slope))


class _WarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flow, align_corners, use_mask):
        a, f = to_pixel_major(x), to_pixel_major(flow)
        out = torch.empty_like(a)
        k_warp(a, f, out, align_corners, use_mask)
        ctx.save_for_backward(a, f)
        ctx.cfg = (align_corners, use_mask)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad):
        a, f = ctx.saved_tensors
        g = to_pixel_major(grad)
        gx = torch.zeros_like(a) if ctx.needs_input_grad[0] else None
        gf = torch.empty_like(f) if ctx.needs_input_grad[1] else None
        k_warp_bwd(a, f, g, gx, gf, *ctx.cfg)
        return (gx.permute(0, 3, 1, 2) if gx is not None else None,
                gf.permute(0, 3, 1, 2) if gf is not None else None, None, None)


def warp(x, flow, align_corners=False, use_mask=True):
    """WarpingLayer_no_div (model/pwc_modules.py:184-207); use_mask=False = tools.torch_warp (utils/tools.py:1274-1304)."""
    _require_cuda(x, flow, lp_ok=True)
    if flow.shape[1] != 2 or flow.shape[0] != x.shape[0] or flow.shape[2:] != x.shape[2:]:
        raise RuntimeError("warp: flow must be [B,2,H,W] matching x")
    if x.dtype in _LP_DTYPES:
        if torch.is_grad_enabled() and (x.requires_grad or flow.requires_grad):
            raise RuntimeError("warp: the fp16 / bf16 storage variant has no backward; train in fp32")
        a, f = _pixel_major_any(x), _pixel_major_any(flow.float())
        out = torch.empty_like(a)
        k_warp_lp(a, f, out, bool(align_corners), use_mask)
        return out.permute(0, 3, 1, 2)
    return _WarpFn.apply(x, flow, bool(align_corners), use_mask)


class _NormFn(torch.autograd.Function):
    """normalize_features for one tensor; the backward differentiates through torch.mean / torch.var like autograd
    does on the reference (model/upflow.py:108-135)."""

    @staticmethod
    def forward(ctx, x):
        a = to_pixel_major(x)
        N, H, W, C = a.shape
        stats = torch.zeros(N, C, 2, dtype=torch.float64, device=x.device)
        k_stats(a, stats)
        out = torch.empty_like(a)
        k_norm_apply(a, stats, out)
        ctx.save_for_backward(a, stats)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad):
        a, stats = ctx.saved_tensors
        g = to_pixel_major(grad)
        gx = torch.empty_like(a)
        k_featnorm_bwd(a, stats, g, gx)
        return gx.permute(0, 3, 1, 2)


def normalize_features(x):
    """Per-image per-channel (x-mean)/sqrt(var+1e-16), unbiased var (model/upflow.py:108-135)."""
    _require_cuda(x)
    return _NormFn.apply(x)


class _ResizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h, w, scale):
        a = to_pixel_major(x)
        N, hi, wi, C = a.shape
        out = _new(N, h, w, C, x)
        k_resize(a, out, scale)
        ctx.cfg = (hi, wi, scale)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad):
        hi, wi, scale = ctx.cfg
        g = to_pixel_major(grad)
        gi = _new(g.shape[0], hi, wi, g.shape[3], grad)
        k_resize_bwd(g, gi, scale)
        return gi.permute(0, 3, 1, 2), None, None, None


def resize_bilinear(x, h, w, flow_rate=False):
    """F.interpolate(bilinear, align_corners=True) [+ u*=w/w_, v*=h/h_]  (model/pwc_modules.py:72-90)."""
    _require_cuda(x)
    a = to_pixel_major(x)
    N, hi, wi, C = a.shape
    if C > 4:
        raise RuntimeError("resize_bilinear: at most 4 channels (flows and masks)")
    scale = None
    if flow_rate:
        if C != 2:
            raise RuntimeError("if_rate=True needs a 2-channel flow")
        scale = (w / wi, h / hi)
    return _ResizeFn.apply(x, h, w, scale)


class _PointwiseFn(torch.autograd.Function):
    """sigmoid with its backward taken from the saved output."""

    @staticmethod
    def forward(ctx, x):
        a = to_pixel_major(x)
        out = torch.empty_like(a)
        k_pointwise(_ext.PW_SIGMOID, a, None, out)
        ctx.save_for_backward(out)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad):
        (out,) = ctx.saved_tensors
        g = to_pixel_major(grad)
        gx = torch.empty_like(out)
        k_pointwise(_ext.PW_SIGMOID_BWD, out, g, gx)
        return gx.permute(0, 3, 1, 2)


class _BlendFn(torch.autograd.Function):
    """out = w*(1-m) + f*m (model/upflow.py:88) on [B,2,H,W], [B,2,H,W], [B,1,H,W]."""

    @staticmethod
    def forward(ctx, w, f, m):
        wa, fa, ma = to_pixel_major(w), to_pixel_major(f), to_pixel_major(m)
        out = torch.empty_like(fa)
        npix = fa.shape[0] * fa.shape[1] * fa.shape[2]
        _ext.check(_lib().upf_blend_fwd(_p(wa), 2, _p(fa), 2, _p(ma), 1, _p(out), 2, npix, _stream()), "blend_fwd")
        ctx.save_for_backward(wa, fa, ma)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad):
        wa, fa, ma = ctx.saved_tensors
        g = to_pixel_major(grad)
        gw, gf, gm = torch.empty_like(wa), torch.empty_like(fa), torch.empty_like(ma)
        npix = fa.shape[0] * fa.shape[1] * fa.shape[2]
        _ext.check(_lib().upf_blend_bwd(_p(wa), 2, _p(fa), 2, _p(ma), 1, _p(g), 2, _p(gw), 2, _p(gf), 2, _p(gm), 1, npix,
                                        _stream()), "blend_bwd")
        return gw.permute(0, 3, 1, 2), gf.permute(0, 3, 1, 2), gm.permute(0, 3, 1, 2)


def sigmoid(x):
    _require_cuda(x)
    return _PointwiseFn.apply(x)


def sgu_blend(flow_init, inter, align_corners=False):
    """flow_up of sgu_model.forward (model/upflow.py:79-88) from flow_init [B,2,H,W] at output resolution and the dense
    block's raw 3-channel output ``inter`` (inter_flow u,v + mask logit), at the same or a lower resolution.
    Without autograd this is ONE fused kernel; when a gradient is needed it is the same arithmetic as a chain of
    differentiable pieces (sigmoid, the two upsamples, the un-masked warp, the blend), each with its own kernels."""
    _require_cuda(flow_init, inter)
    if torch.is_grad_enabled() and (flow_init.requires_grad or inter.requires_grad):
        H, W = flow_init.shape[2:]
        inter_flow, mask = inter[:, :2], sigmoid(inter[:, 2:3])
        if inter.shape[2:] != flow_init.shape[2:]:
            inter_flow = resize_bilinear(inter_flow, H, W, flow_rate=True)
            mask = resize_bilinear(mask, H, W)
        return _BlendFn.apply(warp(flow_init, inter_flow, align_corners, use_mask=False), flow_init, mask)
    f, i = to_pixel_major(flow_init), to_pixel_major(inter)
    out = torch.empty_like(f)
    k_sgu_blend(f, i, out, align_corners)
    return out.permute(0, 3, 1, 2)


# ---------------------------------------------------------------- loss terms of the training step
def _loss_buffers(like):
    ws = torch.empty(int(_lib().upf_loss_workspace_elems()), dtype=torch.float32, device=like.device)
    return ws, torch.empty(2, dtype=torch.float32, device=like.device)


class _RobustLossFn(torch.autograd.Function):
    """photo_loss_multi_type (model/upflow.py:268-290) as one reduction kernel forward, one elementwise kernel
    backward; the mask (if any) gets no gradient."""

    @staticmethod
    def forward(ctx, x, y, mask, kind, q):
        xa, ya = to_pixel_major(x), to_pixel_major(y)
        ma = mask.reshape(-1).contiguous() if mask is not None else None
        N, H, W, C = xa.shape
        ws, out = _loss_buffers(xa)
        _ext.check(_lib().upf_robust_loss_fwd(_p(xa), C, _p(ya), C, _p(ma), 1, _p(ws), _p(out), N * H * W, C, kind,
                                              float(q), _stream()), "robust_loss_fwd")
        ctx.save_for_backward(xa, ya, out, *([ma] if ma is not None else []))
        ctx.kind, ctx.q = kind, float(q)
        return out[0]

    @staticmethod
    def backward(ctx, grad):
        xa, ya, out = ctx.saved_tensors[:3]
        ma = ctx.saved_tensors[3] if len(ctx.saved_tensors) > 3 else None
        N, H, W, C = xa.shape
        g = grad.reshape(1).contiguous()
        _require_cuda(g)
        gx = torch.empty_like(xa) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(ya) if ctx.needs_input_grad[1] else None
        if gx is None and gy is None:
            return None, None, None, None, None
        _ext.check(_lib().upf_robust_loss_bwd(_p(xa), C, _p(ya), C, _p(ma), 1, _p(out), _p(g), _p(gx), C, _p(gy), C,
                                              N * H * W, C, ctx.kind, ctx.q, _stream()), "robust_loss_bwd")
        return (gx.permute(0, 3, 1, 2) if gx is not None else None, gy.permute(0, 3, 1, 2) if gy is not None else None,
                None, None, None)


def robust_loss(x, y, mask=None, kind="abs_robust", q=0.4):
    """The reference's photometric / distillation term (model/upflow.py:268-290) of x, y [N,C,H,W]: mean of
    (|x-y|+0.01)^q ('abs_robust'), ((x-y)^2+1e-6)^q ('charbonnier') or |x-y+1e-6| ('L1'); with mask [N,1,H,W]:
    sum(v*mask)/(sum(mask)+1e-6).  Returns a 0-dim tensor."""
    _require_cuda(x, y, mask)
    if x.shape != y.shape or (mask is not None and mask.numel() != x.shape[0] * x.shape[2] * x.shape[3]):
        raise ValueError("robust_loss: x %s, y %s, mask %s" % (tuple(x.shape), tuple(y.shape),
                                                                None if mask is None else tuple(mask.shape)))
    return _RobustLossFn.apply(x, y, mask, _ext.LOSS_KINDS[kind], q)


class _EdgeSmooth1Fn(torch.autograd.Function):
    """edge_aware_smoothness_order1 (model/upflow.py:198-218); gradient for `pred` only."""

    @staticmethod
    def forward(ctx, img, pred):
        ia, pa = to_pixel_major(img), to_pixel_major(pred)
        N, H, W, Cp = pa.shape
        ws, out = _loss_buffers(pa)
        _ext.check(_lib().upf_edge_smooth1_fwd(_p(ia), ia.shape[3], ia.shape[3], _p(pa), Cp, Cp, _p(ws), _p(out),
                                               N, H, W, _stream()), "edge_smooth1_fwd")
        ctx.save_for_backward(ia, pa)
        return out[0]

    @staticmethod
    def backward(ctx, grad):
        ia, pa = ctx.saved_tensors
        N, H, W, Cp = pa.shape
        g = grad.reshape(1).contiguous()
        _require_cuda(g)
        gp = torch.empty_like(pa)
        _ext.check(_lib().upf_edge_smooth1_bwd(_p(ia), ia.shape[3], ia.shape[3], _p(pa), Cp, Cp, _p(g), _p(gp), Cp,
                                               N, H, W, _stream()), "edge_smooth1_bwd")
        return None, gp.permute(0, 3, 1, 2)


def edge_smooth1(img, pred):
    """First-order edge-aware smoothness of pred [N,Cp,H,W] under image img [N,Ci,H,W] (model/upflow.py:198-218)."""
    _require_cuda(img, pred)
    if img.shape[0] != pred.shape[0] or img.shape[2:] != pred.shape[2:]:
        raise ValueError("edge_smooth1: img %s, pred %s" % (tuple(img.shape), tuple(pred.shape)))
    return _EdgeSmooth1Fn.apply(img, pred)


class _CensusLossFn(torch.autograd.Function):
    """census_loss_torch (utils/loss.py:51-91, abs_robust penalty): grey conversion + one patch-walking reduction
    kernel forward, one gather kernel backward; gradient for the second (warped) image only."""

    @staticmethod
    def forward(ctx, img1, img2, mask, q, max_distance):
        a, b = to_pixel_major(img1), to_pixel_major(img2)
        ma = mask.reshape(-1).contiguous() if mask is not None else None
        N, H, W, _ = a.shape
        grey = torch.empty(N * H * W, 2, dtype=torch.float32, device=a.device)
        dist = torch.empty(N * H * W, dtype=torch.float32, device=a.device)
        ws, out = _loss_buffers(a)
        _ext.check(_lib().upf_census_loss_fwd(_p(a), a.shape[3], _p(b), b.shape[3], _p(ma), 1, _p(grey), _p(dist), _p(ws),
                                              _p(out), N, H, W, int(max_distance), float(q), _stream()), "census_loss_fwd")
        ctx.save_for_backward(grey, dist, out, *([ma] if ma is not None else []))
        ctx.cfg = (N, H, W, float(q), int(max_distance))
        return out[0]

    @staticmethod
    def backward(ctx, grad):
        grey, dist, out = ctx.saved_tensors[:3]
        ma = ctx.saved_tensors[3] if len(ctx.saved_tensors) > 3 else None
        N, H, W, q, d = ctx.cfg
        if not ctx.needs_input_grad[1]:
            return None, None, None, None, None
        g = grad.reshape(1).contiguous()
        _require_cuda(g)
        gimg = torch.empty(N, H, W, 3, dtype=torch.float32, device=grey.device)
        _ext.check(_lib().upf_census_loss_bwd(_p(grey), _p(dist), _p(ma), 1, _p(out), _p(g), _p(gimg), 3, N, H, W, d, q,
                                              _stream()), "census_loss_bwd")
        return None, gimg.permute(0, 3, 1, 2), None, None, None


def census_loss(img1, img1_warp, mask=None, q=0.4, max_distance=3):
    """The reference's census term (utils/loss.py:51-91, abs_robust penalty) of two RGB images [N,3,H,W]; with mask
    [N,1,H,W] (if_use_occ): masked, border of max_distance pixels excluded, normalised by 2*sum(mask)+1e-6; without:
    the mean over all pixels.  Differentiable in img1_warp."""
    _require_cuda(img1, img1_warp, mask)
    if img1.shape != img1_warp.shape or img1.shape[1] != 3 or \
            (mask is not None and mask.numel() != img1.shape[0] * img1.shape[2] * img1.shape[3]):
        raise ValueError("census_loss: img1 %s, img1_warp %s, mask %s" % (tuple(img1.shape), tuple(img1_warp.shape),
                                                                         None if mask is None else tuple(mask.shape)))
    return _CensusLossFn.apply(img1, img1_warp, mask, q, max_distance)


class _BoundaryWarpFn(torch.autograd.Function):
    """tools.boundary_dilated_warp.warp_im (utils/tools.py:350-499): one gather kernel forward, one backward (gradient
    of the flow only)."""

    @staticmethod
    def forward(ctx, image, flow, start):
        im, fl = to_pixel_major(image), to_pixel_major(flow)
        st = start.reshape(-1, 2).float().contiguous()
        N, Hf, Wf, C = im.shape
        _, h, w, _ = fl.shape
        out = torch.empty(N, h, w, C, dtype=torch.float32, device=im.device)
        _ext.check(_lib().upf_boundary_warp_fwd(_p(im), C, C, Hf, Wf, _p(fl), 2, _p(st), _p(out), C, N, h, w, _stream()),
                   "boundary_warp_fwd")
        ctx.save_for_backward(im, fl, st)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad):
        im, fl, st = ctx.saved_tensors
        if not ctx.needs_input_grad[1]:
            return None, None, None
        N, Hf, Wf, C = im.shape
        _, h, w, _ = fl.shape
        g = to_pixel_major(grad)
        gf = torch.empty(N, h, w, 2, dtype=torch.float32, device=im.device)
        _ext.check(_lib().upf_boundary_warp_bwd(_p(im), C, C, Hf, Wf, _p(fl), 2, _p(st), _p(g), C, _p(gf), 2, N, h, w,
                                                _stream()), "boundary_warp_bwd")
        return None, gf.permute(0, 3, 1, 2), None


def boundary_warp(image, flow, start):
    """The reference's boundary-dilated warp (utils/tools.py:350-499): image [N,C,Hf,Wf] (the un-cropped frame), flow
    [N,2,h,w] (of the crop), start [N,2,1,1] (the crop's origin inside the frame, x then y).  Differentiable in flow."""
    _require_cuda(image, flow)
    if not start.is_cuda:
        raise RuntimeError("boundary_warp: start must be a CUDA tensor")
    if image.shape[0] != flow.shape[0] or flow.shape[1] != 2 or start.numel() != 2 * flow.shape[0]:
        raise ValueError("boundary_warp: image %s, flow %s, start %s" % (tuple(image.shape), tuple(flow.shape), tuple(start.shape)))
    return _BoundaryWarpFn.apply(image, flow, start)


class _ConvFn(torch.autograd.Function):
    """conv() of model/pwc_modules.py:10-31 with autograd: the input gradient is the forward kernel on the flipped,
    transposed weights (on the zero-interleaved gradient for stride 2), the weight / bias gradient is
    upf_conv2d_wgrad, the LeakyReLU derivative comes from the saved output."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, dilation, slope, precision):
        Cout, Cin, ks, _ = weight.shape
        tc = precision == _ext.CONV_TF32
        w_simt, w_tc = pack_conv_weight(weight, tc=tc, tc_only=True)
        a = Slice(to_pixel_major(x, ld=(Cin + 3) // 4 * 4), 0, Cin) if (tc and Cin % 4) else Slice(to_pixel_major(x))
        pad = ((ks - 1) * dilation) // 2
        Ho = (a.H + 2 * pad - dilation * (ks - 1) - 1) // stride + 1
        Wo = (a.W + 2 * pad - dilation * (ks - 1) - 1) // stride + 1
        out = _new(a.N, Ho, Wo, Cout, x)
        k_conv(a, w_tc if tc else w_simt, bias.detach().float().contiguous(), out, ks, stride, dilation, slope, None,
               precision | (_ext.CONV_ROUND_OUT if tc and slope != 1.0 else 0))     # hidden activation: TF32-rounded for its consumers
        ctx.save_for_backward(a.buf, out, weight)
        ctx.cfg = (Cin, stride, dilation, slope, precision)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad):
        abuf, out, weight = ctx.saved_tensors
        Cin, stride, dilation, slope, precision = ctx.cfg
        gxb, gw, gb = _conv_backward(Slice(abuf, 0, Cin), Slice(out), weight, Slice(to_pixel_major(grad)), stride, dilation,
                                     slope, precision, ctx.needs_input_grad[1] or ctx.needs_input_grad[2],
                                     ctx.needs_input_grad[0])
        return gxb.permute(0, 3, 1, 2) if gxb is not None else None, gw, gb, None, None, None, None


def _conv_backward(a, act, weight, g, stride, dilation, slope, precision, need_w, need_x):
    """Backward of one conv() call on pixel-major slices.  a: the input slice; act: the saved OUTPUT slice (LeakyReLU
    derivative); g: gradient wrt the output.  Returns (gx buffer or None, grad weight [Cout,Cin,k,k] or None, grad
    bias or None)."""
    Cout, Cin, ks, _ = weight.shape
    tc = precision == _ext.CONV_TF32
    N, Ho, Wo = g.N, g.H, g.W
    ldg = (Cout + 3) // 4 * 4
    if slope != 1.0 or ldg != Cout or g.ld != Cout:
        gp = torch.zeros(N, Ho, Wo, ldg, dtype=torch.float32, device=g.buf.device) if ldg != Cout else \
            torch.empty(N, Ho, Wo, ldg, dtype=torch.float32, device=g.buf.device)
        if slope != 1.0:
            k_pointwise(_ext.PW_LRELU_BWD, act, g, Slice(gp, 0, Cout), slope)
        else:
            k_copy(g, Slice(gp, 0, Cout))
    else:
        gp = g.buf
    gps = Slice(gp, 0, Cout)
    gxb = gw = gb = None
    if need_w:
        gw_t, gb = k_conv_wgrad(a, gps, ks, stride, dilation, want_bias=True, tensor_cores=tc)
        gw = gw_t.reshape(ks, ks, Cin, Cout).permute(3, 2, 0, 1).contiguous()
    if need_x:
        w_simt, w_tc = pack_conv_weight(weight, tc=tc, flip_transpose=True, tc_only=True)   # flipped taps, [Cin, Cout] roles exchanged
        if stride == 1:
            src = gps
        else:
            # adjoint of a strided convolution = stride-1 convolution of the zero-interleaved gradient
            up = torch.zeros(N, a.H, a.W, ldg, dtype=torch.float32, device=g.buf.device)
            up[:, ::stride, ::stride, :][:, :Ho, :Wo] = gp
            src = Slice(up, 0, Cout)
        gxb = _new(a.N, a.H, a.W, Cin, g.buf)
        k_conv(src, w_tc if tc else w_simt, _zero_bias(Cin, g.buf.device), Slice(gxb), ks, 1, dilation, 1.0, None, precision)
    return gxb, gw, gb


class _DenseBlockFn(torch.autograd.Function):
    """FlowEstimatorDense_v2.forward (model/pwc_modules.py:279-286) and the SGU dense block (model/upflow.py:52-60) as
    ONE autograd node on an append-only pixel-major buffer [conv_n | ... | conv1 | x]: convolution i reads the suffix
    that exists so far and writes its output in front of it, so the reference's torch.cat calls (and, backward, the
    slice / add kernels autograd makes of them) disappear.  Backward: see there.  params = (w1, b1, ..., wn, bn, w_last, b_last).  Returns (x_n, conv_last(x_n)) in NCHW views."""

    @staticmethod
    def forward(ctx, x, f_channels, precision, *params):
        B, C, H, W = x.shape
        n = len(f_channels)
        total = C + sum(f_channels)
        ld = (total + 3) // 4 * 4
        buf = torch.zeros(B, H, W, ld, dtype=torch.float32, device=x.device)
        lo = total - C
        k_copy(Slice(to_pixel_major(x)), Slice(buf, lo, C))
        los, precs = [], []
        for i in range(n + 1):
            weight, bias = params[2 * i], params[2 * i + 1]
            prec = _ext.CONV_FP32 if (lo % 4) else precision      # the tensor-core kernels read 16-byte aligned slices
            tc = prec == _ext.CONV_TF32
            w_simt, w_tc = pack_conv_weight(weight, tc=tc, tc_only=True)
            bvec = bias.detach().float().contiguous()
            los.append(lo)
            precs.append(prec)
            if i < n:
                c = f_channels[i]
                k_conv(Slice(buf, lo, total - lo), w_tc if tc else w_simt, bvec, Slice(buf, lo - c, c), 3, 1, 1, LRELU_SLOPE,
                       None, prec | (_ext.CONV_ROUND_OUT if tc else 0))
                lo -= c
            else:
                cout = weight.shape[0]
                out = torch.empty(B, H, W, cout, dtype=torch.float32, device=x.device)
                k_conv(Slice(buf, 0, total), w_tc if tc else w_simt, bvec, Slice(out), 3, 1, 1, 1.0, None, prec)
        ctx.save_for_backward(buf, *params[0::2])
        ctx.cfg = (C, tuple(f_channels), total, tuple(los), tuple(precs), precision)
        return buf[..., :total].permute(0, 3, 1, 2), out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g_x5, g_out):
        # Each channel block of the buffer is read by every LATER convolution, so its gradient is a sum over those
        # convolutions.  Instead of accumulating n nested input gradients (every convolution writing the whole suffix
        # it read: 3x the bytes, and the accumulate is an uncoalesced read in the tensor-core epilogue), the
        # pre-activation output gradients live in a second append-only buffer GP = [gp_1 | ... | gp_n | gp_last], and
        # the gradient of block i is ONE convolution of the suffix behind gp_i with the consumers' weight slices
        # concatenated along K (flipped / transposed), written once, with the incoming x_n gradient as its residual.
        buf, *weights = ctx.saved_tensors
        C, f, total, los, precs, precision = ctx.cfg
        n = len(f)
        B, H, W, _ = buf.shape
        cout = weights[n].shape[0]
        need = ctx.needs_input_grad
        offs = [sum(f[:i]) for i in range(n + 1)]              # offs[i]: where gp of convolution i starts (i == n: conv_last)
        gp_w = offs[n] + cout
        GP = torch.zeros(B, H, W, (gp_w + 3) // 4 * 4, dtype=torch.float32, device=buf.device)
        g5 = to_pixel_major(g_x5) if g_x5 is not None else None
        if g_out is not None:
            k_copy(Slice(to_pixel_major(g_out)), Slice(GP, offs[n], cout))
        grads = [None] * (2 * n + 2)

        # the convolutions' inputs are nested suffixes of buf: transpose it once for all the tensor-core weight gradients
        xt = k_wgrad_planar_input(Slice(buf, 0, total), 3, 1) if any(p == _ext.CONV_TF32 for p in precs) else None

        def param_grads(i, c):
            if need[3 + 2 * i] or need[4 + 2 * i]:
                gw_t, gb = k_conv_wgrad(Slice(buf, los[i], total - los[i]), Slice(GP, offs[i], c), 3, 1, 1, want_bias=True,
                                        tensor_cores=precs[i] == _ext.CONV_TF32,
                                        planar=(xt[0], xt[1], los[i]) if xt is not None else None)
                grads[2 * i] = gw_t.reshape(3, 3, total - los[i], c).permute(3, 2, 0, 1).contiguous()
                grads[2 * i + 1] = gb

        def block_grad(first, s, c):
            """Gradient of buffer channels [s, s+c): consumers are convolutions first..n."""
            wcat = torch.cat([weights[m][:, s - los[m]:s - los[m] + c] for m in range(first, n + 1)], dim=0)
            prec = _ext.CONV_FP32 if (offs[first] % 4) else precision
            tc = prec == _ext.CONV_TF32
            w_simt, w_tc = pack_conv_weight(wcat, tc=tc, flip_transpose=True, tc_only=True)
            gblk = _new(B, H, W, c, buf)
            k_conv(Slice(GP, offs[first], gp_w - offs[first]), w_tc if tc else w_simt, _zero_bias(c, buf.device), Slice(gblk), 3, 1,
                   1, 1.0, Slice(g5, s, c) if g5 is not None else None, prec)
            return gblk

        param_grads(n, cout)
        for i in range(n - 1, -1, -1):
            c, s = f[i], los[i] - f[i]
            gblk = block_grad(i + 1, s, c)
            k_pointwise(_ext.PW_LRELU_BWD, Slice(buf, s, c), Slice(gblk), Slice(GP, offs[i], c), LRELU_SLOPE)
            param_grads(i, c)
        gx = block_grad(0, total - C, C).permute(0, 3, 1, 2) if need[0] else None
        return (gx, None, None) + tuple(grads)


def dense_block(x, f_channels, params, precision=_ext.CONV_FP32):
    """The dense estimator block with autograd (training path): x [B,C,H,W], params = (w1, b1, ..., wn, bn, w_last,
    b_last) as nn.Conv2d parameters.  Returns (x_n [B, C+sum(f), H, W] in the reference's order, conv_last(x_n))."""
    _require_cuda(x, *params)
    if len(params) != 2 * len(f_channels) + 2:
        raise ValueError("dense_block: %d parameters for %d convolutions" % (len(params), len(f_channels) + 1))
    return _DenseBlockFn.apply(x, tuple(int(c) for c in f_channels), int(precision), *params)


def conv2d_autograd(x, weight, bias, stride=1, dilation=1, slope=LRELU_SLOPE, precision=_ext.CONV_FP32):
    """conv() + LeakyReLU with gradients wrt x, weight and bias (training path)."""
    _require_cuda(x, weight)
    return _ConvFn.apply(x, weight, bias, int(stride), int(dilation), float(slope), int(precision))


def conv2d(x, w_packed, bias, cout, ksize, stride=1, dilation=1, slope=LRELU_SLOPE, precision=_ext.CONV_FP32):
    """conv() of model/pwc_modules.py:10-31 on an NCHW tensor with pre-packed weights (pack_conv_weight)."""
    _require_cuda(x)
    C = x.shape[1]
    if (precision & 0xFF) == _ext.CONV_TF32 and C % 4:
        a = Slice(to_pixel_major(x, ld=(C + 3) // 4 * 4), 0, C)     # TMA needs a 16-byte pixel pitch
    else:
        a = Slice(to_pixel_major(x))
    pad = ((ksize - 1) * dilation) // 2
    Ho = (a.H + 2 * pad - dilation * (ksize - 1) - 1) // stride + 1
    Wo = (a.W + 2 * pad - dilation * (ksize - 1) - 1) // stride + 1
    out = _new(a.N, Ho, Wo, cout, x)
    k_conv(a, w_packed, bias, out, ksize, stride, dilation, slope, None, precision)
    return out.permute(0, 3, 1, 2)
