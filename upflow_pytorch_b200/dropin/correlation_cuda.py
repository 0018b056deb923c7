"""`correlation_cuda` -- the module model/correlation_package/correlation.py imports (line 4) -- as a ctypes stub over
libupflow_b200.so, with the pybind module's two functions and calling conventions
(model/correlation_package/correlation_cuda.cc:10-16 forward, :89-96 backward, :169-172 the module definition):

  * the caller passes EMPTY tensors (`input1.new()`, correlation.py:22-24, :35-39); the callee `resize_`s them
    (correlation_cuda.cc:36-42, :107-115) and fills them; rbot1 / rbot2 (the reference's padded NHWC scratch copies)
    are not needed by this kernel and are left empty;
  * work is enqueued on the current CUDA stream (correlation_cuda.cc:76, :158), no synchronisation;
  * returns 1; a failed launch raises RuntimeError("CUDA call failed ...") like AT_ERROR (:81-83).

This file is what INTEGRATION.md section 2 tells a maintainer of the reference to drop next to correlation.py; it is
standalone (ctypes + torch only).  `upflow_pytorch_b200.install_dropin()` also puts it on sys.path, and
tests/test_gpu_kernels.py::test_correlation_cuda_stub_* drives it through a static-method restatement of the reference's
CorrelationFunction.  Supported configuration: the one UPFlow uses (model/upflow.py:561) -- kernel_size 1,
stride1 = stride2 = 1, pad_size == max_displacement (1..6), fp32.
"""
import ctypes
import os

import torch

_LIB_PATH = os.environ.get("UPFLOW_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "lib",
                                                               "libupflow_b200.so")
_lib = None
_P, _I, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError("correlation_cuda: %s not found (build it with `python -m upflow_pytorch_b200.build`)" % _LIB_PATH)
        lib = ctypes.CDLL(_LIB_PATH)
        lib.upf_corr_lrelu_fwd.argtypes = [_P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _F, _I, _P]
        lib.upf_corr_lrelu_fwd_planar.argtypes = [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _P]
        lib.upf_corr_lrelu_bwd.argtypes = [_P, _I, _P, _I, _P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _F, _P]
        lib.upf_nchw_to_nhwc.argtypes = [_P, _P, _I, _I, _I, _I, _I, _P]
        lib.upf_nhwc_to_nchw.argtypes = [_P, _I, _P, _I, _I, _I, _I, _P]
        lib.upf_last_error.restype = ctypes.c_char_p
        _lib = lib
    return _lib


def _chk(rc):
    if rc:
        raise RuntimeError("CUDA call failed: " + _load().upf_last_error().decode())    # AT_ERROR, correlation_cuda.cc:81-83


def _check_config(in1, in2, pad, k, maxd, s1, s2):
    if not (in1.is_cuda and in2.is_cuda):
        raise RuntimeError("correlation_cuda: CUDA tensors expected")
    if in1.dtype != torch.float32 or in2.dtype != torch.float32:
        raise RuntimeError("correlation_cuda (upflow_b200): fp32 only")
    if in1.shape != in2.shape or in1.dim() != 4:
        raise RuntimeError("correlation_cuda: inputs must be two [B,C,H,W] tensors of the same shape")
    if k != 1 or s1 != 1 or s2 != 1 or pad != maxd or not 1 <= maxd <= 6:
        raise RuntimeError("correlation_cuda (upflow_b200): kernel_size=1, stride1=stride2=1, pad_size==max_displacement "
                           "in 1..6 (the configuration of model/upflow.py:561); got k=%d s1=%d s2=%d pad=%d maxd=%d"
                           % (k, s1, s2, pad, maxd))


def _nhwc(t, st):
    """the kernels are pixel-major; the reference hands over NCHW tensors"""
    B, C, H, W = t.shape
    v = t.permute(0, 2, 3, 1)
    if v.is_contiguous():                    # channels_last already
        return v
    o = torch.empty(B, H, W, C, dtype=torch.float32, device=t.device)
    _chk(_load().upf_nchw_to_nhwc(t.contiguous().data_ptr(), o.data_ptr(), C, B, C, H, W, st))
    return o


def forward(input1, input2, rbot1, rbot2, output, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply):
    """correlation_cuda.cc:10-87.  output [B,(2d+1)^2,H,W] = mean_c in1[c,y,x] * in2[c,y+dy,x+dx], zero padded."""
    _check_config(input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2)
    lib = _load()
    with torch.cuda.device_of(input1):
        st = torch.cuda.current_stream().cuda_stream
        B, C, H, W = input1.shape
        D2 = (2 * max_displacement + 1) ** 2
        if (max_displacement in (3, 4) and W % 4 == 0 and H * W >= 16384 and input1.is_contiguous() and input2.is_contiguous()
                and input1.data_ptr() % 16 == 0 and input2.data_ptr() % 16 == 0):
            # the reference's own layout end to end: NCHW operands and NCHW result through TMA, no conversion (corr_planar.cu)
            output.resize_(B, D2, H, W)
            pit = lambda c: (ctypes.c_longlong * 3)(W, H * W, c * H * W)
            _chk(lib.upf_corr_lrelu_fwd_planar(input1.data_ptr(), pit(C), input2.data_ptr(), pit(C), output.data_ptr(), pit(D2),
                                               B, H, W, C, max_displacement, 0, 1.0, 0, st))
            return 1
        a, b = _nhwc(input1, st), _nhwc(input2, st)
        o = torch.empty(B, H, W, D2, dtype=torch.float32, device=input1.device)
        _chk(lib.upf_corr_lrelu_fwd(a.data_ptr(), C, b.data_ptr(), C, o.data_ptr(), D2, B, H, W, C, max_displacement,
                                    None, None, 0, 1.0, 0, st))              # slope 1.0: no activation, flags 0: exact fp32
        output.resize_(B, D2, H, W)                                           # the callee sizes the caller's tensor (:36-42)
        _chk(lib.upf_nhwc_to_nchw(o.data_ptr(), D2, output.data_ptr(), B, D2, H, W, st))
    return 1


def backward(input1, input2, rbot1, rbot2, grad_output, grad_input1, grad_input2, pad_size, kernel_size, max_displacement,
             stride1, stride2, corr_multiply):
    """correlation_cuda.cc:89-167."""
    _check_config(input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2)
    lib = _load()
    with torch.cuda.device_of(input1):
        st = torch.cuda.current_stream().cuda_stream
        B, C, H, W = input1.shape
        D2 = (2 * max_displacement + 1) ** 2
        a, b, g = _nhwc(input1, st), _nhwc(input2, st), _nhwc(grad_output.float(), st)
        g1, g2 = torch.empty_like(a), torch.empty_like(b)
        _chk(lib.upf_corr_lrelu_bwd(a.data_ptr(), C, b.data_ptr(), C, None, 0, g.data_ptr(), D2, g1.data_ptr(), C,
                                    g2.data_ptr(), C, B, H, W, C, max_displacement, 1.0, st))
        for src, dst in ((g1, grad_input1), (g2, grad_input2)):
            dst.resize_(B, C, H, W)
            _chk(lib.upf_nhwc_to_nchw(src.data_ptr(), C, dst.data_ptr(), B, C, H, W, st))
    return 1
