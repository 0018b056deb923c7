"""Parity of every CUDA kernel, called through the C ABI (ctypes), against the
CPU oracle and the golden vectors produced by the reference.  -m gpu."""
import math

import pytest
import torch

from oracle import cpu_oracle as O
from oracle import ref_port as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def upf():
    from upflow_pytorch_b200 import _ext, ops
    _ext.load()
    return ops


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _cuda(t):
    return t.cuda()


def _regen(seed, shape):
    return torch.randn(shape, generator=_g(seed))


def _rn_tf32(t):
    """nearest TF32 value, ties away from zero (cvt.rna.tf32.f32): what the weight packers and the UPF_FLAG_ROUND_TF32
    producers store; tcgen05 kind::tf32 itself truncates whatever it is given"""
    return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


# ------------------------------------------------------------------ correlation
def test_corr_golden(upf, golden):
    """BASELINE config 1 (1x32x64x64, d=4) and the ragged / d=2 / d=6 / C=196 cases, vs Corr_pyTorch outputs."""
    for c in golden("corr"):
        f1 = c["f1"] if c["f1"] is not None else _regen(c["seed"], c["shape"])
        f2 = c["f2"] if c["f2"] is not None else _regen(c["seed"] + 100, c["shape"])
        out = upf.correlation(_cuda(f1), _cuda(f2), c["d"]).cpu()
        assert out.shape == c["out"].shape
        err = (out - c["out"]).abs().max().item()
        assert err <= 1e-5, (c["shape"], c["d"], err)          # fp32 sum of C products
        lr = upf.correlation(_cuda(f1), _cuda(f2), c["d"], leaky_slope=0.1).cpu()
        ref = torch.nn.functional.leaky_relu(c["out"], 0.1)
        assert (lr - ref).abs().max().item() <= 1e-5


@pytest.mark.parametrize("layout", ["nchw", "channels_last"])
@pytest.mark.parametrize("shape,d", [((2, 32, 47, 156), 4), ((1, 64, 24, 78), 4), ((2, 128, 12, 39), 4),
                                     ((1, 96, 33, 70), 4), ((1, 32, 40, 100), 2), ((1, 32, 37, 65), 6),
                                     ((1, 7, 9, 33), 3), ((1, 32, 3, 2), 4),
                                     # > 16384 pixels: the tiled shared-memory kernel (smaller images take the
                                     # one-CTA-per-pixel kernel)
                                     ((1, 32, 130, 140), 4), ((2, 64, 100, 100), 4), ((1, 96, 128, 131), 2),
                                     ((1, 32, 129, 130), 6), ((1, 7, 129, 130), 3), ((1, 196, 128, 129), 4),
                                     ((1, 32, 150, 120), 1), ((1, 36, 140, 120), 5)])
def test_corr_vs_oracle(upf, shape, d, layout):
    f1, f2 = _regen(1, shape), _regen(2, shape)
    a, b = _cuda(f1), _cuda(f2)
    if layout == "channels_last":
        a, b = a.contiguous(memory_format=torch.channels_last), b.contiguous(memory_format=torch.channels_last)
    out = upf.correlation(a, b, d).cpu()
    ref = O.correlation(f1.double(), f2.double(), d).float()
    assert (out - ref).abs().max().item() <= 1e-5


@pytest.mark.parametrize("shape,d", [((2, 32, 64, 96), 4), ((1, 32, 270, 480), 4), ((1, 7, 33, 36), 4), ((2, 196, 12, 40), 4),
                                     ((1, 32, 50, 100), 1), ((1, 20, 70, 132), 2), ((1, 32, 45, 64), 3), ((1, 33, 31, 124), 3),
                                     ((1, 32, 29, 248), 2), ((1, 4, 8, 4), 4), ((3, 64, 47, 156), 4), ((1, 5, 13, 244), 1)])
def test_corr_planar_vs_oracle(upf, shape, d):
    """corr_planar.cu: planar (NCHW) operands by TMA, outer-product register blocking, planar output by TMA stores --
    vs the fp64 oracle; every displacement range it serves (d <= 4), ragged tiles, widths around the 120-column tile, channel counts that are not a multiple of the ring chunk,
    LeakyReLU, the batch shift, and pitched (non-contiguous) operands and output."""
    from upflow_pytorch_b200 import ops
    f1, f2 = _regen(11, shape), _regen(12, shape)
    N, C, H, W = shape
    n = (2 * d + 1) ** 2
    a, b = _cuda(f1), _cuda(f2)
    out = torch.full((N, n, H, W), float("nan"), device="cuda")
    ops.k_corr_planar(a, b, out, d, slope=0.1)
    assert upf.last_kernel() == "corr_planar"
    ref = O.correlation(f1.double(), f2.double(), d, 0.1).float()
    assert (out.cpu() - ref).abs().max().item() <= 1e-5
    # the public operator takes the same route for NCHW inputs (no layout conversion)
    out2 = upf.correlation(a, b, d, leaky_slope=0.1)
    if H * W >= ops.PLANAR_CORR_MIN_PIXELS:
        assert upf.last_kernel() == "corr_planar" and out2.is_contiguous()
        assert torch.equal(out2, out)
    # pitched operands (a window of a larger buffer: row / plane / image pitches all differ from the dense ones) and a
    # pitched output, image n against image (n + 1) % N
    big1 = torch.zeros(N, C + 1, H + 2, W + 8, device="cuda"); big2 = torch.zeros_like(big1)
    bigo = torch.full((N, n + 2, H + 1, W + 8), 7.0, device="cuda")
    v1, v2, vo = big1[:, :C, 1:H + 1, 4:W + 4], big2[:, :C, 1:H + 1, 4:W + 4], bigo[:, 1:n + 1, :H, 4:W + 4]
    v1.copy_(a); v2.copy_(b)
    ops.k_corr_planar(v1, v2, vo, d, f2_shift=1 % N, slope=1.0)
    ref = O.correlation(f1.double(), f2.double()[[(i + 1) % N for i in range(N)]], d).float()
    assert (vo.cpu() - ref).abs().max().item() <= 1e-5
    bigo[:, 1:n + 1, :H, 4:W + 4] = 7.0
    assert (bigo == 7.0).all()          # nothing outside the window was written


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape,d", [((2, 32, 40, 100), 4), ((1, 64, 24, 78), 4), ((1, 7, 9, 33), 3), ((1, 32, 37, 65), 6),
                                     ((1, 36, 30, 40), 2), ((2, 196, 6, 20), 4)])
def test_corr_warp_low_precision_storage(upf, shape, d, dtype):
    """SURVEY 8f rank 4 (the reference dispatches its correlation on Half, correlation_cuda_kernel.cu:352): fp16 / bf16
    STORAGE of the correlation's operands and result and of the warp's source and result, fp32 arithmetic.  Oracle: the
    fp64 routine on the SAME rounded operands; stated tolerance: one rounding of the result to the storage type
    (2^-11 relative for fp16, 2^-8 for bf16) plus the fp32 accumulation error."""
    from upflow_pytorch_b200 import ops
    eps = 2.0 ** -11 if dtype == torch.float16 else 2.0 ** -8
    f1, f2 = _regen(21, shape).to(dtype), _regen(22, shape).to(dtype)
    out = upf.correlation(_cuda(f1), _cuda(f2), d, leaky_slope=0.1)
    assert out.dtype == dtype and upf.last_kernel() == "corr_fwd"
    ref = O.correlation(f1.double(), f2.double(), d, 0.1)
    err = (out.cpu().double() - ref).abs()
    assert (err <= eps * ref.abs() + 2e-6).all(), err.max().item()
    # channels_last input (no layout copy), a channel slice of a wider buffer as the output, the batch shift
    N, C, H, W = shape
    a = _cuda(f1).contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)
    b = _cuda(f2).contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)
    n = (2 * d + 1) ** 2
    X = torch.full((N, H, W, n + 7), 3.0, device="cuda", dtype=dtype)
    ops.k_corr_lp(a, b, X[..., :n], d, f2_shift=N - 1, slope=1.0)
    ref = O.correlation(f1.double(), f2.double()[[(i + N - 1) % N for i in range(N)]], d)
    err = (X[..., :n].permute(0, 3, 1, 2).cpu().double() - ref).abs()
    assert (err <= eps * ref.abs() + 2e-6).all(), err.max().item()
    assert (X[..., n:] == 3.0).all()
    # warp: fp32 flow, half source and result; the mask is the fp32 kernel's (same coordinates), values within one rounding
    flow = (torch.randn(N, 2, H, W, generator=torch.Generator().manual_seed(5)) * 3.0)
    w_lp = upf.warp(_cuda(f1), _cuda(flow))
    w_32 = upf.warp(_cuda(f1.float()), _cuda(flow))
    assert w_lp.dtype == dtype
    assert torch.equal(w_lp == 0, (w_32 == 0) | (w_lp == 0)) and torch.equal((w_32 == 0), (w_32 == 0) & (w_lp == 0))
    err = (w_lp.float() - w_32).abs()
    assert (err <= eps * w_32.abs() + 1e-7).all(), err.max().item()
    with pytest.raises(RuntimeError):
        upf.correlation(_cuda(f1), _cuda(f2.float()), d)


def test_corr_fused_norm_matches_two_step(upf):
    """normalize_features + correlation + LeakyReLU fused (what the engine runs) vs the oracle chain."""
    from upflow_pytorch_b200.ops import Slice
    shape = (2, 64, 24, 78)
    f1 = _regen(3, shape) * 2 + 0.5
    f2 = _regen(4, shape).relu() * 1.5
    a, b = upf.to_pixel_major(_cuda(f1)), upf.to_pixel_major(_cuda(f2))
    s1 = torch.zeros(2, 64, 2, dtype=torch.float64, device="cuda")
    s2 = torch.zeros_like(s1)
    upf.k_stats(a, s1)
    upf.k_stats(b, s2)
    out = torch.empty(2, 24, 78, 81, device="cuda")
    upf.k_corr(a, b, out, 4, s1, s2, slope=0.1)
    ref = O.correlation(O.normalize_features(f1.double()), O.normalize_features(f2.double()), 4, 0.1).float()
    assert (out.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() <= 2e-5
    # batch shift: image n of f1 against image (n+1)%2 of f2
    upf.k_corr(a, b, out, 4, s1, s2, f2_shift=1, slope=0.1)
    ref = O.correlation(O.normalize_features(f1.double()), O.normalize_features(f2.double())[[1, 0]], 4, 0.1).float()
    assert (out.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() <= 2e-5


def test_corr_into_channel_slice(upf):
    """writes exactly its 81 channels of a 576-wide buffer"""
    from upflow_pytorch_b200.ops import Slice
    f1, f2 = _regen(5, (1, 32, 20, 45)), _regen(6, (1, 32, 20, 45))
    X = torch.full((1, 20, 45, 576), 7.0, device="cuda")
    upf.k_corr(upf.to_pixel_major(_cuda(f1)), upf.to_pixel_major(_cuda(f2)), Slice(X, 0, 81), 4, slope=1.0)
    ref = O.correlation(f1, f2, 4)
    assert (X[..., :81].permute(0, 3, 1, 2).cpu() - ref).abs().max().item() <= 1e-5
    assert (X[..., 81:] == 7.0).all()


def test_corr_backward(upf):
    shape, d = (2, 12, 9, 14), 3
    f1 = _regen(7, shape).cuda().requires_grad_()
    f2 = _regen(8, shape).cuda().requires_grad_()
    go = _regen(9, (2, 49, 9, 14))
    upf.correlation(f1, f2, d).backward(go.cuda())
    g1, g2 = O.correlation_backward(f1.detach().cpu().double(), f2.detach().cpu().double(), go.double(), d)
    assert (f1.grad.cpu() - g1.float()).abs().max().item() <= 1e-5
    assert (f2.grad.cpu() - g2.float()).abs().max().item() <= 1e-5


# ------------------------------------------------------------------ warp
def test_warp_golden_mask_bit_exact(upf, golden):
    for w in golden("warp"):
        out = upf.warp(_cuda(w["x"]), _cuda(w["flow"])).cpu()
        ref = w["out"]
        assert torch.equal((ref == 0).all(1), (out == 0).all(1)), w["kind"]     # mask >= 1.0, pixel for pixel
        tol = 2e-6 * max(1.0, ref.abs().max().item())
        assert (out - ref).abs().max().item() <= tol
        nm = upf.warp(_cuda(w["x"]), _cuda(w["flow"]), use_mask=False).cpu()
        assert (nm - w["out_nomask"]).abs().max().item() <= tol


@pytest.mark.parametrize("hw", [(47, 156), (94, 311), (375, 1242), (1, 1), (2, 7)])
@pytest.mark.parametrize("C", [32, 196, 3])
def test_warp_vs_oracle_kitti_sizes(upf, hw, C):
    H, W = hw
    if H * W * C > 375 * 1242 * 32:
        pytest.skip("large")
    x = _regen(11, (1, C, H, W))
    for fl in (_regen(12, (1, 2, H, W)) * 3, torch.randint(-3, 4, (1, 2, H, W), generator=_g(13)).float()):
        out = upf.warp(_cuda(x), _cuda(fl)).cpu()
        ref = O.warp_mask(x, fl)
        assert torch.equal((ref == 0).all(1), (out == 0).all(1))
        assert (out - ref).abs().max().item() <= 4e-6
        out = upf.warp(_cuda(x), _cuda(fl), align_corners=True).cpu()
        ref = O.warp_mask(x, fl, align_corners=True)
        assert torch.equal((ref == 0).all(1), (out == 0).all(1))
        assert (out - ref).abs().max().item() <= 4e-6


def test_warp_fused_moments(upf):
    from upflow_pytorch_b200.ops import Slice
    x = _regen(14, (2, 96, 24, 78)).relu()
    fl = _regen(15, (2, 2, 24, 78)) * 2
    a, f = upf.to_pixel_major(_cuda(x)), upf.to_pixel_major(_cuda(fl))
    out = torch.empty_like(a)
    st = torch.zeros(2, 96, 2, dtype=torch.float64, device="cuda")
    upf.k_warp(a, f, out, False, True, x_shift=1, stats=st)
    ref = O.warp_mask(x[[1, 0]], fl)
    assert (out.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() <= 4e-6
    s = st.cpu()
    assert (s[..., 0] - ref.double().sum((2, 3))).abs().max().item() <= 1e-3
    assert ((s[..., 1] - (ref.double() ** 2).sum((2, 3))).abs() / (1 + s[..., 1].abs())).max().item() <= 1e-5


def test_warp_backward_matches_grid_sample_autograd(upf):
    """gradients vs torch CPU autograd of the op-for-op port (F.grid_sample)"""
    x = _regen(16, (1, 5, 9, 11)).double()
    fl = (_regen(17, (1, 2, 9, 11)) * 1.5).double()
    xr, fr = x.clone().requires_grad_(), fl.clone().requires_grad_()
    go = _regen(18, (1, 5, 9, 11)).double()
    # double-precision reference: the port's formula with explicit align_corners=False
    grid = P._vgrid(fr)
    ref = torch.nn.functional.grid_sample(xr, grid, padding_mode="zeros", align_corners=False)
    ref.backward(go)
    xg, fg = x.float().cuda().requires_grad_(), fl.float().cuda().requires_grad_()
    upf.warp(xg, fg, use_mask=False).backward(go.float().cuda())
    assert (xg.grad.cpu() - xr.grad.float()).abs().max().item() <= 1e-5
    assert (fg.grad.cpu() - fr.grad.float()).abs().max().item() <= 1e-4


# ------------------------------------------------------------------ normalisation / resize / blend
def test_normalize_golden(upf, golden):
    for n in golden("norm"):
        out = upf.normalize_features(_cuda(n["f"])).cpu()
        assert (out - n["out"]).abs().max().item() <= 3e-6


def test_resize_golden(upf, golden):
    for u in golden("upsample"):
        h, w = u["hw"]
        out = upf.resize_bilinear(_cuda(u["x"]), h, w, flow_rate=u["if_rate"]).cpu()
        assert (out - u["out"]).abs().max().item() <= 4e-6


@pytest.mark.parametrize("src,dst", [((94, 311), (375, 1242)), ((6, 20), (12, 39)), ((109, 256), (436, 1024))])
def test_resize_vs_oracle(upf, src, dst):
    x = _regen(19, (2, 2, *src)) * 4
    out = upf.resize_bilinear(_cuda(x), *dst, flow_rate=True).cpu()
    ref = O.upsample2d_flow_as(x, *dst)
    assert (out - ref).abs().max().item() <= 2e-5


def test_sgu_blend_golden(upf, golden):
    s = golden("sgu")
    # teacher-forced with the reference's own inter_flow / (pre-sigmoid) mask: rebuild the logit
    logit = torch.logit(s["inter_mask"].double()).float()
    inter = torch.cat([s["inter_flow"], logit], 1)
    out = upf.sgu_blend(_cuda(s["flow"]), _cuda(inter)).cpu()
    assert (out - s["flow_up"]).abs().max().item() <= 2e-5


def test_sgu_blend_vs_oracle_both_variants(upf):
    flow = _regen(20, (2, 2, 24, 40)) * 3
    inter = torch.cat([_regen(21, (2, 2, 24, 40)) * 1.5, _regen(22, (2, 1, 24, 40)) * 2], 1)
    out = upf.sgu_blend(_cuda(flow), _cuda(inter)).cpu()
    ref = O.sgu_blend(flow, inter[:, :2], torch.sigmoid(inter[:, 2:3]))
    assert (out - ref).abs().max().item() <= 5e-6
    # output-level variant: inter at 24x40, flow at 95x158
    big = _regen(23, (2, 2, 95, 158)) * 5
    out = upf.sgu_blend(_cuda(big), _cuda(inter)).cpu()
    iflow = O.upsample2d_flow_as(inter[:, :2], 95, 158)
    imask = O.resize_bilinear_ac(torch.sigmoid(inter[:, 2:3]), 95, 158)
    ref = O.sgu_blend(big, iflow, imask)
    assert (out - ref).abs().max().item() <= 2e-5


# ------------------------------------------------------------------ convolution
CONV_CASES = [  # (Cin, Cout, k, stride, dil, H, W)
    (115, 128, 3, 1, 1, 12, 39), (563, 2, 3, 1, 1, 12, 20), (128, 96, 3, 1, 8, 24, 30), (96, 64, 3, 1, 16, 24, 30),
    (196, 32, 1, 1, 1, 6, 20), (3, 16, 3, 2, 1, 37, 50), (16, 16, 3, 1, 1, 19, 25), (64, 3, 3, 1, 1, 9, 9),
    (32, 32, 3, 2, 1, 20, 21), (128, 196, 3, 1, 1, 6, 20), (3, 16, 3, 1, 1, 33, 47), (4, 8, 3, 1, 1, 20, 30), (2, 30, 3, 2, 1, 21, 22),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fp32_vs_oracle(upf, case):
    Cin, Cout, k, stride, dil, H, W = case
    x = _regen(30, (2, Cin, H, W))
    w = _regen(31, (Cout, Cin, k, k)) * (2.0 / (Cin * k * k)) ** 0.5
    b = _regen(32, (Cout,)) * 0.1
    wp, _ = upf.pack_conv_weight(_cuda(w))
    out = upf.conv2d(_cuda(x), wp, _cuda(b), Cout, k, stride, dil, 0.1).cpu()
    ref = O.conv2d_direct(x.double(), w.double(), b.double(), dil, stride, 0.1).float()
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 2e-5


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[3] == 1])
def test_conv_tf32_tensor_core_vs_oracle(upf, case):
    """tcgen05 path: operands are truncated to TF32 (10-bit mantissa), fp32 accumulate.  Tolerance: the oracle run
    on TF32-truncated operands must match to fp32 rounding; vs the exact oracle the error is ~2^-10 relative."""
    from upflow_pytorch_b200 import _ext
    Cin, Cout, k, stride, dil, H, W = case
    x = _regen(30, (2, Cin, H, W))
    w = _regen(31, (Cout, Cin, k, k)) * (2.0 / (Cin * k * k)) ** 0.5
    b = _regen(32, (Cout,)) * 0.1
    _, wtc = upf.pack_conv_weight(_cuda(w), tc=True)
    out = upf.conv2d(_cuda(x), wtc, _cuda(b), Cout, k, stride, dil, 0.1, precision=_ext.CONV_TF32).cpu()

    def trunc(t):
        return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
    ref_t = O.conv2d_direct(trunc(x).double(), _rn_tf32(w).double(), b.double(), dil, stride, 0.1).float()
    ref = O.conv2d_direct(x.double(), w.double(), b.double(), dil, stride, 0.1).float()
    err_t = (out - ref_t).abs().max().item()
    err = (out - ref).abs().max().item()
    print("tf32 conv", case, "err vs truncated-operand oracle", err_t, "vs exact", err)
    assert err <= 1e-2
    assert err_t <= 5e-4


def test_dense_block_and_context_golden(golden):
    """a8/a9 through the drop-in modules (reference names) with the reference's own outputs."""
    import upflow_pytorch_b200 as pkg
    pkg.install_dropin()
    from model import pwc_modules
    pwc_modules.set_conv_precision("fp32")
    e = golden("estimator")
    sd = P.det_state_dict(e["wseed"])
    est = pwc_modules.FlowEstimatorDense_v2(115).cuda()
    est.load_state_dict({k[len("flow_estimators."):]: v for k, v in sd.items() if k.startswith("flow_estimators.")})
    with torch.no_grad():
        x5, out = est(e["x"].cuda())
    assert (x5.cpu() - e["x5"]).abs().max().item() <= 3e-5
    assert (out.cpu() - e["out"]).abs().max().item() <= 3e-5
    c = golden("context")
    ctx = pwc_modules.ContextNetwork_v2_(565).cuda()
    ctx.load_state_dict({k[len("context_networks."):]: v for k, v in sd.items() if k.startswith("context_networks.")})
    with torch.no_grad():
        o = ctx(c["x"].cuda())
    assert (o.cpu() - c["out"]).abs().max().item() <= 3e-5


def test_layout_roundtrip(upf):
    x = _regen(40, (2, 37, 13, 29)).cuda()
    a = upf.to_pixel_major(x)
    assert torch.equal(a.permute(0, 3, 1, 2), x)
    assert torch.equal(upf.to_nchw_contiguous(a), x)
    padded = upf.to_pixel_major(x, ld=40)
    assert torch.equal(padded[..., :37].permute(0, 3, 1, 2), x) and (padded[..., 37:] == 0).all()


def test_cpu_tensors_fail_loudly(upf):
    with pytest.raises(RuntimeError):
        upf.correlation(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4, 8, 8), 4)


@pytest.mark.parametrize("case", [(576, 128, 3, 1, 1, 6, 20), (128, 96, 3, 1, 8, 12, 39), (184, 3, 3, 1, 1, 24, 78),
                                  (196, 32, 1, 1, 1, 6, 20), (96, 128, 3, 2, 1, 24, 78), (16, 32, 3, 2, 1, 47, 61),
                                  (128, 196, 3, 2, 1, 12, 39), (128, 128, 3, 1, 2, 6, 20), (64, 32, 3, 1, 1, 12, 39),
                                  (576, 2, 3, 1, 1, 12, 39), (544, 32, 3, 1, 1, 24, 78), (96, 64, 3, 1, 16, 7, 16),
                                  (32, 32, 1, 1, 1, 3, 5)])
def test_conv_tf32_cluster_split_k_and_stride2(upf, case):
    """small grids split the K loop over a thread-block cluster (DSMEM reduction in rank order): same bits on every
    run, fp32-rounding agreement with the TF32-truncated oracle; stride 2 is walked by TMA element strides."""
    from upflow_pytorch_b200 import _ext
    from upflow_pytorch_b200.ops import Slice
    Cin, Cout, k, stride, dil, H, W = case
    x = _regen(50, (2, Cin, H, W))
    w = _regen(51, (Cout, Cin, k, k)) * (2.0 / (Cin * k * k)) ** 0.5
    b = _regen(52, (Cout,)) * 0.1
    _, wtc = upf.pack_conv_weight(_cuda(w), tc=True)

    def trunc(t):
        return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
    ref = O.conv2d_direct(trunc(x).double(), _rn_tf32(w).double(), b.double(), dil, stride, 0.1).float()
    res = _regen(53, tuple(ref.shape))
    ld = (Cin + 3) // 4 * 4
    a = Slice(upf.to_pixel_major(_cuda(x), ld=ld), 0, Cin)
    r = upf.to_pixel_major(_cuda(res))
    outs = []
    for _ in range(2):
        out = torch.full((2, ref.shape[2], ref.shape[3], Cout), float("nan"), device="cuda")
        upf.k_conv(a, wtc, _cuda(b), out, k, stride, dil, 0.1, r, _ext.CONV_TF32)
        outs.append(out.clone())
    assert torch.equal(outs[0], outs[1])
    err = (outs[0].permute(0, 3, 1, 2).cpu() - (ref + res)).abs().max().item()
    assert err <= 5e-4, err
    # the experimental small-grid policies (N narrowing + clusters of <= 8 / <= 16 CTAs, whole-row tiles) give the same result
    # to rounding (another, equally fixed, summation order)
    for cap in (8, 16):
        _ext.load().upf_debug_conv_tc(cap)
        try:
            outc = torch.full((2, ref.shape[2], ref.shape[3], Cout), float("nan"), device="cuda")
            upf.k_conv(a, wtc, _cuda(b), outc, k, stride, dil, 0.1, r, _ext.CONV_TF32)
        finally:
            _ext.load().upf_debug_conv_tc(0)
        assert (outc - outs[0]).abs().max().item() <= 1e-4


@pytest.mark.parametrize("shape,d", [((2, 32, 270, 480), 4), ((1, 16, 256, 512), 3), ((1, 36, 250, 480), 1), ((1, 64, 272, 448), 2),
                                     ((1, 32, 135, 470), 6), ((1, 20, 200, 333), 5), ((2, 32, 94, 311), 4),
                                     # BASELINE config 5's sweep d in {2, 4, 6} at the HD 1/4-res shape
                                     ((1, 32, 270, 480), 2), ((1, 32, 270, 480), 6)])
def test_corr_pipelined_persistent_kernel(upf, shape, d):
    """>= 25 tiles per image: the persistent TMA / warp-specialised kernel (corr_pipe.cu), raw and with fused normalisation +
    LeakyReLU, into a channel slice; and bit-for-bit agreement with the tiled kernel is NOT required (different
    summation order) -- both must meet the same 1e-5 bound against the fp64 oracle."""
    from upflow_pytorch_b200 import _ext
    from upflow_pytorch_b200.ops import Slice
    N, C, H, W = shape
    f1, f2 = _regen(60, shape), _regen(61, shape) * 1.5 + 0.3
    a, b = upf.to_pixel_major(_cuda(f1)), upf.to_pixel_major(_cuda(f2))
    D2 = (2 * d + 1) ** 2
    ref = O.correlation(f1.double(), f2.double(), d).float()
    out = torch.empty(N, H, W, D2, device="cuda")
    upf.k_corr(a, b, out, d, slope=1.0)
    assert (out.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() <= 1e-5
    # fused normalisation + LeakyReLU + batch shift, written into a slice of a wider buffer
    s1 = torch.zeros(N, C, 2, dtype=torch.float64, device="cuda")
    s2 = torch.zeros_like(s1)
    upf.k_stats(a, s1)
    upf.k_stats(b, s2)
    X = torch.full((N, H, W, D2 + 15), 3.0, device="cuda")
    upf.k_corr(a, b, Slice(X, 0, D2), d, s1, s2, f2_shift=N - 1, slope=0.1)
    refn = O.correlation(O.normalize_features(f1.double()), O.normalize_features(f2.double())[[(n + N - 1) % N for n in range(N)]], d, 0.1).float()
    assert (X[..., :D2].permute(0, 3, 1, 2).cpu() - refn).abs().max().item() <= 2e-5
    assert (X[..., D2:] == 3.0).all()
    # the same call through the tiled kernel (debug switch) agrees to rounding
    _ext.load().upf_debug_corr_pipe(0)
    try:
        out2 = torch.empty_like(out)
        upf.k_corr(a, b, out2, d, slope=1.0)
    finally:
        _ext.load().upf_debug_corr_pipe(1)
    assert (out2 - out).abs().max().item() <= 1e-5


@pytest.mark.parametrize("case", [(200, 128, 1, 70, 200), (168, 24, 2, 66, 150), (96, 64, 4, 64, 130), (40, 3, 1, 90, 95),
                                  (16, 32, 1, 50, 70), (100, 48, 1, 40, 61), (544, 32, 1, 47, 156), (160, 16, 2, 47, 100)])
def test_conv_tf32_large_grid_kernels_agree(upf, case):
    """Fine pyramid levels: the same 3x3 convolution through the per-tap kernel (conv_tc.cu, two CTAs per SM), the
    shared-halo kernel (conv_halo.cu) and the linear-window kernel (conv_win.cu; 1, 2 and 4 units per CTA, two MMA
    issuers; one MMA per kernel row with the horizontal taps along N -- the default for Cout <= 64 -- and one per tap)
    -- each must match the TF32-truncated oracle to fp32 rounding, repeat bit for bit, and leave the neighbouring
    channels of the output buffer alone."""
    from upflow_pytorch_b200 import _ext
    from upflow_pytorch_b200.ops import Slice
    lib = _ext.load()
    Cin, Cout, dil, H, W = case
    x = _regen(70, (2, Cin, H, W))
    w = _regen(71, (Cout, Cin, 3, 3)) * (2.0 / (Cin * 9)) ** 0.5
    b = _regen(72, (Cout,)) * 0.1
    _, wtc = upf.pack_conv_weight(_cuda(w), tc=True)

    def trunc(t):
        return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
    ref = O.conv2d_direct(trunc(x).double(), _rn_tf32(w).double(), b.double(), dil, 1, 0.1).float()
    ld = (Cin + 3) // 4 * 4 + 8
    a = Slice(upf.to_pixel_major(_cuda(x), ld=ld), 0, Cin)
    ldo = (Cout + 3) // 4 * 4 + 4
    HALO = (1 << 16) | (128 << 8)
    modes = {"tap": (0, 0, 0), "halo": (0, 0, 1), "win": (3, 0, 0), "win m1": (3, 1, 0), "win m2": (3, 2, 0), "win m4": (3, 4, 0),
             "win one CTA": (3, 16, 0), "win unit split": (3, 8, 0), "win four epilogue warps": (3, 32, 0), "win item per channel block": (19, 0, 0), "win item per channel block m2": (19, 2, 0), "win weight multicast": (3, 64, 0), "win9 weight multicast": (7, 64, 0),
             "win9": (7, 0, 0), "win9 m1": (7, 1, 0), "win9 m2": (7, 2, 0), "win9 m4": (7, 4, 0)}
    try:
        modes["halo four epilogue warps"] = (0, 0, 1, 16)
        modes["win BN=16 kernel-row items"] = (19, 0, 0, 0, 1000 + (1 << 20))   # min_cin >= 1000: channel-block items from (min_cin - 1000) blocks on
        modes["win BN=16 channel-block items"] = (19, 0, 0, 0, 1003)            # (also puts the switch back)
        for name, mode in modes.items():
            wen, fm, hen = mode[:3]
            lib.upf_debug_conv_win(wen, mode[4] if len(mode) > 4 else 0, fm)
            lib.upf_debug_conv_halo(hen, HALO | (mode[3] if len(mode) > 3 else 0))
            outs = []
            for _ in range(3):
                buf = torch.full((2, H, W, ldo), 7.0, device="cuda")
                upf.k_conv(a, wtc, _cuda(b), Slice(buf, 0, Cout), 3, 1, dil, 0.1, None, _ext.CONV_TF32)
                outs.append(buf)
            assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), name
            assert (outs[0][..., Cout:] == 7.0).all(), name
            err = (outs[0][..., :Cout].permute(0, 3, 1, 2).cpu() - ref).abs().max().item()
            assert err <= 5e-4, (name, err)
    finally:
        lib.upf_debug_conv_win(1, 0, 0)
        lib.upf_debug_conv_halo(1, (65 << 16) | (128 << 8))


@pytest.mark.parametrize("case", [(576, 2, 1, 47, 60), (184, 3, 1, 33, 41), (176, 8, 1, 20, 50), (32, 2, 1, 64, 30), (64, 5, 4, 40, 44)])
def test_conv3x3_expand_then_tap_combine(upf, case):
    """3x3 conv with <= 8 outputs as a 1x1 tensor-core conv with 9*Cout outputs + upf_conv3x3_tap_combine (bias,
    LeakyReLU, residual, dilation, zero padding) against the direct oracle on TF32-truncated operands."""
    from upflow_pytorch_b200 import _ext
    from upflow_pytorch_b200.ops import Slice
    Cin, Cout, dil, H, W = case
    x, w, b = _regen(70, (2, Cin, H, W)), _regen(71, (Cout, Cin, 3, 3)) * 0.05, _regen(72, (Cout,)) * 0.1
    res = _regen(73, (2, Cout, H, W))
    trunc = lambda t: (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
    ref = O.conv2d_direct(trunc(x).double(), _rn_tf32(w).double(), b.double(), dilation=dil, leaky_slope=0.1).float() + res
    a = upf.to_pixel_major(_cuda(x))
    wexp = upf.pack_conv_weight(upf.expand_taps_weight(_cuda(w)), tc=True)[1]
    Y = torch.zeros(2, H, W, 9 * 8, device="cuda")
    ys = Slice(Y, 0, 9 * Cout)
    upf.k_conv(a, wexp, torch.zeros(9 * Cout, device="cuda"), ys, 1, 1, 1, 1.0, None, _ext.CONV_TF32)
    out = torch.full((2, H, W, Cout), float("nan"), device="cuda")
    upf.k_tap_combine(ys, _cuda(b), out, dil, 0.1, upf.to_pixel_major(_cuda(res)))
    err = (out.permute(0, 3, 1, 2).cpu() - ref).abs().max().item()
    assert err <= 5e-4, err


# ------------------------------------------------------------------ TF32-rounded producers (UPF_FLAG_ROUND_TF32)
def test_round_tf32_flag_is_exactly_the_rounded_plain_result(upf):
    """Every producer that can store its result pre-rounded for tensor-core consumers must store EXACTLY
    rn_tf32(the un-rounded result): convolution epilogues of all kernel families (SIMT, image conv, cluster split-K,
    window, halo), tap combine, correlation (small / tiled / pipelined), warp, channel copy, SGU blend."""
    from upflow_pytorch_b200 import _ext
    from upflow_pytorch_b200.ops import Slice
    g = _g(77)
    dev = "cuda"

    def conv_pair(N, H, W, Cin, Cout, k, stride, dil, prec):
        x = torch.randn(N, H, W, Cin, generator=g).to(dev)
        w = (torch.randn(Cout, Cin, k, k, generator=g) * 0.1).to(dev)
        b = (torch.randn(Cout, generator=g) * 0.1).to(dev)
        ws, wt = upf.pack_conv_weight(w, tc=prec == _ext.CONV_TF32)
        wk = wt if prec == _ext.CONV_TF32 else ws
        pad = ((k - 1) * dil) // 2
        Ho, Wo = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1, (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
        outs = []
        for flag in (0, _ext.CONV_ROUND_OUT):
            o = torch.zeros(N, Ho, Wo, Cout, device=dev)
            upf.k_conv(Slice(x), wk, b, Slice(o), k, stride, dil, 0.1, None, prec | flag)
            outs.append(o.cpu())
        return outs

    cases = [(2, 20, 24, 3, 16, 3, 2, 1, _ext.CONV_FP32),        # conv_c3
             (1, 9, 11, 20, 12, 3, 1, 1, _ext.CONV_FP32),        # conv_simt
             (2, 12, 20, 64, 96, 3, 1, 8, _ext.CONV_TF32),       # conv_tc (cluster split-K)
             (2, 94, 100, 64, 32, 3, 1, 1, _ext.CONV_TF32),      # conv_win
             (2, 94, 100, 96, 128, 3, 1, 2, _ext.CONV_TF32)]     # conv_halo
    for c in cases:
        plain, rounded = conv_pair(*c)
        assert torch.equal(rounded, _rn_tf32(plain)), c
        assert not torch.equal(rounded, plain)

    # correlation: coarse (corr_small), tiled and pipelined shapes
    for (N, C, H, W) in ((2, 32, 12, 20), (1, 30, 40, 48), (2, 32, 94, 160)):
        f1, f2 = torch.randn(N, H, W, C, generator=g).to(dev), torch.randn(N, H, W, C, generator=g).to(dev)
        o0, o1 = torch.zeros(N, H, W, 81, device=dev), torch.zeros(N, H, W, 81, device=dev)
        upf.k_corr(f1, f2, o0, 4, slope=0.1)
        upf.k_corr(f1, f2, o1, 4, slope=0.1, round_tf32=True)
        assert torch.equal(o1.cpu(), _rn_tf32(o0.cpu())), (N, C, H, W)

    # warp, copy, tap combine, blend
    x = torch.randn(2, 30, 44, 32, generator=g).to(dev)
    fl = (torch.randn(2, 30, 44, 2, generator=g) * 3).to(dev)
    o0, o1 = torch.zeros_like(x), torch.zeros_like(x)
    upf.k_warp(x, fl, o0)
    upf.k_warp(x, fl, o1, round_tf32=True)
    assert torch.equal(o1.cpu(), _rn_tf32(o0.cpu()))
    big = torch.zeros(2, 30, 44, 40, device=dev)
    upf.k_copy(Slice(x, 0, 30), Slice(big, 4, 30), round_tf32=True)
    assert torch.equal(big[..., 4:34].cpu(), _rn_tf32(x[..., :30].cpu()))
    assert big[..., :4].abs().max().item() == 0 and big[..., 34:].abs().max().item() == 0
    upf.k_copy(None, Slice(big, 4, 30))
    assert big.abs().max().item() == 0
    y = torch.randn(2, 30, 44, 72, generator=g).to(dev)
    bias = torch.randn(8, generator=g).to(dev)
    o0, o1 = torch.zeros(2, 30, 44, 8, device=dev), torch.zeros(2, 30, 44, 8, device=dev)
    upf.k_tap_combine(y, bias, o0, 1, 0.1)
    upf.k_tap_combine(y, bias, o1, 1, 0.1, round_tf32=True)
    assert torch.equal(o1.cpu(), _rn_tf32(o0.cpu()))
    inter = torch.randn(2, 30, 44, 4, generator=g).to(dev)
    o0, o1 = torch.zeros(2, 30, 44, 2, device=dev), torch.zeros(2, 30, 44, 2, device=dev)
    slot = torch.full((2, 30, 44, 8), 7.0, device=dev)
    upf.k_sgu_blend(fl, Slice(inter, 0, 3), o0)
    upf.k_sgu_blend(fl, Slice(inter, 0, 3), o1, out_tc=Slice(slot, 2, 4), round_tf32=True)
    assert torch.equal(o0.cpu(), o1.cpu())                                   # the exact output does not change
    assert torch.equal(slot[..., 2:4].cpu(), _rn_tf32(o0.cpu()))
    assert slot[..., 4:6].abs().max().item() == 0 and (slot[..., :2] == 7).all() and (slot[..., 6:] == 7).all()


def test_tc_weight_packing_rounds_to_nearest(upf):
    """the tensor-core weight layouts hold rn_tf32(w): a convolution on them equals the oracle on rounded weights and
    truncated activations to fp32 rounding (and is closer to the exact result than truncated weights would be)"""
    from upflow_pytorch_b200 import _ext
    from upflow_pytorch_b200.ops import Slice
    g = _g(5)
    x, w, b = torch.randn(1, 24, 40, 64, generator=g), torch.randn(32, 64, 3, 3, generator=g) * 0.05, torch.zeros(32)
    trunc = lambda t: (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
    xn = x.permute(0, 3, 1, 2).contiguous()
    ref_rn = O.conv2d_direct(trunc(xn).double(), _rn_tf32(w).double(), b.double(), 1, 1, 1.0).float()
    ref_tr = O.conv2d_direct(trunc(xn).double(), trunc(w).double(), b.double(), 1, 1, 1.0).float()
    for ft in (False, True):
        if ft:
            _, wt = upf.pack_conv_weight(w.cuda(), tc=True, tc_only=True)
        else:
            _, wt = upf.pack_conv_weight(w.cuda(), in_slots=list(range(64)), cin_total=64, tc=True)
        out = torch.zeros(1, 24, 40, 32, device="cuda")
        upf.k_conv(Slice(x.cuda()), wt, b.cuda(), Slice(out), 3, 1, 1, 1.0, None, _ext.CONV_TF32)
        got = out.permute(0, 3, 1, 2).cpu()
        assert (got - ref_rn).abs().max().item() <= 5e-5
        assert (got - ref_tr).abs().max().item() > 2e-4


# ------------------------------------------------------------------ the `correlation_cuda` stub (native boundary b1)
class _RefCorrelationFunction(torch.autograd.Function):
    """model/correlation_package/correlation.py:6-44 restated with static methods (the legacy instance-style Function of
    the reference is rejected by torch >= 1.5): the same calls into `correlation_cuda`, the same empty tensors."""

    @staticmethod
    def forward(ctx, input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply):
        import correlation_cuda
        ctx.save_for_backward(input1, input2)
        ctx.cfg = (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)
        with torch.cuda.device_of(input1):
            rbot1, rbot2, output = input1.new(), input2.new(), input1.new()
            correlation_cuda.forward(input1, input2, rbot1, rbot2, output, *ctx.cfg)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        import correlation_cuda
        input1, input2 = ctx.saved_tensors
        with torch.cuda.device_of(input1):
            rbot1, rbot2 = input1.new(), input2.new()
            grad_input1, grad_input2 = input1.new(), input2.new()
            correlation_cuda.backward(input1, input2, rbot1, rbot2, grad_output.contiguous(), grad_input1, grad_input2, *ctx.cfg)
        return (grad_input1, grad_input2) + (None,) * 6


def test_correlation_cuda_stub_matches_reference_golden(golden):
    """`import correlation_cuda` (upflow_pytorch_b200/dropin/correlation_cuda.py, the file INTEGRATION.md section 2
    installs next to the reference's correlation.py) driven exactly as correlation.py drives the pybind module, against
    the reference's own outputs (tests/golden/corr.pt incl. BASELINE config 1)."""
    import upflow_pytorch_b200
    upflow_pytorch_b200.install_dropin()
    import correlation_cuda
    for c in golden("corr"):
        f1 = c["f1"] if c["f1"] is not None else _regen(c["seed"], c["shape"])
        f2 = c["f2"] if c["f2"] is not None else _regen(c["seed"] + 100, c["shape"])
        d = c["d"]
        out = _RefCorrelationFunction.apply(f1.cuda(), f2.cuda(), d, 1, d, 1, 1, 1)
        assert out.shape == c["out"].shape and out.is_contiguous()
        assert (out.cpu() - c["out"]).abs().max().item() <= 1e-5
    # conventions: returns 1, sizes the caller's empty tensors, rejects what the kernel does not implement
    a, b = torch.randn(1, 8, 6, 7).cuda(), torch.randn(1, 8, 6, 7).cuda()
    o = a.new()
    assert correlation_cuda.forward(a, b, a.new(), b.new(), o, 4, 1, 4, 1, 1, 1) == 1 and o.shape == (1, 81, 6, 7)
    with pytest.raises(RuntimeError):
        correlation_cuda.forward(a, b, a.new(), b.new(), a.new(), 3, 3, 20, 1, 2, 1)      # FlowNet2's setting, not UPFlow's
    with pytest.raises(RuntimeError):
        correlation_cuda.forward(a.cpu(), b.cpu(), a.new(), b.new(), a.new(), 4, 1, 4, 1, 1, 1)


def test_correlation_cuda_stub_backward_vs_autograd():
    """correlation_cuda.backward through the restated CorrelationFunction against autograd of Corr_pyTorch's expression
    (oracle port) in fp64."""
    import upflow_pytorch_b200
    upflow_pytorch_b200.install_dropin()
    g = _g(11)
    f1, f2 = torch.randn(2, 12, 9, 13, generator=g), torch.randn(2, 12, 9, 13, generator=g)
    go = torch.randn(2, 81, 9, 13, generator=g)
    r1, r2 = f1.double().requires_grad_(), f2.double().requires_grad_()
    P.corr_unfold(r1, r2, 4).backward(go.double())
    c1, c2 = f1.cuda().requires_grad_(), f2.cuda().requires_grad_()
    _RefCorrelationFunction.apply(c1, c2, 4, 1, 4, 1, 1, 1).backward(go.cuda())
    assert (c1.grad.cpu().double() - r1.grad).abs().max().item() <= 1e-5
    assert (c2.grad.cpu().double() - r2.grad).abs().max().item() <= 1e-5
