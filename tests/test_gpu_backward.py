"""a11: backward kernels against torch autograd of the reference's own ops on CPU (fp64), and the
whole training-path gradient of the drop-in UPFlow_net against autograd through the op-for-op port.  -m gpu."""
import pytest
import torch
import torch.nn.functional as F

from oracle import cpu_oracle as O
from oracle import ref_port as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def upf():
    from upflow_pytorch_b200 import _ext, ops
    _ext.load()
    return ops


def _rand(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def _check(name, val, tol):
    if not val <= tol:
        raise AssertionError("%s: %.3g > %.3g" % (name, val, tol)) from None
    print("  %s %.3g (tol %.3g)" % (name, val, tol))


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("case", [(20, 12, 3, 1, 1, 13, 18), (16, 8, 3, 2, 1, 13, 18), (16, 8, 3, 2, 1, 14, 20),
                                  (24, 6, 3, 1, 4, 17, 21), (10, 2, 1, 1, 1, 9, 11), (70, 130, 3, 1, 2, 20, 24),
                                  (3, 16, 3, 1, 1, 16, 16), (36, 3, 3, 1, 1, 12, 40)])
@pytest.mark.parametrize("relu", [True, False])
def test_conv_backward_fp32_vs_autograd(upf, case, relu):
    """dgrad (forward kernel on flipped weights; stride 2 on the zero-interleaved gradient), wgrad, bias gradient and
    the LeakyReLU derivative against F.conv2d + F.leaky_relu autograd in fp64.  Tolerance 2e-5 relative to the
    largest gradient entry (fp32 sums over up to 70*9 products / 960 pixels)."""
    from upflow_pytorch_b200 import _ext
    Cin, Cout, k, stride, dil, H, W = case
    x, w, b = _rand(1, 2, Cin, H, W), _rand(2, Cout, Cin, k, k) * 0.2, _rand(3, Cout) * 0.1
    xd, wd, bd = (t.double().requires_grad_() for t in (x, w, b))
    y = F.conv2d(xd, wd, bd, stride=stride, padding=((k - 1) * dil) // 2, dilation=dil)
    y = F.leaky_relu(y, 0.1) if relu else y
    gy = _rand(4, *y.shape)
    y.backward(gy.double())
    xc, wc, bc = (t.cuda().requires_grad_() for t in (x, w, b))
    yc = upf.conv2d_autograd(xc, wc, bc, stride, dil, 0.1 if relu else 1.0, _ext.CONV_FP32)
    _check("y", _rel(yc.detach().cpu(), y.detach()), 2e-5)
    yc.backward(gy.cuda())
    for got, ref, name in ((xc.grad, xd.grad, "dx"), (wc.grad, wd.grad, "dw"), (bc.grad, bd.grad, "db")):
        assert got.shape == ref.shape
        _check(name, _rel(got.cpu(), ref), 2e-5)


@pytest.mark.parametrize("case", [(64, 32, 3, 1, 1, 24, 40), (96, 128, 3, 1, 2, 20, 33), (32, 32, 1, 1, 1, 17, 29), (16, 32, 3, 2, 1, 30, 44),
                                  # few channels, many pixels (K = 65k..68k per cluster of 8)
                                  (8, 16, 3, 1, 1, 128, 256), (16, 8, 1, 1, 1, 160, 200), (3, 16, 3, 1, 2, 96, 320)])
@pytest.mark.parametrize("relu", [False, True])
def test_conv_backward_tf32(upf, case, relu):
    """tensor-core forward, dgrad and wgrad (TF32 operands, fp32 accumulate; stride 2: SIMT fp32 wgrad): 1e-2 relative."""
    from upflow_pytorch_b200 import _ext
    Cin, Cout, k, stride, dil, H, W = case
    slope = 0.1 if relu else 1.0
    x, w, b = _rand(1, 2, Cin, H, W), _rand(2, Cout, Cin, k, k) * 0.1, _rand(3, Cout) * 0.1
    xc, wc, bc = (t.cuda().requires_grad_() for t in (x, w, b))
    yc = upf.conv2d_autograd(xc, wc, bc, stride, dil, slope, _ext.CONV_TF32)
    xd, wd, bd = (t.double().requires_grad_() for t in (x, w, b))
    ypre = F.conv2d(xd, wd, bd, stride=stride, padding=((k - 1) * dil) // 2, dilation=dil)
    # the activation pattern of the TF32 forward (a TF32-sized error flips the sign of ~0.1 % of the outputs; the
    # gradient is checked for the pattern the kernel actually produced)
    y = ypre * torch.where(yc.detach().cpu() > 0, 1.0, slope).double()
    gy = _rand(4, *y.shape)
    y.backward(gy.double())
    yc.backward(gy.cuda())
    _check("y", _rel(yc.detach().cpu(), F.leaky_relu(ypre.detach(), slope)), 1e-2)
    _check("dx", _rel(xc.grad.cpu(), xd.grad), 1e-2)
    _check("dw", _rel(wc.grad.cpu(), wd.grad), 1e-2)
    _check("db", _rel(bc.grad.cpu(), bd.grad), 1e-2)


@pytest.mark.parametrize("shape", [(2, 32, 24, 39), (1, 196, 6, 20), (3, 7, 11, 13)])
def test_normalize_features_backward(upf, shape):
    x = _rand(5, *shape) * 1.7 + 0.4
    g = _rand(6, *shape)
    xd = x.double().requires_grad_()
    P.normalize(xd).backward(g.double())
    xc = x.cuda().requires_grad_()
    upf.normalize_features(xc).backward(g.cuda())
    _check("dx", _rel(xc.grad.cpu(), xd.grad), 1e-5)


@pytest.mark.parametrize("case", [(2, 12, 39, 24, 78, True), (2, 94, 311, 375, 1242, True), (1, 47, 156, 188, 621, False),
                                  (1, 1, 7, 5, 7, False), (2, 24, 78, 24, 78, True), (2, 4, 13, 256, 832, True), (2, 3, 5, 64, 96, False)])
def test_resize_backward(upf, case):
    C, h, w, H, W, rate = case
    C = 2 if rate else 1
    x = _rand(7, 2, C, h, w)
    g = _rand(8, 2, C, H, W)
    xd = x.double().requires_grad_()
    P.upsample2d_flow_as(xd, H, W, if_rate=rate).backward(g.double())
    xc = x.cuda().requires_grad_()
    upf.resize_bilinear(xc, H, W, flow_rate=rate).backward(g.cuda())
    _check("dx", _rel(xc.grad.cpu(), xd.grad), 3e-5)


@pytest.mark.parametrize("lowres", [False, True])
def test_sgu_blend_backward(upf, lowres):
    """the differentiable chain (sigmoid, upsamples, un-masked warp, blend) against model/upflow.py:79-88 in autograd;
    forward of the chain == the fused inference kernel."""
    H, W = 36, 52
    h, w = (9, 13) if lowres else (H, W)
    flow_init, inter = _rand(9, 2, 2, H, W) * 2.0, _rand(10, 2, 3, h, w)
    g = _rand(11, 2, 2, H, W)
    fd, idd = flow_init.double().requires_grad_(), inter.double().requires_grad_()
    inter_flow, mask = idd[:, :2], torch.sigmoid(idd[:, 2:3])
    if lowres:
        inter_flow = P.upsample2d_flow_as(inter_flow, H, W, if_rate=True)
        mask = P.upsample2d_flow_as(mask, H, W)
    ref = P.torch_warp(fd, inter_flow) * (1 - mask) + fd * mask
    ref.backward(g.double())
    fc, ic = flow_init.cuda().requires_grad_(), inter.cuda().requires_grad_()
    out = upf.sgu_blend(fc, ic)
    out.backward(g.cuda())
    with torch.no_grad():
        fused = upf.sgu_blend(flow_init.cuda(), inter.cuda())
    _check("chain vs fused kernel", (out.detach() - fused).abs().max().item(), 2e-6)
    _check("forward", _rel(out.detach().cpu(), ref.detach()), 1e-5)
    # the warp gradient is discontinuous where a sample crosses a pixel boundary: compare away from fp32-vs-fp64 flips
    for got, refg, name in ((fc.grad, fd.grad, "dflow_init"), (ic.grad, idd.grad, "dinter")):
        bad = ((got.cpu().double() - refg).abs() > 1e-4 * refg.abs().max()).float().mean().item()
        _check(name + " fraction of entries off by > 1e-4", bad, 2e-3)


def test_training_path_gradients_vs_port():
    """forward_2_frame_v3 with autograd (every op an autograd node on this library's kernels) against autograd through
    the op-for-op CPU port of the reference, same weights, same smooth loss, robust-mask diagnostic on both sides
    (tests/test_gpu_engine.py explains the `mask >= 1.0` noise).  All 80 parameter tensors receive a gradient."""
    import upflow_pytorch_b200
    from upflow_pytorch_b200 import ops
    upflow_pytorch_b200.install_dropin()
    from model.upflow import UPFlow_net
    conf = UPFlow_net.config()
    conf.update({"if_norm_before_cost_volume": True, "norm_moments_across_channels": False,
                 "norm_moments_across_images": False, "if_sgu_upsample": True})
    net = conf()
    sd = P.det_state_dict(5)
    net.load_state_dict(sd)
    net = net.cuda().train()
    net.conv_precision = "fp32"
    im1, im2 = O.synthetic_pair(64, 96, seed=77)

    def loss_of(f, b):
        return (f ** 2).mean() + (b - 0.5).abs().mean() + (f[:, :, 1:] - f[:, :, :-1]).abs().mean()

    P.MASK_THRESHOLD = 0.9999
    ops.DIAG_MASK_THRESHOLD = 0.9999
    try:
        sdr = {k: v.clone().requires_grad_() for k, v in sd.items()}
        rf, rb, _ = P.forward_2_frame(im1, im2, sdr)
        ref_loss = loss_of(rf, rb)
        ref_loss.backward()
        f, b, flows = net.forward_2_frame_v3(im1.cuda(), im2.cuda())
        loss = loss_of(f, b)
        loss.backward()
    finally:
        P.MASK_THRESHOLD = 1.0
        ops.DIAG_MASK_THRESHOLD = None
    assert abs(loss.item() - ref_loss.item()) <= 1e-4 * abs(ref_loss.item())
    worst = 0.0
    n = 0
    for name, p in net.named_parameters():
        assert p.grad is not None, name
        ref = sdr[name].grad
        err = (p.grad.cpu() - ref).norm().item() / max(ref.norm().item(), 1e-12)
        worst = max(worst, err)
        n += 1
    print("training-path gradient: %d tensors, worst relative L2 error %.3g" % (n, worst))
    assert n == 80
    assert worst <= 2e-2, worst


def _dropin_net(conf_dict, wseed, precision):
    import upflow_pytorch_b200
    upflow_pytorch_b200.install_dropin()
    from model.upflow import UPFlow_net
    conf = UPFlow_net.config()
    conf.update(conf_dict)
    net = conf()
    net.load_state_dict(P.det_state_dict(wseed), strict=bool(conf_dict.get("if_sgu_upsample")))   # no SGU: its 20 tensors are absent
    net.conv_precision = precision
    return net.cuda().train()


def test_training_step_vs_reference_golden(golden):
    """UPFlow_net.forward(if_loss=True) + backward against the REFERENCE's own training step executed on CPU
    (tests/golden/train_step.pt, oracle/make_golden_train.py).  The reference's `mask >= 1.0` makes the forward noisy
    at the 1e-2 px level (tests/test_gpu_engine.py), so losses are pinned to 2 %, gradient norms to 15 % and the
    direction of the stored gradients to a cosine of 0.97."""
    g = golden("train_step")
    net = _dropin_net(g["conf"], g["wseed"], "fp32")
    im1, im2 = O.synthetic_pair(*g["hw"], seed=g["pair_seed"], batch=g["batch"])
    out = net({"im1": im1.cuda(), "im2": im2.cuda(), "if_loss": True})
    for k in ("photo_loss", "smooth_loss", "msd_loss"):
        _check(k, abs(out[k].item() - g[k]) / abs(g[k]), 2e-2)
    from upflow_pytorch_b200.train import total_loss
    loss = total_loss(out)
    _check("loss", abs(loss.item() - g["loss"]) / g["loss"], 2e-2)
    loss.backward()
    worst = 0.0
    for name, p in net.named_parameters():
        assert p.grad is not None, name
        worst = max(worst, abs(p.grad.norm().item() - g["grad_norm"][name]) / max(g["grad_norm"][name], 1e-12))
    _check("worst gradient-norm deviation over 80 tensors", worst, 0.15)
    params = dict(net.named_parameters())
    for name, ref in g["grads"].items():
        got = params[name].grad.cpu().flatten()
        cos = torch.dot(got, ref.flatten()) / (got.norm() * ref.norm())
        _check("1 - cos(grad %s)" % name, 1.0 - cos.item(), 0.03)


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_trainer_steps_reduce_the_loss(precision):
    """scripts/simple_train.py's loop (Adam amsgrad, photo + smooth + msd) on one synthetic pair: the loss falls."""
    from upflow_pytorch_b200.train import Trainer
    conf = {"if_norm_before_cost_volume": True, "norm_moments_across_channels": False, "norm_moments_across_images": False,
            "if_sgu_upsample": True, "if_use_boundary_warp": False, "multi_scale_distillation_weight": 0.01}
    net = _dropin_net(conf, 11, precision)
    tr = Trainer(net, lr=2e-4)
    im1, im2 = O.synthetic_pair(64, 96, seed=5, batch=2)
    batch = {"im1": im1.cuda(), "im2": im2.cuda()}
    losses = [tr.train_step(batch).item() for _ in range(6)]
    print("  losses", ["%.4f" % v for v in losses])
    assert all(v == v for v in losses)
    assert losses[-1] < losses[0]
    assert tr.grads.numel == 3494549


def test_training_step_default_loss_config_vs_reference_golden(golden):
    """scripts/simple_train.py's default loss configuration -- boundary-dilated warp on the un-cropped frames plus the
    census term -- against the reference's own step (tests/golden/train_step_boundary_census.pt)."""
    g = golden("train_step_boundary_census")
    net = _dropin_net(g["conf"], g["wseed"], "fp32")
    hw, start = g["hw"], g["start"]
    raw1, raw2 = O.synthetic_pair(hw[0] + 16, hw[1] + 16, seed=g["pair_seed"], batch=g["batch"])
    crop = lambda t: torch.stack([t[b, :, int(start[b, 1]):int(start[b, 1]) + hw[0], int(start[b, 0]):int(start[b, 0]) + hw[1]]
                                  for b in range(g["batch"])])
    out = net({"im1": crop(raw1).cuda(), "im2": crop(raw2).cuda(), "im1_raw": raw1.cuda(), "im2_raw": raw2.cuda(),
               "start": start.cuda(), "if_loss": True})
    for k in ("photo_loss", "smooth_loss", "census_loss", "msd_loss"):
        _check(k, abs(out[k].item() - g[k]) / abs(g[k]), 2e-2)
    from upflow_pytorch_b200.train import total_loss
    loss = total_loss(out)
    _check("loss", abs(loss.item() - g["loss"]) / g["loss"], 2e-2)
    loss.backward()
    worst = 0.0
    for name, p in net.named_parameters():
        if name in g["grad_norm"]:
            worst = max(worst, abs(p.grad.norm().item() - g["grad_norm"][name]) / max(g["grad_norm"][name], 1e-12))
    _check("worst gradient-norm deviation", worst, 0.15)


def test_trainer_cuda_graph_matches_eager():
    """the captured forward + backward (Trainer(use_cuda_graph=True)) follows the eager trainer step for step
    (scatter-add atomics in the warp backward make the two differ in the last bits only)."""
    from upflow_pytorch_b200.train import Trainer
    conf = {"if_norm_before_cost_volume": True, "norm_moments_across_channels": False, "norm_moments_across_images": False,
            "if_sgu_upsample": True, "if_use_boundary_warp": False, "multi_scale_distillation_weight": 0.01}
    im1, im2 = O.synthetic_pair(64, 96, seed=5, batch=2)
    batch = {"im1": im1.cuda(), "im2": im2.cuda()}
    losses = []
    for graphed in (False, True):
        tr = Trainer(_dropin_net(conf, 11, "fp32"), lr=2e-4, use_cuda_graph=graphed)
        losses.append([tr.train_step(batch).item() for _ in range(4)])
    print("  eager  ", ["%.5f" % v for v in losses[0]])
    print("  graphed", ["%.5f" % v for v in losses[1]])
    # same weights and batch: the first step agrees to rounding; the trajectories then drift apart at the rate the
    # training itself amplifies last-bit differences (measured 8e-6, 6e-4, 3e-4 relative on steps 2-4)
    for i, (a, b) in enumerate(zip(*losses)):
        _check("loss difference, step %d" % (i + 1), abs(a - b) / abs(a), 1e-5 if i == 0 else 2e-2)


def _torch_loss_terms():
    """The drop-in's torch expressions of the two loss terms (the kernels' A/B partner; pinned to the reference by
    the golden training steps above)."""
    import upflow_pytorch_b200
    upflow_pytorch_b200.install_dropin()
    from model.upflow import network_tools
    return network_tools


@pytest.mark.parametrize("kind", ["abs_robust", "charbonnier", "L1"])
@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("shape", [(2, 3, 37, 53), (1, 2, 6, 20), (4, 2, 256, 832)])
def test_robust_loss_kernels_vs_torch(upf, kind, masked, shape):
    """upf_robust_loss_fwd/bwd (csrc/loss.cu) against the torch expression of photo_loss_multi_type
    (model/upflow.py:268-290) in fp64: value to 5e-6 relative, both gradients to 2e-5 of their largest entry.
    Inputs arrive channels_last and NCHW-contiguous, like the training path's."""
    nt = _torch_loss_terms()
    N, C, H, W = shape
    x = _rand(1, N, C, H, W).cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    y = _rand(2, N, C, H, W).cuda().requires_grad_()
    mask = (torch.rand(N, 1, H, W, generator=torch.Generator().manual_seed(3)) > 0.4).float().cuda()
    got = upf.robust_loss(x, y, mask if masked else None, kind, 0.4)
    gx, gy = torch.autograd.grad(got * 3.0, (x, y))
    xd, yd = x.detach().double().requires_grad_(), y.detach().double().requires_grad_()
    nt.use_loss_kernels = False
    try:
        want = nt.photo_loss_multi_type(xd, yd, mask.double(), kind, 0.4, photo_loss_use_occ=masked)
    finally:
        nt.use_loss_kernels = True
    wx, wy = torch.autograd.grad(want * 3.0, (xd, yd))
    _check("value", abs(got.item() - want.item()) / abs(want.item()), 5e-6)
    _check("grad x", _rel(gx, wx), 2e-5)
    _check("grad y", _rel(gy, wy), 2e-5)
    # only one side needs a gradient (the distillation label, the source image)
    (gy2,) = torch.autograd.grad(upf.robust_loss(x.detach(), y, mask if masked else None, kind, 0.4) * 3.0, (y,))
    assert torch.equal(gy2, gy)
    # deterministic
    assert upf.robust_loss(x, y, mask if masked else None, kind, 0.4).item() == got.item()


@pytest.mark.parametrize("shape", [(2, 37, 53), (1, 2, 2), (4, 256, 832), (1, 375, 1242)])
def test_edge_smooth1_kernels_vs_torch(upf, shape):
    """upf_edge_smooth1_fwd/bwd against the torch expression of edge_aware_smoothness_order1 (model/upflow.py:198-218)
    in fp64.  The flow is quantised so that some neighbouring differences are exactly zero (sign(0) = 0)."""
    nt = _torch_loss_terms()
    N, H, W = shape
    img = torch.rand(N, 3, H, W, generator=torch.Generator().manual_seed(4)).cuda()
    flow = (_rand(5, N, 2, H, W) * 2).round().div(2).cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    got = upf.edge_smooth1(img, flow)
    (gf,) = torch.autograd.grad(got * 2.0, (flow,))
    fd = flow.detach().double().requires_grad_()
    nt.use_loss_kernels = False
    try:
        want = nt.edge_aware_smoothness_order1(img.double(), fd)
    finally:
        nt.use_loss_kernels = True
    (wf,) = torch.autograd.grad(want * 2.0, (fd,))
    _check("value", abs(got.item() - want.item()) / abs(want.item()), 5e-6)
    _check("grad", _rel(gf, wf), 2e-5)
    assert upf.edge_smooth1(img, flow).item() == got.item()


def test_loss_kernels_reject_bad_arguments(upf):
    x = torch.zeros(1, 2, 4, 4, device="cuda")
    with pytest.raises(ValueError):
        upf.robust_loss(x, torch.zeros(1, 2, 4, 5, device="cuda"))
    with pytest.raises(KeyError):
        upf.robust_loss(x, x, None, "SSIM")
    with pytest.raises(RuntimeError):
        upf.edge_smooth1(torch.zeros(1, 3, 1, 4, device="cuda"), torch.zeros(1, 2, 1, 4, device="cuda"))
    with pytest.raises(RuntimeError):
        upf.robust_loss(x.cpu(), x.cpu())


def test_training_step_loss_kernels_match_torch_terms():
    """The whole training step with the fused loss kernels against the same step with the torch expressions: every
    loss term to 1e-5 relative, every parameter gradient to 1e-3 of its norm (same forward, same backward kernels;
    only the loss branch differs)."""
    nt = _torch_loss_terms()
    from upflow_pytorch_b200.train import total_loss
    conf = {"if_norm_before_cost_volume": True, "norm_moments_across_channels": False, "norm_moments_across_images": False,
            "if_sgu_upsample": True, "if_use_boundary_warp": False, "multi_scale_distillation_weight": 0.01,
            "photo_loss_use_occ": True}
    im1, im2 = O.synthetic_pair(64, 96, seed=8, batch=2)
    res = {}
    for use in (True, False):
        net = _dropin_net(conf, 13, "fp32")
        nt.use_loss_kernels = use
        try:
            out = net({"im1": im1.cuda(), "im2": im2.cuda(), "if_loss": True})
            total_loss(out).backward()
        finally:
            nt.use_loss_kernels = True
        res[use] = ({k: out[k].item() for k in ("photo_loss", "smooth_loss", "msd_loss")},
                    {n: p.grad.clone() for n, p in net.named_parameters()})
    for k, v in res[False][0].items():
        _check(k, abs(res[True][0][k] - v) / abs(v), 1e-5)
    worst = max(((res[True][1][n] - g).norm() / g.norm().clamp_min(1e-20)).item() for n, g in res[False][1].items())
    _check("worst relative gradient difference", worst, 1e-3)


@pytest.mark.parametrize("shape", [(20, 12, 3), (2, 3, 3), (16, 3, 3), (7, 5, 1), (196, 128, 3)])
def test_repack_conv_weight_kernel(upf, shape):
    """upf_repack_conv_weight (one launch per convolution call of the training path) against the permute / flip /
    transpose it replaces: bit-exact, padding columns zero."""
    Cout, Cin, k = shape
    w = _rand(9, Cout, Cin, k, k).cuda()
    got, _ = upf.pack_conv_weight(w)
    pad = lambda c: (c + 3) // 4 * 4
    want = torch.zeros(k * k, Cin, pad(Cout), device="cuda")
    want[:, :, :Cout] = w.permute(2, 3, 1, 0).reshape(k * k, Cin, Cout)
    assert torch.equal(got, want)
    got_d, _ = upf.pack_conv_weight(w, flip_transpose=True)
    w_d = w.flip(2, 3).transpose(0, 1).contiguous()                    # [Cin, Cout, k, k]: Cout and Cin exchanged
    want_d = torch.zeros(k * k, Cout, pad(Cin), device="cuda")
    want_d[:, :, :Cin] = w_d.permute(2, 3, 1, 0).reshape(k * k, Cout, Cin)
    assert torch.equal(got_d, want_d)


@pytest.mark.parametrize("case", [("fp32", 115, (128, 128, 96, 64, 32), 2, 2, 24, 40), ("tf32", 115, (128, 128, 96, 64, 32), 2, 2, 24, 40),
                                  ("fp32", 64, (32, 32, 32, 16, 8), 3, 2, 20, 28), ("tf32", 64, (32, 32, 32, 16, 8), 3, 1, 64, 96),
                                  ("fp32", 13, (8, 6), 2, 1, 9, 11)])
def test_dense_block_node_vs_reference_data_flow(case):
    """ops.dense_block (one autograd node on an append-only buffer; backward: one K-concatenated input-gradient
    convolution per channel block) against the reference's own data flow -- torch.cat after every conv (model/pwc_modules.py:279-286)
    with one autograd node per conv: same forward values, gradients of the input and of all 2(n+1) parameters to 2e-5
    (fp32) / 3e-3 (TF32: the two paths sum the same products in a different order) of their largest entry."""
    import upflow_pytorch_b200
    upflow_pytorch_b200.install_dropin()
    from model import pwc_modules
    precision, ch_in, f, cout, B, H, W = case
    pwc_modules.set_conv_precision(precision)
    torch.manual_seed(3)
    blk = pwc_modules.FlowEstimatorDense_v2(ch_in, f_channels=f, out_channel=cout).cuda().train()
    for p_ in blk.parameters():
        if p_.dim() == 1:
            torch.nn.init.normal_(p_, std=0.1)
    x = _rand(1, B, ch_in, H, W).cuda().requires_grad_()
    total = ch_in + sum(f)
    r5, ro = _rand(2, B, total, H, W).cuda(), _rand(3, B, cout, H, W).cuda()
    res = {}
    try:
        for fused in (True, False):
            pwc_modules._DenseBlock.fused_training_block = fused
            blk.zero_grad(set_to_none=True)
            x5, out = blk(x)
            assert x5.shape == (B, total, H, W) and out.shape == (B, cout, H, W)
            ((x5 * r5).sum() + (out * ro).sum()).backward()
            res[fused] = (x5.detach().clone(), out.detach().clone(), x.grad.clone(),
                          {n: p_.grad.clone() for n, p_ in blk.named_parameters()})
            x.grad = None
    finally:
        pwc_modules._DenseBlock.fused_training_block = True
        pwc_modules.set_conv_precision("fp32")
    tol_f, tol_g = (1e-6, 2e-5) if precision == "fp32" else (1e-5, 3e-3)
    _check("x5", _rel(res[True][0], res[False][0]), tol_f)
    _check("conv_last", _rel(res[True][1], res[False][1]), tol_f)
    _check("grad x", _rel(res[True][2], res[False][2]), tol_g)
    for n in res[False][3]:
        _check("grad " + n, _rel(res[True][3][n], res[False][3][n]), tol_g)


@pytest.mark.parametrize("shape", [(20, 12, 3), (2, 3, 3), (7, 5, 1), (196, 128, 3), (115, 226, 3)])
@pytest.mark.parametrize("flip", [False, True])
def test_repack_conv_weight_tc_kernel_matches_two_step_packing(upf, shape, flip):
    """upf_repack_conv_weight_tc (one launch) == upf_repack_conv_weight followed by upf_conv_tc_pack_weights, bit-exact."""
    Cout, Cin, k = shape
    w = _rand(10, Cout, Cin, k, k).cuda()
    _, two_step = upf.pack_conv_weight(w, tc=True, flip_transpose=flip)
    none, one = upf.pack_conv_weight(w, tc=True, flip_transpose=flip, tc_only=True)
    assert none is None and torch.equal(one, two_step)


@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("shape", [(2, 20, 31, 3), (1, 64, 96, 3), (1, 9, 12, 1)])
def test_census_loss_kernels_vs_torch(upf, masked, shape):
    """upf_census_loss_fwd/bwd (csrc/loss.cu) against the torch expression of census_loss_torch (utils/loss.py:51-91,
    pinned to the reference by tests/golden/loss_ops.pt) in fp64: value to 1e-5 relative, gradient of the warped image
    to 1e-4 of its largest entry."""
    import upflow_pytorch_b200
    upflow_pytorch_b200.install_dropin()
    from utils.loss import loss_functions
    N, H, W, d = shape
    gen = torch.Generator().manual_seed(6)
    a = torch.rand(N, 3, H, W, generator=gen).cuda()
    b = (a.cpu() + 0.1 * torch.randn(N, 3, H, W, generator=gen)).cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    mask = (torch.rand(N, 1, H, W, generator=gen) > 0.3).float().cuda()
    got = upf.census_loss(a, b, mask if masked else None, 0.4, d)
    (gb,) = torch.autograd.grad(got * 2.0, (b,))
    bd = b.detach().double().requires_grad_()
    loss_functions.use_loss_kernels = False
    try:
        want = loss_functions.census_loss_torch(a.double(), bd, mask.double(), 0.4, False, masked, True, max_distance=d)
    finally:
        loss_functions.use_loss_kernels = True
    (wb,) = torch.autograd.grad(want * 2.0, (bd,))
    _check("value", abs(got.item() - want.item()) / abs(want.item()), 1e-5)
    _check("grad", _rel(gb, wb), 1e-4)
    assert upf.census_loss(a, b, mask if masked else None, 0.4, d).item() == got.item()


def test_boundary_warp_kernels_vs_torch(upf):
    """upf_boundary_warp_fwd/bwd against the torch expression of tools.boundary_dilated_warp.warp_im (utils/tools.py:350-499,
    pinned to the reference by tests/golden/loss_ops.pt) in fp64, with samples inside, on the border of and far outside
    the un-cropped frame.  The flow is a multiple of 1/64 plus 1/128: exact in fp32 and never on a pixel boundary, so
    both sides pick the same corners."""
    import upflow_pytorch_b200
    upflow_pytorch_b200.install_dropin()
    from utils.tools import tools
    N, Hf, Wf, h, w = 2, 40, 56, 24, 32
    gen = torch.Generator().manual_seed(12)
    frame = torch.rand(N, 3, Hf, Wf, generator=gen).cuda()
    start = torch.tensor([[5.0, 7.0], [16.0, 9.0]]).reshape(N, 2, 1, 1).cuda()
    flow = ((torch.randn(N, 2, h, w, generator=gen) * 10 * 64).round() / 64 + 1.0 / 128).cuda()
    flow = flow.contiguous(memory_format=torch.channels_last).requires_grad_()
    got = upf.boundary_warp(frame, flow, start)
    r = _rand(13, N, 3, h, w).cuda()
    (gf,) = torch.autograd.grad((got * r).sum(), (flow,))
    fd = flow.detach().double().requires_grad_()
    tools.boundary_dilated_warp.use_kernel = False
    try:
        want = tools.boundary_dilated_warp.warp_im(frame, fd, start.double())
    finally:
        tools.boundary_dilated_warp.use_kernel = True
    (wf,) = torch.autograd.grad((want * r.double()).sum(), (fd,))
    outside = ((fd[:, 0] + start[:, 0].double() + torch.arange(w, device="cuda")) < 0).float().mean().item()
    assert 0.02 < outside < 0.5, outside                       # the clamped-corner branch is exercised
    _check("warp", _rel(got, want), 2e-5)
    _check("grad flow", _rel(gf, wf), 2e-4)
    # through the drop-in entry point
    again = tools.boundary_dilated_warp.warp_im(frame, flow, start)
    assert torch.equal(again, got)


@pytest.mark.parametrize("level", [(4, 128, 416), (8, 64, 208), (8, 16, 52), (8, 4, 13)])
def test_wgrad_tensor_core_operand_geometry_vs_simt(upf, level):
    """The tensor-core weight gradient over the training step's convolution shapes: blocked, pre-swizzled planar operands
    landed by bulk copies, 1..8 k blocks per ring slot (small operand tiles travel several at a time), one resident CTA
    with a deep ring for 128-wide N tiles -- against the fp32 SIMT weight gradient (TF32 operands: 2e-3 relative L2).
    (32 -> 2 at 4x128x416 is the shape whose last A block read past its ring slot before the slot was sized for it.)"""
    from upflow_pytorch_b200.ops import Slice
    N, h, w = level
    g = torch.Generator().manual_seed(3)
    convs = [(16, 16, 3, 1), (32, 2, 3, 1), (32, 32, 1, 1), (64, 32, 3, 1), (565, 128, 3, 1), (128, 96, 3, 8), (96, 64, 3, 16),
             (243, 128, 3, 1), (531, 32, 3, 1), (184, 3, 3, 1)]
    for (cin, cout, ks, dil) in convs:
        if h * w > 30000 and cin > 32:
            continue
        X = torch.randn(N, h, w, (cin + 3) // 4 * 4, generator=g).cuda()
        G = torch.randn(N, h, w, (cout + 3) // 4 * 4, generator=g).cuda()
        xs, gs = Slice(X, 0, cin), Slice(G, 0, cout)
        gw, gb = upf.k_conv_wgrad(xs, gs, ks, 1, dil, want_bias=True, tensor_cores=True)
        rw, rb = upf.k_conv_wgrad(xs, gs, ks, 1, dil, want_bias=True, tensor_cores=False)
        rel = ((gw - rw).norm() / rw.norm()).item()
        assert rel <= 2e-3, ((cin, cout, ks, dil), rel)
        assert (gb - rb).abs().max().item() <= 1e-3 * rb.abs().max().item() + 1e-4
