"""conv_chain.cu (upf_conv_chain_fwd): a chain of dependent convolutions in one persistent launch must equal the same
layers launched one by one through upf_conv2d_fwd (another, equally fixed, order of the fp32 partial sums over K) and the
fp64 oracle on the TF32 operands; bitwise reproducible; every cluster size / grid size.  -m gpu."""
import pytest
import torch

from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def upf():
    from upflow_pytorch_b200 import _ext, ops
    _ext.load()
    return ops


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _rn_tf32(t):
    return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def _dense_chain(upf, N, H, W, widths, x_ch, dils, seed, last_cout=2, k_last=3):
    """An append-only dense block like FlowEstimatorDense_v2 followed by a small head with a residual and a rounded copy:
    layer i reads channels [0, x_ch + sum(widths[:i])) and writes the next widths[i]; the head reads everything."""
    from upflow_pytorch_b200.ops import Slice
    g = _g(seed)
    ld = (x_ch + sum(widths) + 3) // 4 * 4
    X0 = torch.zeros(N, H, W, ld)
    X0[..., :x_ch] = _rn_tf32(torch.randn(N, H, W, x_ch, generator=g))
    specs = []
    c = x_ch
    for i, wd in enumerate(widths):
        w = torch.randn(wd, c, 3, 3, generator=g) * (2.0 / (c * 9)) ** 0.5
        b = torch.randn(wd, generator=g) * 0.1
        specs.append((w, b, c, wd, 3, dils[i], 0.1))
        c += wd
    w = torch.randn(last_cout, c, k_last, k_last, generator=g) * (2.0 / (c * k_last * k_last)) ** 0.5
    b = torch.randn(last_cout, generator=g) * 0.1
    specs.append((w, b, c, last_cout, k_last, 1, 1.0))
    res = torch.randn(N, H, W, last_cout, generator=g)
    return X0, specs, res, ld


def _run(upf, X0, specs, res, chain, x_ch):
    from upflow_pytorch_b200 import _ext
    from upflow_pytorch_b200.ops import Slice
    X = X0.cuda().clone()
    N, H, W, ld = X.shape
    last_cout = specs[-1][3]
    head = torch.full((N, H, W, last_cout), float("nan"), device="cuda")
    copy = torch.full((N, H, W, 8), float("nan"), device="cuda")
    r = res.cuda()
    packed = [(upf.pack_conv_weight(w.cuda(), tc=True)[1], b.cuda()) for (w, b, *_r) in specs]
    layers = []
    c = x_ch
    for i, (w, b, cin, cout, k, dil, slope) in enumerate(specs):
        wtc, bb = packed[i]
        last = i == len(specs) - 1
        out = Slice(head, 0, cout) if last else Slice(X, c, cout)
        if chain:
            layers.append(upf.chain_layer(Slice(X, 0, cin), wtc, bb, out, k, dil, slope, r if last else None, round_tf32=not last,
                                          out2=Slice(copy, 0, cout) if last else None))
        else:
            upf.k_conv(Slice(X, 0, cin), wtc, bb, out, k, 1, dil, slope, r if last else None,
                       _ext.CONV_TF32 | (0 if last else _ext.CONV_ROUND_OUT))
        c += cout
    if chain:
        upf.k_conv_chain(layers)
    else:
        upf.k_copy(Slice(head, 0, last_cout), Slice(copy, 0, last_cout), round_tf32=True)
    torch.cuda.synchronize()
    return X.cpu(), head.cpu(), copy[..., :last_cout].cpu()


def _oracle(X0, specs, res, x_ch):
    """fp64 convolutions on the TF32 operands, every hidden activation rounded to TF32 like the kernels store it"""
    X = X0.clone()
    c = x_ch
    for i, (w, b, cin, cout, k, dil, slope) in enumerate(specs):
        x = X[..., :cin].permute(0, 3, 1, 2).double()
        y = O.conv2d_direct(x, _rn_tf32(w).double(), b.double(), dil, 1, slope).float().permute(0, 2, 3, 1)
        if i == len(specs) - 1:
            return X, y + res
        X[..., c:c + cout] = _rn_tf32(y)
        c += cout


CASES = [  # (N, H, W, x_ch, widths, dils, last_cout, k_last)
    (2, 6, 20, 128, (128, 128, 96, 64, 32), (1, 1, 1, 1, 1), 2, 3),       # the estimator at 1/64 of a KITTI frame
    (2, 12, 39, 64, (32, 32, 32, 16, 8), (1, 1, 1, 1, 1), 3, 3),          # the SGU block at 1/32
    (2, 24, 78, 100, (128, 96), (2, 4), 2, 3),                            # dilated, 30 pixel tiles (several rounds)
    (1, 5, 7, 36, (16, 48), (8, 16), 5, 1),                               # tiny map, big dilations, 1x1 head, odd widths
    (3, 9, 33, 20, (8, 8, 8), (1, 2, 1), 1, 3),
]


@pytest.mark.parametrize("case", CASES)
def test_chain_equals_layer_by_layer_and_oracle(upf, case):
    N, H, W, x_ch, widths, dils, last_cout, k_last = case
    X0, specs, res, ld = _dense_chain(upf, N, H, W, widths, x_ch, dils, 7, last_cout, k_last)
    Xs, hs, cs = _run(upf, X0, specs, res, False, x_ch)
    Xc, hc, cc = _run(upf, X0, specs, res, True, x_ch)
    Xc2, hc2, cc2 = _run(upf, X0, specs, res, True, x_ch)
    assert torch.equal(Xc, Xc2) and torch.equal(hc, hc2) and torch.equal(cc, cc2)       # bitwise reproducible
    assert not torch.isnan(hc).any() and not torch.isnan(cc).any()
    # hidden activations are stored TF32-rounded: a last-bit difference of the fp32 sum may move one by a TF32 ulp (2^-11 relative)
    scale = Xs.abs().max().item()
    assert (Xc - Xs).abs().max().item() <= 1.5e-3 * scale
    assert (hc - hs).abs().max().item() <= 2e-3
    assert torch.equal(cc, _rn_tf32(hc))
    Xo, ho = _oracle(X0, specs, res, x_ch)
    assert (Xc - Xo).abs().max().item() <= 1.5e-3 * scale
    assert (hc - ho).abs().max().item() <= 2e-3
    # mean agreement is at fp32-rounding level
    assert (hc - ho).abs().mean().item() <= 1e-4


@pytest.mark.parametrize("cs,ncl", [(8, 1), (8, 3), (4, 0), (4, 5), (2, 0), (1, 0), (1, 7)])
def test_chain_cluster_and_grid_sizes(upf, cs, ncl):
    """any cluster size (K split) and any number of clusters (items per cluster: one ... many rounds) gives the layer-by-layer
    result; a single cluster walks every item of every layer alone"""
    from upflow_pytorch_b200 import _ext
    N, H, W, x_ch, widths, dils, last_cout, k_last = CASES[2]
    X0, specs, res, ld = _dense_chain(upf, N, H, W, widths, x_ch, dils, 11, last_cout, k_last)
    Xs, hs, cs_ = _run(upf, X0, specs, res, False, x_ch)
    _ext.check(_ext.load().upf_debug_conv_chain(cs, ncl), "debug_conv_chain")
    try:
        Xc, hc, cc = _run(upf, X0, specs, res, True, x_ch)
    finally:
        _ext.load().upf_debug_conv_chain(8, 0)
    assert (Xc - Xs).abs().max().item() <= 1.5e-3 * Xs.abs().max().item()
    assert (hc - hs).abs().max().item() <= 2e-3


def test_chain_back_to_back_launches(upf):
    """the grid barrier's words are cleared by the last CTA of a launch: many chains in a row on one stream"""
    N, H, W, x_ch, widths, dils, last_cout, k_last = CASES[1]
    X0, specs, res, ld = _dense_chain(upf, N, H, W, widths, x_ch, dils, 13, last_cout, k_last)
    first = _run(upf, X0, specs, res, True, x_ch)
    for _ in range(20):
        again = _run(upf, X0, specs, res, True, x_ch)
        assert all(torch.equal(a, b) for a, b in zip(first, again))


def test_chain_argument_validation(upf):
    from upflow_pytorch_b200 import _ext
    lib = _ext.load()
    assert lib.upf_conv_chain_fwd(None, 1, 1, 4, 4, None) != 0
    L = _ext.ChainLayer()
    arr = (_ext.ChainLayer * 1)(L)
    import ctypes
    assert lib.upf_conv_chain_fwd(ctypes.cast(arr, ctypes.c_void_p), 1, 1, 4, 4, None) != 0          # null pointers
    assert lib.upf_conv_chain_fwd(ctypes.cast(arr, ctypes.c_void_p), 17, 1, 4, 4, None) != 0         # too many layers
    assert lib.upf_debug_conv_chain(3, 0) != 0


def test_engine_with_and_without_chains():
    """the whole two-frame forward at KITTI size: coarse levels as persistent chains vs one launch per convolution"""
    import bench
    from upflow_pytorch_b200.engine import DecoderEngine
    H, W, B = 375, 1242, 1
    sd = {k: v.cuda() for k, v in bench.make_weights().items()}
    im1, im2 = bench.synth_inputs(B, H, W, 1234)
    flows = []
    for chain in (False, True):
        eng = DecoderEngine(sd, precision="tf32", mask_threshold=0.9999)
        eng.chain = chain
        with torch.no_grad():
            f, b, lv = eng.forward(im1.cuda(), im2.cuda())
        torch.cuda.synchronize()
        flows.append((f.clone(), b.clone(), [x.clone() for pair in lv for x in pair]))
    epe = (flows[0][0] - flows[1][0]).pow(2).sum(1).sqrt()
    print("chain vs per-layer: mean EPE %.3g px, max %.3g px" % (epe.mean().item(), epe.max().item()))
    assert epe.mean().item() <= 2e-3
    # the graph replays it bit-identically
    eng = DecoderEngine(sd, precision="tf32", mask_threshold=0.9999)
    eng.chain = True
    with torch.no_grad():
        g = eng.capture(B, H, W)
    outs = []
    for _ in range(3):
        g(im1.cuda(), im2.cuda())
        torch.cuda.synchronize()
        outs.append(g.flow_f.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert torch.equal(outs[0], flows[1][0])
