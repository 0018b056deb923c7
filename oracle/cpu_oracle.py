"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the UPFlow decoder hot path.

This file is the oracle the CUDA kernels are checked against.  Only tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import it; the product package never does and fails
loudly when its CUDA library is missing.

Every function restates one reference routine from first principles (explicit
index arithmetic on CPU tensors -- no F.grid_sample / F.interpolate /
F.unfold) and cites the reference lines it follows (paths relative to
/root/reference).  Parity status: the reference ships NO tests and NO golden
vectors, so the restatement is pinned by executing the
unmodified reference itself: ``tests/test_oracle_vs_reference.py`` (runs where
/root/reference is mounted) and the committed fixtures in ``tests/golden/``
produced by ``oracle/make_golden.py`` from the shimmed reference.

All tensors are NCHW float32 unless ``dtype=torch.float64`` is requested (the
fp64 variants bound the rounding of the fp32 reference itself).
"""
import math

import torch

LRELU_SLOPE = 0.1  # nn.LeakyReLU(0.1): model/pwc_modules.py:29, model/upflow.py:342


# --------------------------------------------------------------------------
# a1/a2/a3  cost volume  (utils/pytorch_correlation.py:27-50;
#           model/correlation_package/correlation_cuda_kernel.cu:60-108;
#           LeakyReLU at model/upflow.py:563-564)
# --------------------------------------------------------------------------
def correlation(f1, f2, max_disp=4, leaky_slope=None):
    """out[b,(dy+d)*(2d+1)+(dx+d),y,x] = mean_c f1[b,c,y,x]*f2[b,c,y+dy,x+dx],
    f2 zero outside the image; dy is the slow displacement index."""
    B, C, H, W = f1.shape
    d = max_disp
    D = 2 * d + 1
    f2p = torch.zeros(B, C, H + 2 * d, W + 2 * d, dtype=f2.dtype)
    f2p[:, :, d:d + H, d:d + W] = f2
    out = torch.empty(B, D * D, H, W, dtype=f1.dtype)
    for iy in range(D):
        for ix in range(D):
            out[:, iy * D + ix] = (f1 * f2p[:, :, iy:iy + H, ix:ix + W]).sum(1) / C
    if leaky_slope is not None:
        out = torch.where(out < 0, out * leaky_slope, out)
    return out


def correlation_backward(f1, f2, grad_out, max_disp=4):
    """Gradients of ``correlation`` (no LeakyReLU) wrt f1 and f2
    (correlation_cuda_kernel.cu:116-300)."""
    B, C, H, W = f1.shape
    d = max_disp
    D = 2 * d + 1
    f2p = torch.zeros(B, C, H + 2 * d, W + 2 * d, dtype=f2.dtype)
    f2p[:, :, d:d + H, d:d + W] = f2
    g1 = torch.zeros_like(f1)
    g2p = torch.zeros_like(f2p)
    for iy in range(D):
        for ix in range(D):
            g = grad_out[:, iy * D + ix].unsqueeze(1) / C
            g1 += g * f2p[:, :, iy:iy + H, ix:ix + W]
            g2p[:, :, iy:iy + H, ix:ix + W] += g * f1
    return g1, g2p[:, :, d:d + H, d:d + W].contiguous()


# --------------------------------------------------------------------------
# a4  bilinear warp + validity mask (model/pwc_modules.py:184-207) and the
#     mask-free variant tools.torch_warp (utils/tools.py:1284-1304).
#     grid_sample arithmetic: ATen native/cuda/GridSampler.cuh:23-31 (unnormalise)
#     and the bilinear corner weights of grid_sampler_2d.
# --------------------------------------------------------------------------
def _sample_coords(flow, align_corners):
    """Pixel-space sampling position (ix, iy) for every output pixel.

    The reference first maps ``x+u`` to [-1,1] with three separately rounded
    fp32 ops (pwc_modules.py:195-198) and grid_sample maps it back.  The
    round trip is NOT the identity in fp32 and decides the ``mask >= 1.0``
    comparison, so it is restated op for op."""
    B, _, H, W = flow.shape
    dt = flow.dtype
    xs = torch.arange(W, dtype=dt).view(1, 1, W).expand(B, H, W)
    ys = torch.arange(H, dtype=dt).view(1, H, 1).expand(B, H, W)
    gx = (2.0 * (xs + flow[:, 0])) / max(W - 1, 1) - 1.0
    gy = (2.0 * (ys + flow[:, 1])) / max(H - 1, 1) - 1.0
    if align_corners:
        ix = ((gx + 1.0) / 2) * (W - 1)
        iy = ((gy + 1.0) / 2) * (H - 1)
    else:
        # ((g+1)*size-1)/2 is evaluated as ONE fused multiply-add by both the
        # ATen CPU vector kernel and nvcc (-fmad): emulate the single rounding
        # in float64 (exact product of two fp32 numbers) -> fp32.
        if dt == torch.float32:
            ix = ((gx + 1.0).double() * (W / 2) - 0.5).float()
            iy = ((gy + 1.0).double() * (H / 2) - 0.5).float()
        else:
            ix = (gx + 1.0) * (W / 2) - 0.5
            iy = (gy + 1.0) * (H / 2) - 0.5
    return ix, iy


def _bilinear_gather(x, ix, iy):
    """zeros-padding bilinear sample of x[B,C,H,W] at (ix,iy)[B,H',W'];
    returns (sample, sum of in-bounds corner weights)."""
    B, C, H, W = x.shape
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    x1 = x0 + 1
    y1 = y0 + 1
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    out = torch.zeros(B, C, *ix.shape[1:], dtype=x.dtype)
    wsum = torch.zeros_like(ix)
    xf = x.reshape(B, C, H * W)
    for xx, yy, ww in ((x0, y0, w_nw), (x1, y0, w_ne), (x0, y1, w_sw), (x1, y1, w_se)):
        inb = (xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)
        xi = xx.clamp(0, W - 1).long()
        yi = yy.clamp(0, H - 1).long()
        lin = (yi * W + xi).view(B, 1, -1).expand(B, C, -1)
        v = torch.gather(xf, 2, lin).view(B, C, *ix.shape[1:])
        wz = torch.where(inb, ww, torch.zeros_like(ww))
        out = out + v * wz.unsqueeze(1)
        wsum = wsum + wz
    return out, wsum


def warp_mask(x, flow, align_corners=False):
    """WarpingLayer_no_div.forward (pwc_modules.py:184-207): sample x at
    pixel+flow, zero every pixel whose in-bounds corner weights sum < 1.0."""
    ix, iy = _sample_coords(flow, align_corners)
    out, wsum = _bilinear_gather(x, ix, iy)
    mask = (wsum >= 1.0).to(x.dtype).unsqueeze(1)
    return out * mask


def warp_mask_with_mask(x, flow, align_corners=False):
    ix, iy = _sample_coords(flow, align_corners)
    out, wsum = _bilinear_gather(x, ix, iy)
    mask = (wsum >= 1.0).to(x.dtype).unsqueeze(1)
    return out * mask, mask


def torch_warp(x, flow, align_corners=False):
    """tools.torch_warp (utils/tools.py:1274-1304): same sampling, no mask."""
    ix, iy = _sample_coords(flow, align_corners)
    return _bilinear_gather(x, ix, iy)[0]


# --------------------------------------------------------------------------
# a5  per-image per-channel normalisation (model/upflow.py:94-137 with
#     moments_across_channels=False, moments_across_images=False, test.py:24-26)
# --------------------------------------------------------------------------
def normalize_features(f):
    B, C, H, W = f.shape
    n = H * W
    mean = f.sum(dim=(2, 3), keepdim=True) / n
    cen = f - mean
    var = (cen * cen).sum(dim=(2, 3), keepdim=True) / (n - 1)   # unbiased, torch.var default
    std = torch.sqrt(var + 1e-16)
    return cen / std


# --------------------------------------------------------------------------
# a6  bilinear resize (align_corners=True) with flow rescale
#     (model/pwc_modules.py:77-90; ATen native/cuda/UpSample.cuh:100-124)
# --------------------------------------------------------------------------
def _axis_taps(n_in, n_out, dtype):
    scale = (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0
    scale = torch.tensor(scale, dtype=dtype)
    src = scale * torch.arange(n_out, dtype=dtype)
    i0 = src.long()                      # truncation == floor (src >= 0)
    i1 = i0 + (i0 < n_in - 1).long()
    l1 = src - i0.to(dtype)
    l0 = 1.0 - l1
    return i0, i1, l0, l1


def resize_bilinear_ac(x, h, w):
    """F.interpolate(x, [h, w], mode='bilinear', align_corners=True)."""
    B, C, H, W = x.shape
    y0, y1, ly0, ly1 = _axis_taps(H, h, x.dtype)
    x0, x1, lx0, lx1 = _axis_taps(W, w, x.dtype)
    top = x[:, :, y0]
    bot = x[:, :, y1]
    lx0 = lx0.view(1, 1, 1, w)
    lx1 = lx1.view(1, 1, 1, w)
    ly0 = ly0.view(1, 1, h, 1)
    ly1 = ly1.view(1, 1, h, 1)
    return ly0 * (lx0 * top[..., x0] + lx1 * top[..., x1]) + ly1 * (lx0 * bot[..., x0] + lx1 * bot[..., x1])


def upsample2d_flow_as(flow, h, w, if_rate=True):
    """pwc_modules.py:77-90: resize, then u *= w/w_in, v *= h/h_in (ratio of
    SIZES, python floats rounded to fp32 at the multiply)."""
    _, C, h_in, w_in = flow.shape
    res = resize_bilinear_ac(flow, h, w)
    if if_rate:
        assert C == 2
        res = torch.stack([res[:, 0] * (w / w_in), res[:, 1] * (h / h_in)], dim=1)
    return res


# --------------------------------------------------------------------------
# a8/a9  conv + LeakyReLU stacks (model/pwc_modules.py:10-31, :252-286, :398-412)
# --------------------------------------------------------------------------
def conv2d_direct(x, weight, bias, dilation=1, stride=1, leaky_slope=None):
    """3x3 (or 1x1) cross-correlation with zero padding ((k-1)*dil)//2, restated
    as a sum over taps of shifted 1x1 contractions (no F.conv2d)."""
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    ph = ((kh - 1) * dilation) // 2
    pw = ((kw - 1) * dilation) // 2
    Ho = (H + 2 * ph - dilation * (kh - 1) - 1) // stride + 1
    Wo = (W + 2 * pw - dilation * (kw - 1) - 1) // stride + 1
    xp = torch.zeros(B, Cin, H + 2 * ph, W + 2 * pw, dtype=x.dtype)
    xp[:, :, ph:ph + H, pw:pw + W] = x
    out = torch.zeros(B, Cout, Ho, Wo, dtype=x.dtype)
    for ky in range(kh):
        for kx in range(kw):
            win = xp[:, :, ky * dilation: ky * dilation + (Ho - 1) * stride + 1: stride,
                     kx * dilation: kx * dilation + (Wo - 1) * stride + 1: stride]
            out += torch.einsum("bchw,oc->bohw", win, weight[:, :, ky, kx])
    out += bias.view(1, -1, 1, 1)
    if leaky_slope is not None:
        out = torch.where(out < 0, out * leaky_slope, out)
    return out


def dense_block(x, params, prefix):
    """FlowEstimatorDense_v2 / sgu FlowEstimatorDense_temp forward
    (pwc_modules.py:279-286, model/upflow.py:52-60): every conv output is
    PREPENDED to its input; conv_last has no activation."""
    for name in ("conv1", "conv2", "conv3", "conv4", "conv5"):
        y = conv2d_direct(x, params[f"{prefix}.{name}.0.weight"], params[f"{prefix}.{name}.0.bias"],
                          leaky_slope=LRELU_SLOPE)
        x = torch.cat([y, x], dim=1)
    out = conv2d_direct(x, params[f"{prefix}.conv_last.0.weight"], params[f"{prefix}.conv_last.0.bias"])
    return x, out


CONTEXT_DILATIONS = (1, 2, 4, 8, 16, 1, 1)  # pwc_modules.py:401-409


def context_network(x, params, prefix="context_networks"):
    for i, dil in enumerate(CONTEXT_DILATIONS):
        x = conv2d_direct(x, params[f"{prefix}.convs.{i}.0.weight"], params[f"{prefix}.convs.{i}.0.bias"],
                          dilation=dil, leaky_slope=None if i == 6 else LRELU_SLOPE)
    return x


# --------------------------------------------------------------------------
# a7  self-guided upsample (model/upflow.py:71-89)
# --------------------------------------------------------------------------
def sgu_blend(flow_init, inter_flow, inter_mask, align_corners=False):
    """flow_up = torch_warp(flow_init, inter_flow)*(1-m) + flow_init*m
    (model/upflow.py:88); inter_mask is already sigmoided."""
    return torch_warp(flow_init, inter_flow, align_corners) * (1 - inter_mask) + flow_init * inter_mask


def sgu_forward(flow_init, f1, f2, params, output_level_flow=None, align_corners=False):
    h, w = f1.shape[2:]
    if flow_init.shape[2] != h or flow_init.shape[3] != w:
        flow_init = upsample2d_flow_as(flow_init, h, w)
    f2w = warp_mask(f2, flow_init, align_corners)
    _, x_out = dense_block(torch.cat([f1, f2w], 1), params, "sgi_model.dense_estimator_mask")
    inter_flow = x_out[:, :2]
    inter_mask = torch.sigmoid(x_out[:, 2:3])
    if output_level_flow is not None:
        H, W = output_level_flow.shape[2:]
        inter_flow = upsample2d_flow_as(inter_flow, H, W)
        inter_mask = resize_bilinear_ac(inter_mask, H, W)
        flow_init = output_level_flow
    return sgu_blend(flow_init, inter_flow, inter_mask, align_corners)


# --------------------------------------------------------------------------
# a10  per-level decoder and the two-frame driver
#      (model/upflow.py:535-573, :494-533)
# --------------------------------------------------------------------------
def decode_level(level, flow_1, flow_2, x1, x1_1x1, x2, x2_1x1, params, sgu=True, align_corners=False,
                 taps=None):
    h, w = x1.shape[2:]
    flow_1_up = upsample2d_flow_as(flow_1, h, w)
    flow_2_up = upsample2d_flow_as(flow_2, h, w)
    if level == 0:
        x2_warp, x1_warp = x2, x1
    else:
        if sgu:
            flow_1_up = sgu_forward(flow_1_up, x1_1x1, x2_1x1, params, align_corners=align_corners)
            flow_2_up = sgu_forward(flow_2_up, x2_1x1, x1_1x1, params, align_corners=align_corners)
        x2_warp = warp_mask(x2, flow_1_up, align_corners)
        x1_warp = warp_mask(x1, flow_2_up, align_corners)
    n1, n2w = normalize_features(x1), normalize_features(x2_warp)
    n2, n1w = normalize_features(x2), normalize_features(x1_warp)
    corr_1 = correlation(n1, n2w, 4, LRELU_SLOPE)
    corr_2 = correlation(n2, n1w, 4, LRELU_SLOPE)
    x5_1, res_1 = dense_block(torch.cat([corr_1, x1_1x1, flow_1_up], 1), params, "flow_estimators")
    x5_2, res_2 = dense_block(torch.cat([corr_2, x2_1x1, flow_2_up], 1), params, "flow_estimators")
    fine_1 = context_network(torch.cat([x5_1, flow_1_up + res_1], 1), params)
    fine_2 = context_network(torch.cat([x5_2, flow_2_up + res_2], 1), params)
    if taps is not None:
        taps.append(dict(level=level, flow_1_up=flow_1_up, flow_2_up=flow_2_up, x2_warp=x2_warp,
                         corr_1=corr_1, x5_1=x5_1, res_1=res_1, fine_1=fine_1))
    return flow_1_up, flow_2_up, res_1 + fine_1, res_2 + fine_2


def feature_pyramid(x, params, prefix="feature_pyramid_extractor"):
    """FeatureExtractor.forward (pwc_modules.py:136-142), coarsest first."""
    feats = []
    for l in range(6):
        x = conv2d_direct(x, params[f"{prefix}.convs.{l}.0.0.weight"], params[f"{prefix}.convs.{l}.0.0.bias"],
                          stride=2, leaky_slope=LRELU_SLOPE)
        x = conv2d_direct(x, params[f"{prefix}.convs.{l}.1.0.weight"], params[f"{prefix}.convs.{l}.1.0.bias"],
                          leaky_slope=LRELU_SLOPE)
        feats.append(x)
    return feats[::-1]


def sgu_output_conv(x, params, prefix="sgi_model.upsample_output_conv"):
    """model/upflow.py:66-69: 3->16, 16->16 s2, 16->32, 32->32 s2."""
    for i, s in enumerate((1, 2, 1, 2)):
        x = conv2d_direct(x, params[f"{prefix}.{i}.0.weight"], params[f"{prefix}.{i}.0.bias"], stride=s,
                          leaky_slope=LRELU_SLOPE)
    return x


def forward_2_frame(im1, im2, params, sgu=True, align_corners=False, taps=None):
    """UPFlow_net.forward_2_frame_v3 (model/upflow.py:494-533)."""
    p1 = feature_pyramid(im1, params) + [im1]
    p2 = feature_pyramid(im2, params) + [im2]
    B, _, h0, w0 = p1[0].shape
    flow_f = torch.zeros(B, 2, h0, w0, dtype=im1.dtype)
    flow_b = torch.zeros(B, 2, h0, w0, dtype=im1.dtype)
    flows = []
    for level in range(5):
        x1, x2 = p1[level], p2[level]
        w_, b_ = params[f"conv_1x1.{level}.0.weight"], params[f"conv_1x1.{level}.0.bias"]
        x1_1x1 = conv2d_direct(x1, w_, b_, leaky_slope=LRELU_SLOPE)
        x2_1x1 = conv2d_direct(x2, w_, b_, leaky_slope=LRELU_SLOPE)
        flow_f, flow_b, res_f, res_b = decode_level(level, flow_f, flow_b, x1, x1_1x1, x2, x2_1x1, params,
                                                    sgu=sgu, align_corners=align_corners, taps=taps)
        flow_f = flow_f + res_f
        flow_b = flow_b + res_b
        flows.append([flow_f, flow_b])
    H, W = im1.shape[2:]
    flow_f_out = upsample2d_flow_as(flow_f, H, W)
    flow_b_out = upsample2d_flow_as(flow_b, H, W)
    if sgu:
        g1 = sgu_output_conv(im1, params)
        g2 = sgu_output_conv(im2, params)
        flow_f_out = sgu_forward(flow_f, g1, g2, params, output_level_flow=flow_f_out, align_corners=align_corners)
        flow_b_out = sgu_forward(flow_b, g2, g1, params, output_level_flow=flow_b_out, align_corners=align_corners)
    return flow_f_out, flow_b_out, flows[::-1]


# --------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------
def synthetic_pair(H, W, seed=1234, batch=1):
    """Smooth textured pair with known flow (u,v)=(-3,+2): im1 = base[8:,8:],
    im2 = base[6:,11:] of a bicubically upsampled random field."""
    g = torch.Generator().manual_seed(seed)
    lo = torch.rand(batch, 3, H // 8 + 4, W // 8 + 4, generator=g)
    base = torch.nn.functional.interpolate(lo, size=(H + 16, W + 16), mode="bicubic", align_corners=False) - 0.5
    im1 = base[:, :, 8:8 + H, 8:8 + W].contiguous()
    im2 = base[:, :, 6:6 + H, 11:11 + W].contiguous()
    return im1, im2


def epe(a, b):
    return torch.sqrt(((a - b) ** 2).sum(1)).mean().item()
