"""TEST / BASELINE INFRASTRUCTURE ONLY -- op-for-op CPU port of the reference's
decoder hot path, used (a) as the CPU baseline that ``bench.py`` times on the GPU
box's host cores (``cpu_baseline.kind == "port"``: the Python reference itself
cannot travel to the box) and (b) as the full-size oracle for end-to-end parity.

Unlike ``cpu_oracle.py`` (a first-principles restatement with explicit index
arithmetic) this port issues the SAME library calls in the SAME order as the
reference does on the installed torch -- ``F.conv2d``, ``F.grid_sample`` with
the default ``align_corners``, ``F.interpolate(align_corners=True)``, the
``F.unfold`` correlation -- so on CPU it is bit-identical to the reference
(checked in tests/test_oracle_vs_reference.py) and exactly as fast: timing it
is timing the reference's own CPU path.  It is functional (weights come in as a
state-dict with the reference's key names).

The product package never imports this file.
"""
import torch
import torch.nn.functional as F

SLOPE = 0.1
# 1.0 is the reference (model/pwc_modules.py:206).  Tests may set 0.9999 on BOTH sides (this port and the CUDA
# engine) to demonstrate parity without the 1-ulp mask flips; never changed for baselines or golden vectors.
MASK_THRESHOLD = 1.0


def _conv(x, sd, key, stride=1, dilation=1, relu=True):
    # conv(): Conv2d(pad=((k-1)*dil)//2) + LeakyReLU(0.1)   model/pwc_modules.py:10-31
    w = sd[key + ".0.weight"]
    k = w.shape[-1]
    y = F.conv2d(x, w, sd[key + ".0.bias"], stride=stride, padding=((k - 1) * dilation) // 2, dilation=dilation)
    return F.leaky_relu(y, SLOPE) if relu else y


def corr_unfold(in1, in2, d=4):
    # Corr_pyTorch.forward, kernel_size=1   utils/pytorch_correlation.py:27-50
    bz, cn, hei, wid = in1.shape
    f1 = F.unfold(in1, kernel_size=1, padding=0, stride=1)
    f2 = F.unfold(in2, kernel_size=1, padding=0, stride=1)
    sk = f2.shape[1]
    f2_ = f2.reshape(bz, sk, hei, wid).reshape(bz * sk, hei, wid).unsqueeze(1)
    f2 = F.unfold(f2_, kernel_size=(hei, wid), padding=d, stride=1)
    _, kn, wn = f2.shape
    f2_2 = f2.reshape(bz, sk, kn, wn).transpose(1, 3).transpose(2, 3)
    res = (f2_2 * f1.unsqueeze(1)).mean(dim=2)
    return res.reshape(bz, wn, hei, wid)


def _vgrid(flow):
    # mesh + flow -> [-1,1]   model/pwc_modules.py:186-199, utils/tools.py:1284-1299
    B, _, H, W = flow.shape
    # (the reference builds the mesh on the host and moves it with .cuda() when x.is_cuda, pwc_modules.py:192-193)
    xx = torch.arange(0, W, device=flow.device).view(1, -1).repeat(H, 1).view(1, 1, H, W).repeat(B, 1, 1, 1)
    yy = torch.arange(0, H, device=flow.device).view(-1, 1).repeat(1, W).view(1, 1, H, W).repeat(B, 1, 1, 1)
    vgrid = torch.cat((xx, yy), 1).float() + flow
    vgrid[:, 0] = 2.0 * vgrid[:, 0] / max(W - 1, 1) - 1.0
    vgrid[:, 1] = 2.0 * vgrid[:, 1] / max(H - 1, 1) - 1.0
    return vgrid.permute(0, 2, 3, 1)


def warp_mask(x, flow):
    # WarpingLayer_no_div   model/pwc_modules.py:184-207
    vgrid = _vgrid(flow)
    xw = F.grid_sample(x, vgrid, padding_mode="zeros", align_corners=False)
    mask = F.grid_sample(torch.ones_like(x), vgrid, align_corners=False)
    return xw * (mask >= MASK_THRESHOLD).float()


def torch_warp(x, flow):
    # tools.torch_warp   utils/tools.py:1274-1304
    return F.grid_sample(x, _vgrid(flow), padding_mode="zeros", align_corners=False)


def upsample2d_flow_as(x, h, w, if_rate=False):
    # model/pwc_modules.py:77-90
    res = F.interpolate(x, [h, w], mode="bilinear", align_corners=True)
    if if_rate:
        h_, w_ = x.shape[2:]
        u, v = res.chunk(2, dim=1)
        res = torch.cat([u * (w / w_), v * (h / h_)], dim=1)
    return res


# (if_norm_before_cost_volume, norm_moments_across_channels, norm_moments_across_images) of UPFlow_net.config
# (model/upflow.py:311-313).  test.py:24-26 runs (True, False, False); the class default is (False, True, True).
NORM_MODE = (True, False, False)


def normalize_pair(fa, fb):
    # network_tools.normalize_features((fa, fb), normalize=True, center=True, ...)   model/upflow.py:94-137, :549-555
    if_norm, across_ch, across_img = NORM_MODE
    if not if_norm:
        return fa, fb
    axes = [1, 2, 3] if across_ch else [2, 3]
    means = [torch.mean(f, dim=axes, keepdim=True) for f in (fa, fb)]
    variances = [torch.var(f, dim=axes, keepdim=True) for f in (fa, fb)]
    if across_img:
        means = [torch.mean(torch.stack(means, dim=0), dim=(0,))] * 2
        variances = [torch.var(torch.stack(variances, dim=0), dim=(0,))] * 2
    stds = [torch.sqrt(v + 1e-16) for v in variances]
    return (fa - means[0]) / stds[0], (fb - means[1]) / stds[1]


def normalize(f):
    # network_tools.normalize_features, per image per channel   model/upflow.py:108-135
    mean = torch.mean(f, dim=[2, 3], keepdim=True)
    var = torch.var(f, dim=[2, 3], keepdim=True)
    return (f - mean) / torch.sqrt(var + 1e-16)


def dense(x, sd, prefix):
    # FlowEstimatorDense_v2.forward   model/pwc_modules.py:279-286
    for n in ("conv1", "conv2", "conv3", "conv4", "conv5"):
        x = torch.cat([_conv(x, sd, f"{prefix}.{n}"), x], dim=1)
    return x, _conv(x, sd, f"{prefix}.conv_last", relu=False)


def context(x, sd, prefix="context_networks"):
    # ContextNetwork_v2_.forward   model/pwc_modules.py:401-412
    for i, dil in enumerate((1, 2, 4, 8, 16, 1, 1)):
        x = _conv(x, sd, f"{prefix}.convs.{i}", dilation=dil, relu=i != 6)
    return x


def sgu(flow_init, f1, f2, sd, output_level_flow=None):
    # sgu_model.forward   model/upflow.py:71-89
    h, w = f1.shape[2:]
    if flow_init.shape[2] != h or flow_init.shape[3] != w:
        flow_init = upsample2d_flow_as(flow_init, h, w, if_rate=True)
    f2w = warp_mask(f2, flow_init)
    _, x_out = dense(torch.cat((f1, f2w), dim=1), sd, "sgi_model.dense_estimator_mask")
    inter_flow = x_out[:, :2]
    inter_mask = torch.sigmoid(x_out[:, 2:3])
    if output_level_flow is not None:
        H, W = output_level_flow.shape[2:]
        inter_flow = upsample2d_flow_as(inter_flow, H, W, if_rate=True)
        inter_mask = upsample2d_flow_as(inter_mask, H, W)
        flow_init = output_level_flow
    flow_up = torch_warp(flow_init, inter_flow) * (1 - inter_mask) + flow_init * inter_mask
    return flow_up, inter_flow, inter_mask


def decode_level(level, flow_1, flow_2, x1, x1_1x1, x2, x2_1x1, sd, use_sgu=True, taps=None):
    # UPFlow_net.decode_level_res   model/upflow.py:535-573
    h, w = x1.shape[2:]
    f1u = upsample2d_flow_as(flow_1, h, w, if_rate=True)
    f2u = upsample2d_flow_as(flow_2, h, w, if_rate=True)
    if level == 0:
        x2w, x1w = x2, x1
    else:
        if use_sgu:
            f1u = sgu(f1u, x1_1x1, x2_1x1, sd)[0]
            f2u = sgu(f2u, x2_1x1, x1_1x1, sd)[0]
        x2w = warp_mask(x2, f1u)
        x1w = warp_mask(x1, f2u)
    n1, n2w = normalize_pair(x1, x2w)
    n2, n1w = normalize_pair(x2, x1w)
    c1 = F.leaky_relu(corr_unfold(n1, n2w), SLOPE)
    c2 = F.leaky_relu(corr_unfold(n2, n1w), SLOPE)
    x5_1, r1 = dense(torch.cat([c1, x1_1x1, f1u], dim=1), sd, "flow_estimators")
    x5_2, r2 = dense(torch.cat([c2, x2_1x1, f2u], dim=1), sd, "flow_estimators")
    fine1 = context(torch.cat([x5_1, f1u + r1], dim=1), sd)
    fine2 = context(torch.cat([x5_2, f2u + r2], dim=1), sd)
    if taps is not None:
        taps.append(dict(level=level, flow_1_up=f1u, flow_2_up=f2u, x2_warp=x2w, n1=n1, n2w=n2w, corr_1=c1,
                         x5_1=x5_1, res_1=r1, fine_1=fine1))
    return f1u, f2u, r1 + fine1, r2 + fine2


def pyramid(x, sd, prefix="feature_pyramid_extractor"):
    # FeatureExtractor.forward   model/pwc_modules.py:136-142
    out = []
    for l in range(6):
        x = _conv(x, sd, f"{prefix}.convs.{l}.0", stride=2)
        x = _conv(x, sd, f"{prefix}.convs.{l}.1")
        out.append(x)
    return out[::-1]


def output_conv(x, sd, prefix="sgi_model.upsample_output_conv"):
    # model/upflow.py:66-69
    for i, s in enumerate((1, 2, 1, 2)):
        x = _conv(x, sd, f"{prefix}.{i}", stride=s)
    return x


def forward_2_frame(im1, im2, sd, use_sgu=True, taps=None):
    # UPFlow_net.forward_2_frame_v3   model/upflow.py:494-533
    p1 = pyramid(im1, sd) + [im1]
    p2 = pyramid(im2, sd) + [im2]
    B, _, h0, w0 = p1[0].shape
    flow_f = torch.zeros(B, 2, h0, w0, device=im1.device)
    flow_b = torch.zeros(B, 2, h0, w0, device=im1.device)
    levels = []
    for l in range(5):
        levels.append((p1[l], _conv(p1[l], sd, f"conv_1x1.{l}"), p2[l], _conv(p2[l], sd, f"conv_1x1.{l}")))
    flows = []
    for l, (x1, x1a, x2, x2a) in enumerate(levels):
        flow_f, flow_b, rf, rb = decode_level(l, flow_f, flow_b, x1, x1a, x2, x2a, sd, use_sgu, taps)
        flow_f = flow_f + rf
        flow_b = flow_b + rb
        flows.append([flow_f, flow_b])
    H, W = im1.shape[2:]
    out_f = upsample2d_flow_as(flow_f, H, W, if_rate=True)
    out_b = upsample2d_flow_as(flow_b, H, W, if_rate=True)
    if use_sgu:
        g1 = output_conv(im1, sd)
        g2 = output_conv(im2, sd)
        out_f = sgu(flow_f, g1, g2, sd, output_level_flow=out_f)[0]
        out_b = sgu(flow_b, g2, g1, sd, output_level_flow=out_b)[0]
    return out_f, out_b, flows[::-1]


# --- deterministic weights: independent of module construction order ---------
REF_SHAPES = None


def reference_param_shapes():
    """name -> shape of the 80 tensors of UPFlow_net(if_sgu_upsample=True)
    (model/upflow.py:329-361)."""
    shapes = {}

    def add(key, cin, cout, k=3):
        shapes[key + ".0.weight"] = (cout, cin, k, k)
        shapes[key + ".0.bias"] = (cout,)

    chs = [3, 16, 32, 64, 96, 128, 196]
    for l in range(6):
        add(f"feature_pyramid_extractor.convs.{l}.0", chs[l], chs[l + 1])
        add(f"feature_pyramid_extractor.convs.{l}.1", chs[l + 1], chs[l + 1])
    n = 115
    for name, c in zip(("conv1", "conv2", "conv3", "conv4", "conv5"), (128, 128, 96, 64, 32)):
        add(f"flow_estimators.{name}", n, c)
        n += c
    add("flow_estimators.conv_last", n, 2)
    cin = 565
    for i, c in enumerate((128, 128, 128, 96, 64, 32, 2)):
        add(f"context_networks.convs.{i}", cin, c)
        cin = c
    for l, c in enumerate((196, 128, 96, 64, 32)):
        add(f"conv_1x1.{l}", c, 32, k=1)
    n = 64
    for name, c in zip(("conv1", "conv2", "conv3", "conv4", "conv5"), (32, 32, 32, 16, 8)):
        add(f"sgi_model.dense_estimator_mask.{name}", n, c)
        n += c
    add("sgi_model.dense_estimator_mask.conv_last", n, 3)
    for i, (a, b) in enumerate(((3, 16), (16, 16), (16, 32), (32, 32))):
        add(f"sgi_model.upsample_output_conv.{i}", a, b)
    return shapes


def det_state_dict(seed=0, shapes=None, bias_scale=0.02, head_gain=0.1):
    """Deterministic MSRA-like weights keyed by NAME (crc32 of the key seeds a
    private CPU generator), so any module tree with the reference's key names
    gets identical tensors regardless of construction order.  The flow heads
    (``conv_last``, ``context_networks.convs.6``) are scaled by ``head_gain``
    so an untrained net produces flows of a few pixels, not tens."""
    import zlib
    shapes = shapes or reference_param_shapes()
    sd = {}
    for k in sorted(shapes):
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(k.encode())) % (2 ** 31))
        shp = shapes[k]
        if k.endswith("weight"):
            fan_in = shp[1] * shp[2] * shp[3]
            sd[k] = torch.randn(shp, generator=g) * (2.0 / fan_in) ** 0.5
            if "conv_last" in k or "context_networks.convs.6" in k:
                sd[k] = sd[k] * head_gain
        else:
            sd[k] = torch.randn(shp, generator=g) * bias_scale
    return sd
