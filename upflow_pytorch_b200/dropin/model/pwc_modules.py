"""Drop-in for the reference's model/pwc_modules.py: same names, signatures,
return values and state-dict keys, with the arithmetic routed to
libupflow_b200.so.  Only the symbols model/upflow.py imports
(model/upflow.py:7-8) are provided; the legacy variants it never instantiates
(WarpingLayer, FlowEstimatorDense, _v3, ContextNetwork, _v2, Occ*) are not.
"""
from __future__ import absolute_import, division, print_function

import logging

import torch
import torch.nn as nn

from upflow_pytorch_b200 import _ext, ops
from upflow_pytorch_b200.ops import Slice
from utils.tools import tools

_PRECISION = {"mode": _ext.CONV_FP32}


def set_conv_precision(mode):
    """'fp32' (SIMT, strict) or 'tf32' (tcgen05 tensor cores) for module-level convs."""
    _PRECISION["mode"] = {"fp32": _ext.CONV_FP32, "tf32": _ext.CONV_TF32, "tf32x3": _ext.CONV_TF32}[mode]


class _ConvBlock(nn.Sequential):
    """Conv2d + LeakyReLU(0.1) as ONE fused kernel.  Child '0' is a plain
    nn.Conv2d so parameter names stay `<name>.0.weight/.bias`."""

    def __init__(self, conv2d, is_relu):
        mods = [conv2d] + ([nn.LeakyReLU(0.1, inplace=True)] if is_relu else [])
        super().__init__(*mods)
        self.is_relu = is_relu
        self._packed = None

    def packed(self, precision):
        c = self[0]
        key = (c.weight._version, c.weight.data_ptr(), c.bias._version, precision)
        if self._packed is None or self._packed[0] != key:
            tc = precision == _ext.CONV_TF32 and c.stride[0] == 1
            w, w_tc = ops.pack_conv_weight(c.weight, tc=tc)
            self._packed = (key, w_tc if tc else w, c.bias.detach().float().contiguous(), tc)
        return self._packed[1:]

    def forward(self, x):
        c = self[0]
        if torch.is_grad_enabled() and (x.requires_grad or c.weight.requires_grad or c.bias.requires_grad):
            # training: autograd node whose backward is dgrad (forward kernel on flipped weights) + upf_conv2d_wgrad
            return ops.conv2d_autograd(x, c.weight, c.bias, c.stride[0], c.dilation[0], 0.1 if self.is_relu else 1.0,
                                       _PRECISION["mode"])
        w, b, tc = self.packed(_PRECISION["mode"])
        # hidden activations of tensor-core convolutions are stored rounded to the nearest TF32 value (the MMA truncates)
        prec = (_ext.CONV_TF32 | (_ext.CONV_ROUND_OUT if self.is_relu else 0)) if tc else _ext.CONV_FP32
        return ops.conv2d(x, w, b, c.out_channels, c.kernel_size[0], c.stride[0], c.dilation[0],
                          0.1 if self.is_relu else 1.0, prec)


def conv(in_planes, out_planes, kernel_size=3, stride=1, dilation=1, isReLU=True, if_IN=False, IN_affine=False,
         if_BN=False):
    """conv() of model/pwc_modules.py:10-49 (the IN/BN variants are never used by UPFlow)."""
    if if_IN or if_BN:
        raise NotImplementedError("InstanceNorm/BatchNorm conv variants are not used by UPFlow_net")
    if kernel_size not in (1, 3):
        raise NotImplementedError("kernel_size must be 1 or 3")
    return _ConvBlock(nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, dilation=dilation,
                                padding=((kernel_size - 1) * dilation) // 2, bias=True), isReLU)


def initialize_msra(modules):
    logging.info("Initializing MSRA")
    for layer in modules:
        if isinstance(layer, (nn.Conv2d, nn.ConvTranspose2d)):
            nn.init.kaiming_normal_(layer.weight)
            if layer.bias is not None:
                nn.init.constant_(layer.bias, 0)


def upsample2d_as(inputs, target_as, mode="bilinear"):
    _, _, h, w = target_as.size()
    return ops.resize_bilinear(inputs, h, w)


def upsample2d_flow_as(inputs, target_as, mode="bilinear", if_rate=False):
    """model/pwc_modules.py:77-90."""
    if mode != "bilinear":
        raise NotImplementedError("only bilinear")
    _, _, h, w = target_as.size()
    return ops.resize_bilinear(inputs, h, w, flow_rate=if_rate)


def upsample_flow(inputs, target_size=None, target_flow=None, mode="bilinear"):
    """model/pwc_modules.py:93-104."""
    if target_size is not None:
        h, w = target_size
    elif target_flow is not None:
        _, _, h, w = target_flow.size()
    else:
        raise ValueError('wrong input')
    return ops.resize_bilinear(inputs, h, w, flow_rate=True)


class FeatureExtractor(nn.Module):
    """model/pwc_modules.py:122-142."""

    def __init__(self, num_chs, if_end_relu=True, if_end_norm=False):
        super(FeatureExtractor, self).__init__()
        self.num_chs = num_chs
        self.convs = nn.ModuleList()
        for l, (ch_in, ch_out) in enumerate(zip(num_chs[:-1], num_chs[1:])):
            self.convs.append(nn.Sequential(conv(ch_in, ch_out, stride=2),
                                            conv(ch_out, ch_out, isReLU=if_end_relu, if_IN=if_end_norm)))

    def forward(self, x):
        feature_pyramid = []
        for c in self.convs:
            x = c(x)
            feature_pyramid.append(x)
        return feature_pyramid[::-1]


class WarpingLayer_no_div(nn.Module):
    """model/pwc_modules.py:179-207: one fused kernel instead of mesh + 2 grid_sample + compare + mul."""

    def __init__(self):
        super(WarpingLayer_no_div, self).__init__()

    def forward(self, x, flow):
        return ops.warp(x, flow, align_corners=False, use_mask=True)


class _DenseBlock(tools.abstract_model):
    """Shared body of FlowEstimatorDense_v2 (model/pwc_modules.py:250-286) and the SGU block
    (model/upflow.py:24-60): convs run into one append-only pixel-major buffer, then x5 is assembled in the
    reference's channel order [conv5, conv4, conv3, conv2, conv1, x]."""

    fused_training_block = True

    def _build(self, ch_in, f_channels, out_channel):
        N = ch_in
        for i, c in enumerate(f_channels):
            setattr(self, "conv%d" % (i + 1), conv(N, c))
            N += c
        self.conv_last = conv(N, out_channel, isReLU=False)
        self._ch_in, self._f = ch_in, tuple(f_channels)
        return N

    def forward(self, x):
        ops._require_cuda(x)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            # training: one autograd node over the same append-only buffer as below (ops.dense_block); the reference's
            # own data flow, torch.cat after every conv (model/pwc_modules.py:279-286), stays as the A/B partner
            if not self.fused_training_block:
                for i in range(len(self._f)):
                    x = torch.cat([getattr(self, "conv%d" % (i + 1))(x), x], dim=1)
                return x, self.conv_last(x)
            params = []
            for blk in [getattr(self, "conv%d" % (i + 1)) for i in range(len(self._f))] + [self.conv_last]:
                params += [blk[0].weight, blk[0].bias]
            return ops.dense_block(x, self._f, params, _PRECISION["mode"])
        B, C, H, W = x.shape
        total = C + sum(self._f)
        ld = (total + 3) // 4 * 4
        buf = torch.zeros(B, H, W, ld, dtype=torch.float32, device=x.device)
        # reference order, newest first: [conv5 | conv4 | ... | conv1 | x]
        offs = []
        o = total
        o -= C
        x_off = o
        ops.k_copy(Slice(ops.to_pixel_major(x)), Slice(buf, x_off, C))
        lo = x_off
        for i, c in enumerate(self._f):
            blk = getattr(self, "conv%d" % (i + 1))
            w, b, tc = blk.packed(_ext.CONV_FP32 if (lo % 4) else _PRECISION["mode"])
            new_lo = lo - c
            ops.k_conv(Slice(buf, lo, total - lo), w, b, Slice(buf, new_lo, c), 3, 1, 1, 0.1, None,
                       (_ext.CONV_TF32 | _ext.CONV_ROUND_OUT) if tc else _ext.CONV_FP32)
            lo = new_lo
        w, b, tc = self.conv_last.packed(_ext.CONV_FP32 if (lo % 4) else _PRECISION["mode"])
        cout = self.conv_last[0].out_channels
        out = torch.empty(B, H, W, cout, dtype=torch.float32, device=x.device)
        ops.k_conv(Slice(buf, 0, total), w, b, Slice(out), 3, 1, 1, 1.0, None, _ext.CONV_TF32 if tc else _ext.CONV_FP32)
        x5 = buf[..., :total].permute(0, 3, 1, 2)
        return x5, out.permute(0, 3, 1, 2)


class FlowEstimatorDense_v2(_DenseBlock):

    def __init__(self, ch_in, f_channels=(128, 128, 96, 64, 32), out_channel=2):
        super(FlowEstimatorDense_v2, self).__init__()
        self.n_channels = self._build(ch_in, f_channels, out_channel)


class ContextNetwork_v2_(nn.Module):
    """model/pwc_modules.py:396-412."""

    def __init__(self, ch_in, f_channels=(128, 128, 128, 96, 64, 32, 2)):
        super(ContextNetwork_v2_, self).__init__()
        self.convs = nn.Sequential(
            conv(ch_in, f_channels[0], 3, 1, 1),
            conv(f_channels[0], f_channels[1], 3, 1, 2),
            conv(f_channels[1], f_channels[2], 3, 1, 4),
            conv(f_channels[2], f_channels[3], 3, 1, 8),
            conv(f_channels[3], f_channels[4], 3, 1, 16),
            conv(f_channels[4], f_channels[5], 3, 1, 1),
            conv(f_channels[5], f_channels[6], isReLU=False)
        )

    def forward(self, x):
        return self.convs(x)
