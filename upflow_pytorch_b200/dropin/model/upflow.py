"""Drop-in for the reference's model/upflow.py: `network_tools.sgu_model`,
`network_tools.normalize_features` and `UPFlow_net` (config / update / forward /
forward_2_frame_v3 / decode_level_res / self_guided_upsample / load_model) with
the reference's names, signatures, output dictionary and state-dict keys
(SURVEY.md sections 3.5, 8b).

Inference (`if_loss=False`, the path test.py drives, test.py:40-47) runs on
`upflow_pytorch_b200.engine.DecoderEngine` -- both flow directions stacked,
fused kernels, no torch.cat.  With autograd enabled (training,
scripts/simple_train.py:140-146) the same forward runs module by module, every
op an autograd node whose forward AND backward are this library's kernels
(ops.py).  The loss branch (model/upflow.py:394-491) is elementwise torch around
the library warp: smoothness, photometric (plain or boundary-dilated warp), census
and multi-scale-distillation terms.
"""
from __future__ import absolute_import, division, print_function

import collections

import torch
import torch.nn as nn

from model.correlation_package.correlation import Correlation
from model.pwc_modules import (ContextNetwork_v2_, FeatureExtractor, FlowEstimatorDense_v2, WarpingLayer_no_div,
                               _DenseBlock, conv, initialize_msra, upsample2d_flow_as, upsample_flow)
from upflow_pytorch_b200 import ops
from upflow_pytorch_b200.engine import DecoderEngine
from utils.loss import loss_functions
from utils.pytorch_correlation import Corr_pyTorch
from utils.tools import tools


class network_tools():
    class sgu_model(tools.abstract_model):
        """Self-guided upsample (model/upflow.py:20-92)."""

        def __init__(self):
            super(network_tools.sgu_model, self).__init__()

            class FlowEstimatorDense_temp(_DenseBlock):
                def __init__(self, ch_in, f_channels=(128, 128, 96, 64, 32), ch_out=2):
                    super(FlowEstimatorDense_temp, self).__init__()
                    self.num_feature_channel = self._build(ch_in, f_channels, ch_out)

            self.warping_layer = WarpingLayer_no_div()
            self.dense_estimator_mask = FlowEstimatorDense_temp(64, f_channels=(32, 32, 32, 16, 8), ch_out=3)
            self.upsample_output_conv = nn.Sequential(conv(3, 16, kernel_size=3, stride=1, dilation=1),
                                                      conv(16, 16, stride=2),
                                                      conv(16, 32, kernel_size=3, stride=1, dilation=1),
                                                      conv(32, 32, stride=2), )

        def forward(self, flow_init, feature_1, feature_2, output_level_flow=None):
            h, w = flow_init.shape[2:]
            h_f, w_f = feature_1.shape[2:]
            if h != h_f or w != w_f:
                flow_init = upsample2d_flow_as(flow_init, feature_1, mode="bilinear", if_rate=True)
            feature_2_warp = self.warping_layer(feature_2, flow_init)
            _, x_out = self.dense_estimator_mask(torch.cat((feature_1, feature_2_warp), dim=1))
            if output_level_flow is not None:
                flow_init = output_level_flow
            # sigmoid, the two upsamples, torch_warp and the blend are one kernel (upf_sgu_blend)
            flow_up = ops.sgu_blend(flow_init, x_out)
            inter_flow = x_out[:, :2, :, :]
            inter_mask = ops.sigmoid(x_out[:, 2:3, :, :])
            if output_level_flow is not None:
                inter_flow = upsample2d_flow_as(inter_flow, output_level_flow, mode="bilinear", if_rate=True)
                inter_mask = upsample2d_flow_as(inter_mask, output_level_flow, mode="bilinear")
            return flow_init, flow_up, inter_flow, inter_mask

        def output_conv(self, x):
            return self.upsample_output_conv(x)

    @classmethod
    def normalize_features(cls, feature_list, normalize, center, moments_across_channels=True,
                           moments_across_images=True):
        """model/upflow.py:94-137.  Per-image per-channel moments (the shipped configuration, test.py:24-26) are the
        library's fused statistics + apply kernels with their own backward.  The pooled moment modes (the class
        defaults, :312-313) and the normalize / center switches exist on this module-level (training) path as the
        reference's own expressions on the CUDA tensors -- autograd differentiates them as it does there; the fused
        inference engine serves every mode with kernels (upf_featnorm_combine)."""
        if normalize and center and not moments_across_channels and not moments_across_images:
            return [ops.normalize_features(f) for f in feature_list]
        axes = [1, 2, 3] if moments_across_channels else [2, 3]
        means = [torch.mean(f, dim=axes, keepdim=True) for f in feature_list]
        variances = [torch.var(f, dim=axes, keepdim=True) for f in feature_list]
        if moments_across_images:
            means = [torch.mean(torch.stack(means, dim=0), dim=(0,))] * len(feature_list)
            variances = [torch.var(torch.stack(variances, dim=0), dim=(0,))] * len(feature_list)
        stds = [torch.sqrt(v + 1e-16) for v in variances]
        if center:
            feature_list = [f - m for f, m in zip(feature_list, means)]
        if normalize:
            feature_list = [f / sd for f, sd in zip(feature_list, stds)]
        return list(feature_list)

    # ---- loss terms of the training step (model/upflow.py:198-290).  They run AFTER the decoder on full-resolution
    # 2-/3-channel tensors.  The photometric / distillation term and the first-order smoothness term are fused kernels
    # (csrc/loss.cu); the torch expressions below remain for what the kernels do not take
    # (a mask or an image that itself needs a gradient, the second-order term) and as the A/B partner
    # (`network_tools.use_loss_kernels = False`, tests/test_gpu_backward.py).
    use_loss_kernels = True

    @classmethod
    def edge_aware_smoothness_order1(cls, img, pred):
        """model/upflow.py:198-218 (note the reference's naming: `gradient_x` differences ROWS)."""
        def d_rows(t):
            return t[:, :, :-1, :] - t[:, :, 1:, :]

        def d_cols(t):
            return t[:, :, :, :-1] - t[:, :, :, 1:]
        if cls.use_loss_kernels and pred.is_cuda and not img.requires_grad and min(pred.shape[2:]) >= 2:
            return ops.edge_smooth1(img.float(), pred)           # one kernel forward, one backward (csrc/loss.cu)
        w_r = torch.exp(-torch.mean(torch.abs(d_rows(img)), 1, keepdim=True))
        w_c = torch.exp(-torch.mean(torch.abs(d_cols(img)), 1, keepdim=True))
        return torch.mean(torch.abs(d_rows(pred)) * w_r) + torch.mean(torch.abs(d_cols(pred)) * w_c)

    @classmethod
    def edge_aware_smoothness_order2(cls, img, pred):
        """model/upflow.py:220-245."""
        def d_rows(t, s=1):
            return t[:, :, :-s, :] - t[:, :, s:, :]

        def d_cols(t, s=1):
            return t[:, :, :, :-s] - t[:, :, :, s:]
        w_r = torch.exp(-torch.mean(torch.abs(d_rows(img, 2)), 1, keepdim=True))
        w_c = torch.exp(-torch.mean(torch.abs(d_cols(img, 2)), 1, keepdim=True))
        return torch.mean(torch.abs(d_rows(d_rows(pred))) * w_r) + torch.mean(torch.abs(d_cols(d_cols(pred))) * w_c)

    @classmethod
    def flow_smooth_delta(cls, flow, if_second_order=False):
        """model/upflow.py:247-266."""
        def grad(x):
            return x[:, :, :, 1:] - x[:, :, :, :-1], x[:, :, 1:] - x[:, :, :-1]
        dx, dy = grad(flow)
        loss = dx.abs().mean() + dy.abs().mean()
        if if_second_order:
            dx2, dxdy = grad(dx)
            dydx, dy2 = grad(dy)
            loss = loss + dx2.abs().mean() + dxdy.abs().mean() + dydx.abs().mean() + dy2.abs().mean()
        return loss

    @classmethod
    def photo_loss_multi_type(cls, x, y, occ_mask, photo_loss_type='abs_robust', photo_loss_delta=0.4,
                              photo_loss_use_occ=False):
        """model/upflow.py:268-290 (the SSIM variant is not provided)."""
        if photo_loss_type not in ('abs_robust', 'charbonnier', 'L1'):
            raise NotImplementedError('photo_loss type %s' % photo_loss_type)
        mask = occ_mask if photo_loss_use_occ else None
        if cls.use_loss_kernels and x.is_cuda and x.shape == y.shape and (mask is None or (
                not mask.requires_grad and mask.numel() == x.shape[0] * x.shape[2] * x.shape[3])):
            return ops.robust_loss(x, y, mask, photo_loss_type, photo_loss_delta)   # csrc/loss.cu
        if photo_loss_type == 'abs_robust':
            loss_diff = (torch.abs(x - y) + 0.01).pow(photo_loss_delta)
        elif photo_loss_type == 'charbonnier':
            loss_diff = ((x - y) ** 2 + 1e-6).pow(photo_loss_delta)
        elif photo_loss_type == 'L1':
            loss_diff = torch.abs(x - y + 1e-6)
        else:
            raise NotImplementedError('photo_loss type %s' % photo_loss_type)
        if photo_loss_use_occ:
            return torch.sum(loss_diff * occ_mask) / (torch.sum(occ_mask) + 1e-6)
        return torch.mean(loss_diff)


class UPFlow_net(tools.abstract_model):
    class config(tools.abstract_config):
        def __init__(self):
            # identical attribute set and defaults to model/upflow.py:293-323
            self.occ_type = 'for_back_check'
            self.alpha_1 = 0.1
            self.alpha_2 = 0.5
            self.occ_check_obj_out_all = 'obj'
            self.stop_occ_gradient = False
            self.smooth_level = 'final'
            self.smooth_type = 'edge'
            self.smooth_order_1_weight = 1
            self.smooth_order_2_weight = 0
            self.photo_loss_type = 'abs_robust'
            self.photo_loss_delta = 0.4
            self.photo_loss_use_occ = False
            self.photo_loss_census_weight = 0
            self.if_norm_before_cost_volume = False
            self.norm_moments_across_channels = True
            self.norm_moments_across_images = True
            self.multi_scale_distillation_weight = 0
            self.multi_scale_distillation_style = 'upup'
            self.multi_scale_distillation_occ = True
            self.if_froze_pwc = False
            self.input_or_sp_input = 1
            self.if_use_boundary_warp = True
            self.if_sgu_upsample = False
            self.if_use_cor_pytorch = False

        def __call__(self, ):
            return UPFlow_net(self)

    # extra, non-reference knobs of the B200 build (class attributes so `config` stays identical)
    conv_precision = "tf32"      # 'tf32' = tcgen05 tensor cores (what cuDNN does by default), 'tf32x3' = three TF32 passes
                                 # per conv (fp32-class results, 3.8x the time), 'fp32' = strict SIMT (18x the time)
    use_cuda_graph = True        # replay one captured graph per input shape (inference has no host decisions)

    def __init__(self, conf: config):
        super(UPFlow_net, self).__init__()
        self.conf = conf
        self.search_range = 4
        self.num_chs = [3, 16, 32, 64, 96, 128, 196]
        self.estimator_f_channels = (128, 128, 96, 64, 32)
        self.context_f_channels = (128, 128, 128, 96, 64, 32, 2)
        self.output_level = 4
        self.num_levels = 7
        self.leakyRELU = nn.LeakyReLU(0.1, inplace=True)
        self.feature_pyramid_extractor = FeatureExtractor(self.num_chs)
        self.warping_layer = WarpingLayer_no_div()
        self.dim_corr = (self.search_range * 2 + 1) ** 2
        self.num_ch_in = self.dim_corr + 32 + 2
        self.flow_estimators = FlowEstimatorDense_v2(self.num_ch_in, f_channels=self.estimator_f_channels)
        self.context_networks = ContextNetwork_v2_(self.flow_estimators.n_channels + 2,
                                                   f_channels=self.context_f_channels)
        self.conv_1x1 = nn.ModuleList([conv(196, 32, kernel_size=1, stride=1, dilation=1),
                                       conv(128, 32, kernel_size=1, stride=1, dilation=1),
                                       conv(96, 32, kernel_size=1, stride=1, dilation=1),
                                       conv(64, 32, kernel_size=1, stride=1, dilation=1),
                                       conv(32, 32, kernel_size=1, stride=1, dilation=1)])
        self.occ_check_model_ls = []
        self.correlation_pytorch = Corr_pyTorch(pad_size=self.search_range, kernel_size=1,
                                                max_displacement=self.search_range, stride1=1, stride2=1)
        self.sgi_model = network_tools.sgu_model() if self.conf.if_sgu_upsample else None
        self.occ_check_model = tools.occ_check_model(occ_type=self.conf.occ_type, occ_alpha_1=self.conf.alpha_1,
                                                     occ_alpha_2=self.conf.alpha_2,
                                                     obj_out_all=self.conf.occ_check_obj_out_all)
        initialize_msra(self.modules())
        if self.conf.if_froze_pwc:
            self.froze_PWC()
        self._engine = None
        self._engine_key = None
        self._graphs = {}

    # ------------------------------------------------------------------ engine plumbing
    def _get_engine(self):
        params = list(self.parameters())
        key = (self.conv_precision, params[0].device, tuple(p._version for p in params),
               tuple(p.data_ptr() for p in params), bool(self.conf.if_norm_before_cost_volume),
               bool(self.conf.norm_moments_across_channels), bool(self.conf.norm_moments_across_images))
        if self._engine is None or self._engine_key != key:
            occ = None
            if self.conf.occ_type == 'for_back_check':
                occ = (self.conf.alpha_1, self.conf.alpha_2, self.conf.occ_check_obj_out_all)
            self._engine = DecoderEngine(self.state_dict(), device=params[0].device, precision=self.conv_precision,
                                         use_sgu=bool(self.conf.if_sgu_upsample), occ=occ,
                                         norm=(self.conf.if_norm_before_cost_volume, self.conf.norm_moments_across_channels,
                                               self.conf.norm_moments_across_images))
            self._engine_key = key
            self._graphs = {}
        return self._engine

    def set_lane(self, lane: int):
        """Select the workspace / graph set the next inference calls use (engine.DecoderEngine.lane).  Calls on different
        lanes may be in flight on different CUDA streams at the same time (pipeline.PipelinedInference(lanes=2))."""
        self._get_engine().lane = int(lane)

    def forward(self, input_dict: dict):
        """model/upflow.py:370-392 (inference branch)."""
        im1_ori, im2_ori = input_dict['im1'], input_dict['im2']
        if input_dict['if_loss'] and self.conf.input_or_sp_input != 1:
            im1, im2 = input_dict['im1_sp'], input_dict['im2_sp']
        else:
            im1, im2 = im1_ori, im2_ori
        output_dict = {}
        fused = self._infer(im1, im2, want_flows=False) if not input_dict['if_loss'] else None
        if fused is not None:
            # inference: flows AND occlusion masks come out of the fused engine (one captured graph)
            flow_f, flow_b, flows, occ_fw, occ_bw = fused
        else:
            flow_f, flow_b, flows = self.forward_2_frame_v3(im1, im2, if_loss=input_dict['if_loss'])
            occ_fw, occ_bw = self.occ_check_model(flow_f=flow_f, flow_b=flow_b)
        output_dict['flow_f_out'] = flow_f
        output_dict['flow_b_out'] = flow_b
        output_dict['occ_fw'] = occ_fw
        output_dict['occ_bw'] = occ_bw
        if input_dict['if_loss']:
            self._losses(input_dict, output_dict, im1_ori, im2_ori, flow_f, flow_b, flows, occ_fw, occ_bw)
        return output_dict

    def _losses(self, input_dict, output_dict, im1_ori, im2_ori, flow_f, flow_b, flows, occ_fw, occ_bw):
        """model/upflow.py:394-491: smoothness, photometric (plain or boundary-dilated warp), census and multi-scale
        distillation terms -- elementwise torch around the library warp."""
        conf = self.conf
        nt = network_tools
        if conf.smooth_level == 'final':
            s_flow_f, s_flow_b, s_im1, s_im2 = flow_f, flow_b, im1_ori, im2_ori
        elif conf.smooth_level == '1/4':
            s_flow_f, s_flow_b = flows[0]
            th, tw = s_flow_f.shape[2:]
            s_im1 = torch.nn.functional.interpolate(im1_ori, (th, tw), mode='area')
            s_im2 = torch.nn.functional.interpolate(im2_ori, (th, tw), mode='area')
        else:
            raise ValueError('wrong smooth level choosed: %s' % conf.smooth_level)
        smooth_loss = 0
        for weight, second in ((conf.smooth_order_1_weight, False), (conf.smooth_order_2_weight, True)):
            if weight > 0:
                if conf.smooth_type == 'edge':
                    fn = nt.edge_aware_smoothness_order2 if second else nt.edge_aware_smoothness_order1
                    smooth_loss = smooth_loss + weight * fn(img=s_im1, pred=s_flow_f) + weight * fn(img=s_im2, pred=s_flow_b)
                elif conf.smooth_type == 'delta':
                    smooth_loss = smooth_loss + weight * nt.flow_smooth_delta(s_flow_f, second) \
                        + weight * nt.flow_smooth_delta(s_flow_b, second)
                else:
                    raise ValueError('wrong smooth_type: %s' % conf.smooth_type)
        output_dict['smooth_loss'] = smooth_loss
        if conf.if_use_boundary_warp:
            im1_s, im2_s, start_s = input_dict['im1_raw'], input_dict['im2_raw'], input_dict['start']
            im1_warp = tools.boundary_dilated_warp.warp_im(im2_s, flow_f, start_s)
            im2_warp = tools.boundary_dilated_warp.warp_im(im1_s, flow_b, start_s)
        else:
            im1_warp = tools.torch_warp(im2_ori, flow_f)
            im2_warp = tools.torch_warp(im1_ori, flow_b)
        if conf.stop_occ_gradient:
            occ_fw, occ_bw = occ_fw.clone().detach(), occ_bw.clone().detach()
        kw = dict(photo_loss_type=conf.photo_loss_type, photo_loss_delta=conf.photo_loss_delta,
                  photo_loss_use_occ=conf.photo_loss_use_occ)
        output_dict['photo_loss'] = nt.photo_loss_multi_type(im1_ori, im1_warp, occ_fw, **kw) \
            + nt.photo_loss_multi_type(im2_ori, im2_warp, occ_bw, **kw)
        output_dict['im1_warp'] = im1_warp
        output_dict['im2_warp'] = im2_warp
        if conf.photo_loss_census_weight > 0:
            ckw = dict(q=conf.photo_loss_delta, charbonnier_or_abs_robust=False, if_use_occ=conf.photo_loss_use_occ, averge=True)
            output_dict['census_loss'] = conf.photo_loss_census_weight * (
                loss_functions.census_loss_torch(img1=im1_ori, img1_warp=im1_warp, mask=occ_fw, **ckw)
                + loss_functions.census_loss_torch(img1=im2_ori, img1_warp=im2_warp, mask=occ_bw, **ckw))
        else:
            output_dict['census_loss'] = None
        if conf.multi_scale_distillation_weight > 0:
            label_f, label_b = flow_f.clone().detach(), flow_b.clone().detach()
            terms = []
            for scale_fw, scale_bw in flows:
                if conf.multi_scale_distillation_style == 'down':
                    lf = upsample_flow(label_f, target_flow=scale_fw)
                    of = torch.nn.functional.interpolate(occ_fw, list(scale_fw.shape[2:]), mode='nearest')
                    lb = upsample_flow(label_b, target_flow=scale_bw)
                    ob = torch.nn.functional.interpolate(occ_bw, list(scale_bw.shape[2:]), mode='nearest')
                elif conf.multi_scale_distillation_style == 'upup':
                    lf, of, lb, ob = label_f, occ_fw, label_b, occ_bw
                    scale_fw = upsample_flow(scale_fw, target_flow=lf)
                    scale_bw = upsample_flow(scale_bw, target_flow=lb)
                else:
                    raise ValueError('wrong multi_scale_distillation_style: %s' % conf.multi_scale_distillation_style)
                terms.append(nt.photo_loss_multi_type(scale_fw, lf, of, 'abs_robust',
                                                      photo_loss_use_occ=conf.multi_scale_distillation_occ))
                terms.append(nt.photo_loss_multi_type(scale_bw, lb, ob, 'abs_robust',
                                                      photo_loss_use_occ=conf.multi_scale_distillation_occ))
            output_dict['msd_loss'] = conf.multi_scale_distillation_weight * sum(terms)
        else:
            output_dict['msd_loss'] = None

    def forward_2_frame_v3(self, x1_raw, x2_raw, if_loss=False):
        """model/upflow.py:494-533 on the fused engine; outputs are fresh tensors."""
        if not x1_raw.is_cuda:
            raise RuntimeError("UPFlow_net (upflow_pytorch_b200) runs on CUDA only: move the model and the inputs "
                               "with .cuda(); there is no CPU path")
        if torch.is_grad_enabled() and (x1_raw.requires_grad or any(p.requires_grad for p in self.parameters())):
            return self._forward_2_frame_modules(x1_raw, x2_raw)
        f, b, flows, _, _ = self._infer(x1_raw, x2_raw, want_flows=True)
        return f, b, flows

    def _infer(self, x1_raw, x2_raw, want_flows):
        """The fused engine (CUDA graph per input shape).  Returns fresh tensors (flow_f, flow_b, flows or None, occ_fw,
        occ_bw), or None when autograd needs the module-level path."""
        if not x1_raw.is_cuda:
            raise RuntimeError("UPFlow_net (upflow_pytorch_b200) runs on CUDA only: move the model and the inputs "
                               "with .cuda(); there is no CPU path")
        if torch.is_grad_enabled() and (x1_raw.requires_grad or any(p.requires_grad for p in self.parameters())):
            return None
        eng = self._get_engine()
        if self.use_cuda_graph:
            key = tuple(x1_raw.shape) + (eng.lane,)
            g = self._graphs.get(key)
            if g is None:
                lanes = 1 + max([k[-1] for k in self._graphs] + [eng.lane])
                if len(self._graphs) >= 8 * lanes:         # 8 input shapes per lane
                    torch.cuda.synchronize(x1_raw.device)  # other lanes' replays may still be running on these buffers
                    self._graphs.clear()
                    eng.release_workspaces()       # the graphs' buffers go with them
                g = self._graphs[key] = eng.capture(x1_raw.shape[0], x1_raw.shape[2], x1_raw.shape[3])
            f, b = g(x1_raw, x2_raw)
            flows, occ = g.flows, g.occ
        else:
            if len(eng._ws) >= 24:
                eng.release_workspaces()
            f, b, flows = eng.forward(x1_raw.float(), x2_raw.float())
            occ = eng.last_occ
        if occ is None:
            f, b = f.clone(), b.clone()
            occ_fw, occ_bw = self.occ_check_model(flow_f=f, flow_b=b)
        else:
            B = f.shape[0]
            fo = eng.flow_out_buffer(x1_raw.shape).clone()          # ONE copy of [2B,H,W,2] instead of two
            oc = occ.clone()
            fo, oc = fo.permute(0, 3, 1, 2), oc.permute(0, 3, 1, 2)
            f, b, occ_fw, occ_bw = fo[:B], fo[B:], oc[:B], oc[B:]
        flows = [[a.clone(), c.clone()] for a, c in flows] if want_flows else None
        return f, b, flows, occ_fw, occ_bw

    def _forward_2_frame_modules(self, x1_raw, x2_raw):
        """Training path: model/upflow.py:494-533 module by module, every op an autograd node backed by the
        library's forward AND backward kernels (ops.py)."""
        from model import pwc_modules
        pwc_modules.set_conv_precision(self.conv_precision)
        x1_pyramid = self.feature_pyramid_extractor(x1_raw) + [x1_raw]
        x2_pyramid = self.feature_pyramid_extractor(x2_raw) + [x2_raw]
        b_size, _, h0, w0 = x1_pyramid[0].shape
        flow_f = torch.zeros(b_size, 2, h0, w0, dtype=torch.float32, device=x1_raw.device)
        flow_b = torch.zeros_like(flow_f)
        levels = []
        for l, (x1, x2) in enumerate(zip(x1_pyramid, x2_pyramid)):
            levels.append((x1, self.conv_1x1[l](x1), x2, self.conv_1x1[l](x2)))
            if l == self.output_level:
                break
        flows = []
        for level, (x1, x1_1by1, x2, x2_1by1) in enumerate(levels):
            flow_f, flow_b, res_f, res_b = self.decode_level_res(level, flow_f, flow_b, x1, x1_1by1, x2, x2_1by1,
                                                                 x1_raw, x2_raw)
            flow_f = flow_f + res_f
            flow_b = flow_b + res_b
            flows.append([flow_f, flow_b])
        flow_f_out = upsample2d_flow_as(flow_f, x1_raw, mode="bilinear", if_rate=True)
        flow_b_out = upsample2d_flow_as(flow_b, x1_raw, mode="bilinear", if_rate=True)
        if self.conf.if_sgu_upsample:
            f1 = self.sgi_model.output_conv(x1_raw)
            f2 = self.sgi_model.output_conv(x2_raw)
            flow_f_out = self.self_guided_upsample(flow_f, f1, f2, output_level_flow=flow_f_out)
            flow_b_out = self.self_guided_upsample(flow_b, f2, f1, output_level_flow=flow_b_out)
        return flow_f_out, flow_b_out, flows[::-1]

    def decode_level_res(self, level, flow_1, flow_2, feature_1, feature_1_1x1, feature_2, feature_2_1x1, img_ori_1,
                         img_ori_2):
        """model/upflow.py:535-573, module by module (both directions)."""
        flow_1_up = upsample2d_flow_as(flow_1, feature_1, mode="bilinear", if_rate=True)
        flow_2_up = upsample2d_flow_as(flow_2, feature_2, mode="bilinear", if_rate=True)
        if level == 0:
            feature_2_warp, feature_1_warp = feature_2, feature_1
        else:
            if self.conf.if_sgu_upsample:
                flow_1_up = self.self_guided_upsample(flow_1_up, feature_1_1x1, feature_2_1x1)
                flow_2_up = self.self_guided_upsample(flow_2_up, feature_2_1x1, feature_1_1x1)
            feature_2_warp = self.warping_layer(feature_2, flow_1_up)
            feature_1_warp = self.warping_layer(feature_1, flow_2_up)
        if self.conf.if_norm_before_cost_volume:
            feature_1, feature_2_warp = network_tools.normalize_features(
                (feature_1, feature_2_warp), normalize=True, center=True,
                moments_across_channels=self.conf.norm_moments_across_channels,
                moments_across_images=self.conf.norm_moments_across_images)
            feature_2, feature_1_warp = network_tools.normalize_features(
                (feature_2, feature_1_warp), normalize=True, center=True,
                moments_across_channels=self.conf.norm_moments_across_channels,
                moments_across_images=self.conf.norm_moments_across_images)
        # both correlation back ends of the reference (:557-562) are the same fused kernel here, LeakyReLU included
        out_corr_relu_1 = ops.correlation(feature_1, feature_2_warp, self.search_range, leaky_slope=0.1)
        out_corr_relu_2 = ops.correlation(feature_2, feature_1_warp, self.search_range, leaky_slope=0.1)
        feature_int_1, flow_res_1 = self.flow_estimators(torch.cat([out_corr_relu_1, feature_1_1x1, flow_1_up], dim=1))
        feature_int_2, flow_res_2 = self.flow_estimators(torch.cat([out_corr_relu_2, feature_2_1x1, flow_2_up], dim=1))
        flow_fine_1 = self.context_networks(torch.cat([feature_int_1, flow_1_up + flow_res_1], dim=1))
        flow_fine_2 = self.context_networks(torch.cat([feature_int_2, flow_2_up + flow_res_2], dim=1))
        return flow_1_up, flow_2_up, flow_res_1 + flow_fine_1, flow_res_2 + flow_fine_2

    def froze_PWC(self):
        for m in (self.feature_pyramid_extractor, self.flow_estimators, self.context_networks, self.conv_1x1):
            for param in m.parameters():
                param.requires_grad = False

    def self_guided_upsample(self, flow_up_bilinear, feature_1, feature_2, output_level_flow=None):
        _, out_flow, _, _ = self.sgi_model(flow_up_bilinear, feature_1, feature_2, output_level_flow=output_level_flow)
        return out_flow
