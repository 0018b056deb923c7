"""The decoder engine: UPFlow's two-frame forward (forward_2_frame_v3,
model/upflow.py:494-533, with decode_level_res :535-573 and sgu_model.forward
:71-89) re-hosted on the library's kernels.

What changes relative to the reference's orchestration (results do not):
  * the forward and the backward flow directions share weights and have no data
    dependence inside a level, so they are STACKED in one batch of N = 2B
    images ([im1.., im2..]); "the other image" is a batch shift of B that the
    warp / correlation kernels apply while addressing (no swapped copies);
  * activations are pixel-major (NHWC).  Each dense block lives in ONE
    append-only buffer per level (X: 576 channels for the flow estimator, S:
    192 for the SGU block); convolutions read a channel prefix and write their
    output slice in place, so none of the 302 torch.cat calls of a KITTI
    forward exist.  The reference PREPENDS new features
    (model/pwc_modules.py:280-284); the fixed permutation between the two
    orders is folded into the packed weights once, at load time;
  * LeakyReLU, bias, residual flow additions, feature normalisation, the
    validity mask and the flow rescale are fused into the producing kernels;
  * no host-side tensor construction, no host<->device copies and no
    synchronisation inside the forward: it is capturable in a CUDA graph.

X buffer channel map (estimator input order of the reference is
[corr 81 | f_1x1 32 | flow 2], model/upflow.py:565):
    0..80 corr | 81..112 f_1x1 | 113..114 flow_up | 115..116 flow_up+flow_res
    | 117..127 zero | 128 conv1 | 256 conv2 | 384 conv3 | 480 conv4 | 544 conv5

Precision "tf32": tcgen05 kind::tf32 reads the top 19 bits of an fp32 operand,
i.e. it TRUNCATES.  Every buffer a tensor-core convolution reads is therefore
written already ROUNDED TO THE NEAREST TF32 value by its producer (convolution
epilogues of hidden layers, the correlation, the SGU warp, the flow slots), and
the weights are rounded when packed: unbiased operand error 2^-12 instead of a
one-sided 2^-11.  The two flow slots of X are rounded COPIES; the exact flows
the warp / blend / residual additions need live in their own small buffers
("fu", "flow2").  With the shipped checkpoint at KITTI size this is the
difference between 3.8e-3 px (truncation) and 5.6e-4 px (rounding) mean EPE
against the fp32 reference (robust-mask diagnostic).
"""
import os

import torch

from . import _ext, ops
from .ops import Slice

NUM_CHS = (196, 128, 96, 64, 32)          # decoder levels 0..4 (1/64 .. 1/4), model/upflow.py:336
EST_CH = (128, 128, 96, 64, 32)           # model/upflow.py:338
CTX_CH = (128, 128, 128, 96, 64, 32, 2)   # model/upflow.py:339
CTX_DIL = (1, 2, 4, 8, 16, 1, 1)          # model/pwc_modules.py:401-409
SGU_CH = (32, 32, 32, 16, 8)              # model/upflow.py:62
X_LD = 576
X_CORR, X_F1X1, X_FLOW, X_FLOW2 = 0, 81, 113, 115
X_OFF = (128, 256, 384, 480, 544)         # conv1..conv5 output offsets
S_LD = 192
S_OFF = (64, 96, 128, 160, 176)
SLOPE = 0.1


def _dense_slots(x_width, x_slots, out_offsets, out_channels, k):
    """Slots, in the append-only buffer, of the reference's input channels of
    the k-th conv (k=0..5) of a dense block: reference order is
    [conv_k-1 out, ..., conv_1 out, x]."""
    slots = []
    for j in range(k - 1, -1, -1):
        slots += list(range(out_offsets[j], out_offsets[j] + out_channels[j]))
    slots += list(x_slots)
    return slots


CHAIN_MAX_PIXELS = 4096      # feature maps (N*h*w) up to this size run their dense blocks as ONE persistent launch (conv_chain.cu)
EXPAND_MAX_COUT = 0          # 3x3 convs with at most this many outputs run as 1x1-expand + tap-combine (0 = never: since conv_win.cu puts the horizontal taps along N, one N = 48 MMA per kernel row beats the expansion -- KITTI forward 2.519 -> 2.444 ms, profiles/r2_ab_expand.txt; tests and A/B runs set it) ...
EXPAND_MIN_PIXELS = 5000     # ... on images with at least this many pixels (below, the extra launch costs more)


class ConvSpec:
    __slots__ = ("w", "w_tc", "bias", "cin", "cout", "k", "stride", "dil", "slope", "w_exp", "zero_bias", "w_lo_tc", "zero_b",
                 "hidden")

    def __init__(self, weight, bias, stride=1, dil=1, relu=True, in_slots=None, cin_total=None, tc=True, x3=False):
        self.cout, _, self.k, _ = weight.shape
        self.cin = cin_total or weight.shape[1]
        self.stride, self.dil = stride, dil
        self.slope = SLOPE if relu else 1.0
        self.hidden = relu            # a hidden activation: only convolutions (and, for the encoder features, the
                                      # correlation / warp) read it -- stored TF32-rounded under precision "tf32"
        # tensor cores for everything but the 3-channel image convs (K = 27: one quarter-empty K block per tap)
        self.w, self.w_tc = ops.pack_conv_weight(weight, in_slots, cin_total, tc=tc and stride in (1, 2) and self.cin >= 16)
        self.bias = bias.detach().float().contiguous()
        # few output channels: Y = 1x1 conv with 9*Cout outputs (one N<=80 MMA per K step instead of nine N=16 ones),
        # then upf_conv3x3_tap_combine gathers the nine shifted slices
        # 3xTF32: the low part of the weights (w - w truncated to TF32), packed like w
        self.w_lo_tc = None
        if x3 and self.w_tc is not None:
            w32 = weight.detach().float().contiguous()
            w_hi = ((w32.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)      # what the packer stores: nearest TF32
            self.w_lo_tc = ops.pack_conv_weight(w32 - w_hi, in_slots, cin_total, tc=True)[1]
            self.zero_b = torch.zeros(self.cout, dtype=torch.float32, device=weight.device)
        self.w_exp = None
        if not x3 and self.w_tc is not None and self.k == 3 and stride == 1 and self.cout <= EXPAND_MAX_COUT:
            self.w_exp = ops.pack_conv_weight(ops.expand_taps_weight(weight), in_slots, cin_total, tc=True)[1]
            self.zero_bias = torch.zeros(9 * self.cout, dtype=torch.float32, device=weight.device)


class DecoderEngine:
    def __init__(self, state_dict, device="cuda", precision="tf32", align_corners=False, use_sgu=True,
                 mask_threshold=1.0, occ=None, norm=(True, False, False)):
        """state_dict: the reference's parameter names.
        norm = (if_norm_before_cost_volume, norm_moments_across_channels, norm_moments_across_images) of
        UPFlow_net.config (model/upflow.py:311-313); test.py runs (True, False, False), the class default is
        (False, True, True)."""
        if not torch.cuda.is_available():
            raise RuntimeError("DecoderEngine needs a CUDA device: the decoder path has no CPU implementation")
        _ext.load()
        self.device = torch.device(device)
        if precision not in ("tf32", "fp32", "tf32x3"):
            raise ValueError("precision must be 'tf32', 'tf32x3' or 'fp32'")
        self.precision = precision
        self.tc = precision in ("tf32", "tf32x3")
        # 'tf32x3': every tensor-core convolution as three TF32 passes (hi*hi + lo*hi + hi*lo) accumulated before the
        # activation -- fp32-class results (2^-21) at a third of the tensor-core rate; every convolution input keeps a
        # "low part" twin buffer (x - x truncated to TF32) that the producing kernel's output is split into
        self.x3 = precision == "tf32x3"
        self.rnd = precision == "tf32"       # producers round what tensor-core convolutions will read
        self._lo_bufs = {}
        self.align_corners = bool(align_corners)
        self.use_sgu = use_sgu
        # 1.0 = the reference's `mask >= 1.0` (model/pwc_modules.py:206); 0.9999 = diagnostic robust mask
        self.mask = True if mask_threshold == 1.0 else float(mask_threshold)
        self._ws = {}
        self.norm, self.norm_ch, self.norm_img = (bool(v) for v in norm)
        # (alpha_1, alpha_2, 'all'|'obj'|'out'): also produce the forward/backward consistency masks that
        # UPFlow_net.forward returns (tools.occ_check_model, model/upflow.py:386) -- one more launch in the graph
        self.occ = occ
        self.last_occ = None
        self.overlap = True        # image-only work on a side stream (forward())
        # Workspace set in use.  Two forwards whose graphs replay CONCURRENTLY (pipeline.PipelinedInference(lanes=2):
        # one pair's latency-bound coarse levels fill the SMs another pair's fine levels leave idle) must not share
        # buffers: each lane has its own workspaces, scratch and captured graphs.
        self.lane = 0
        # coarse levels: estimator + context network (13 convolutions) and the SGU block (6) as one launch each
        self.chain = precision == "tf32" and os.environ.get("UPF_CHAIN", "0") == "1"
        self.load_weights(state_dict)

    # ------------------------------------------------------------ weights
    def load_weights(self, sd):
        g = lambda k: sd[k].detach().to(self.device, torch.float32)
        tc = self.tc

        def spec(key, **kw):
            return ConvSpec(g(key + ".0.weight"), g(key + ".0.bias"), tc=tc, x3=self.x3, **kw)

        self.enc = []
        for l in range(6):
            self.enc.append((spec(f"feature_pyramid_extractor.convs.{l}.0", stride=2),
                             spec(f"feature_pyramid_extractor.convs.{l}.1")))
        self.conv1x1 = [spec(f"conv_1x1.{l}") for l in range(5)]
        # flow estimator: input x = X[0:115]
        x_slots = list(range(115))
        self.est = []
        names = ("conv1", "conv2", "conv3", "conv4", "conv5", "conv_last")
        widths = (128,) + tuple(o + c for o, c in zip(X_OFF, EST_CH))      # prefix each conv reads
        for k, name in enumerate(names):
            self.est.append(spec(f"flow_estimators.{name}", relu=(k < 5),
                                 in_slots=_dense_slots(115, x_slots, X_OFF, EST_CH, k), cin_total=widths[k]))
        # context network: input = [x5 (563) | flow_up+flow_res (2)]
        ctx0_slots = _dense_slots(115, x_slots, X_OFF, EST_CH, 5) + [X_FLOW2, X_FLOW2 + 1]
        self.ctx = []
        for i in range(7):
            kw = dict(dil=CTX_DIL[i], relu=(i < 6))
            if i == 0:
                kw.update(in_slots=ctx0_slots, cin_total=X_LD)
            self.ctx.append(spec(f"context_networks.convs.{i}", **kw))
        self.sgu = None
        if self.use_sgu:
            s_slots = list(range(64))
            swidths = (64,) + tuple(o + c for o, c in zip(S_OFF, SGU_CH))
            self.sgu = [spec(f"sgi_model.dense_estimator_mask.{name}", relu=(k < 5),
                             in_slots=_dense_slots(64, s_slots, S_OFF, SGU_CH, k), cin_total=swidths[k])
                        for k, name in enumerate(names)]
            self.outconv = [spec(f"sgi_model.upsample_output_conv.{i}", stride=s) for i, s in enumerate((1, 2, 1, 2))]

    # ------------------------------------------------------------ helpers
    def _scratch(self, N, H, W, C):
        side = getattr(self, "_side", None)
        on_side = side is not None and torch.cuda.current_stream() == side
        key = ("scratch", N, H, W, C, on_side, self.lane)  # one per stream: the side stream overlaps the main one
        buf = self._ws.get(key)
        if buf is None:
            buf = self._ws[key] = torch.zeros(N, H, W, C, dtype=torch.float32, device=self.device)
        return buf

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def _lo(self, buf):
        """the low-part twin of a buffer (tf32x3 mode)"""
        lo = self._lo_bufs.get(id(buf))
        if lo is None:
            lo = torch.zeros_like(buf)
            self._lo_bufs[id(buf)] = (lo, buf)            # keep `buf` alive: the key is its id
            return lo
        return lo[0]

    def _split(self, sl):
        """tf32x3: refresh the low part of a slice some non-convolution kernel just wrote"""
        if self.x3:
            ops.k_act_split(None, sl, Slice(self._lo(sl.buf), sl.c0, sl.C))

    def _conv(self, cs, x, out, residual=None):
        use_tc = self.tc and cs.w_tc is not None
        if self.x3 and use_tc:
            xlo = Slice(self._lo(x.buf), x.c0, x.C)
            tmp = Slice(self._scratch(out.N, out.H, out.W, (cs.cout + 3) // 4 * 4), 0, cs.cout)
            args = (cs.k, cs.stride, cs.dil, 1.0)
            ops.k_conv(xlo, cs.w_tc, cs.zero_b, tmp, *args, None, _ext.CONV_TF32)          # lo * w_hi
            ops.k_conv(x, cs.w_lo_tc, cs.zero_b, tmp, *args, tmp, _ext.CONV_TF32)          # + hi * w_lo
            ops.k_conv(x, cs.w_tc, cs.bias, tmp, *args, tmp, _ext.CONV_TF32)               # + hi * w_hi + bias
            ops.k_act_split(tmp, out, Slice(self._lo(out.buf), out.c0, out.C), cs.slope, residual)
            return
        if self.x3:
            ops.k_conv(x, cs.w, cs.bias, out, cs.k, cs.stride, cs.dil, cs.slope, residual, _ext.CONV_FP32)   # exact SIMT
            self._split(out)
            return
        rnd = self.rnd and cs.hidden
        if use_tc and cs.w_exp is not None and x.H * x.W >= EXPAND_MIN_PIXELS:
            Y = self._scratch(x.N, x.H, x.W, 9 * EXPAND_MAX_COUT)
            ys = Slice(Y, 0, 9 * cs.cout)
            ops.k_conv(x, cs.w_exp, cs.zero_bias, ys, 1, 1, 1, 1.0, None, _ext.CONV_TF32)
            ops.k_tap_combine(ys, cs.bias, out, cs.dil, cs.slope, residual, round_tf32=rnd)
            return
        ops.k_conv(x, cs.w_tc if use_tc else cs.w, cs.bias, out, cs.k, cs.stride, cs.dil, cs.slope, residual,
                   (_ext.CONV_TF32 if use_tc else _ext.CONV_FP32) | (_ext.CONV_ROUND_OUT if rnd else 0))

    def _workspace(self, B, H, W):
        key = (B, H, W, self.lane)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        dev = self.device
        N = 2 * B
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        sizes = [(H, W)]
        for _ in range(6):
            h, w = sizes[-1]
            sizes.append(((h + 1) // 2, (w + 1) // 2))       # stride-2 pad-1 3x3 conv
        ws = {"sizes": sizes, "N": N}
        # encoder features: full-res input [N,H,W,4(3 used)], level tensors (two convs each)
        ws["im"] = z(N, H, W, 4)
        chs = (16, 32, 64, 96, 128, 196)
        ws["enc_a"] = [z(N, *sizes[l + 1], chs[l]) for l in range(6)]
        ws["enc_b"] = [z(N, *sizes[l + 1], chs[l]) for l in range(6)]
        lv = []
        # every (sum, sum^2) accumulator of a forward lives in ONE arena, cleared by one memset per forward
        arena = torch.zeros(2 * N * sum(NUM_CHS) * 2, dtype=torch.float64, device=dev)
        ws["stats_arena"] = arena
        pooled = self.norm and (self.norm_ch or self.norm_img)      # the pooled moment modes go through upf_featnorm_combine
        off = 0
        for l in range(5):
            h, w = sizes[6 - l]
            n = N * NUM_CHS[l] * 2
            d = {"hw": (h, w), "X": z(N, h, w, X_LD), "T0": z(N, h, w, 128), "T1": z(N, h, w, 128),
                 "flow": z(N, h, w, 2), "xw": z(N, h, w, NUM_CHS[l]),
                 "fu": z(N, h, w, 2), "flow2": z(N, h, w, 2),        # exact flow_up / flow_up + flow_res (X holds TF32 copies)
                 "stats_own": arena[off:off + n].view(N, NUM_CHS[l], 2),
                 "stats_w": arena[off + n:off + 2 * n].view(N, NUM_CHS[l], 2)}
            if pooled:
                d["stats_ca"] = torch.zeros(N, NUM_CHS[l], 2, dtype=torch.float64, device=dev)
                d["stats_cb"] = torch.zeros(N, NUM_CHS[l], 2, dtype=torch.float64, device=dev)
            off += 2 * n
            if self.use_sgu and l > 0:
                d["S"] = z(N, h, w, S_LD)
                d["inter"] = z(N, h, w, 4)
                d["flow_bil"] = z(N, h, w, 2)
            lv.append(d)
        ws["levels"] = lv
        if self.use_sgu:
            h4, w4 = sizes[2]
            ws["oc_a"] = z(N, H, W, 16)
            ws["oc_b"] = z(N, *sizes[1], 16)
            ws["oc_c"] = z(N, *sizes[1], 32)
            ws["S_out"] = z(N, h4, w4, S_LD)
            ws["inter_out"] = z(N, h4, w4, 4)
            ws["flow_full_bil"] = z(N, H, W, 2)
        ws["flow_out"] = z(N, H, W, 2)
        if self.occ is not None:
            ws["occ"] = z(N, H, W, 1)
        self._ws[key] = ws
        return ws

    def _use_chain(self, buf):
        return self.chain and buf.shape[0] * buf.shape[1] * buf.shape[2] <= CHAIN_MAX_PIXELS

    def _chain_layer(self, cs, x, out, residual=None, out2=None):
        return ops.chain_layer(x, cs.w_tc, cs.bias, out, cs.k, cs.dil, cs.slope, residual, self.rnd and cs.hidden, out2)

    def _sgu_dense(self, S, inter):
        """FlowEstimatorDense_temp (model/upflow.py:24-60) on the S buffer."""
        if self._use_chain(S):
            layers = [self._chain_layer(self.sgu[k], Slice(S, 0, self.sgu[k].cin), Slice(S, S_OFF[k], SGU_CH[k])) for k in range(5)]
            layers.append(self._chain_layer(self.sgu[5], Slice(S, 0, self.sgu[5].cin), Slice(inter, 0, 3)))
            ops.k_conv_chain(layers)
            return
        for k in range(5):
            self._conv(self.sgu[k], Slice(S, 0, self.sgu[k].cin), Slice(S, S_OFF[k], SGU_CH[k]))
        self._conv(self.sgu[5], Slice(S, 0, self.sgu[5].cin), Slice(inter, 0, 3))

    # ------------------------------------------------------------ forward
    def encode(self, ws, im1, im2):
        """FeatureExtractor (model/pwc_modules.py:122-142) on [im1; im2]."""
        B = im1.shape[0]
        im = ws["im"]
        for i, x in enumerate((im1, im2)):
            v = x.permute(0, 2, 3, 1)
            if x.is_contiguous():
                _ext.check(_ext.load().upf_nchw_to_nhwc(ops._p(x), ops._p(im, i * B * im.shape[1] * im.shape[2] * 4), 4,
                                                        B, 3, x.shape[2], x.shape[3], ops._stream()), "nchw_to_nhwc")
            else:
                ops.k_copy(Slice(v.contiguous()), Slice(im[i * B:(i + 1) * B], 0, 3))
        x = Slice(im, 0, 3)
        for l in range(6):
            self._conv(self.enc[l][0], x, Slice(ws["enc_a"][l]))
            self._conv(self.enc[l][1], Slice(ws["enc_a"][l]), Slice(ws["enc_b"][l]))
            x = Slice(ws["enc_b"][l])
        return ws["enc_b"]

    def forward(self, im1, im2, taps=None):
        """Returns (flow_f_out, flow_b_out, flows) like forward_2_frame_v3:
        NCHW-shaped [B,2,H,W] views; flows = per-level [fw, bw], finest first."""
        ops._require_cuda(im1, im2)
        B, _, H, W = im1.shape
        ws = self._workspace(B, H, W)
        N = 2 * B
        ac = self.align_corners
        if self.norm:
            ws["stats_arena"].zero_()                  # all feature / warp moment accumulators of this forward
        feats = self.encode(ws, im1, im2)              # index l -> 1/2^(l+1); decoder level L uses feats[5-L]
        # Work that depends on the images only -- the 1x1 adapters and feature statistics of levels 1..4 and
        # sgi_model.output_conv (two full-resolution convolutions) -- runs on a SIDE STREAM while the main stream
        # walks the coarse levels, which are a latency-bound chain of small launches that leaves most SMs idle.
        # Inside a captured graph the fork / join events become parallel branches.
        main = torch.cuda.current_stream()
        side = self._side_stream() if self.overlap else main     # overlap=False: one stream (per-kernel profiling)
        ev_adapters, ev_outconv = torch.cuda.Event(), torch.cuda.Event()
        if side is not main:
            side.wait_stream(main)
        with torch.cuda.stream(side):
            for L in range(1, 5):
                d = ws["levels"][L]
                F = Slice(feats[5 - L])
                self._conv(self.conv1x1[L], F, Slice(d["X"], X_F1X1, 32))
                if self.norm:
                    ops.k_stats(F, d["stats_own"])
            ev_adapters.record(side)
            if self.use_sgu:
                # sgi_model.output_conv on both images (model/upflow.py:66-69, :527-528)
                self._conv(self.outconv[0], Slice(ws["im"], 0, 3), Slice(ws["oc_a"]))
                self._conv(self.outconv[1], Slice(ws["oc_a"]), Slice(ws["oc_b"]))
                self._conv(self.outconv[2], Slice(ws["oc_b"]), Slice(ws["oc_c"]))
                self._conv(self.outconv[3], Slice(ws["oc_c"]), Slice(ws["S_out"], 0, 32))
            ev_outconv.record(side)
        prev_flow = None
        flows = []
        for L in range(5):
            d = ws["levels"][L]
            h, w = d["hw"]
            F = Slice(feats[5 - L])
            C = NUM_CHS[L]
            X = d["X"]
            flow_up = Slice(d["fu"])                    # exact; X[113:117] = (its TF32 copy, 0, 0) for the convolutions
            x_flow = Slice(X, X_FLOW, 4)
            if L == 0:
                # 1x1 adapter (model/upflow.py:508-513) straight into its estimator slot
                self._conv(self.conv1x1[L], F, Slice(X, X_F1X1, 32))
                if self.norm:
                    ops.k_stats(F, d["stats_own"])
            elif L == 1:
                main.wait_event(ev_adapters)
            if L == 0:
                # upsampling a zero flow (model/upflow.py:504-505, :536): "fu" is zero from allocation on and nothing
                # writes it; the X slot is cleared every forward (its second half is rewritten below)
                ops.k_copy(None, x_flow)
                self._split(x_flow)
            else:
                ph, pw = ws["levels"][L - 1]["hw"]
                if self.use_sgu:
                    bil = Slice(d["flow_bil"])
                    ops.k_resize(prev_flow, bil, (w / pw, h / ph))
                    S = d["S"]
                    ops.k_copy(Slice(X, X_F1X1, 32), Slice(S, 0, 32))
                    ops.k_warp(Slice(X, X_F1X1, 32), bil, Slice(S, 32, 32), ac, self.mask, x_shift=B, round_tf32=self.rnd)
                    self._split(Slice(S, 0, 64))
                    self._sgu_dense(S, d["inter"])
                    ops.k_sgu_blend(bil, Slice(d["inter"], 0, 3), flow_up, ac, out_tc=x_flow, round_tf32=self.rnd)
                else:
                    ops.k_resize(prev_flow, flow_up, (w / pw, h / ph))
                    ops.k_copy(None, Slice(X, X_FLOW2, 2))
                    ops.k_copy(flow_up, Slice(X, X_FLOW, 2), round_tf32=self.rnd)
                self._split(x_flow)
            # feature statistics (model/upflow.py:549-555: computed above / on the side stream), warp + its
            # statistics (:546-547)
            # if_norm_before_cost_volume (model/upflow.py:549-555): off = the raw features are correlated
            if L == 0:
                s1 = s2 = d["stats_own"] if self.norm else None
                f2, shift = F, B
            else:
                ops.k_warp(F, flow_up, Slice(d["xw"]), ac, self.mask, x_shift=B, stats=d["stats_w"] if self.norm else None)
                s1, s2 = (d["stats_own"], d["stats_w"]) if self.norm else (None, None)
                f2, shift = Slice(d["xw"]), 0
            if self.norm and (self.norm_ch or self.norm_img):
                ops.k_norm_combine(s1, s2, d["stats_ca"], d["stats_cb"], N, C, h * w, shift, self.norm_ch, self.norm_img)
                s1, s2 = d["stats_ca"], d["stats_cb"]
            ops.k_corr(F, f2, Slice(X, X_CORR, 81), 4, s1, s2, f2_shift=shift, slope=SLOPE, round_tf32=self.rnd)
            self._split(Slice(X, X_CORR, 81))
            flow2 = Slice(d["flow2"])
            bufs = (d["T0"], d["T1"])
            if self._use_chain(X):
                # estimator, flow_up + flow_res (exact into "flow2", TF32 copy into the context input slot) and the context
                # network: 13 dependent convolutions, one persistent launch
                layers = [self._chain_layer(self.est[k], Slice(X, 0, self.est[k].cin), Slice(X, X_OFF[k], EST_CH[k])) for k in range(5)]
                layers.append(self._chain_layer(self.est[5], Slice(X, 0, X_LD), flow2, residual=flow_up, out2=Slice(X, X_FLOW2, 2)))
                t_in = Slice(X, 0, X_LD)
                for i in range(6):
                    t_out = Slice(bufs[i % 2], 0, CTX_CH[i])
                    layers.append(self._chain_layer(self.ctx[i], t_in, t_out))
                    t_in = t_out
                layers.append(self._chain_layer(self.ctx[6], t_in, Slice(d["flow"]), residual=flow2))
                ops.k_conv_chain(layers)
            else:
                # dense flow estimator (model/pwc_modules.py:279-286)
                for k in range(5):
                    self._conv(self.est[k], Slice(X, 0, self.est[k].cin), Slice(X, X_OFF[k], EST_CH[k]))
                # flow_up + flow_res (model/upflow.py:567): exact into "flow2", TF32 copy into the context input slot
                self._conv(self.est[5], Slice(X, 0, X_LD), flow2, residual=flow_up)
                ops.k_copy(flow2, Slice(X, X_FLOW2, 2), round_tf32=self.rnd)
                self._split(Slice(X, X_FLOW2, 2))
                # context network (model/pwc_modules.py:401-412); last conv adds (flow_up + flow_res): :569-572, :519
                t_in = Slice(X, 0, X_LD)
                for i in range(6):
                    t_out = Slice(bufs[i % 2], 0, CTX_CH[i])
                    self._conv(self.ctx[i], t_in, t_out)
                    t_in = t_out
                self._conv(self.ctx[6], t_in, Slice(d["flow"]), residual=flow2)
            prev_flow = Slice(d["flow"])
            flows.append(d["flow"])
            if taps is not None:
                taps.append({"level": L, "flow_up": d["fu"].permute(0, 3, 1, 2).clone(),
                             "corr": X[..., :81].permute(0, 3, 1, 2).clone(),
                             "xw": d["xw"].permute(0, 3, 1, 2).clone() if L > 0 else None,
                             "flow": d["flow"].permute(0, 3, 1, 2).clone()})
        # ---- 1/4 -> full resolution (model/upflow.py:522-530)
        h4, w4 = ws["levels"][4]["hw"]
        out = Slice(ws["flow_out"])
        if self.use_sgu:
            bil = Slice(ws["flow_full_bil"])
            ops.k_resize(prev_flow, bil, (W / w4, H / h4))
            S = ws["S_out"]
            main.wait_event(ev_outconv)                # output_conv features (side stream)
            # the flow handed to sgu_model here is the 1/4-res flow itself (already at feature size, :73-75)
            ops.k_warp(Slice(S, 0, 32), prev_flow, Slice(S, 32, 32), ac, self.mask, x_shift=B, round_tf32=self.rnd)
            self._split(Slice(S, 32, 32))
            self._sgu_dense(S, ws["inter_out"])
            ops.k_sgu_blend(bil, Slice(ws["inter_out"], 0, 3), out, ac)
        else:
            ops.k_resize(prev_flow, out, (W / w4, H / h4))
            main.wait_event(ev_outconv)
        self.last_occ = None
        if self.occ is not None:
            ops.k_occ_check(out, Slice(ws["occ"]), self.occ[0], self.occ[1], self.occ[2], ac)
            self.last_occ = ws["occ"]
        fo = ws["flow_out"].permute(0, 3, 1, 2)
        lvl = [[f[:B].permute(0, 3, 1, 2), f[B:].permute(0, 3, 1, 2)] for f in flows]
        return fo[:B], fo[B:], lvl[::-1]

    def release_workspaces(self):
        """Drop every per-shape workspace (hundreds of MB each at KITTI size), scratch and low-part twin buffer.
        UPFlow_net calls this when it drops its captured graphs, so evaluating a dataset with many image sizes does not
        grow GPU memory without bound; graphs captured on the old workspaces must be dropped with them."""
        self._ws.clear()
        self._lo_bufs.clear()

    def flow_out_buffer(self, shape):
        """the [2B,H,W,2] buffer the last forward of this input shape wrote (forward flows first)"""
        return self._workspace(shape[0], shape[2], shape[3])["flow_out"]

    # ------------------------------------------------------------ CUDA graph
    def capture(self, B, H, W):
        """Capture the whole two-frame forward for one input shape in a CUDA graph (the forward has no host
        decisions, allocations or synchronisation).  Returns a GraphedForward."""
        im1 = torch.zeros(B, 3, H, W, dtype=torch.float32, device=self.device)
        im2 = torch.zeros_like(im1)
        self.forward(im1, im2)                         # eager warm-up: workspace, tensor maps, func attributes
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        n0 = _ext.launch_count()
        with torch.cuda.graph(graph):
            out = self.forward(im1, im2)
        return GraphedForward(graph, im1, im2, out, _ext.launch_count() - n0, self.last_occ)


class GraphedForward:
    """Replay handle: copy a pair into (im1, im2), call replay(), read flow_f / flow_b (views of the workspace)."""

    def __init__(self, graph, im1, im2, out, launches, occ=None):
        self.graph, self.im1, self.im2 = graph, im1, im2
        self.occ = occ                                 # [2B,H,W,1] visibility masks (forward directions first) or None
        self.flow_f, self.flow_b, self.flows = out
        self.launches = launches                       # library kernels per replay

    def replay(self):
        self.graph.replay()

    def __call__(self, im1, im2):
        self.im1.copy_(im1, non_blocking=True)
        self.im2.copy_(im2, non_blocking=True)
        self.graph.replay()
        return self.flow_f, self.flow_b
