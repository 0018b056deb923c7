"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's TRAINING forward (UPFlow_net.forward with if_loss=True,
model/upflow.py:370-491) as an op-for-op torch port on top of oracle/ref_port.py: occlusion check, edge-aware
smoothness, photometric term, multi-scale distillation.  Differentiable with torch autograd on any device, so it serves

  * as the "reference GPU path" anchor of the training step in bench.py (eager PyTorch + cuDNN + autograd on the same
    GPU, `reference_gpu_path` of the `train` record), and
  * as a second, independent statement of the loss branch next to the drop-in's (tests/test_oracle_golden.py checks it
    against tests/golden/train_step.pt, produced by the reference's own step).

Covers the configuration of BASELINE config 4 and of that golden: plain torch_warp photometric term (no
boundary-dilated warp), no census term.  The product never imports this file.
"""
import torch
import torch.nn.functional as F

from oracle import ref_port as P


def _mag(x):
    # length_sq_v0   utils/tools.py:556-561  (sum of |component|)
    return torch.sum(torch.pow(x ** 2, 0.5), dim=1, keepdim=True)


def _outgoing(flow):
    # tools.occ_check_model.torch_outgoing_occ_check   utils/tools.py:641-668
    B, _, H, W = flow.shape
    xx = torch.arange(0, W, device=flow.device).view(1, -1).repeat(H, 1).view(1, 1, H, W).repeat(B, 1, 1, 1).float()
    yy = torch.arange(0, H, device=flow.device).view(-1, 1).repeat(1, W).view(1, 1, H, W).repeat(B, 1, 1, 1).float()
    fx, fy = torch.split(flow, 1, 1)
    px, py = xx + fx, yy + fy
    m = torch.ones_like(px)
    m[px > W - 1] = 0
    m[px < 0] = 0
    m[py > H - 1] = 0
    m[py < 0] = 0
    return m.float()


def occ_check(flow_f, flow_b, alpha_1=0.1, alpha_2=0.5, obj_out_all="obj"):
    # tools.occ_check_model.__call__ / _forward_backward_occ_check   utils/tools.py:519-588
    mag = _mag(flow_f) + _mag(flow_b)
    diff_f = flow_f + P.torch_warp(flow_b, flow_f)
    diff_b = flow_b + P.torch_warp(flow_f, flow_b)
    thr = alpha_1 * mag + alpha_2
    occ_f, occ_b = (_mag(diff_f) < thr).float(), (_mag(diff_b) < thr).float()
    if obj_out_all == "all":
        return occ_f, occ_b
    out_f, out_b = _outgoing(flow_f), _outgoing(flow_b)
    if obj_out_all == "out":
        return out_f, out_b

    def obj(occ, out):
        # torch_get_obj_occ_check   utils/tools.py:670-677
        m = torch.zeros_like(occ)
        m[occ == 1] = 1
        m[out == 0] = 1
        return m
    return obj(occ_f, out_f), obj(occ_b, out_b)


def edge_smooth1(img, pred):
    # network_tools.edge_aware_smoothness_order1   model/upflow.py:198-218
    gx = lambda t: t[:, :, :-1, :] - t[:, :, 1:, :]
    gy = lambda t: t[:, :, :, :-1] - t[:, :, :, 1:]
    wx = torch.exp(-torch.mean(torch.abs(gx(img)), 1, keepdim=True))
    wy = torch.exp(-torch.mean(torch.abs(gy(img)), 1, keepdim=True))
    return torch.mean(torch.abs(gx(pred)) * wx) + torch.mean(torch.abs(gy(pred)) * wy)


def edge_smooth2(img, pred):
    # network_tools.edge_aware_smoothness_order2   model/upflow.py:220-245
    gx = lambda t, s=1: t[:, :, :-s, :] - t[:, :, s:, :]
    gy = lambda t, s=1: t[:, :, :, :-s] - t[:, :, :, s:]
    wx = torch.exp(-torch.mean(torch.abs(gx(img, 2)), 1, keepdim=True))
    wy = torch.exp(-torch.mean(torch.abs(gy(img, 2)), 1, keepdim=True))
    return torch.mean(torch.abs(gx(gx(pred))) * wx) + torch.mean(torch.abs(gy(gy(pred))) * wy)


def robust(x, y, occ, use_occ, q=0.4):
    # network_tools.photo_loss_multi_type, 'abs_robust'   model/upflow.py:268-290
    d = (torch.abs(x - y) + 0.01).pow(q)
    if use_occ:
        return torch.sum(d * occ) / (torch.sum(occ) + 1e-6)
    return torch.mean(d)


def upsample_flow(x, h, w):
    # model/pwc_modules.py:93-104 (out of place: the reference's in-place scaling of an interpolate result is the same)
    h_, w_ = x.shape[2:]
    res = F.interpolate(x, [h, w], mode="bilinear", align_corners=True)
    return torch.cat([res[:, 0:1] * (w / w_), res[:, 1:2] * (h / h_)], dim=1)


def training_loss(im1, im2, sd, smooth1_weight=1.0, smooth2_weight=0.0, photo_use_occ=False, msd_weight=0.01, msd_occ=True,
                  alpha_1=0.1, alpha_2=0.5, obj_out_all="obj", photo_delta=0.4):
    """UPFlow_net.forward(if_loss=True) (model/upflow.py:370-491) for smooth_level='final', smooth_type='edge',
    photo_loss_type='abs_robust', if_use_boundary_warp=False, census weight 0, msd style 'upup'.  sd: name -> tensor
    (leaf tensors requiring grad for training).  Returns the terms and their sum (Loss_manager.compute_loss,
    scripts/simple_train.py:45-53)."""
    flow_f, flow_b, flows = P.forward_2_frame(im1, im2, sd)
    occ_f, occ_b = occ_check(flow_f, flow_b, alpha_1, alpha_2, obj_out_all)
    out = {"flow_f_out": flow_f, "flow_b_out": flow_b, "occ_fw": occ_f, "occ_bw": occ_b}
    smooth = 0
    if smooth1_weight > 0:
        smooth = smooth + smooth1_weight * edge_smooth1(im1, flow_f) + smooth1_weight * edge_smooth1(im2, flow_b)
    if smooth2_weight > 0:
        smooth = smooth + smooth2_weight * edge_smooth2(im1, flow_f) + smooth2_weight * edge_smooth2(im2, flow_b)
    out["smooth_loss"] = smooth
    im1_warp, im2_warp = P.torch_warp(im2, flow_f), P.torch_warp(im1, flow_b)
    out["photo_loss"] = robust(im1, im1_warp, occ_f, photo_use_occ, photo_delta) + robust(im2, im2_warp, occ_b, photo_use_occ, photo_delta)
    loss = out["photo_loss"] + smooth
    if msd_weight > 0:
        lf, lb = flow_f.clone().detach(), flow_b.clone().detach()
        H, W = lf.shape[2:]
        terms = []
        for sf, sb in flows:
            terms.append(robust(upsample_flow(sf, H, W), lf, occ_f, msd_occ))
            terms.append(robust(upsample_flow(sb, H, W), lb, occ_b, msd_occ))
        out["msd_loss"] = msd_weight * sum(terms)
        loss = loss + out["msd_loss"]
    out["loss"] = loss
    return out
