"""KITTI evaluation helpers: the metrics test.py reports (README.md:10 of the reference)
and the 16-bit KITTI flow PNG codec, without tensorflow / pypng.

    flow_error_avg, outlier_pct   <- kitti_flow.Evaluation_bench (dataset/kitti_dataset.py:464-499)
    read_png_flow                 <- kitti_train.read_png_flow (dataset/kitti_dataset.py:130-149; pypng there)
    write_kitti_png_file          <- tools.write_kitti_png_file (utils/tools.py:1516-1525)

Host-side utilities around the decoder (tensors in, Python floats / files out); torch and OpenCV only.
"""
import numpy as np
import torch


def _euclidean(t):
    return torch.sqrt(torch.sum(t ** 2, dim=(1,), keepdim=True))


def flow_error_avg(flow_1, flow_2, mask):
    """Average end-point error over the masked pixels; [N,2,H,W], [N,2,H,W], [N,1,H,W]."""
    diff = _euclidean(flow_1 - flow_2) * mask
    return torch.sum(diff) / (torch.sum(mask) + 1e-6)


def outlier_pct(gt_flow, predflow, mask, threshold=3.0, relative=0.05):
    """KITTI Fl: percentage of masked pixels whose error exceeds max(threshold px, relative * |gt|)."""
    diff = _euclidean(gt_flow - predflow) * mask
    thr = torch.tensor(threshold).type_as(gt_flow)
    if relative is not None:
        outliers = diff > torch.max(thr, _euclidean(gt_flow) * relative)
    else:
        outliers = diff > thr
    return torch.sum(outliers) / torch.sum(mask) * 100


def evaluate(predflow, occ_flow, occ_mask, noc_flow, noc_mask):
    """(EPE all, Fl all, EPE noc, EPE occ) of one batch, as Evaluation_bench.__call__ accumulates them
    (dataset/kitti_dataset.py:433-452)."""
    epe_all = flow_error_avg(predflow, occ_flow, occ_mask)
    f1 = outlier_pct(occ_flow, predflow, occ_mask)
    epe_noc = flow_error_avg(predflow, noc_flow, noc_mask)
    epe_occ = flow_error_avg(predflow, occ_flow, occ_mask - noc_mask)
    return epe_all.item(), f1.item(), epe_noc.item(), epe_occ.item()


def read_png_flow(fpath):
    """KITTI 16-bit flow PNG -> (flow [2,H,W] float64, valid [1,H,W] uint8); channels R,G = u,v as
    (value - 2^15) / 64, B = valid."""
    import cv2
    bgr = cv2.imread(fpath, cv2.IMREAD_UNCHANGED)
    if bgr is None or bgr.dtype != np.uint16 or bgr.ndim != 3 or bgr.shape[2] != 3:
        raise ValueError("%s is not a 3-channel 16-bit PNG" % fpath)
    rgb = bgr[:, :, ::-1]
    flow = (rgb[:, :, 0:2].astype('float64') - 2 ** 15) / 64.0
    mask = np.uint8(rgb[:, :, 2:3])
    return np.transpose(flow, [2, 0, 1]), np.transpose(mask, [2, 0, 1])


def write_kitti_png_file(flow_fn, flow_data, mask_data=None):
    """flow_data [H,W,2] (u,v) -> KITTI 16-bit PNG; OpenCV writes B,G,R = valid, v, u."""
    import cv2
    flow_img = np.zeros((flow_data.shape[0], flow_data.shape[1], 3), dtype=np.uint16)
    if mask_data is None:
        mask_data = np.ones([flow_data.shape[0], flow_data.shape[1]], dtype=np.uint16)
    flow_img[:, :, 2] = (flow_data[:, :, 0] * 64.0 + 2 ** 15).astype(np.uint16)
    flow_img[:, :, 1] = (flow_data[:, :, 1] * 64.0 + 2 ** 15).astype(np.uint16)
    flow_img[:, :, 0] = mask_data[:, :]
    if not cv2.imwrite(flow_fn, flow_img):
        raise IOError("could not write %s" % flow_fn)
