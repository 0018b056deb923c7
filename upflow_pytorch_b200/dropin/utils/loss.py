"""Drop-in for the part of the reference's utils/loss.py the model uses (model/upflow.py:447-455):
`loss_functions.photo_loss_function` and `loss_functions.census_loss_torch`.  Loss-side ops on 3-channel images
(SURVEY.md section 8f rank 2): the census term runs as fused kernels (csrc/loss.cu) for CUDA tensors; the torch
expressions are what the kernels are tested against (`loss_functions.use_loss_kernels = False`) and serve the cases the
kernels do not take (charbonnier penalty, a first image or mask that needs a gradient, CPU tensors in the oracle tests)."""
import torch
import torch.nn.functional as F


class loss_functions():
    use_loss_kernels = True

    @classmethod
    def photo_loss_function(cls, diff, mask, q, charbonnier_or_abs_robust, if_use_occ, averge=True):
        """utils/loss.py:17-49 (note the reference's factor 2 on the mask sum in the occlusion-aware branches)."""
        if charbonnier_or_abs_robust:
            if if_use_occ:
                p = ((diff) ** 2 + 1e-6).pow(q) * mask
                p, ap = (p.mean(), mask.mean()) if averge else (p.sum(), mask.sum())
                return p / (ap * 2 + 1e-6)
            p = ((diff) ** 2 + 1e-8).pow(q)
            return p.mean() if averge else p.sum()
        d = (torch.abs(diff) + 0.01).pow(q)
        if if_use_occ:
            return torch.sum(d * mask) / (torch.sum(mask) * 2 + 1e-6)
        return d.mean() if averge else d.sum()

    @classmethod
    def census_loss_torch(cls, img1, img1_warp, mask, q, charbonnier_or_abs_robust, if_use_occ, averge=True, max_distance=3):
        """utils/loss.py:51-91: soft ternary census transform over a (2d+1)^2 patch of the grey image, soft Hamming
        distance, border mask.  The reference extracts the patch with a one-hot 49-channel conv2d; here the 49 shifted
        copies are slices of the zero-padded grey image (same values)."""
        if cls.use_loss_kernels and img1.is_cuda and not charbonnier_or_abs_robust and averge and not img1.requires_grad \
                and mask is not None and not mask.requires_grad and img1.shape[1] == 3 and 1 <= max_distance <= 8:
            from upflow_pytorch_b200 import ops
            return ops.census_loss(img1.float(), img1_warp, mask if if_use_occ else None, q, max_distance)   # csrc/loss.cu
        d = max_distance
        n = 2 * d + 1

        def ternary(image):
            R, G, B = torch.split(image, 1, 1)
            grey = 0.2989 * R + 0.5870 * G + 0.1140 * B
            H, W = grey.shape[2:]
            pad = F.pad(grey, [d, d, d, d])
            patches = torch.cat([pad[:, :, i:i + H, j:j + W] for i in range(n) for j in range(n)], dim=1)
            t = patches - grey
            return t / torch.sqrt(0.81 + t ** 2)

        def hamming(t1, t2):
            dist = (t1 - t2) ** 2
            return torch.sum(dist / (0.1 + dist), 1, keepdim=True)

        dist = hamming(ternary(img1), ternary(img1_warp))
        inner = torch.ones(mask.shape[0], mask.shape[1], mask.shape[2] - 2 * d, mask.shape[3] - 2 * d,
                           dtype=torch.float32, device=mask.device)
        transform_mask = F.pad(inner, [d, d, d, d])
        return cls.photo_loss_function(diff=dist, mask=mask * transform_mask, q=q,
                                       charbonnier_or_abs_robust=charbonnier_or_abs_robust, if_use_occ=if_use_occ,
                                       averge=averge)
