"""CPU tier: the closed-form gradients the loss kernels of csrc/loss.cu implement (gather forms), restated in plain
Python loops and checked against torch autograd of the reference-pinned torch expressions in fp64.  The kernels
themselves are checked against the same expressions on the GPU (tests/test_gpu_backward.py)."""
import torch

import upflow_pytorch_b200

upflow_pytorch_b200.install_dropin()
from model.upflow import network_tools          # noqa: E402
from utils.loss import loss_functions          # noqa: E402
from utils.tools import tools                  # noqa: E402

GREY = torch.tensor([0.2989, 0.5870, 0.1140], dtype=torch.float64)


def test_census_gather_gradient_formula():
    """census_bwd_kernel: a pixel collects its role as the centre of its own patch and as a neighbour in the patches
    around it; coefficient d(term)/d(dist) = q (dist+0.01)^(q-1) m / denominator (utils/loss.py:51-91, :44-46)."""
    torch.manual_seed(0)
    H, W, d, q = 7, 8, 2, 0.4
    a = torch.rand(1, 3, H, W, dtype=torch.float64)
    b = (a + 0.1 * torch.randn(1, 3, H, W, dtype=torch.float64)).requires_grad_()
    mask = (torch.rand(1, 1, H, W) > 0.3).double()
    g1 = (a[0] * GREY[:, None, None]).sum(0)
    g2 = (b.detach()[0] * GREY[:, None, None]).sum(0)

    def at(g, y, x):
        return g[y, x].item() if 0 <= y < H and 0 <= x < W else 0.0

    def tau(v):
        return v / (0.81 + v * v) ** 0.5

    def pair(n1, n2, c1, c2):                      # h'(u) * tau'(v2): census_pair_grad
        v1, v2 = n1 - c1, n2 - c2
        u = tau(v1) - tau(v2)
        return (0.2 * u / (0.1 + u * u) ** 2) * (0.81 / (0.81 + v2 * v2) ** 1.5)

    offs = [(dy, dx) for dy in range(-d, d + 1) for dx in range(-d, d + 1)]
    dist = [[sum((lambda u: u * u / (0.1 + u * u))(tau(at(g1, y + dy, x + dx) - at(g1, y, x)) - tau(at(g2, y + dy, x + dx) - at(g2, y, x)))
                 for dy, dx in offs) for x in range(W)] for y in range(H)]
    for masked in (False, True):
        want = loss_functions.census_loss_torch(a, b, mask, q, False, masked, True, max_distance=d)
        (wb,) = torch.autograd.grad(want, (b,))

        def m(y, x):
            if not masked:
                return 1.0
            return mask[0, 0, y, x].item() if (d <= y < H - d and d <= x < W - d) else 0.0
        sm = sum(m(y, x) for y in range(H) for x in range(W))
        inv = 1.0 / (2 * sm + 1e-6) if masked else 1.0 / (H * W)
        value = inv * sum((dist[y][x] + 0.01) ** q * m(y, x) for y in range(H) for x in range(W))
        assert abs(value - want.item()) <= 1e-12

        def coef(y, x):
            return inv * q * m(y, x) * (dist[y][x] + 0.01) ** (q - 1)
        for y in range(H):
            for x in range(W):
                acc = 0.0
                for dy, dx in offs:
                    acc += coef(y, x) * pair(at(g1, y + dy, x + dx), at(g2, y + dy, x + dx), at(g1, y, x), at(g2, y, x))
                    cy, cx = y - dy, x - dx
                    if 0 <= cy < H and 0 <= cx < W:
                        acc -= coef(cy, cx) * pair(at(g1, y, x), at(g2, y, x), at(g1, cy, cx), at(g2, cy, cx))
                for k in range(3):
                    assert abs(acc * GREY[k].item() - wb[0, k, y, x].item()) <= 1e-12, (masked, y, x, k)


def test_edge_smoothness_gather_gradient_formula():
    """edge_smooth1_bwd_kernel: a pixel collects the (at most four) differences it takes part in, each weighted by
    exp(-mean_c |d img|) and the mean's 1/count (model/upflow.py:198-218)."""
    torch.manual_seed(1)
    N, H, W = 2, 5, 6
    img = torch.rand(N, 3, H, W, dtype=torch.float64)
    pred = ((torch.randn(N, 2, H, W, dtype=torch.float64) * 2).round() / 2).requires_grad_()    # exact ties: sign(0) = 0
    network_tools.use_loss_kernels = False
    try:
        want = network_tools.edge_aware_smoothness_order1(img, pred)
    finally:
        network_tools.use_loss_kernels = True
    (wg,) = torch.autograd.grad(want, (pred,))
    inv_r, inv_c = 1.0 / (N * 2 * (H - 1) * W), 1.0 / (N * 2 * H * (W - 1))
    sgn = lambda v: (v > 0) - (v < 0)
    p = pred.detach()
    for n in range(N):
        for y in range(H):
            for x in range(W):
                wr = lambda yy: torch.exp(-(img[n, :, yy, x] - img[n, :, yy + 1, x]).abs().mean()).item()
                wc = lambda xx: torch.exp(-(img[n, :, y, xx] - img[n, :, y, xx + 1]).abs().mean()).item()
                for c in range(2):
                    v = p[n, c, y, x].item()
                    g = 0.0
                    if y < H - 1:
                        g += sgn(v - p[n, c, y + 1, x].item()) * wr(y) * inv_r
                    if y > 0:
                        g -= sgn(p[n, c, y - 1, x].item() - v) * wr(y - 1) * inv_r
                    if x < W - 1:
                        g += sgn(v - p[n, c, y, x + 1].item()) * wc(x) * inv_c
                    if x > 0:
                        g -= sgn(p[n, c, y, x - 1].item() - v) * wc(x - 1) * inv_c
                    assert abs(g - wg[n, c, y, x].item()) <= 1e-12, (n, c, y, x)


def test_boundary_warp_flow_gradient_formula():
    """bdwarp_bwd_kernel: floor and clamp are piecewise constant, so d/du = sum_k g_k [(y1c-y)(Ic-Ia) + (y-y0c)(Id-Ib)]
    and d/dv = sum_k g_k [(x1c-x)(Ib-Ia) + (x-x0c)(Id-Ic)] with the CLAMPED corners (utils/tools.py:383-470)."""
    torch.manual_seed(2)
    N, Hf, Wf, h, w = 1, 9, 11, 5, 6
    frame = torch.rand(N, 3, Hf, Wf).double()                 # fp32-representable: the reference casts the frame to float
    start = torch.tensor([[2.0, 3.0]], dtype=torch.float64).reshape(N, 2, 1, 1)
    flow = ((torch.randn(N, 2, h, w, dtype=torch.float64) * 4 * 64).round() / 64 + 1.0 / 128).requires_grad_()
    r = torch.randn(N, 3, h, w, dtype=torch.float64)
    tools.boundary_dilated_warp.use_kernel = False
    try:
        out = tools.boundary_dilated_warp.warp_im(frame, flow, start)
    finally:
        tools.boundary_dilated_warp.use_kernel = True
    (wg,) = torch.autograd.grad((out * r).sum(), (flow,))
    clamp = lambda v, hi: min(max(v, 0.0), float(hi))
    import math
    outside = 0
    for y in range(h):
        for x in range(w):
            fx = x + 2.0 + flow[0, 0, y, x].item()
            fy = y + 3.0 + flow[0, 1, y, x].item()
            outside += fx < 0 or fx > Wf - 1 or fy < 0 or fy > Hf - 1
            x0, y0 = math.floor(fx), math.floor(fy)
            x0c, x1c, y0c, y1c = clamp(x0, Wf - 1), clamp(x0 + 1, Wf - 1), clamp(y0, Hf - 1), clamp(y0 + 1, Hf - 1)
            Ia, Ib = frame[0, :, int(y0c), int(x0c)], frame[0, :, int(y1c), int(x0c)]
            Ic, Id = frame[0, :, int(y0c), int(x1c)], frame[0, :, int(y1c), int(x1c)]
            gu = (r[0, :, y, x] * ((y1c - fy) * (Ic - Ia) + (fy - y0c) * (Id - Ib))).sum().item()
            gv = (r[0, :, y, x] * ((x1c - fx) * (Ib - Ia) + (fx - x0c) * (Id - Ic))).sum().item()
            assert abs(gu - wg[0, 0, y, x].item()) <= 1e-12 and abs(gv - wg[0, 1, y, x].item()) <= 1e-12, (y, x)
    assert outside > 0                                         # the clamped branch was exercised


def test_robust_term_gradient_formula():
    """robust_loss_bwd_kernel: q (|d|+0.01)^(q-1) sign(d) mask / (sum(mask)+1e-6) (model/upflow.py:268-290)."""
    torch.manual_seed(3)
    x = torch.randn(2, 3, 4, 5, dtype=torch.float64).requires_grad_()
    y = torch.randn(2, 3, 4, 5, dtype=torch.float64)
    mask = (torch.rand(2, 1, 4, 5) > 0.4).double()
    network_tools.use_loss_kernels = False
    try:
        want = network_tools.photo_loss_multi_type(x, y, mask, 'abs_robust', 0.4, photo_loss_use_occ=True)
    finally:
        network_tools.use_loss_kernels = True
    (wg,) = torch.autograd.grad(want, (x,))
    d = (x - y).detach()
    got = 0.4 * (d.abs() + 0.01) ** (0.4 - 1) * d.sign() * mask / (mask.sum() + 1e-6)
    assert (got - wg).abs().max().item() <= 1e-12
