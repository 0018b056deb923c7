"""Two GPUs, NCCL: a 2-rank data-parallel Trainer step against the 1-rank step on the same global batch (SURVEY section
4 iv, section 8e).  Runs only where two GPUs are visible (`gpurun --gpus 2`); world-size-2 host logic on CPU is covered
by tests/test_train_host.py with gloo."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

CONF = {"if_norm_before_cost_volume": True, "norm_moments_across_channels": False, "norm_moments_across_images": False,
        "if_sgu_upsample": True, "if_use_boundary_warp": False, "multi_scale_distillation_weight": 0.01}


def _net(precision, msd=0.01):
    sys.path.insert(0, ROOT)
    import upflow_pytorch_b200
    from oracle import ref_port as P
    upflow_pytorch_b200.install_dropin()
    from model.upflow import UPFlow_net
    conf = UPFlow_net.config()
    conf.update(dict(CONF, multi_scale_distillation_weight=msd))
    net = conf()
    net.load_state_dict(P.det_state_dict(11))
    net.conv_precision = precision
    return net.cuda().train()


def _batch():
    from oracle import cpu_oracle as O
    im1, im2 = O.synthetic_pair(64, 96, seed=5, batch=4)
    return {"im1": im1, "im2": im2}


def _worker(rank, world, port, q, msd):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        from upflow_pytorch_b200.train import Trainer, shard_batch
        net = _net("fp32", msd)
        tr = Trainer(net, lr=1e-4)
        shard = {k: v.cuda() for k, v in shard_batch(_batch(), rank, world).items()}
        loss = tr.train_step(shard)
        flat_params = torch.cat([p.detach().flatten() for p in tr.grads.params])
        q.put((rank, loss.item(), tr.grads.flat.cpu(), flat_params.cpu()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("msd,rel_tol,cos_tol", [(0.0, 1e-4, 0.9999999), (0.01, 5e-2, 0.999)])
def test_two_rank_trainer_step_equals_one_rank_step_on_the_global_batch(msd, rel_tol, cos_tol):
    """Gradients after the NCCL all-reduce (sum / world size) are identical on both ranks and equal the full-batch
    gradient of one rank up to SURVEY 8e's caveat: the photometric and smoothness terms are means over all pixels of the
    batch (mean of the shard losses = global loss), the distillation term is a MASKED mean, sum(l*m) / (sum(m) + 1e-6)
    (model/upflow.py:160-161), which is not linear in the shards -- measured on 2xB200: 1.2e-4 relative on the loss,
    1.8e-2 relative L2 / cos 0.99984 on the gradient with it, and fp32 summation order only without it."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 2000 + (7 if msd else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, msd)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r, loss, grads, params = q.get(timeout=600)
        res[r] = (loss, grads, params)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert torch.equal(res[0][1], res[1][1]), "ranks disagree after the all-reduce"
    assert torch.equal(res[0][2], res[1][2]), "replicas diverged after one Adam step"
    # one rank, the whole batch
    from upflow_pytorch_b200.train import Trainer
    net = _net("fp32", msd)
    tr = Trainer(net, lr=1e-4)
    loss1 = tr.train_step({k: v.cuda() for k, v in _batch().items()}).item()
    g1, g2 = tr.grads.flat.cpu().double(), res[0][1].double()
    rel = ((g1 - g2).norm() / g1.norm()).item()
    cos = (torch.dot(g1, g2) / (g1.norm() * g2.norm())).item()
    mean_shard_loss = 0.5 * (res[0][0] + res[1][0])
    print("  2-rank vs 1-rank: loss %.6f vs %.6f, gradient rel L2 %.3g, cos %.6f" % (mean_shard_loss, loss1, rel, cos))
    assert abs(mean_shard_loss - loss1) <= 1e-3 * abs(loss1)
    assert rel <= rel_tol and cos >= cos_tol
