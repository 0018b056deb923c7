"""Host-side data-parallel logic (upflow_pytorch_b200/train.py) on CPU with the gloo backend, world size 2:
batch sharding, the flat gradient buffer and its single all-reduce.  The network itself is CUDA-only, so a small CPU
module stands in for it here; the GPU tier trains the real one."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from upflow_pytorch_b200 import train


def _model():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.LeakyReLU(0.1), torch.nn.Conv2d(4, 2, 3, padding=1))


def _batch():
    g = torch.Generator().manual_seed(4)
    return {"im1": torch.randn(4, 3, 8, 10, generator=g), "im2": torch.randn(4, 3, 8, 10, generator=g), "if_loss": True}


def _loss(net, b):
    # per-image mean, then batch mean: equal shards => mean of shard losses == full-batch loss
    return ((net(b["im1"]) - net(b["im2"])) ** 2).mean()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _model()
        fg = train.FlatGradients(net.parameters())
        shard = train.shard_batch(_batch(), rank, world)
        assert shard["im1"].shape[0] == 2 and shard["if_loss"] is True
        fg.zero()
        _loss(net, shard).backward()
        nbytes = fg.all_reduce_mean()
        assert nbytes == fg.numel * 4
        q.put((rank, fg.flat.clone()))
    finally:
        dist.destroy_process_group()


def test_shard_batch_and_flat_gradients_single_process():
    b = _batch()
    s0, s1 = train.shard_batch(b, 0, 2), train.shard_batch(b, 1, 2)
    assert torch.equal(torch.cat([s0["im1"], s1["im1"]]), b["im1"])
    with pytest.raises(ValueError):
        train.shard_batch(b, 0, 3)
    net = _model()
    fg = train.FlatGradients(net.parameters())
    _loss(net, b).backward()
    assert fg.flat.abs().sum() > 0
    o = 0
    for p in fg.params:                                   # grads are views of the flat buffer, in parameter order
        assert p.grad.data_ptr() == fg.flat.data_ptr() + 4 * o
        assert torch.equal(p.grad.flatten(), fg.flat[o:o + p.numel()])
        o += p.numel()
    opt = torch.optim.Adam(fg.params, lr=1e-3)
    opt.zero_grad(set_to_none=True)
    fg.zero()
    assert all(p.grad is not None for p in fg.params) and fg.flat.abs().sum() == 0
    assert fg.all_reduce_mean() == fg.numel * 4           # no process group: a no-op


def test_two_rank_allreduce_equals_full_batch_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    net = _model()
    fg = train.FlatGradients(net.parameters())
    _loss(net, _batch()).backward()
    assert torch.equal(res[0], res[1])
    assert (res[0] - fg.flat).abs().max().item() <= 1e-6


def test_total_loss_sums_present_terms():
    out = {"photo_loss": torch.tensor([1.0, 3.0]), "smooth_loss": torch.tensor(0.5), "census_loss": None, "msd_loss": None}
    assert train.total_loss(out).item() == 2.5
