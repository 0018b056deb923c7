"""Data-parallel training step of UPFlow_net (BASELINE config 4).

The reference trains single-GPU (scripts/simple_train.py:118-146: Adam(amsgrad), lr 1e-4, weight decay 1e-4,
``loss = sum of the loss terms``, ``loss.backward()``, ``optimizer.step()``).  Here image pairs are the unit of
parallelism: one process per GPU holds a replica, takes a contiguous slice of the batch, runs forward + backward on
this library's kernels, and the ONLY collective is one all-reduce (sum) of the flat fp32 gradient buffer
(3,494,549 elements, 13.98 MB) over NCCL, divided by the world size.  Every ``param.grad`` is a view into that
buffer, so there is no gather/scatter copy around the collective.

Host logic only (torch.distributed plumbing); all arithmetic of the network lives in libupflow_b200.so.
"""
import torch
import torch.distributed as dist

LOSS_TERMS = ("photo_loss", "smooth_loss", "census_loss", "msd_loss")   # Loss_manager.compute_loss, simple_train.py:45-53


def shard_batch(batch, rank, world_size):
    """Contiguous slice [rank*B/G, (rank+1)*B/G) of every batched tensor of an input dict (B % G == 0)."""
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v) and v.dim() > 0:
            B = v.shape[0]
            if B % world_size:
                raise ValueError("batch size %d is not divisible by the world size %d" % (B, world_size))
            per = B // world_size
            out[k] = v[rank * per:(rank + 1) * per]
        else:
            out[k] = v
    return out


class FlatGradients:
    """One contiguous fp32 buffer holding every parameter's gradient; ``param.grad`` are views into it."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero(self):
        self.flat.zero_()
        o = 0
        for p in self.params:                      # re-attach (an optimizer's zero_grad(set_to_none=True) drops them)
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * o:
                p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def all_reduce_mean(self, group=None):
        """Sum over ranks, divide by the world size.  Returns the bytes moved per rank (algorithmic)."""
        if dist.is_available() and dist.is_initialized():
            ws = dist.get_world_size(group)
            if ws > 1:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(ws)
        return self.numel * 4


def total_loss(output_dict):
    """Sum of the loss terms present (Loss_manager.compute_loss, scripts/simple_train.py:45-53)."""
    loss = 0
    for name in LOSS_TERMS:
        v = output_dict.get(name)
        if v is not None and not (isinstance(v, (int, float)) and v == 0):
            loss = loss + v.mean()
    return loss


class Trainer:
    """scripts/simple_train.py:118-146, data-parallel.

    use_cuda_graph=True captures zero-grad + forward + losses + backward of one batch SHAPE in a CUDA graph (the
    eager step issues ~3300 launches from Python, 30 % of its wall time); the gradient all-reduce and the Adam update
    stay outside the graph.  The batch tensors are copied into the graph's static inputs on every step."""

    def __init__(self, net, lr=1e-4, weight_decay=1e-4, group=None, use_cuda_graph=False):
        self.net = net
        self.group = group
        self.grads = FlatGradients(net.parameters())
        self.optimizer = torch.optim.Adam(self.grads.params, lr=lr, amsgrad=True, weight_decay=weight_decay)
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}
        self.graph_launches = 0

    def _forward_backward(self, batch):
        self.grads.zero()
        out = self.net(batch)
        loss = total_loss(out)
        loss.backward()
        return loss.detach()

    def _graphed(self, batch):
        key = tuple((k, tuple(v.shape)) for k, v in sorted(batch.items()) if torch.is_tensor(v))
        g = self._graphs.get(key)
        if g is None:
            static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                      # warm-up off the capture stream (torch.cuda.graphs recipe)
                for _ in range(2):
                    self._forward_backward(static)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            from . import _ext
            graph = torch.cuda.CUDAGraph()
            n0 = _ext.launch_count()
            with torch.cuda.graph(graph):
                loss = self._forward_backward(static)
            self.graph_launches = _ext.launch_count() - n0     # library kernels inside one replay
            g = self._graphs[key] = (graph, static, loss)
        graph, static, loss = g
        for k, v in batch.items():
            if torch.is_tensor(v):
                static[k].copy_(v, non_blocking=True)
        graph.replay()
        return loss

    def train_step(self, batch):
        """batch: the LOCAL shard (dict with im1, im2 [, im1_raw, im2_raw, start]).  Returns the local loss tensor."""
        self.net.train()
        batch = dict(batch)
        batch["if_loss"] = True
        loss = self._graphed(batch) if self.use_cuda_graph else self._forward_backward(batch)
        self.grads.all_reduce_mean(self.group)
        self.optimizer.step()
        return loss.clone() if self.use_cuda_graph else loss
