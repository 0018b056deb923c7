"""Host <-> device pipelining around ``UPFlow_net(input_dict)`` for throughput serving.

The reference's evaluation loop (dataset/kitti_dataset.py:394-450 with tools.data_prefetcher, utils/tools.py:166-212)
overlaps the host->device copy of the NEXT pair with the forward of the current one on a side stream.  This is the same
idea for the drop-in model, both directions:

    pipe = PipelinedInference(net)               # or PipelinedInference(net, lanes=2)
    for im1_host, im2_host in pairs:             # pinned host tensors [B,3,H,W]
        done = pipe.submit(im1_host, im2_host)   # -> the flow of the pair submitted `lanes` calls ago, or None
    rest = pipe.drain()                          # the flows not handed out yet, in order (flush(): the last one)

* two copy streams (an in-order stream would park the H2D of pair k+1 behind the D2H of flow k, which waits for forward
  k): H2D of pair k+1 while the compute stream runs pair k; D2H of pair k's flow into a pinned staging buffer while pair
  k+1 computes;
* compute: ``net({'im1','im2','if_loss': False})`` -- the public call, unchanged (one CUDA-graph replay);
* ``lanes`` compute streams: pair k runs on lane k % lanes with that lane's own workspaces and captured graph
  (``UPFlow_net.set_lane``), so with lanes = 2 the graphs of two consecutive pairs replay CONCURRENTLY: the three coarse
  pyramid levels of a pair are a dependent chain of ~60 small launches that occupies a fraction of the 148 SMs, and the
  other pair's kernels run next to them.  Every pair is still one complete forward with its own inputs and result; only
  the order in which the GPU interleaves two pairs' kernels changes, not one bit of either result;
* the caller receives a result ``lanes`` submits later and may read it until the next-but-one submit.

Nothing here touches the arithmetic: plumbing only (streams, events, pinned buffers).
"""
import torch


class PipelinedInference:
    def __init__(self, net, device=None, lanes=1):
        assert lanes >= 1
        self.net = net
        self.lanes = int(lanes)
        self.depth = self.lanes + 2          # slots: `lanes` pairs in flight, one result with the caller, one spare
        self.device = torch.device(device) if device is not None else next(net.parameters()).device
        self.h2d = torch.cuda.Stream(device=self.device)
        self.copy = torch.cuda.Stream(device=self.device)          # device -> host
        self.computes = [torch.cuda.Stream(device=self.device) for _ in range(self.lanes)]
        self.compute = self.computes[0]
        ev = lambda: [torch.cuda.Event() for _ in range(self.depth)]
        self._dev_in = [None] * self.depth      # device input slots (im1, im2)
        self._host_out = [None] * self.depth    # pinned result slots
        self._h2d, self._done, self._d2h = ev(), ev(), ev()
        self._free = ev()                       # the compute stream has consumed input slot s
        self._k = 0
        self._pending = []                      # slots whose results have not been handed out yet, oldest first

    def _slot(self, s, im1, im2):
        if self._dev_in[s] is None or self._dev_in[s][0].shape != im1.shape:
            self._dev_in[s] = (torch.empty(im1.shape, dtype=torch.float32, device=self.device),
                               torch.empty(im2.shape, dtype=torch.float32, device=self.device))
            B, _, H, W = im1.shape
            self._host_out[s] = torch.empty(B, 2, H, W, dtype=torch.float32).pin_memory()
        return self._dev_in[s]

    def submit(self, im1_host, im2_host):
        """Enqueue one pair (pinned host tensors).  Returns the forward flow [B,2,H,W] of the pair submitted ``lanes``
        calls ago, on the host (valid until the next-but-one submit), or None for the first ``lanes`` calls."""
        k = self._k
        s = k % self.depth
        lane = k % self.lanes
        compute = self.computes[lane]
        a, b = self._slot(s, im1_host, im2_host)
        with torch.cuda.stream(self.h2d):
            if k >= self.depth:
                self.h2d.wait_event(self._free[s])             # the forward that read this input slot has finished with it
            a.copy_(im1_host, non_blocking=True)
            b.copy_(im2_host, non_blocking=True)
            self._h2d[s].record(self.h2d)
        with torch.cuda.stream(compute), torch.no_grad():
            compute.wait_event(self._h2d[s])
            if k >= self.depth:
                compute.wait_event(self._d2h[s])               # result slot s has left for the host
            if self.lanes > 1:
                self.net.set_lane(lane)
            out = self.net({"im1": a, "im2": b, "if_loss": False})
            self._free[s].record(compute)
            flow = out["flow_f_out"]
            self._done[s].record(compute)
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(self._done[s])
            self._host_out[s].copy_(flow, non_blocking=True)
            flow.record_stream(self.copy)
            self._d2h[s].record(self.copy)
        self._pending.append(s)
        self._k += 1
        if len(self._pending) <= self.lanes:
            return None
        prev = self._pending.pop(0)
        self._d2h[prev].synchronize()                          # the caller reads pair k-lanes while `lanes` pairs run
        return self._host_out[prev]

    def drain(self):
        """Wait for every submitted pair; returns the flows not handed out yet, oldest first."""
        out = []
        while self._pending:
            prev = self._pending.pop(0)
            self._d2h[prev].synchronize()
            out.append(self._host_out[prev])
        if self.lanes > 1:
            self.net.set_lane(0)
        return out

    def flush(self):
        """Wait for every submitted pair and return the LAST flow (or None).  With lanes = 1 at most one flow is
        pending, so nothing is lost; with more lanes use drain() to receive all of them."""
        rest = self.drain()
        return rest[-1] if rest else None
