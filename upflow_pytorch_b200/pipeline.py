"""Host <-> device pipelining around ``UPFlow_net(input_dict)`` for throughput serving.

The reference's evaluation loop (dataset/kitti_dataset.py:394-450 with tools.data_prefetcher, utils/tools.py:166-212)
overlaps the host->device copy of the NEXT pair with the forward of the current one on a side stream.  This is the same
idea for the drop-in model, both directions:

    pipe = PipelinedInference(net)
    for im1_host, im2_host in pairs:             # pinned host tensors [B,3,H,W]
        done = pipe.submit(im1_host, im2_host)   # -> the PREVIOUS pair's flow (pinned host tensor) or None
    last = pipe.flush()

* two copy streams (an in-order stream would park the H2D of pair k+1 behind the D2H of flow k, which waits for forward
  k): H2D of pair k+1 while the compute stream runs pair k; D2H of pair k's flow into one of two pinned staging buffers
  while pair k+1 computes;
* compute stream: ``net({'im1','im2','if_loss': False})`` -- the public call, unchanged (one CUDA-graph replay);
* the caller receives a result one submit later (depth-2 pipeline) and may read it until the next-but-one submit.

Nothing here touches the arithmetic: plumbing only (streams, events, pinned buffers).
"""
import torch


class PipelinedInference:
    def __init__(self, net, device=None):
        self.net = net
        self.device = torch.device(device) if device is not None else next(net.parameters()).device
        self.h2d = torch.cuda.Stream(device=self.device)
        self.copy = torch.cuda.Stream(device=self.device)          # device -> host
        self.compute = torch.cuda.Stream(device=self.device)
        self._dev_in = [None, None]         # two device input slots (im1, im2)
        self._host_out = [None, None]       # two pinned result slots
        self._h2d = [torch.cuda.Event(), torch.cuda.Event()]
        self._done = [torch.cuda.Event(), torch.cuda.Event()]
        self._d2h = [torch.cuda.Event(), torch.cuda.Event()]
        self._free = [torch.cuda.Event(), torch.cuda.Event()]      # the compute stream has consumed input slot s
        self._k = 0
        self._pending = None                # slot whose result has not been handed out yet

    def _slot(self, s, im1, im2):
        if self._dev_in[s] is None or self._dev_in[s][0].shape != im1.shape:
            self._dev_in[s] = (torch.empty(im1.shape, dtype=torch.float32, device=self.device),
                               torch.empty(im2.shape, dtype=torch.float32, device=self.device))
            B, _, H, W = im1.shape
            self._host_out[s] = torch.empty(B, 2, H, W, dtype=torch.float32).pin_memory()
        return self._dev_in[s]

    def submit(self, im1_host, im2_host):
        """Enqueue one pair (pinned host tensors).  Returns the previous pair's forward flow [B,2,H,W] on the host
        (valid until the next-but-one submit), or None for the first call."""
        s = self._k & 1
        a, b = self._slot(s, im1_host, im2_host)
        with torch.cuda.stream(self.h2d):
            if self._k >= 2:
                self.h2d.wait_event(self._free[s])             # the forward that read this input slot has finished with it
            a.copy_(im1_host, non_blocking=True)
            b.copy_(im2_host, non_blocking=True)
            self._h2d[s].record(self.h2d)
        with torch.cuda.stream(self.compute), torch.no_grad():
            self.compute.wait_event(self._h2d[s])
            if self._k >= 2:
                self.compute.wait_event(self._d2h[s])          # result slot s has left for the host
            out = self.net({"im1": a, "im2": b, "if_loss": False})
            self._free[s].record(self.compute)
            flow = out["flow_f_out"]
            self._done[s].record(self.compute)
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(self._done[s])
            self._host_out[s].copy_(flow, non_blocking=True)
            flow.record_stream(self.copy)
            self._d2h[s].record(self.copy)
        prev, self._pending = self._pending, s
        self._k += 1
        if prev is None:
            return None
        self._d2h[prev].synchronize()                          # the caller reads pair k-1 while pair k runs
        return self._host_out[prev]

    def flush(self):
        """Wait for the last submitted pair and return its flow (or None)."""
        prev, self._pending = self._pending, None
        if prev is None:
            return None
        self._d2h[prev].synchronize()
        return self._host_out[prev]
