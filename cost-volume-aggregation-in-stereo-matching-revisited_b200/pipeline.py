"""Host-side streaming of stereo pairs through the hot path: H2D copies of pair i+1 overlap the kernels of pair i.

The reference's driver (`my_img.py:96-102`) does `img.cuda()` -> `model(...)` -> `.cpu()` serially per pair; on a
B200 the 87 MB of fp32 feature maps per KITTI pair cost ~1.7 ms of PCIe time, 40 % of the 4.4 ms the kernels need.
`HotPathPipeline` keeps `depth` device-side input slots and two CUDA streams (copy-in, compute + copy-out) tied
together with events, so that in steady state the copies are free.  torch is used for the pinned/device buffers,
streams and events only.
"""
import torch


class HotPathPipeline:
    def __init__(self, net, depth=2):
        self.net = net
        self.depth = depth
        self.dev = next(net.parameters()).device
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.compute_stream = torch.cuda.Stream(device=self.dev)
        self.slots = [None] * depth            # device input tensors per slot
        self.out_host = [None] * depth         # pinned output buffers per slot
        self.h2d_done = [torch.cuda.Event() for _ in range(depth)]
        self.slot_free = [torch.cuda.Event() for _ in range(depth)]
        self.out_done = [torch.cuda.Event() for _ in range(depth)]
        self._used = [False] * depth

    def _ensure(self, slot, host_feats):
        if self.slots[slot] is None or any(a.shape != b.shape for a, b in zip(self.slots[slot], host_feats)):
            self.slots[slot] = [torch.empty(t.shape, dtype=torch.float32, device=self.dev) for t in host_feats]
            self.out_host[slot] = None

    def submit(self, i, host_feats):
        """Queue pair i (5 pinned fp32 host tensors: gwc_l, gwc_r, cat_l, cat_r, g).  Returns the slot index; the
        results land in pinned host buffers, valid after `wait(slot)`."""
        slot = i % self.depth
        self._ensure(slot, host_feats)
        with torch.cuda.stream(self.copy_stream):
            if self._used[slot]:
                self.copy_stream.wait_event(self.slot_free[slot])      # kernels of the previous user are done
            for d, h in zip(self.slots[slot], host_feats):
                d.copy_(h, non_blocking=True)
            self.h2d_done[slot].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.h2d_done[slot])
            with torch.no_grad():
                pred4, pv = self.net.hot_path(*self.slots[slot])
            self.slot_free[slot].record(self.compute_stream)
            if self.out_host[slot] is None:
                self.out_host[slot] = (torch.empty(pred4.shape, dtype=torch.float32).pin_memory(),
                                       torch.empty(pv.shape, dtype=torch.float32).pin_memory())
            self.out_host[slot][0].copy_(pred4, non_blocking=True)
            self.out_host[slot][1].copy_(pv, non_blocking=True)
            self.out_done[slot].record(self.compute_stream)
        self._used[slot] = True
        return slot

    def wait(self, slot):
        self.out_done[slot].synchronize()
        return self.out_host[slot]

    def run(self, host_sets):
        """Stream a list of host feature sets; returns the (pred4, prob_volume2) of the LAST pair (pinned buffers)."""
        slot = 0
        for i, hs in enumerate(host_sets):
            slot = self.submit(i, hs)
        return self.wait(slot)
