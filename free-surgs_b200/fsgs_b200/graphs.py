"""CUDA-graph capture of a whole render step (forward + loss + backward).

The host side of one fused frame -- PyTorch dispatch, autograd, ctypes marshalling, ~25 kernel launches -- costs
about as much wall time as the GPU needs for the frame (~0.8 ms vs ~1.1 ms at 500k Gaussians).  Free-SurGS' loops
read a scalar back every iteration (`rgb_loss.item()` / `flow_loss.item()`, train.py:191-192), so after every
synchronisation the GPU idles until the host has issued the next frame's first kernels.  The tracking loop runs
50 iterations per frame with identical shapes: the classic case for a CUDA graph.

    step = GraphedStep(lambda: my_step(static_inputs))    # runs my_step eagerly a few times, then captures it
    out = step.replay()                                   # re-launches the captured kernels: no Python in between

``my_step`` must be a pure function of tensors that live at fixed addresses (update them in place between replays:
``static_target.copy_(new_target)``; parameters updated in place by the optimiser qualify) and must set the
gradients it produces to ``None`` at its start (so that the backward allocates them from the graph's memory pool,
the usual whole-network-capture recipe).  It may call ``fsgs_b200.render`` any number of times.

What cannot be captured is the read-back of the per-frame instance count that sizes the binning buffer; in
capture the library runs in fixed-capacity mode (``FSGS_FLAG_FIXED_CAPACITY``): the buffer is sized from the
instance count seen during the eager warm-up times ``headroom``.  If a later replay produces more instances (the
pose moved a lot, Gaussians were added) the binning / compositing kernels skip themselves; ``overflowed()``
detects that (one small device->host read) and ``recapture()`` rebuilds the graph with a larger capacity.
"""
from __future__ import annotations

import ctypes
from typing import Callable, List

import torch

from . import _lib
from . import frame_render
from .rasterizer import _POOL


class GraphedStep:
    def __init__(self, fn: Callable[[], object], warmup: int = 3, headroom: float = 1.25, device=None):
        self.fn, self.warmup, self.headroom = fn, int(warmup), float(headroom)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.graph, self.outputs, self._captured, self.capacity = None, None, [], 0
        self._frozen = []          # frame_render.FrozenModel buffers the captured forwards read
        self._capture()

    def _capture(self) -> None:
        dev = self.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        # the capacity must hold the LARGEST frame of the step (fn may render several frames; every fused forward
        # records its instance count in the per-thread maximum, reset here)
        frame_render._TLS.max_instances = 0
        with torch.cuda.stream(side):                      # warm-up off the default stream, as graph capture requires
            for _ in range(max(self.warmup, 1)):
                self.fn()
        most = int(getattr(frame_render._TLS, "max_instances", 0))
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # the warm-up ran on a stream nobody will use again: its scratch buffers (~300 MB at 500k Gaussians) sit in
        # the workspace pool under that stream's key -- hand them back to the allocator instead of stranding them
        _POOL.drop_stream(dev.index, side.cuda_stream)
        self.capacity = max(int(most * self.headroom) + 4096, int(self.capacity * self.headroom))
        _lib.check(_lib.lib().fsgs_set_instance_capacity(dev.index, self.capacity))
        del frame_render._CAPTURED[:]
        del frame_render._CAPTURED_FROZEN[:]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = self.fn()
        self._captured = list(frame_render._CAPTURED)
        self._frozen = list(dict.fromkeys(frame_render._CAPTURED_FROZEN))
        del frame_render._CAPTURED[:]
        del frame_render._CAPTURED_FROZEN[:]

    def replay(self):
        for fm in self._frozen:        # tracking against a frozen model: re-evaluate the pose-independent rows (in
            fm.refresh()               # place, eagerly) if the model was written to since -- a few version compares
        self.graph.replay()
        return self.outputs

    def release(self) -> None:
        """Destroy the captured graph (and drop its outputs / private memory pool).  Call this before
        ``torch.distributed.destroy_process_group()`` when the step contains a collective: NCCL keeps a
        communicator alive for as long as a captured graph references it, and tearing the process group down
        first waits on that reference forever."""
        if self.graph is not None:
            torch.cuda.synchronize(self.device)
            self.graph.reset()
        self.graph, self.outputs, self._captured = None, None, []

    def _counters(self):
        """(instances, longest tile list) of every captured fused forward in the LAST replay (synchronises)."""
        out = []
        for img, _cap, W, H, _bin_cap in self._captured:
            off = (ctypes.c_size_t * 6)()
            _lib.lib().fsgs_img_offsets(W, H, off)
            c = img[off[5]:off[5] + 24].view(torch.int64).tolist()
            out.append((int(c[0]), int(c[2])))
        return out

    def instance_counts(self) -> List[int]:
        """Instance count of every captured fused forward in the LAST replay (synchronises)."""
        return [n for n, _ in self._counters()]

    def overflowed(self) -> bool:
        """True if the last replay produced more instances than the captured capacity (its binning / compositing
        kernels then skipped themselves).  Also raises if the device watchdog fired (a captured forward cannot
        read that flag back itself)."""
        # more instances than the binning buffer holds, or a tile list longer than its bin (the counting pass drops
        # the keys into fixed-stride per-tile bins sized from the warm-up's longest list)
        over = any(n > cap or (bin_cap and longest > bin_cap)
                   for (n, longest), (_img, cap, _w, _h, bin_cap) in zip(self._counters(), self._captured))
        rc = _lib.lib().fsgs_watchdog_flag(self.device.index, 1)
        if rc < 0:
            _lib.check(rc)
        if rc == 1:
            _lib.check(-5)
        return over

    def recapture(self) -> None:
        """Rebuild the graph (after ``overflowed()``, or when tensor shapes changed)."""
        self.capacity = max(self.capacity, max(self.instance_counts(), default=0))
        self._capture()
