"""Frame data-parallelism for the hot path.

The reference is single-process / single-GPU; what it already does is sum the losses of two
views before one backward (train.py:236-259).  The multi-GPU form of that semantic: the Gaussian
model is replicated, rank g renders frame(s) g of the batch with its own pose column, and the
Gaussian gradients are SUM-all-reduced (one process per GPU, NCCL over NVLink; ``gloo`` on CPU
for the host-logic tests).  Pose gradients are local to the rank that owns the frame -- no
exchange.  The densification statistics the reference accumulates per iteration
(gaussian_model.py:678-681, gaussian_renderer/__init__.py:77-80) are reduced with the matching
operator so that ``densify_and_prune`` stays bit-identical on every rank.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist

PARAM_KEYS = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation")


def shard_frames(frames: Sequence[int], world_size: int, rank: int) -> List[int]:
    """Round-robin assignment of a frame batch to ranks (rank g gets frames g, g+W, ...)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return [f for i, f in enumerate(frames) if i % world_size == rank]


def _is_dist(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def _shared_flat(tensors):
    """If the gradient tensors tile one contiguous float32 storage exactly (as the fused backward
    allocates them), return a flat alias of that storage, else None."""
    try:
        if any(t.dtype != torch.float32 or not t.is_contiguous() for t in tensors):
            return None
        st = tensors[0].untyped_storage()
        if any(t.untyped_storage().data_ptr() != st.data_ptr() for t in tensors):
            return None
        spans = sorted((t.storage_offset(), t.storage_offset() + t.numel()) for t in tensors)
        if spans[0][0] != 0 or any(a[1] != b[0] for a, b in zip(spans, spans[1:])):
            return None
        total = spans[-1][1]
        if total * 4 != st.nbytes():
            return None
        return torch.empty(0, dtype=torch.float32, device=tensors[0].device).set_(st, 0, (total,))
    except Exception:  # noqa: BLE001 -- any exotic tensor subclass: just take the generic path
        return None


def allreduce_gaussian_grads(params: dict, group=None, bucket: bool = True) -> int:
    """Sum the gradients of the six Gaussian parameter tensors over the ranks, in place.
    A rank that rendered no frame (or whose tensor got no gradient) contributes zeros.
    ``bucket=True`` packs the 59 floats/Gaussian into one flat buffer -> a single collective
    (236 B/Gaussian: 118 MB at 500k) instead of six.  Returns the number of bytes reduced."""
    tensors = []
    for k in PARAM_KEYS:
        p = params[k]
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        tensors.append(p.grad)
    nbytes = sum(t.numel() * t.element_size() for t in tensors)
    if not _is_dist(group):
        return nbytes
    flat = _shared_flat(tensors)
    if flat is not None:          # the fused backward hands out views of one buffer: one collective, no copies
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return nbytes
    if bucket:
        flat = torch.cat([t.reshape(-1) for t in tensors])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for t in tensors:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()
    else:
        handles = [dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True) for t in tensors]
        for h in handles:
            h.wait()
    return nbytes


class _NvlinkExchange:
    """The rows of the frame-parallel exchange in a symmetric buffer + the hand-written two-shot all-reduce over
    NVLink / NVSwitch (``fsgs_exchange_rows``: multimem.ld_reduce / multimem.st through the switch when the fabric
    offers a multicast object, peer loads / stores otherwise) between two stream-ordered cross-GPU barriers.
    torch's symmetric-memory module only supplies the plumbing: the allocation mapped into every rank, the multicast
    address, the barrier kernels."""

    def __init__(self, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem = symm_mem
        self.group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.buf, self.hdl, self.peers = None, None, None
        self.multicast = 0
        self.cap = 0                 # floats per region: the buffer holds [rows | slice sums of the pull-gather form]
        self.sum_peers = None

    def alloc(self, n_floats: int, device) -> torch.Tensor:
        """Collective on first use / growth (every rank reaches its first fused backward at the same point)."""
        import ctypes
        need = (int(n_floats) + 3) // 4 * 4
        if self.buf is None or self.cap < need or self.buf.device != device:
            cap = max(need, int(self.cap * 1.5) // 4 * 4)
            self.cap = cap
            self.buf = self.symm_mem.empty(2 * cap, dtype=torch.float32, device=device)
            self.buf.zero_()
            self.hdl = self.symm_mem.rendezvous(self.buf, self.group)
            self.multicast = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
            # With multimem a GPU's links carry ~(1 + 1/N) x the payload in each direction whatever N is, with peer
            # loads / stores 2 (N-1)/N x.  Measured on B200 / NVSwitch, 28 MB of rows (tools/symm_probe.py):
            #   N = 2: peer 61 us, multimem 98 us;  N = 4: 87 / 100 us;  N = 8: 102 / 99 us  (ncclAllReduce: 72 / 110 / 138 us)
            # -> in-switch reduction from 8 ranks on.  FSGS_EXCHANGE_MULTICAST=0/1 forces one of them (A/B).
            import os
            force = os.environ.get("FSGS_EXCHANGE_MULTICAST")
            if force is not None:
                self.multicast = self.multicast if force == "1" else 0
            elif self.world < 8:
                self.multicast = 0
            ptrs = [int(x) for x in self.hdl.buffer_ptrs]
            self.peers = (ctypes.c_void_p * self.world)(*ptrs)
            self.sum_peers = (ctypes.c_void_p * self.world)(*[p + 4 * cap for p in ptrs])
        return self.buf[:n_floats]

    def reduce(self, flat: torch.Tensor) -> None:
        """SUM ``flat`` (a slice of the symmetric buffer) over the ranks, in place, ordered on the current stream."""
        import ctypes
        from . import _lib
        off = (flat.data_ptr() - self.buf.data_ptr()) // 4
        assert off % 4 == 0 and flat.device == self.buf.device
        n4 = (flat.numel() + 3) // 4                    # (the buffer is padded to whole float4 words, zero-filled)
        stream = torch.cuda.current_stream(flat.device).cuda_stream
        self.hdl.barrier(channel=0)                     # every rank's rows are written
        _lib.check(_lib.lib().fsgs_exchange_rows(ctypes.c_void_p(self.multicast) if self.multicast else None, self.peers,
                                                 self.world, self.rank, off // 4, n4, ctypes.c_void_p(stream)))
        self.hdl.barrier(channel=1)                     # every slice is summed and written back everywhere

    def expand_pull_gather(self, st, P, first, count, xyz, cam_center, rows, grads, stream) -> None:
        """Two-shot with its second half riding on the expansion (4+ ranks): reduce-scatter of the rows into the sum
        region of every slice's owner (``fsgs_exchange_rows_scatter``), then ``fsgs_compact_grad_expand_peers`` fetches
        every word from its owner while it writes the SH gradients.  Two barriers, as the plain two-shot form."""
        import ctypes
        from . import _lib
        assert rows.data_ptr() == self.buf.data_ptr() and first == 0 and count == P, "one Gaussian range only"
        p = lambda x: ctypes.c_void_p(x.data_ptr())
        n4 = (P * 14 + 3) // 4
        self.hdl.barrier(channel=0)                     # every rank's rows are written
        _lib.check(_lib.lib().fsgs_exchange_rows_scatter(
            ctypes.c_void_p(self.multicast) if self.multicast else None, self.peers, self.world, self.rank, 0, n4,
            self.cap // 4, ctypes.c_void_p(stream)))
        self.hdl.barrier(channel=1)                     # every slice's sums sit in its owner's sum region
        _lib.check(_lib.lib().fsgs_compact_grad_expand_peers(
            ctypes.byref(st), P, 0, P, p(xyz), p(cam_center), self.sum_peers, self.world, 1, 0, n4, p(grads["xyz"]),
            p(grads["f_dc"]), p(grads["f_rest"]), p(grads["opacity"]), p(grads["scaling"]), p(grads["rotation"]),
            ctypes.c_void_p(stream)))

    def expand(self, st, P, first, count, xyz, cam_center, rows, grads, stream) -> None:
        """One-shot exchange (small rank counts): ``fsgs_compact_grad_expand_peers`` pulls rows [first, first+count)
        from every rank's buffer, adds them in rank order and expands them into ``grads`` -- collective and consumer
        in one kernel.  Between two cross-GPU barriers: all rows written / all ranks done reading."""
        import ctypes
        from . import _lib
        assert rows.data_ptr() == self.buf.data_ptr(), "the rows must start the symmetric buffer"
        p = lambda x: ctypes.c_void_p(x.data_ptr())
        self.hdl.barrier(channel=0)
        _lib.check(_lib.lib().fsgs_compact_grad_expand_peers(
            ctypes.byref(st), P, first, count, p(xyz), p(cam_center), self.peers, self.world, 0, 0, 0, p(grads["xyz"]),
            p(grads["f_dc"]), p(grads["f_rest"]), p(grads["opacity"]), p(grads["scaling"]), p(grads["rotation"]),
            ctypes.c_void_p(stream)))
        self.hdl.barrier(channel=1)


def enable_frame_parallel(group=None, check_cam_center: torch.Tensor = None, chunks: int = 1,
                          exchange: str = "nccl") -> None:
    """Fold the gradient exchange into the fused render's backward (the fast path; replaces a later call to
    ``allreduce_gaussian_grads``).  In Free-SurGS every SH-coefficient gradient of a Gaussian is
    ``basis_k(dir) * gc`` with ``gc`` the clamp-masked colour gradient, and neither ``dir = normalize(xyz -
    cam_center)`` nor the mask depends on the frame (the SH view origin is frozen at the first camera,
    gaussian_model.py:317 / pose_optimizer.py:603).  The ranks therefore SUM-reduce 14 floats per Gaussian
    (xyz, opacity, scaling, rotation, gc: 56 B) instead of 59 (236 B) and expand the SH gradients locally from the
    reduced ``gc`` (``fsgs_sh_grad_expand``).  After ``loss.backward()`` every rank holds the summed gradients of all
    frames; pose gradients stay local.  ``check_cam_center``: if given, asserts that all ranks use the same SH view
    origin (the precondition).  ``exchange``: "nccl" = ``ncclAllReduce`` of the rows; "nvlink" = the library's own
    two-shot all-reduce over NVLink / NVSwitch on a symmetric buffer (``fsgs_exchange_rows``; multimem through the
    switch where available).  ``chunks``: the per-Gaussian backward kernel runs in that many Gaussian ranges and range k
    is exchanged + expanded on a side stream while range k+1 is computed (measured on 2 x B200 with NCCL: 4 ranges
    1.25 ms/step, 1 range 1.15 -- four small collectives cost more latency than the overlap hides; default 1)."""
    from . import frame_render
    if not _is_dist(group):
        frame_render.set_grad_reducer(None)
        return
    if check_cam_center is not None:
        c = check_cam_center.detach().float().reshape(-1).clone()
        ref = c.clone()
        dist.broadcast(ref, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if not torch.equal(c, ref):
            raise ValueError("frame-parallel SH-gradient exchange needs the same cam_center (SH view origin) on every rank")
    if exchange == "nvlink":
        import os
        xch = _NvlinkExchange(group)
        # two ranks: the one-shot form (rank sum folded into the expansion kernel; same link traffic as the two-shot
        # kernel at N = 2, one kernel less).  FSGS_EXCHANGE_ONE_SHOT=0/1 forces either (A/B; 1 is usable up to 8 ranks).
        # four ranks and more: the two-shot kernel.  (Its variant with the all-gather half riding on the expansion kernel,
        # "pull_gather", was measured at N = 4: exchange 0.045 + expansion 0.057 ms against 0.073 + 0.027 -- no gain, step
        # 1.101 vs 1.083 ms; kept as an option.)  FSGS_EXCHANGE_FORM=two_shot|one_shot|pull_gather forces a form (A/B,
        # tests); FSGS_EXCHANGE_ONE_SHOT=0/1 is the older spelling of the first two.
        form = os.environ.get("FSGS_EXCHANGE_FORM")
        if form is None:
            force = os.environ.get("FSGS_EXCHANGE_ONE_SHOT")
            form = ("one_shot" if force == "1" else "two_shot") if force is not None else \
                   ("one_shot" if xch.world == 2 else "two_shot")
        if form not in ("two_shot", "one_shot", "pull_gather"):
            raise ValueError(f"FSGS_EXCHANGE_FORM must be two_shot, one_shot or pull_gather, got {form!r}")
        fused = {"two_shot": None, "one_shot": xch.expand, "pull_gather": xch.expand_pull_gather}[form]
        frame_render.set_grad_reducer(xch.reduce, chunks=chunks if fused is None else 1, alloc=xch.alloc, expand=fused)
        _STATE["exchange"] = xch
        return
    if exchange != "nccl":
        raise ValueError(f"exchange must be 'nccl' or 'nvlink', got {exchange!r}")
    frame_render.set_grad_reducer(lambda flat: dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group), chunks=chunks)


_STATE = {"exchange": None}


def disable_frame_parallel() -> None:
    from . import frame_render
    frame_render.set_grad_reducer(None)
    _STATE["exchange"] = None


COMPACT_LAYOUT = (("_rotation", 4), ("_xyz", 3), ("_scaling", 3), ("_opacity", 1), ("gc", 3))   # 14 floats / Gaussian


def expand_sh_grads_reference(gc: torch.Tensor, xyz: torch.Tensor, cam_center: torch.Tensor, sh_degree: int, eval_basis):
    """Plain-torch statement of what ``fsgs_sh_grad_expand`` computes (used by the CPU tests of the protocol):
    dL/dfeatures[P,16,3] = basis(dir)[P,16,1] * gc[P,1,3], zero above the active degree.
    ``eval_basis(deg, dirs) -> [P,16]`` supplies the real SH basis."""
    d = xyz - cam_center.reshape(1, 3)
    d = d / d.norm(dim=1, keepdim=True)
    B = eval_basis(sh_degree, d)
    nb = (sh_degree + 1) ** 2
    B = torch.cat([B[:, :nb], torch.zeros(B.shape[0], 16 - nb, dtype=B.dtype)], dim=1) if B.shape[1] >= nb else B
    full = B[:, :, None] * gc[:, None, :]
    return full[:, :1, :], full[:, 1:, :]


def allreduce_densification_stats(variables: dict, group=None) -> None:
    """``xyz_gradient_accum`` and ``denom`` are sums over the frames seen, ``max_radii2D`` a max."""
    if not _is_dist(group):
        return
    dist.all_reduce(variables["xyz_gradient_accum"], op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(variables["denom"], op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(variables["max_radii2D"], op=dist.ReduceOp.MAX, group=group)


def dp_render_step(render_fn, poses, pc, frames: Iterable[int], loss_fn, group=None):
    """One data-parallel step: render this rank's frames, sum their losses, one backward, then the
    gradient all-reduce.  ``render_fn(poses, frame, pc, gs_grad, cam_grad)`` is ``fsgs_b200.render``;
    ``loss_fn(frame, render_pkg) -> scalar``.  Returns (local loss or None, list of render packages)."""
    from . import frame_render
    folded = frame_render._GRAD_REDUCER["fn"] is not None    # enable_frame_parallel: the exchange happens in backward
    frames = list(frames)
    if folded and _is_dist(group):
        n = torch.tensor([len(frames)], dtype=torch.int64)
        lo, hi = n.clone(), n.clone()
        if dist.get_backend(group) == "nccl":
            lo, hi = lo.cuda(), hi.cuda()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        if int(lo) != int(hi):
            raise ValueError("frame-parallel mode runs one collective per rendered frame: every rank needs the same "
                             f"number of frames per step (got between {int(lo)} and {int(hi)})")
    pkgs, loss = [], None
    for f in frames:
        pkg = render_fn(poses, f, pc, gs_grad=True, cam_grad=True)
        l = loss_fn(f, pkg)
        loss = l if loss is None else loss + l
        pkgs.append(pkg)
    if loss is not None:
        loss.backward()
    if not folded:
        allreduce_gaussian_grads(pc.params, group=group)
    return loss, pkgs
