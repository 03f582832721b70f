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
    pkgs, loss = [], None
    for f in frames:
        pkg = render_fn(poses, f, pc, gs_grad=True, cam_grad=True)
        l = loss_fn(f, pkg)
        loss = l if loss is None else loss + l
        pkgs.append(pkg)
    if loss is not None:
        loss.backward()
    allreduce_gaussian_grads(pc.params, group=group)
    return loss, pkgs
