"""Densification bookkeeping and model surgery around the rasteriser, plus the checkpoint
tuples of the two model objects (8f N4) -- host-side mirrors of the reference so that loops with densification
(BASELINE.json configs[3]) can be driven without importing /root/reference:

  * ``add_densification_stats``      scene/gaussian_model.py:678-681   (also foldable into the fused backward:
                                     set ``pc.fold_densification_stats = True`` and ``fsgs_b200.render`` hands the
                                     accumulators to ``fsgs_render_backward_ex``, which updates them in
                                     ``k_preprocess_fused_bwd`` -- no [P,3] gradient read-back, no boolean gathers)
  * ``densify_and_prune``            scene/gaussian_model.py:656-676 = clone (:644-654) + split (:620-642) + prune,
                                     with the optimiser-state surgery of :523-603
  * ``capture`` / ``restore``        scene/gaussian_model.py:86-116, scene/pose_optimizer.py:472-487 -- the same
                                     tuples, field for field, so a checkpoint written by either side loads in the other

Plain PyTorch on the device (infrequent tensor indexing: every 300 iterations in train.py:305-316); nothing here is
on the per-frame path.  Pinned to the reference's own functions by tests/golden/ref_densify.npz
(oracle/make_golden_ref_densify.py imports /root/reference unmodified and records its results on seeded inputs).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
import torch.nn as nn

PARAM_NAMES = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation")


def add_densification_stats(variables: dict, viewspace_point_tensor, update_filter) -> None:
    """accum[f] += ||viewspace_points.grad[f]||_2 (all three components), denom[f] += 1, without the boolean-mask
    gather / scatter of the reference (and its host sync)."""
    g = viewspace_point_tensor.grad
    f = update_filter.to(g.dtype).unsqueeze(-1)
    variables['xyz_gradient_accum'] += torch.norm(g, dim=-1, keepdim=True) * f
    variables['denom'] += f


def quaternion_rotation(r: torch.Tensor) -> torch.Tensor:
    """``build_rotation`` (utils/general_utils.py:204-225): normalise (w,x,y,z), return [N,3,3]."""
    q = r / torch.sqrt((r * r).sum(dim=1, keepdim=True))
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=1)
    return R.view(-1, 3, 3)


def _swap_param(optimizer, group, new_value: torch.Tensor, state_fn: Callable[[torch.Tensor], torch.Tensor]):
    """Replace the single parameter of an optimiser group by ``new_value`` and carry its Adam moments over through
    ``state_fn`` (the reference's cat_tensors_to_optimizer / _prune_optimizer, gaussian_model.py:523-585)."""
    old = group["params"][0]
    state = optimizer.state.get(old, None)
    new = nn.Parameter(new_value.requires_grad_(True))
    if state is not None:
        state["exp_avg"] = state_fn(state["exp_avg"])
        state["exp_avg_sq"] = state_fn(state["exp_avg_sq"])
        del optimizer.state[old]
        optimizer.state[new] = state
    group["params"][0] = new
    return new


def _groups(optimizer):
    return [g for g in optimizer.param_groups if len(g["params"]) == 1]


def append_points(pc, new: Dict[str, torch.Tensor]) -> None:
    """``densification_postfix`` (:587-606): concatenate the new Gaussians to every parameter (zero Adam moments for
    them) and RESET the three densification statistics to zeros of the new size."""
    for group in _groups(pc.optimizer):
        ext = new[group["name"]]
        pc.params[group["name"]] = _swap_param(
            pc.optimizer, group, torch.cat((group["params"][0], ext), dim=0),
            lambda m, ext=ext: torch.cat((m, torch.zeros_like(ext)), dim=0))
    n, dev = pc.params["_xyz"].shape[0], pc.params["_xyz"].device
    pc.variables['xyz_gradient_accum'] = torch.zeros((n, 1), device=dev)
    pc.variables['denom'] = torch.zeros((n, 1), device=dev)
    pc.variables['max_radii2D'] = torch.zeros(n, device=dev)


def prune_points(pc, mask: torch.Tensor) -> None:
    """``prune_points`` (:545-560): drop the Gaussians where ``mask`` is True, moments and statistics with them."""
    keep = ~mask
    for group in _groups(pc.optimizer):
        pc.params[group["name"]] = _swap_param(pc.optimizer, group, group["params"][0][keep], lambda m: m[keep])
    for k in ('xyz_gradient_accum', 'denom', 'max_radii2D'):
        pc.variables[k] = pc.variables[k][keep]


def densify_and_clone(pc, grads: torch.Tensor, grad_threshold: float) -> None:
    """Small Gaussians with a large screen-space gradient are duplicated in place (:644-654)."""
    small = torch.max(pc.get_scaling, dim=1).values <= pc.variables['scene_radius'] * 0.01
    sel = (torch.norm(grads, dim=-1) >= grad_threshold) & small
    append_points(pc, {k: pc.params[k][sel] for k in PARAM_NAMES})


def densify_and_split(pc, grads: torch.Tensor, grad_threshold: float, N: int = 2) -> None:
    """Large Gaussians with a large gradient are replaced by N samples of themselves, 1.6 N times smaller (:620-642)."""
    n = pc.params["_xyz"].shape[0]
    dev = pc.params["_xyz"].device
    padded = torch.zeros(n, device=dev)
    padded[:grads.shape[0]] = grads.squeeze()
    large = torch.max(pc.get_scaling, dim=1).values > pc.variables['scene_radius'] * 0.01
    sel = (padded >= grad_threshold) & large
    stds = pc.get_scaling[sel].repeat(N, 1)
    samples = torch.normal(mean=torch.zeros((stds.size(0), 3), device=dev), std=stds)
    rots = quaternion_rotation(pc.params['_rotation'][sel]).repeat(N, 1, 1)
    new = {"_xyz": torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + pc.params["_xyz"][sel].repeat(N, 1),
           "_scaling": torch.log(pc.get_scaling[sel].repeat(N, 1) / (0.8 * N)),
           "_rotation": pc.params['_rotation'][sel].repeat(N, 1),
           "_features_dc": pc.params['_features_dc'][sel].repeat(N, 1, 1),
           "_features_rest": pc.params['_features_rest'][sel].repeat(N, 1, 1),
           "_opacity": pc.params['_opacity'][sel].repeat(N, 1)}
    append_points(pc, new)
    prune_points(pc, torch.cat((sel, torch.zeros(N * int(sel.sum()), device=dev, dtype=torch.bool))))


def densify_and_prune(pc, max_grad: float, min_opacity: float, max_screen_size: Optional[float], gs_mask=None) -> int:
    """``GaussianModel.densify_and_prune`` (:656-676).  ``pc`` needs ``params``, ``variables`` (incl. ``scene_radius``),
    ``optimizer`` (one group per parameter, named like it) and the activation properties.  Returns the number pruned.
    (As in the reference the screen-size criterion sees the statistics the clone / split steps have just reset.)"""
    accum = pc.variables['xyz_gradient_accum'].reshape(-1, 1)
    denom = pc.variables['denom'].reshape(-1, 1)
    grads = accum / denom                   # NaN where a Gaussian was never seen: every comparison below is False
    densify_and_clone(pc, grads, max_grad)
    densify_and_split(pc, grads, max_grad)
    prune = (pc.get_opacity < min_opacity).squeeze()
    if max_screen_size:
        big_vs = pc.variables['max_radii2D'] > max_screen_size
        big_ws = pc.get_scaling.max(dim=1).values > 0.1 * pc.variables['scene_radius']
        prune = prune | big_vs | big_ws
    prune_points(pc, prune)
    return int(prune.sum())


def reset_opacity(pc) -> None:
    """``reset_opacity`` (gaussian_model.py:452-456 + replace_tensor_to_optimizer :501-521): activated opacities
    capped at 0.01, the parameter's Adam moments zeroed."""
    capped = torch.min(pc.get_opacity, torch.ones_like(pc.get_opacity) * 0.01)
    new = torch.log(capped / (1 - capped))
    for group in _groups(pc.optimizer):
        if group["name"] == "_opacity":
            pc.params["_opacity"] = _swap_param(pc.optimizer, group, new.detach().clone(), torch.zeros_like)


# ---- optimiser + checkpoint tuples ---------------------------------------------------------------------------
def training_setup(pc, lr: Dict[str, float], eps: float = 1e-15) -> torch.optim.Optimizer:
    """One Adam group per parameter, named like it, in the reference's order (gaussian_model.py:382-409)."""
    groups = [{'params': [pc.params[k]], 'lr': float(lr[k]), 'name': k} for k in PARAM_NAMES]
    pc.optimizer = torch.optim.Adam(groups, lr=0.0, eps=eps)
    return pc.optimizer


def capture(pc):
    """``GaussianModel.capture`` (:86-100): the 12-tuple train.py saves as chkpnt*.pth (train.py:371-373)."""
    return (pc.active_sh_degree, pc.params['_xyz'], pc.params['_features_dc'], pc.params['_features_rest'],
            pc.params['_scaling'], pc.params['_rotation'], pc.params['_opacity'], pc.variables['max_radii2D'],
            pc.variables['xyz_gradient_accum'], pc.variables['denom'], pc.optimizer.state_dict(),
            getattr(pc, "spatial_lr_scale", 0))


def restore(pc, model_args, lr: Dict[str, float]) -> None:
    """``GaussianModel.restore`` (:102-116): unpack the tuple, rebuild the optimiser, load its state."""
    (pc.active_sh_degree, pc.params['_xyz'], pc.params['_features_dc'], pc.params['_features_rest'], pc.params['_scaling'],
     pc.params['_rotation'], pc.params['_opacity'], pc.variables['max_radii2D'], pc.variables['xyz_gradient_accum'],
     pc.variables['denom'], opt_dict, pc.spatial_lr_scale) = model_args
    training_setup(pc, lr)
    pc.optimizer.load_state_dict(opt_dict)


def capture_poses(poses, optimizer):
    """``PoseModel.capture`` (pose_optimizer.py:472-479): (optimizer state, r, t, pred_w2c, intrinsic)."""
    return (optimizer.state_dict(), poses.pose_param_net.r, poses.pose_param_net.t, poses.record_data['pred_w2c'],
            poses.record_data['intrinsic'])


def restore_poses(poses, model_args) -> None:
    """``PoseModel.restore`` (:481-487) -- like the reference it does not reload the optimiser state."""
    (_opt_dict, poses.pose_param_net.r, poses.pose_param_net.t, poses.record_data['pred_w2c'],
     poses.record_data['intrinsic']) = model_args
