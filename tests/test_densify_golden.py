"""fsgs_b200/densify.py (densification statistics, densify_and_prune with its optimiser surgery, reset_opacity, the
checkpoint tuple) against vectors recorded from the REFERENCE's own GaussianModel on the same seeded inputs
(tests/golden/ref_densify.npz + ref_chkpnt_gaussians.pth, written by oracle/make_golden_ref_densify.py in the build
container).  CPU, bit-exact: both sides are the same sequence of PyTorch operations."""
import os
import types

import numpy as np
import torch

from fsgs_b200 import densify, model
from oracle import make_golden_ref_densify as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _lrs():
    o = G.OPT
    return {"_xyz": o["position_lr_init"] * 5.0, "_features_dc": o["feature_lr"], "_features_rest": o["feature_lr"] / 20.0,
            "_opacity": o["opacity_lr"], "_scaling": o["scaling_lr"], "_rotation": o["rotation_lr"]}


def _model_after_one_step():
    inp, grads, accum, denom, max_radii, vs_grad, vis = G.seeded_inputs()
    pc = model.SplatModel({k: v.clone() for k, v in inp.items()}, active_sh_degree=2)
    densify.training_setup(pc, _lrs())
    for k, p in pc.params.items():
        p.grad = grads[k].clone()
    pc.optimizer.step()
    pc.optimizer.zero_grad(set_to_none=True)
    return pc, (accum, denom, max_radii, vs_grad, vis)


def test_add_densification_stats_matches_the_reference():
    gold = np.load(os.path.join(GOLD, "ref_densify.npz"))
    pc, (_, _, _, vs_grad, vis) = _model_after_one_step()
    vs = types.SimpleNamespace(grad=vs_grad)
    densify.add_densification_stats(pc.variables, vs, vis)
    densify.add_densification_stats(pc.variables, vs, ~vis | (vs_grad[:, 0] > 0))
    assert np.array_equal(pc.variables['xyz_gradient_accum'].numpy(), gold["stats_accum"])
    assert np.array_equal(pc.variables['denom'].numpy(), gold["stats_denom"])


def test_densify_and_prune_and_reset_opacity_match_the_reference():
    gold = np.load(os.path.join(GOLD, "ref_densify.npz"))
    pc, (accum, denom, max_radii, _, _) = _model_after_one_step()
    for k, p in pc.params.items():
        assert np.array_equal(p.detach().numpy(), gold["pre" + k]), k          # same Adam step as the reference's
    pc.variables.update(xyz_gradient_accum=accum.clone(), denom=denom.clone(), max_radii2D=max_radii.clone(),
                        scene_radius=torch.tensor(G.SCENE_RADIUS))
    torch.manual_seed(G.SEED_SPLIT)
    densify.densify_and_prune(pc, G.MAX_GRAD, G.MIN_OPACITY, G.MAX_SCREEN)
    assert pc.params["_xyz"].shape[0] == gold["post_xyz"].shape[0] != 400
    for k, p in pc.params.items():
        assert np.array_equal(p.detach().numpy(), gold["post" + k]), k
        st = pc.optimizer.state[p]
        assert np.array_equal(st["exp_avg"].numpy(), gold["post_exp_avg" + k]), k
        assert np.array_equal(st["exp_avg_sq"].numpy(), gold["post_exp_avg_sq" + k]), k
    for k in ("xyz_gradient_accum", "denom", "max_radii2D"):
        assert np.array_equal(pc.variables[k].numpy(), gold["post_var_" + k]), k
    densify.reset_opacity(pc)
    assert np.array_equal(pc.params["_opacity"].detach().numpy(), gold["reset_opacity"])
    assert np.array_equal(pc.optimizer.state[pc.params["_opacity"]]["exp_avg"].numpy(), gold["reset_opacity_exp_avg"])
    # the optimiser still steps on the surgically replaced parameters
    for p in pc.params.values():
        p.grad = torch.ones_like(p)
    pc.optimizer.step()


def test_checkpoint_written_by_the_reference_loads_and_round_trips(tmp_path):
    """train.py:371-373 saves ``(gaussians.capture(), iteration)``; train.py:106-113 restores it.  The reference's own
    checkpoint must restore into the mirror, and the mirror's capture must be the same tuple field for field."""
    gold = np.load(os.path.join(GOLD, "ref_densify.npz"))
    (ref_tuple, it) = torch.load(os.path.join(GOLD, "ref_chkpnt_gaussians.pth"), weights_only=False)
    assert it == 1234 and len(ref_tuple) == 12
    pc = model.SplatModel({k: torch.zeros(1, *s) for k, s in (("_xyz", (3,)), ("_features_dc", (1, 3)), ("_features_rest", (15, 3)),
                                                             ("_opacity", (1,)), ("_scaling", (3,)), ("_rotation", (4,)))})
    densify.restore(pc, ref_tuple, _lrs())
    assert pc.active_sh_degree == 2 and pc.spatial_lr_scale == 5.0
    for k, p in pc.params.items():
        assert np.array_equal(p.detach().numpy(), gold["pre" + k]), k
    ours, (accum, denom, max_radii, _, _) = _model_after_one_step()
    ours.spatial_lr_scale = 5.0
    ours.variables.update(xyz_gradient_accum=accum.clone(), denom=denom.clone(), max_radii2D=max_radii.clone())
    path = tmp_path / "chkpnt1234.pth"
    torch.save((densify.capture(ours), 1234), path)
    (mine, _) = torch.load(path, weights_only=False)
    for a, b in zip(mine, ref_tuple):
        if torch.is_tensor(b):
            assert torch.equal(a.detach(), b.detach())
        elif isinstance(b, dict):            # optimizer.state_dict(): same groups, same moments
            assert [g["name"] for g in a["param_groups"]] == [g["name"] for g in b["param_groups"]]
            assert [g["lr"] for g in a["param_groups"]] == [g["lr"] for g in b["param_groups"]]
            for sa, sb in zip(a["state"].values(), b["state"].values()):
                assert torch.equal(sa["exp_avg"], sb["exp_avg"]) and torch.equal(sa["exp_avg_sq"], sb["exp_avg_sq"])
        else:
            assert a == b
    # and the restored model keeps training: the optimiser state came along
    for p in pc.params.values():
        p.grad = torch.ones_like(p)
    pc.optimizer.step()


def test_pose_checkpoint_tuple_round_trips(tmp_path):
    """PoseModel.capture / restore (pose_optimizer.py:472-487): (optimizer state, r, t, pred_w2c, intrinsic)."""
    K = [[258.75, 0, 160.0], [0, 258.75, 128.0], [0, 0, 1]]
    poses = model.FramePoses(3, K, 320, 256, device="cpu")
    poses.set_pose(1, (1.0, 0.01, -0.02, 0.03), (0.1, 0.2, 0.3))
    poses.record_data['pred_w2c'][1] = poses.get_pose(1).detach().numpy()
    opt = torch.optim.Adam([{'params': poses.pose_param_net.r, 'lr': 0.01}, {'params': poses.pose_param_net.t, 'lr': 0.01}],
                           lr=0.001, eps=1e-15)
    path = tmp_path / "poses10.pth"
    torch.save((densify.capture_poses(poses, opt), 10), path)
    (tup, it) = torch.load(path, weights_only=False)
    assert it == 10 and len(tup) == 5
    other = model.FramePoses(3, K, 320, 256, device="cpu")
    densify.restore_poses(other, tup)
    assert torch.equal(other.pose_param_net.r.detach(), poses.pose_param_net.r.detach())
    assert torch.equal(other.get_pose(1).detach(), poses.get_pose(1).detach())
    assert np.array_equal(other.record_data['pred_w2c'], poses.record_data['pred_w2c'])
