"""Config 3 in miniature: Free-SurGS' progressive joint pose + Gaussian optimisation (reference
train.py:318-345 ``progressive_run`` -> ``tracking`` :154-210 -> ``mapping`` :213-295) on a short synthetic
sequence, driven only through the drop-in surface (``render``, ``LearnPose`` poses, the PyTorch losses).

The real ``scared_demo`` sequence and the reference's ``train.py`` cannot travel to the GPU box, so the loop is
restated here with the same structure: frame 0's pose is the anchor, each new frame starts from the previous
frame's estimate (pose_optimizer.py:512-515), is tracked with the Gaussian model frozen (pose gradient only --
the library's pose-only backward), then the model is refined on (a random keyframe, the new frame) with the pose
detached.  Reported like the reference's evaluation: PSNR of the re-rendered frames and the absolute
trajectory error of the recovered camera centres."""
import math
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))

from fsgs_b200 import _lib  # noqa: E402
from fsgs_b200.synth import frame_pose_params, make_scene  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _psnr(a, b):
    return float(-10.0 * torch.log10(((a - b) ** 2).mean().clamp_min(1e-12)))


def _cam_centre(Rt):
    Rt = Rt.detach().double().cpu()
    return (-Rt[:3, :3].T @ Rt[:3, 3]).numpy()


def test_progressive_tracking_and_mapping_on_a_synthetic_sequence():
    from fsgs_b200 import frame_render as render
    from fsgs_b200 import model
    from fsgs_b200.losses import rgb_loss_func_fused as rgb_loss_func

    N, P, W, H = 4, 20000, 320, 256
    sc = make_scene(P, W, H, size_mult=2.0, seed=7)
    cam = sc.camera
    K = [[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1]]

    # ---- ground truth: the scene seen from N poses ------------------------------------------------------
    poses_gt = model.FramePoses(N, K, W, H, device=DEV)
    settings = poses_gt.setup_camera(np.eye(4))
    for k in range(N):
        poses_gt.set_pose(k, *frame_pose_params(k))
    pc_gt = model.SplatModel({k: v.to(DEV) for k, v in sc.params.items()}, cam=settings, requires_grad=False)
    with torch.no_grad():
        targets = [render.render(poses_gt, k, pc_gt, gs_grad=False, cam_grad=False)["render"].clone() for k in range(N)]
        Rt_gt = [poses_gt.get_pose(k).clone() for k in range(N)]

    # ---- learner: wrong colours, unknown poses except the anchor frame -----------------------------------
    g = torch.Generator().manual_seed(1)
    params = {k: v.clone() for k, v in sc.params.items()}
    params["_features_dc"] = params["_features_dc"] + 0.25 * torch.randn(params["_features_dc"].shape, generator=g)
    poses = model.FramePoses(N, K, W, H, device=DEV)
    poses.setup_camera(np.eye(4))
    poses.set_pose(0, *frame_pose_params(0))
    pc = model.SplatModel({k: v.to(DEV) for k, v in params.items()}, cam=settings)
    with torch.no_grad():
        psnr0 = [_psnr(render.render(poses_gt, k, pc, gs_grad=False, cam_grad=False)["render"], targets[k]) for k in range(N)]

    opt_g = torch.optim.Adam([{"params": [pc.params["_features_dc"]], "lr": 1e-2},
                              {"params": [pc.params["_opacity"]], "lr": 1e-2}], eps=1e-15)

    def mapping(frames, iters):
        for v in pc.params.values():
            v.requires_grad_(True)
        for _ in range(iters):
            pc.zero_grad()
            loss = 0.0
            for f in frames:
                out = render.render(poses, f, pc, gs_grad=True, cam_grad=False)          # train.py:245-249
                loss = loss + rgb_loss_func(out["render"], targets[f])
            loss.backward()
            opt_g.step()

    def tracking(t, iters):
        for v in pc.params.values():                      # frozen model: only dL/dpose is computed
            v.requires_grad_(False)
        # (a fresh Adam per frame: the other frames' columns have zero gradient and zero state, so they do not move)
        opt_p = torch.optim.Adam([poses.pose_param_net.r, poses.pose_param_net.t], lr=1e-3, eps=1e-15)
        for _ in range(iters):
            opt_p.zero_grad(set_to_none=True)
            out = render.render(poses, t, pc, gs_grad=False, cam_grad=True)              # train.py:167-171
            loss = (out["render"] - targets[t]).abs().mean()
            loss.backward()
            opt_p.step()

    keyframes = [0]
    mapping([0], 40)
    _lib.profile_enable(True)
    err_init, err_final = [], []
    try:
        for t in range(1, N):
            with torch.no_grad():                         # pose_optimizer.py:512-515
                poses.pose_param_net.r[..., t] = poses.pose_param_net.r[..., t - 1]
                poses.pose_param_net.t[..., t] = poses.pose_param_net.t[..., t - 1]
            err_init.append(float(np.linalg.norm(_cam_centre(poses.get_pose(t)) - _cam_centre(Rt_gt[t]))))
            tracking(t, 120)
            err_final.append(float(np.linalg.norm(_cam_centre(poses.get_pose(t)) - _cam_centre(Rt_gt[t]))))
            kf = keyframes[int(torch.randint(len(keyframes), (1,), generator=g))]
            mapping([kf, t], 20)
            keyframes.append(t)
        prof = _lib.profile_collect()
    finally:
        _lib.profile_enable(False)
    # tracking ran on the pose-only kernels, mapping on the general ones
    assert prof["k_preprocess_pose_bwd"][1] == 120 * (N - 1)
    assert prof["k_preprocess_fused_bwd"][1] == 2 * 20 * (N - 1)

    with torch.no_grad():
        psnr1 = [_psnr(render.render(poses, k, pc, gs_grad=False, cam_grad=False)["render"], targets[k]) for k in range(N)]
    ate = math.sqrt(sum(float(np.sum((_cam_centre(poses.get_pose(k)) - _cam_centre(Rt_gt[k])) ** 2)) for k in range(N)) / N)
    step = float(np.linalg.norm(_cam_centre(Rt_gt[1]) - _cam_centre(Rt_gt[0])))
    print(f"sequence: PSNR {np.mean(psnr0):.2f} -> {np.mean(psnr1):.2f} dB, ATE {ate:.2e} (inter-frame step {step:.2e}), "
          f"per-frame centre error before/after tracking {err_init} / {err_final}")
    assert np.mean(err_final) < 0.5 * np.mean(err_init), (err_init, err_final)
    assert ate < 0.5 * step, (ate, step)
    assert np.mean(psnr1) > np.mean(psnr0) + 2.0, (psnr0, psnr1)
