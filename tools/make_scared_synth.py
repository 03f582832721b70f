#!/usr/bin/env python
"""Write a synthetic sequence in the on-disk format Free-SurGS' loader reads (the real
``scared_demo`` is a Google-Drive download, unavailable offline).

Format, from ``PoseModel.__init__`` (reference scene/pose_optimizer.py:355-414):

    <root>/input/{scene}_{data}_{tag}_{img}.png                 RGB frames, sorted by name
    <root>/poses/{scene}_{data}/frame_{img}.json                {"camera-pose": 4x4, "camera-calibration": {"KL": 3x3}}
    <root>/flow/flow_fw_{name}.npz, flow_bw_{name}.npz          ['pred'] float32 [1,2,H,W]  (all frames but the last)
    <root>/monodep/depth_{name}.npz                             ['pred'] float32 [H,W]      INVERSE depth

``KL`` is the intrinsic matrix at 1280x1024; the loader rescales it to the image size (:413-414).  The loader
normalises 1/pred to [0.5, 1.5] per frame (:407), so the scene is built with depths in about that range.

The frames are renderings of the "endo-synth" Gaussian scene (fsgs_b200/synth.py) from the poses
``frame_pose_params(k)``: colour and expected depth come out of the library's own rasteriser (or, with
``renderer="oracle"``, the CPU oracle -- for CPU tests of this writer), optical flow is the exact rigid flow
of that depth under the relative pose, the "monocular depth" is the exact inverse depth and the pose files hold
the ground-truth world->camera matrices relative to frame 0 (train.py pins frame 0 to the identity, train.py:41).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "free-surgs_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from fsgs_b200.synth import frame_pose_params, make_scene, pose_matrix  # noqa: E402


def _render_frame_fsgs(sc, Rt, device):
    """-> (rgb [3,H,W] with white background, expected depth [H,W], coverage [H,W]) on the CPU."""
    from fsgs_b200 import model
    from fsgs_b200.rasterizer import GaussianRasterizer
    poses, pc = model.scene_to_device(sc, device)
    with torch.no_grad():
        xyz = pc.params["_xyz"]
        Rt_d = Rt.to(device)
        means_cam = model.transform_to_frame(xyz, Rt_d)
        feats = pc.get_features
        shs_view = feats.transpose(1, 2).reshape(-1, 3, 16)
        d = xyz - poses.cam_center[None]
        d = d / d.norm(dim=1, keepdim=True)
        colors = torch.clamp_min(model.eval_sh(3, shs_view, d) + 0.5, 0.0)
        kw = dict(means3D=means_cam, means2D=torch.zeros_like(xyz), opacities=pc.get_opacity, scales=pc.get_scaling,
                  rotations=pc.get_rotation)
        rgb, _, _ = GaussianRasterizer(pc.cam)(colors_precomp=colors, **kw)
        black = pc.cam._replace(bg=torch.zeros(3, device=device))
        cov, _, dsum = GaussianRasterizer(black)(colors_precomp=torch.ones_like(colors), **kw)
    return rgb.cpu(), dsum[0].cpu(), cov[0].cpu()


def _render_frame_oracle(sc, Rt):
    from oracle import c_oracle
    from oracle import render_oracle as R
    dt = torch.float32
    p = {k: v.to(dt) for k, v in sc.params.items()}
    with torch.no_grad():
        means_cam = R.transform_to_frame(p["_xyz"], Rt.to(dt))
        feats = torch.cat((p["_features_dc"], p["_features_rest"]), dim=1)
        colors = R.sh_colors(p["_xyz"], feats, 3, sc.camera.campos.to(dt))
        kw = dict(means2D=torch.zeros_like(means_cam), opacities=torch.sigmoid(p["_opacity"]),
                  scales=torch.exp(p["_scaling"]), rotations=torch.nn.functional.normalize(p["_rotation"]))
        rgb, _, _, _ = c_oracle.rasterize(means_cam, st=sc.camera, colors_precomp=colors, **kw)
        import copy
        black = copy.copy(sc.camera)
        black.bg = torch.zeros(3)
        cov, _, dsum, _ = c_oracle.rasterize(means_cam, st=black, colors_precomp=torch.ones_like(colors), **kw)
    return rgb, dsum[0], cov[0]


def rigid_flow(depth, K, T_ab):
    """Flow [2,H,W] (x, y) that takes pixels of frame a (with depth `depth`) to frame b; T_ab = w2c_b @ c2w_a."""
    H, W = depth.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    z = depth.double()
    X = torch.stack([(xs - cx) / fx * z, (ys - cy) / fy * z, z, torch.ones_like(z)], dim=0).reshape(4, -1)
    Xb = (T_ab.double() @ X)[:3]
    u = fx * Xb[0] / Xb[2] + cx
    v = fy * Xb[1] / Xb[2] + cy
    return torch.stack([u.reshape(H, W) - xs, v.reshape(H, W) - ys], dim=0).float()


def sequence_pose(k, motion_scale=4.0):
    """World->camera of frame k: the test pose plus k * motion_scale * the per-frame increment of
    fsgs_b200.synth (motion_scale 4: ~0.5 degree and 1.8 % of the scene depth per frame, a few pixels of flow at
    320x256 -- the order of the SCARED sequences' inter-frame motion after the loader's depth normalisation)."""
    from fsgs_b200.synth import DELTA_POSE_Q, DELTA_POSE_T, TEST_POSE_Q, TEST_POSE_T
    q = tuple(a + k * motion_scale * b for a, b in zip(TEST_POSE_Q, DELTA_POSE_Q))
    t = tuple(a + k * motion_scale * b for a, b in zip(TEST_POSE_T, DELTA_POSE_T))
    return pose_matrix(q, t, dtype=torch.float64)


def write_sequence(root, n_frames=8, W=320, H=256, P=20000, m=2.0, seed=0, device="cuda", renderer="fsgs",
                   scene_id="1", data_id="1", motion_scale=4.0):
    """Returns a dict with the ground truth (w2c relative to frame 0 [N,4,4], K at image size, depths [N,H,W])."""
    from PIL import Image
    for sub in ("input", os.path.join("poses", f"{scene_id}_{data_id}"), "flow", "monodep"):
        os.makedirs(os.path.join(root, sub), exist_ok=True)
    sc = make_scene(P, W, H, size_mult=m, seed=seed)
    cam = sc.camera
    K = torch.tensor([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1]], dtype=torch.float64)
    KL = K.clone()
    KL[0] *= 1280.0 / W
    KL[1] *= 1024.0 / H
    Rts = [sequence_pose(k, motion_scale) for k in range(n_frames)]
    rel = [Rt @ torch.inverse(Rts[0]) for Rt in Rts]                 # world := camera frame of frame 0
    names, depths = [], []
    for k in range(n_frames):
        if renderer == "fsgs":
            rgb, dsum, cov = _render_frame_fsgs(sc, Rts[k].float(), device)
        else:
            rgb, dsum, cov = _render_frame_oracle(sc, Rts[k].float())
        depth = torch.where(cov > 0.5, dsum / cov.clamp(min=1e-6), torch.full_like(dsum, float("nan")))
        depth = torch.where(torch.isnan(depth), torch.nanmean(depth), depth)     # uncovered pixels: the mean depth
        depths.append(depth)
        name = f"{scene_id}_{data_id}_frame_{k:06d}"
        names.append(name)
        img = (rgb.clamp(0, 1).permute(1, 2, 0).numpy() * 255.0 + 0.5).astype(np.uint8)
        Image.fromarray(img).save(os.path.join(root, "input", name + ".png"))
        np.savez(os.path.join(root, "monodep", f"depth_{name}.npz"), pred=(1.0 / depth).numpy().astype(np.float32))
        with open(os.path.join(root, "poses", f"{scene_id}_{data_id}", f"frame_{k:06d}.json"), "w") as f:
            json.dump({"camera-pose": rel[k].tolist(), "camera-calibration": {"KL": KL.tolist()}}, f)
    for k in range(n_frames - 1):
        fw = rigid_flow(depths[k], K, rel[k + 1] @ torch.inverse(rel[k]))
        bw = rigid_flow(depths[k + 1], K, rel[k] @ torch.inverse(rel[k + 1]))
        np.savez(os.path.join(root, "flow", f"flow_fw_{names[k]}.npz"), pred=fw[None].numpy())
        np.savez(os.path.join(root, "flow", f"flow_bw_{names[k]}.npz"), pred=bw[None].numpy())
    return {"w2c": torch.stack(rel), "K": K, "depths": torch.stack(depths), "names": names}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("root")
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--size", default="320x256")
    ap.add_argument("--P", type=int, default=20000)
    ap.add_argument("--m", type=float, default=2.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--renderer", default="fsgs", choices=["fsgs", "oracle"])
    a = ap.parse_args()
    W, H = (int(x) for x in a.size.split("x"))
    gt = write_sequence(a.root, a.frames, W, H, a.P, a.m, a.seed, renderer=a.renderer)
    print("wrote", a.root, "frames", a.frames, f"{W}x{H}", "depth range", float(gt["depths"].min()), float(gt["depths"].max()))


if __name__ == "__main__":
    main()
