"""Synthetic "endo-synth" scene generator.

Everything is generated on the CPU from a seeded ``torch.Generator`` in float32 so that the
oracle (CPU), the CUDA path (GPU box) and ``bench.py`` all see bit-identical inputs.

The camera mirrors ``PoseModel.setup_camera`` (reference ``scene/pose_optimizer.py:600-633``):
SCARED-like pinhole intrinsics rescaled as in ``scene/pose_optimizer.py:413-414``, identity
rasteriser view, OpenGL-style projection with near 0.01 / far 100, white background.
The Gaussian parameters are stored in the *raw* (pre-activation) form that
``scene/gaussian_model.py:118-138`` applies its activations to.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Tuple

import torch

TEST_POSE_Q = (1.0, 0.01, -0.02, 0.015)   # (w, x, y, z), normalised by LearnPose.forward
TEST_POSE_T = (0.01, -0.005, 0.02)
# per-frame increment for multi-frame sequences (config 5): pose_k = test pose (+) k * delta
DELTA_POSE_Q = (0.0, 0.002, 0.003, -0.001)
DELTA_POSE_T = (0.004, 0.001, -0.002)


def quat_to_rot(q: torch.Tensor) -> torch.Tensor:
    """(w,x,y,z) -> 3x3, same convention as ``LearnPose.q2rot`` (pose_optimizer.py:843-860)."""
    q = q / q.norm()
    r, x, y, z = q[0], q[1], q[2], q[3]
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)]),
        torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)]),
        torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]),
    ])


def pose_matrix(q, t, dtype=torch.float32) -> torch.Tensor:
    """4x4 world->camera ``Rt`` (``LearnPose.getWorld2View2``, pose_optimizer.py:862-877)."""
    q = torch.as_tensor(q, dtype=torch.float64)
    t = torch.as_tensor(t, dtype=torch.float64)
    Rt = torch.eye(4, dtype=torch.float64)
    Rt[:3, :3] = quat_to_rot(q)
    Rt[:3, 3] = t
    return Rt.to(dtype)


def frame_pose_params(k: int) -> Tuple[Tuple[float, ...], Tuple[float, ...]]:
    """Raw (un-normalised) quaternion + translation of frame ``k`` of the synthetic sequence."""
    q = tuple(a + k * b for a, b in zip(TEST_POSE_Q, DELTA_POSE_Q))
    t = tuple(a + k * b for a, b in zip(TEST_POSE_T, DELTA_POSE_T))
    return q, t


@dataclass
class SynthCamera:
    """Plain-data twin of ``GaussianRasterizationSettings`` as built by ``setup_camera``."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor            # [3]
    scale_modifier: float
    viewmatrix: torch.Tensor    # [1,4,4]  (w2c transposed = column-major)
    projmatrix: torch.Tensor    # [1,4,4]
    sh_degree: int
    campos: torch.Tensor        # [3]
    prefiltered: bool = False
    debug: bool = False
    fx: float = 0.0
    fy: float = 0.0
    cx: float = 0.0
    cy: float = 0.0


def make_camera(width: int, height: int, w2c: torch.Tensor | None = None,
                near: float = 0.01, far: float = 100.0) -> SynthCamera:
    """Restates ``PoseModel.setup_camera`` (pose_optimizer.py:600-633) for the synthetic intrinsics."""
    fx = fy = 1035.0 * width / 1280.0
    cx, cy = width / 2.0, height / 2.0
    if w2c is None:
        w2c = torch.eye(4, dtype=torch.float32)
    w2c = w2c.float()
    cam_center = torch.inverse(w2c)[:3, 3].contiguous()
    w2c_t = w2c.unsqueeze(0).transpose(1, 2)
    opengl_proj = torch.tensor([[2 * fx / width, 0.0, -(width - 2 * cx) / width, 0.0],
                                [0.0, 2 * fy / height, -(height - 2 * cy) / height, 0.0],
                                [0.0, 0.0, far / (far - near), -(far * near) / (far - near)],
                                [0.0, 0.0, 1.0, 0.0]]).float().unsqueeze(0).transpose(1, 2)
    full_proj = w2c_t.bmm(opengl_proj)
    return SynthCamera(
        image_height=height, image_width=width,
        tanfovx=width / (2 * fx), tanfovy=height / (2 * fy),
        bg=torch.ones(3, dtype=torch.float32), scale_modifier=1.0,
        viewmatrix=w2c_t.contiguous(), projmatrix=full_proj.contiguous(),
        sh_degree=0, campos=cam_center, prefiltered=False, debug=False,
        fx=fx, fy=fy, cx=cx, cy=cy)


@dataclass
class SynthScene:
    P: int
    width: int
    height: int
    size_mult: float
    seed: int
    camera: SynthCamera
    params: Dict[str, torch.Tensor]          # raw parameters, reference names
    pose_q: torch.Tensor                     # [4] raw quaternion (w,x,y,z)
    pose_t: torch.Tensor                     # [3]
    active_sh_degree: int = 3
    grads_out: Dict[str, torch.Tensor] = field(default_factory=dict)  # fixed upstream gradients

    def Rt(self, dtype=torch.float32) -> torch.Tensor:
        return pose_matrix(self.pose_q.tolist(), self.pose_t.tolist(), dtype)


def make_scene(P: int, width: int, height: int, size_mult: float = 2.0, seed: int = 0,
               frame: int = 0, with_upstream_grads: bool = True) -> SynthScene:
    """SURVEY.md 8d generator.  Gaussians lie on a smooth depth-[0.3,2] surface seen by the
    test pose; ``size_mult`` (m) scales splat footprints.  ``frame`` selects the pose of a
    multi-frame sequence (the Gaussians themselves are placed from frame 0's pose)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    cam = make_camera(width, height)
    W, H = float(width), float(height)

    def rnd(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32)

    def uni(*shape):
        return torch.rand(*shape, generator=g, dtype=torch.float32)

    u = (uni(P) * 1.1 - 0.05) * W
    v = (uni(P) * 1.1 - 0.05) * H
    z = 1.0 + 0.4 * torch.sin(2 * math.pi * u / W) * torch.cos(2 * math.pi * v / H) + 0.05 * rnd(P)
    z = z.clamp(0.3, 2.0)
    mean_cam = torch.stack([(u - cam.cx) * z / cam.fx, (v - cam.cy) * z / cam.fy, z], dim=1)

    q0, t0 = frame_pose_params(0)
    Rt0 = pose_matrix(q0, t0, torch.float64)
    Rinv = Rt0[:3, :3].T
    xyz = ((mean_cam.double() - Rt0[:3, 3]) @ Rinv.T).float()   # world mean = Rt^-1 * mean

    base = size_mult * (z / cam.fx) * math.sqrt(H * W / P)
    scales = base[:, None] * torch.exp(0.5 * rnd(P, 3))
    rot = rnd(P, 4)
    rot = rot / rot.norm(dim=1, keepdim=True)
    opacity_logit = 1.5 * rnd(P, 1)
    f_dc = 0.5 * rnd(P, 1, 3)
    f_rest = 0.1 * rnd(P, 15, 3)

    params = {
        "_xyz": xyz.contiguous(),
        "_features_dc": f_dc.contiguous(),
        "_features_rest": f_rest.contiguous(),
        "_opacity": opacity_logit.contiguous(),            # get_opacity = sigmoid
        "_scaling": torch.log(scales).contiguous(),        # get_scaling = exp
        "_rotation": rot.contiguous(),                     # get_rotation = normalize
    }
    q, t = frame_pose_params(frame)
    scene = SynthScene(P=P, width=width, height=height, size_mult=size_mult, seed=seed,
                       camera=cam, params=params,
                       pose_q=torch.tensor(q, dtype=torch.float32),
                       pose_t=torch.tensor(t, dtype=torch.float32))
    if with_upstream_grads:
        g2 = torch.Generator(device="cpu").manual_seed(seed + 7919)
        scene.grads_out = {
            "G_rgb": torch.randn(3, height, width, generator=g2, dtype=torch.float32),
            "G_dep": torch.randn(height, width, generator=g2, dtype=torch.float32),
        }
    return scene
