"""Host-side mirrors of the reference objects that sit either side of the rasteriser, so the
path can be driven (tests, bench, multi-GPU) without importing /root/reference.  Names, shapes
and semantics follow the reference:

  * ``LearnPose``          -- scene/pose_optimizer.py:755-877 (r [1,4,N] quaternion (w,x,y,z), t [3,N])
  * ``FramePoses``         -- the slice of ``PoseModel`` the renderer touches: ``get_pose``,
                              ``cam_center``, ``setup_camera`` (pose_optimizer.py:600-638)
  * ``SplatModel``         -- the slice of ``GaussianModel`` the renderer touches: ``params``,
                              ``variables``, ``cam``, activations (gaussian_model.py:40-138)
  * ``transform_to_frame`` -- pose_optimizer.py:960-989
  * ``eval_sh``            -- utils/sh_utils.py:57-112

These are plain PyTorch (device plumbing around the library call), not a compute fallback: the
rasterisation itself always goes through ``libfsgs_raster.so``.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .rasterizer import GaussianRasterizationSettings

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def eval_sh(deg: int, sh: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    assert 0 <= deg <= 3
    assert sh.shape[-1] >= (deg + 1) ** 2
    result = C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        result = result - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            result = (result + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5]
                      + C2[2] * (2.0 * zz - xx - yy) * sh[..., 6] + C2[3] * xz * sh[..., 7]
                      + C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                result = (result + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10]
                          + C3[2] * y * (4 * zz - xx - yy) * sh[..., 11]
                          + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
                          + C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + C3[5] * z * (xx - yy) * sh[..., 14]
                          + C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return result


def transform_to_frame(means3D, viewmatrix, gaussians_grad=True, camera_grad=True):
    rel_w2c = viewmatrix if camera_grad else viewmatrix.detach()
    pts = means3D if gaussians_grad else means3D.detach()
    pts_ones = torch.ones(pts.shape[0], 1, device=pts.device, dtype=pts.dtype)
    pts4 = torch.cat((pts, pts_ones), dim=1)
    return (rel_w2c @ pts4.T).T[:, :3]


class _PoseFn(torch.autograd.Function):
    """LearnPose.forward as two one-thread library kernels (fsgs_pose_forward / _backward) instead of
    the ~150 element-wise PyTorch launches the reference's q2rot + autograd take per frame."""

    @staticmethod
    def forward(ctx, r, t, cam_id):
        import ctypes
        from . import _lib
        from .rasterizer import _on_device
        rc_, tc_ = r.detach().contiguous(), t.detach().contiguous()
        n = rc_.shape[-1]
        Rt = torch.empty(4, 4, dtype=torch.float32, device=r.device)
        s = ctypes.c_void_p(torch.cuda.current_stream(r.device).cuda_stream)
        with _on_device(r.device):
            _lib.check(_lib.lib().fsgs_pose_forward(ctypes.c_void_p(rc_.data_ptr()), ctypes.c_void_p(tc_.data_ptr()),
                                                    int(cam_id), int(n), ctypes.c_void_p(Rt.data_ptr()), s))
        ctx.save_for_backward(rc_)
        ctx.cam_id, ctx.n, ctx.shapes = int(cam_id), int(n), (r.shape, t.shape)
        return Rt

    @staticmethod
    def backward(ctx, g):
        import ctypes
        from . import _lib
        (rc_,) = ctx.saved_tensors
        g = g.contiguous().float()
        dr = torch.empty(ctx.shapes[0], dtype=torch.float32, device=g.device)
        dt = torch.empty(ctx.shapes[1], dtype=torch.float32, device=g.device)
        from .rasterizer import _on_device
        s = ctypes.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)
        with _on_device(g.device):
            _lib.check(_lib.lib().fsgs_pose_backward(ctypes.c_void_p(rc_.data_ptr()), ctx.cam_id, ctx.n,
                                                     ctypes.c_void_p(g.data_ptr()), ctypes.c_void_p(dr.data_ptr()),
                                                     ctypes.c_void_p(dt.data_ptr()), s))
        return dr, dt, None


class LearnPose(nn.Module):
    """Per-frame learnable pose: quaternion ``r[1,4,N]`` (w,x,y,z) + translation ``t[3,N]``."""

    def __init__(self, num_cams: int, device="cuda"):
        super().__init__()
        self.num_cams = num_cams
        r = torch.zeros(1, 4, num_cams, device=device)
        r[0, 0, :] = 1.0
        self.r = nn.Parameter(r.contiguous())
        self.t = nn.Parameter(torch.zeros(3, num_cams, device=device))

    @staticmethod
    def q2rot(q):
        norm = torch.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])
        q = q / norm[:, None]
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        rot = torch.stack([
            1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
            2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
            2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1)
        return rot.view(-1, 3, 3)

    def forward(self, cam_id: int) -> torch.Tensor:
        cam_id = int(cam_id)
        if self.r.is_cuda:
            return _PoseFn.apply(self.r, self.t, cam_id)
        # CPU tensors (host-logic tests only): the reference's own formulation
        r = F.normalize(self.r[..., cam_id])
        t = self.t[..., cam_id]
        R = self.q2rot(r)[0]
        top = torch.cat([R, t[:, None]], dim=1)
        bottom = torch.tensor([[0.0, 0.0, 0.0, 1.0]], device=top.device, dtype=top.dtype)
        return torch.cat([top, bottom], dim=0)


def setup_camera(w2c, K, w: int, h: int, device="cuda", near=0.01, far=100.0) -> GaussianRasterizationSettings:
    """``PoseModel.setup_camera`` (pose_optimizer.py:600-633): returns (settings, cam_center)."""
    w2c = torch.as_tensor(np.asarray(w2c), dtype=torch.float32, device=device)
    cam_center = torch.inverse(w2c)[:3, 3]
    w2c_t = w2c.unsqueeze(0).transpose(1, 2)
    fx, fy, cx, cy = float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2])
    opengl_proj = torch.tensor([[2 * fx / w, 0.0, -(w - 2 * cx) / w, 0.0],
                                [0.0, 2 * fy / h, -(h - 2 * cy) / h, 0.0],
                                [0.0, 0.0, far / (far - near), -(far * near) / (far - near)],
                                [0.0, 0.0, 1.0, 0.0]], device=device).float().unsqueeze(0).transpose(1, 2)
    full_proj = w2c_t.bmm(opengl_proj)
    cam = GaussianRasterizationSettings(
        image_height=h, image_width=w, tanfovx=w / (2 * fx), tanfovy=h / (2 * fy),
        bg=torch.tensor([1, 1, 1], dtype=torch.float32, device=device), scale_modifier=1.0,
        viewmatrix=w2c_t.contiguous(), projmatrix=full_proj.contiguous(), sh_degree=0, campos=cam_center,
        prefiltered=False, debug=False)
    return cam, cam_center


class FramePoses:
    """What ``render`` needs from ``PoseModel``: ``get_pose(i)``, ``cam_center``."""

    def __init__(self, num_cams: int, K, w: int, h: int, device="cuda"):
        self.pose_param_net = LearnPose(num_cams, device=device)
        self.record_data = {"intrinsic": np.asarray(K, dtype=np.float64), "image_width": w, "image_height": h,
                            "pred_w2c": np.tile(np.eye(4), (num_cams, 1, 1))}
        self.device = device
        self.cam_center = torch.zeros(3, device=device)

    def setup_camera(self, w2c):
        cam, self.cam_center = setup_camera(w2c, self.record_data["intrinsic"], self.record_data["image_width"],
                                            self.record_data["image_height"], device=self.device)
        return cam

    def get_pose(self, timestep: int, record: bool = False) -> torch.Tensor:
        """The reference also copies the pose to the host here (pose_optimizer.py:637, one sync per
        call); ``record=True`` reproduces that, the default leaves the stream asynchronous."""
        update_pose = self.pose_param_net.forward(timestep)
        if record:
            self.record_data['pred_w2c'][timestep] = update_pose.detach().cpu().numpy()
        return update_pose

    def set_pose(self, k: int, q, t) -> None:
        with torch.no_grad():
            self.pose_param_net.r[0, :, k] = torch.as_tensor(q, dtype=torch.float32, device=self.device)
            self.pose_param_net.t[:, k] = torch.as_tensor(t, dtype=torch.float32, device=self.device)


class SplatModel:
    """What ``render`` needs from ``GaussianModel``: raw ``params`` (reference names), activations,
    ``variables`` bookkeeping, the single rasteriser settings ``cam``."""

    def __init__(self, params: Dict[str, torch.Tensor], cam: Optional[GaussianRasterizationSettings] = None,
                 active_sh_degree: int = 3, max_sh_degree: int = 3, requires_grad: bool = True):
        self.params = {k: (v.detach().clone().requires_grad_(requires_grad)) for k, v in params.items()}
        P = self.params['_xyz'].shape[0]
        dev = self.params['_xyz'].device
        self.variables = {'max_radii2D': torch.zeros(P, device=dev), 'xyz_gradient_accum': torch.zeros(P, 1, device=dev),
                          'denom': torch.zeros(P, 1, device=dev)}
        self.cam = cam
        self.active_sh_degree = active_sh_degree
        self.max_sh_degree = max_sh_degree

    get_xyz = property(lambda self: self.params['_xyz'])
    get_scaling = property(lambda self: torch.exp(self.params['_scaling']))
    get_rotation = property(lambda self: F.normalize(self.params['_rotation']))
    get_opacity = property(lambda self: torch.sigmoid(self.params['_opacity']))
    get_features = property(lambda self: torch.cat((self.params['_features_dc'], self.params['_features_rest']), dim=1))

    def add_densification_stats(self, viewspace_point_tensor, update_filter):
        """gaussian_model.py:678-681."""
        grad = viewspace_point_tensor.grad[update_filter]
        self.variables['xyz_gradient_accum'][update_filter] += torch.norm(grad, dim=-1, keepdim=True)
        self.variables['denom'][update_filter] += 1

    def zero_grad(self):
        for v in self.params.values():
            v.grad = None


def scene_to_device(scene, device="cuda"):
    """Build (FramePoses, SplatModel) on ``device`` from a ``synth.SynthScene``."""
    cam = scene.camera
    K = [[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1]]
    poses = FramePoses(1, K, scene.width, scene.height, device=device)
    settings = poses.setup_camera(np.eye(4))
    poses.set_pose(0, scene.pose_q.tolist(), scene.pose_t.tolist())
    model = SplatModel({k: v.to(device) for k, v in scene.params.items()}, cam=settings,
                       active_sh_degree=scene.active_sh_degree)
    return poses, model
