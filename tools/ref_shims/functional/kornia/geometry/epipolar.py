"""kornia.geometry.epipolar -- the three functions scene/pose_optimizer.py:20-22 imports, with kornia's
documented semantics (shapes: R [B,3,3], t [B,3,1], K [B,3,3], points [B,N,2], F [B,3,3])."""
import torch


def _cross_matrix(t):
    tx, ty, tz = t[..., 0, 0], t[..., 1, 0], t[..., 2, 0]
    z = torch.zeros_like(tx)
    return torch.stack([torch.stack([z, -tz, ty], -1), torch.stack([tz, z, -tx], -1), torch.stack([-ty, tx, z], -1)], -2)


def relative_camera_motion(R1, t1, R2, t2):
    R = R2 @ R1.transpose(-2, -1)
    t = t2 - R @ t1
    return R, t


def essential_from_Rt(R1, t1, R2, t2):
    """E = [t]_x R of the motion from camera 1 to camera 2 (extrinsics given as world->camera)."""
    R, t = relative_camera_motion(R1, t1, R2, t2)
    return _cross_matrix(t) @ R


def fundamental_from_essential(E_mat, K1, K2):
    """F = K2^-T E K1^-1."""
    return torch.inverse(K2).transpose(-2, -1) @ E_mat @ torch.inverse(K1)


def _homog(p):
    return torch.cat([p, torch.ones_like(p[..., :1])], dim=-1)


def sampson_epipolar_distance(pts1, pts2, Fm, squared=True, eps=1e-8):
    """First-order geometric error of x2^T F x1 = 0:  (x2^T F x1)^2 / (|F x1|_{xy}^2 + |F^T x2|_{xy}^2)."""
    if pts1.shape[-1] == 2:
        pts1 = _homog(pts1)
    if pts2.shape[-1] == 2:
        pts2 = _homog(pts2)
    F_t = Fm.transpose(-2, -1)
    line1_in_2 = pts1 @ F_t            # (F x1)^T per point
    line2_in_1 = pts2 @ Fm             # (F^T x2)^T per point
    numerator = (pts2 * line1_in_2).sum(dim=-1).pow(2)
    denominator = line1_in_2[..., :2].norm(2, dim=-1).pow(2) + line2_in_1[..., :2].norm(2, dim=-1).pow(2)
    out = numerator / denominator
    if squared:
        return out
    return (out + eps).sqrt()
