"""kornia.geometry.linalg -- rigid 4x4 helpers imported by utils/geometry_utils.py:14."""
import torch


def compose_transformations(trans_01, trans_12):
    rmat_02 = trans_01[..., :3, :3] @ trans_12[..., :3, :3]
    tvec_02 = trans_01[..., :3, :3] @ trans_12[..., :3, 3:] + trans_01[..., :3, 3:]
    out = torch.zeros_like(trans_01)
    out[..., :3, :3] = rmat_02
    out[..., :3, 3:] = tvec_02
    out[..., 3, 3] = 1.0
    return out


def inverse_transformation(trans_12):
    rmat_21 = trans_12[..., :3, :3].transpose(-2, -1)
    tvec_21 = -rmat_21 @ trans_12[..., :3, 3:]
    out = torch.zeros_like(trans_12)
    out[..., :3, :3] = rmat_21
    out[..., :3, 3:] = tvec_21
    out[..., 3, 3] = 1.0
    return out
