"""kornia.geometry.conversions -- rotation matrix -> quaternion (w, x, y, z), the one function the
reference names (scene/pose_optimizer.py:529, utils/general_utils.py:121; both on paths train.py leaves off)."""
import torch


def rotation_matrix_to_quaternion(R, eps=1e-8):
    R = R.reshape(-1, 3, 3)
    m00, m01, m02 = R[:, 0, 0], R[:, 0, 1], R[:, 0, 2]
    m10, m11, m12 = R[:, 1, 0], R[:, 1, 1], R[:, 1, 2]
    m20, m21, m22 = R[:, 2, 0], R[:, 2, 1], R[:, 2, 2]
    tr = m00 + m11 + m22
    out = torch.zeros(R.shape[0], 4, dtype=R.dtype, device=R.device)
    for i in range(R.shape[0]):
        if tr[i] > 0:
            s = torch.sqrt(tr[i] + 1.0 + eps) * 2
            q = (0.25 * s, (m21[i] - m12[i]) / s, (m02[i] - m20[i]) / s, (m10[i] - m01[i]) / s)
        elif m00[i] > m11[i] and m00[i] > m22[i]:
            s = torch.sqrt(1.0 + m00[i] - m11[i] - m22[i] + eps) * 2
            q = ((m21[i] - m12[i]) / s, 0.25 * s, (m01[i] + m10[i]) / s, (m02[i] + m20[i]) / s)
        elif m11[i] > m22[i]:
            s = torch.sqrt(1.0 + m11[i] - m00[i] - m22[i] + eps) * 2
            q = ((m02[i] - m20[i]) / s, (m01[i] + m10[i]) / s, 0.25 * s, (m12[i] + m21[i]) / s)
        else:
            s = torch.sqrt(1.0 + m22[i] - m00[i] - m11[i] + eps) * 2
            q = ((m10[i] - m01[i]) / s, (m02[i] + m20[i]) / s, (m12[i] + m21[i]) / s, 0.25 * s)
        out[i] = torch.stack([torch.as_tensor(v, dtype=R.dtype, device=R.device) for v in q])
    return out
