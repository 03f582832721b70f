"""CPU tests of the dependency stand-ins the config-3 run executes (tools/ref_shims/functional): the epipolar
functions must have kornia's semantics -- Free-SurGS' rigid mask is thresholded Sampson distance
(scene/pose_optimizer.py:732-746, train.py:158-162) -- and the SSIM stand-in scikit-image's definition."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_shims  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tools", "ref_shims", "functional"))
from kornia.geometry import conversions, epipolar, linalg  # noqa: E402
from skimage.metrics import structural_similarity  # noqa: E402


def _two_views(n=200, seed=0):
    g = torch.Generator().manual_seed(seed)
    K = torch.tensor([[260.0, 0, 160.0], [0, 255.0, 128.0], [0, 0, 1.0]]).double()
    X = torch.cat([torch.rand(n, 2, generator=g).double() * 2 - 1, torch.rand(n, 1, generator=g).double() + 2.0], 1)

    def rot(ax, ay):
        cx, sx, cy, sy = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay)
        Rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]).double()
        Ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]).double()
        return Rx @ Ry
    R1, t1 = rot(0.02, -0.01), torch.tensor([[0.01], [0.02], [0.0]]).double()
    R2, t2 = rot(-0.03, 0.05), torch.tensor([[0.20], [-0.05], [0.03]]).double()
    proj = lambda R, t: ((K @ (R @ X.T + t)).T)
    p1, p2 = proj(R1, t1), proj(R2, t2)
    return K, (R1, t1), (R2, t2), p1[:, :2] / p1[:, 2:], p2[:, :2] / p2[:, 2:]


def test_sampson_distance_vanishes_on_true_correspondences_and_grows_with_the_error():
    K, (R1, t1), (R2, t2), x1, x2 = _two_views()
    E = epipolar.essential_from_Rt(R1[None], t1[None], R2[None], t2[None])
    F = epipolar.fundamental_from_essential(E, K[None], K[None])
    d0 = epipolar.sampson_epipolar_distance(x1[None], x2[None], F)
    assert d0.shape == (1, x1.shape[0]) and float(d0.max()) < 1e-16
    # the epipolar constraint itself
    h = lambda p: torch.cat([p, torch.ones_like(p[:, :1])], 1)
    assert float(((h(x2) @ F[0]) * h(x1)).sum(1).abs().max()) < 1e-10
    # a 2-pixel error across the epipolar line: first-order geometric distance ~ (2 px)^2 * sin^2(angle) <= 4
    noisy = x2 + torch.tensor([0.0, 2.0]).double()
    d1 = epipolar.sampson_epipolar_distance(x1[None], noisy[None], F)
    assert 0.05 < float(d1.median()) <= 4.0 + 1e-6
    assert torch.allclose(epipolar.sampson_epipolar_distance(x1[None], noisy[None], F, squared=False) ** 2, d1, atol=1e-6)


def test_rigid_transform_helpers_and_quaternion_conversion():
    g = torch.Generator().manual_seed(1)
    q = torch.nn.functional.normalize(torch.randn(5, 4, generator=g).double(), dim=1)
    q = q * torch.sign(q[:, :1])
    w, x, y, z = q.unbind(1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z),
                     2 * (y * z - w * x), 2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).view(5, 3, 3)
    got = conversions.rotation_matrix_to_quaternion(R)                                      # (w, x, y, z), up to sign
    assert float(torch.minimum((got - q).abs().amax(1), (got + q).abs().amax(1)).max()) < 1e-6
    T = torch.eye(4).double().repeat(5, 1, 1)
    T[:, :3, :3], T[:, :3, 3] = R, torch.randn(5, 3, generator=g).double()
    assert torch.allclose(linalg.compose_transformations(T, linalg.inverse_transformation(T)), torch.eye(4).double().expand(5, 4, 4), atol=1e-12)


def test_structural_similarity_stand_in():
    g = np.random.default_rng(0)
    a = g.random((40, 52, 3))
    assert abs(structural_similarity(a, a, data_range=1, channel_axis=2) - 1.0) < 1e-12
    b = np.clip(a + 0.1 * g.standard_normal(a.shape), 0, 1)
    s = structural_similarity(a, b, data_range=1, multichannel=True, channel_axis=2)      # the reference's call
    assert 0.2 < s < 0.99
    # direct evaluation of scikit-image's definition at one interior pixel of one channel (7x7 window, N/(N-1))
    x, y = a[10:17, 20:27, 0], b[10:17, 20:27, 0]
    ux, uy, n = x.mean(), y.mean(), 49
    vx, vy, vxy = x.var() * n / (n - 1), y.var() * n / (n - 1), ((x - ux) * (y - uy)).mean() * n / (n - 1)
    want = ((2 * ux * uy + 1e-4) * (2 * vxy + 9e-4)) / ((ux ** 2 + uy ** 2 + 1e-4) * (vx + vy + 9e-4))
    full = structural_similarity(a[10:17, 20:27, 0], b[10:17, 20:27, 0], data_range=1)      # 7x7 image: one window
    assert abs(full - want) < 1e-12


def test_install_never_shadows_an_installed_package():
    done = ref_shims.install()
    assert "scipy" not in done["functional"] + done["inert"] and "torch" not in done["functional"] + done["inert"]
    import PIL  # noqa: F401  (real)
    assert not isinstance(sys.modules["PIL"], ref_shims._InertModule)
