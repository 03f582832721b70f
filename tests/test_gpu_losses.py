"""The library's fused image loss (fsgs_rgb_loss_forward / _backward: L1 + SSIM of the reference's
``rgb_loss_func``, utils/loss_utils.py:47-96) against

  * the reference's own outputs (tests/golden/ref_python_half.npz: values AND the gradient w.r.t. the image), and
  * the PyTorch formulation (fsgs_b200.losses.rgb_loss_func, itself pinned to the reference by
    tests/test_losses_golden.py) evaluated in float64 on the same inputs.

Tolerances: 1e-5 abs on the loss, 1e-4 relative (L2) on the gradient -- or three times the float32 PyTorch
formulation's own distance from float64 where that is larger (sigma^2 = E[x^2] - mu^2 cancels in float32 for both)."""
import os
import sys
import time

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from parity import rel_err  # noqa: E402

from fsgs_b200 import _lib  # noqa: E402
from fsgs_b200 import losses as L  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_python_half.npz"))


def _truth(img, gt, lam, mask, scale):
    x = img.double().detach().requires_grad_(True)
    m = None if mask is None else mask.double() if mask.dtype != torch.bool else mask
    loss = L.rgb_loss_func(x, gt.double(), lam, m) * scale
    loss.backward()
    return float(loss), x.grad


def _torch32(img, gt, lam, mask, scale):
    x = img.detach().clone().requires_grad_(True)
    loss = L.rgb_loss_func(x, gt, lam, mask) * scale
    loss.backward()
    return float(loss), x.grad


def _fused(img, gt, lam, mask, scale):
    x = img.detach().clone().requires_grad_(True)
    loss = L.rgb_loss_func_fused(x, gt, lam, mask) * scale
    loss.backward()
    return float(loss), x.grad


@pytest.mark.parametrize("shape", [(3, 152, 200), (3, 256, 320), (1, 40, 33), (3, 16, 16), (3, 7, 300)])
@pytest.mark.parametrize("mask_kind", ["none", "bool_hw", "bool_1hw", "float_chw"])
def test_fused_rgb_loss_matches_the_pytorch_formulation(shape, mask_kind):
    C, H, W = shape
    g = torch.Generator().manual_seed(H * 1000 + W)
    # smooth image + noise (flat regions stress the sigma^2 cancellation), target = perturbed image
    yy, xx = torch.meshgrid(torch.linspace(0, 3, H), torch.linspace(0, 4, W), indexing="ij")
    base = 0.5 + 0.4 * torch.sin(xx + 0.5 * yy)[None].expand(C, H, W)
    img = (base + 0.05 * torch.randn(C, H, W, generator=g)).clamp(0, 1).to(DEV)
    gt = (base + 0.10 * torch.randn(C, H, W, generator=g)).clamp(0, 1).to(DEV)
    mask = {"none": None,
            "bool_hw": (torch.rand(H, W, generator=g) > 0.3).to(DEV),
            "bool_1hw": (torch.rand(1, H, W, generator=g) > 0.3).to(DEV),
            "float_chw": torch.rand(C, H, W, generator=g).to(DEV)}[mask_kind]
    lam, scale = 0.2, 5.0
    l64, g64 = _truth(img, gt, lam, mask, scale)
    l32, g32 = _torch32(img, gt, lam, mask, scale)
    lf, gf = _fused(img, gt, lam, mask, scale)
    assert abs(lf - l64) <= 1e-5 * scale, (lf, l64, l32)
    tol = max(1e-4, 3.0 * rel_err(g32.double(), g64))
    assert rel_err(gf.double(), g64) <= tol, (rel_err(gf.double(), g64), rel_err(g32.double(), g64))
    assert torch.isfinite(gf).all()


def test_fused_rgb_loss_matches_the_reference_golden_vectors():
    T = lambda k: torch.from_numpy(GOLD[k]).to(DEV)
    img, gt, msk = T("loss_img"), T("loss_gt"), T("loss_mask")
    close = lambda a, b: abs(float(a) - float(b)) <= 1e-5 * max(1.0, abs(float(b)))
    assert close(L.rgb_loss_func_fused(img, gt), T("loss_rgb"))
    assert close(L.rgb_loss_func_fused(img, gt, mask=msk), T("loss_rgb_masked"))
    # the golden gradient is that of  5 * rgb_loss + depth terms  w.r.t. the image (sub-sampled 8x8)
    x = img.clone().requires_grad_(True)
    (L.rgb_loss_func_fused(x, gt) * 5.0).backward()
    assert rel_err(x.grad[:, ::8, ::8].cpu(), torch.from_numpy(GOLD["loss_dimg"])) < 1e-4
    # aux outputs of the forward: (loss, L1, SSIM)
    out = L._RgbLossFused.apply(img, gt, None, 0.2)
    assert close(out, T("loss_rgb"))


def test_fused_rgb_loss_is_deterministic_and_needs_no_gradient_buffers_without_grad():
    g = torch.Generator().manual_seed(3)
    img, gt = torch.rand(3, 300, 420, generator=g).to(DEV), torch.rand(3, 300, 420, generator=g).to(DEV)
    a = [float(L.rgb_loss_func_fused(img, gt)) for _ in range(3)]
    assert a[0] == a[1] == a[2]
    x = img.clone().requires_grad_(True)
    grads = []
    for _ in range(2):
        x.grad = None
        L.rgb_loss_func_fused(x, gt).backward()
        grads.append(x.grad.clone())
    assert torch.equal(grads[0], grads[1])
    with torch.no_grad():
        assert float(L.rgb_loss_func_fused(x, gt)) == a[0]


def test_fused_rgb_loss_full_frame_and_speed():
    """1280x1024: value / gradient against the float32 PyTorch formulation, and the two timed side by side (the
    numbers are printed, and only a generous 2x is asserted so that a busy box cannot fail the suite)."""
    C, H, W = 3, 1024, 1280
    g = torch.Generator().manual_seed(9)
    img, gt = torch.rand(C, H, W, generator=g).to(DEV), torch.rand(C, H, W, generator=g).to(DEV)
    mask = (torch.rand(1, H, W, generator=g) > 0.2).to(DEV)
    l32, g32 = _torch32(img, gt, 0.2, mask, 1.0)
    lf, gf = _fused(img, gt, 0.2, mask, 1.0)
    assert abs(lf - l32) <= 2e-5
    assert rel_err(gf, g32) <= 2e-4

    def bench(fn, n=10):
        for _ in range(3):
            fn(img, gt, 0.2, mask, 1.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn(img, gt, 0.2, mask, 1.0)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3
    t_torch, t_fused = bench(_torch32), bench(_fused)
    _lib.profile_enable(True)
    for _ in range(5):
        _fused(img, gt, 0.2, mask, 1.0)
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    k_f, k_b = prof["k_rgb_loss_fwd"][0] / prof["k_rgb_loss_fwd"][1], prof["k_rgb_loss_bwd"][0] / prof["k_rgb_loss_bwd"][1]
    print(f"rgb_loss fwd+bwd at 1280x1024: PyTorch {t_torch:.3f} ms, fused {t_fused:.3f} ms "
          f"(kernels: forward+reduce {k_f:.3f} ms, backward {k_b:.3f} ms)")
    assert t_fused < 0.5 * t_torch
