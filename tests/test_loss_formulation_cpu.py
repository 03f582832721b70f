"""The closed forms the fused loss kernels implement (free-surgs_b200/csrc/fsgs_kernels_loss.cuh), restated in torch
float64 on the CPU and checked against autograd of the reference formulations (fsgs_b200.losses, which
tests/test_losses_golden.py pins to the reference's utils/loss_utils.py).  The kernels themselves are compared with
the same formulations on the GPU (tests/test_gpu_losses.py, tests/test_gpu_zz_pearson_loss.py); this file keeps the
derivations honest where no GPU is available."""
import math

import pytest
import torch
import torch.nn.functional as F

from fsgs_b200 import losses as L


@pytest.mark.parametrize("shape,masked", [((3, 37, 45), False), ((1, 16, 50), True), ((3, 9, 9), True)])
def test_ssim_l1_backward_is_three_convolutions_of_per_pixel_derivatives(shape, masked):
    """k_rgb_loss_fwd / k_rgb_loss_bwd: with a = G*x, b = G*y, p = G*x^2, q = G*y^2, r = G*xy and
    S = A1 A2 / (B1 B2), the gradient is  dL/dx = (1-l)/N sign(x-y) - l/N [G*Da + 2x G*Dp + y G*Dr]  with
    Da = dS/da, Dp = dS/dp, Dr = dS/dr (G symmetric, zero padding), times the mask."""
    C, H, W = shape
    g = torch.Generator().manual_seed(H + W)
    img = torch.rand(C, H, W, generator=g, dtype=torch.float64).requires_grad_(True)
    gt = torch.rand(C, H, W, generator=g, dtype=torch.float64)
    mask = (torch.rand(1, H, W, generator=g) > 0.3) if masked else None
    lam = 0.2
    loss = L.rgb_loss_func(img, gt, lam, mask)
    loss.backward()

    m = torch.ones(1, H, W, dtype=torch.float64) if mask is None else mask.double()
    x, y = img.detach() * m, gt * m
    w = L._gaussian_window(11, 1.5, C, x)
    conv = lambda t: F.conv2d(t, w, padding=5, groups=C)
    a, b, p, q, r = conv(x), conv(y), conv(x * x), conv(y * y), conv(x * y)
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    A1, A2 = 2 * a * b + C1, 2 * (r - a * b) + C2
    B1, B2 = a * a + b * b + C1, (p - a * a) + (q - b * b) + C2
    S = A1 * A2 / (B1 * B2)
    Da = (2 * b * (A2 - A1) - S * 2 * a * (B2 - B1)) / (B1 * B2)       # exactly the kernel's expressions
    Dp = -S / B2
    Dr = 2 * A1 / (B1 * B2)
    N = C * H * W
    value = (1 - lam) * (x - y).abs().sum() / N + lam * (1 - S.sum() / N)
    grad = m * ((1 - lam) / N * torch.sign(x - y) - lam / N * (conv(Da) + 2 * x * conv(Dp) + y * conv(Dr)))
    assert abs(float(value) - float(loss.detach())) < 1e-13
    assert float((grad - img.grad).norm() / img.grad.norm()) < 1e-12


def test_separable_window_equals_the_references_outer_product_window():
    """The kernels apply the 1-D window twice; the reference builds the 2-D window as an outer product in float32."""
    w2 = L._gaussian_window(11, 1.5, 1, torch.zeros(1))[0, 0]
    g1 = torch.tensor([math.exp(-((i - 5) ** 2) / (2 * 1.5 ** 2)) for i in range(11)], dtype=torch.float32)
    g1 = g1 / g1.sum()
    assert torch.allclose(torch.outer(g1, g1), w2, atol=1e-9)
    assert abs(float(g1.sum()) - 1.0) < 1e-6


@pytest.mark.parametrize("shape", [(40, 50), (3, 7), (256, 384)])
def test_pearson_closed_form_from_five_raw_sums(shape):
    """k_pearson_sums / k_pearson_finish / k_pearson_bwd: loss and both gradients from
    (sum x, sum y, sum x^2, sum y^2, sum xy) -- unbiased std, eps added to the std."""
    g = torch.Generator().manual_seed(shape[0])
    x = torch.rand(*shape, generator=g, dtype=torch.float64).requires_grad_(True)
    y = (0.7 * x.detach() + 0.3 * torch.rand(*shape, generator=g, dtype=torch.float64)).requires_grad_(True)
    loss = L.pearson_depth_loss(x, y)
    loss.backward()
    xd, yd = x.detach().flatten(), y.detach().flatten()
    N = float(xd.numel())
    sx_, sy_, sxx, syy, sxy = xd.sum(), yd.sum(), (xd * xd).sum(), (yd * yd).sum(), (xd * yd).sum()
    mx, my = sx_ / N, sy_ / N
    qx, qy = ((sxx - N * mx * mx) / (N - 1)).sqrt(), ((syy - N * my * my) / (N - 1)).sqrt()
    c = sxy - N * mx * my
    sx, sy = qx + 1e-6, qy + 1e-6
    assert abs(float(1 - c / (N * sx * sy)) - float(loss.detach())) < 1e-12
    k1, ky, kx = 1 / (N * sx * sy), c / (N * sx * sy * sy * (N - 1) * qy), c / (N * sx * sx * sy * (N - 1) * qx)
    gy = -(k1 * (xd - mx) - ky * (yd - my))
    gx = -(k1 * (yd - my) - kx * (xd - mx))
    assert float((gy - y.grad.flatten()).norm() / y.grad.norm()) < 1e-10
    assert float((gx - x.grad.flatten()).norm() / x.grad.norm()) < 1e-10
