"""The library's fused Pearson depth loss (fsgs_pearson_forward / _backward; reference utils/loss_utils.py:98-109)
against the float64 PyTorch formulation (fsgs_b200.losses.pearson_depth_loss, pinned to the reference by
tests/test_losses_golden.py) and the reference's golden value.  1e-5 abs on the loss, 1e-4 relative on the gradients."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from parity import rel_err  # noqa: E402

from fsgs_b200 import losses as L  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_python_half.npz"))


@pytest.mark.parametrize("shape", [(256, 384), (1024, 1280), (37, 5), (3, 7)])      # (n = 2 is degenerate: |corr| = 1, zero gradient)
@pytest.mark.parametrize("which", ["target", "both"])
def test_fused_pearson_depth_loss_matches_the_pytorch_formulation(shape, which):
    g = torch.Generator().manual_seed(shape[0] + shape[1])
    a = (0.5 + torch.rand(*shape, generator=g)).to(DEV)
    b = (0.8 * a.cpu() + 0.4 * torch.rand(*shape, generator=g) + 0.3).to(DEV)
    x64, y64 = a.double().requires_grad_(which == "both"), b.double().requires_grad_(True)
    (L.pearson_depth_loss(x64, y64) * 0.05).backward()
    x, y = a.clone().requires_grad_(which == "both"), b.clone().requires_grad_(True)
    loss = L.pearson_depth_loss_fused(x, y)
    (loss * 0.05).backward()
    assert abs(float(loss) - float(L.pearson_depth_loss(a.double(), b.double()))) <= 1e-5
    assert rel_err(y.grad.double(), y64.grad) <= 1e-4
    if which == "both":
        assert rel_err(x.grad.double(), x64.grad) <= 1e-4
    else:
        assert x.grad is None


def test_fused_pearson_matches_the_reference_golden_value_and_is_deterministic():
    a, b = torch.from_numpy(GOLD["loss_dep_a"]).to(DEV), torch.from_numpy(GOLD["loss_dep_b"]).to(DEV)
    v = [float(L.pearson_depth_loss_fused(a, b)) for _ in range(3)]
    assert v[0] == v[1] == v[2]
    assert abs(v[0] - float(GOLD["loss_pearson"])) <= 1e-5


@pytest.mark.parametrize("shape,box,p_corr", [((1024, 1280), 128, 0.5), ((256, 384), 128, 0.5), ((300, 517), 64, 0.9)])
@pytest.mark.parametrize("which", ["target", "both"])
def test_fused_local_pearson_loss_matches_the_batched_formulation(shape, box, p_corr, which):
    """fsgs_local_pearson_forward / _backward (reference utils/loss_utils.py:112-127) against the batched PyTorch
    formulation in float64 (fsgs_b200.losses.local_pearson_loss, pinned to the reference's golden value on the CPU by
    tests/test_losses_golden.py).  Same seed -> the same two torch.randint draws -> the same patches."""
    g = torch.Generator().manual_seed(shape[0] * 7 + shape[1])
    a = (0.5 + torch.rand(*shape, generator=g)).to(DEV)
    b = (0.8 * a.cpu() + 0.4 * torch.rand(*shape, generator=g) + 0.3).to(DEV)
    x64, y64 = a.double().requires_grad_(which == "both"), b.double().requires_grad_(True)
    torch.manual_seed(99)
    ref = L.local_pearson_loss(x64, y64, box, p_corr)
    (ref * 0.15).backward()
    x, y = a.clone().requires_grad_(which == "both"), b.clone().requires_grad_(True)
    torch.manual_seed(99)
    loss = L.local_pearson_loss_fused(x, y, box, p_corr)
    (loss * 0.15).backward()
    assert abs(float(loss) - float(ref)) <= 1e-5
    assert rel_err(y.grad.double(), y64.grad) <= 1e-4
    assert int((y.grad != 0).sum()) > 0 and int((y.grad == 0).sum()) > 0        # only the patches get a gradient
    if which == "both":
        assert rel_err(x.grad.double(), x64.grad) <= 1e-4
    else:
        assert x.grad is None


def test_fused_pearson_on_two_elements_is_finite():
    """n = 2 is degenerate (|corr| = 1 exactly, the analytic gradient is 0): the kernels must return the loss to
    1e-5 and a gradient that is zero to an absolute 1e-4 -- no NaN / inf from the vanishing variance terms."""
    for vals in ((1.0, 2.0, 0.5, 0.9), (1.0, 2.0, 0.9, 0.5), (3.0, 3.5, 10.0, 10.25)):
        a = torch.tensor(vals[:2], device=DEV).view(1, 2)
        b = torch.tensor(vals[2:], device=DEV).view(1, 2).requires_grad_(True)
        loss = L.pearson_depth_loss_fused(a, b)
        loss.backward()
        want = float(L.pearson_depth_loss(a.double(), b.detach().double()))
        assert abs(float(loss) - want) <= 1e-5
        assert torch.isfinite(b.grad).all() and float(b.grad.abs().max()) <= 1e-4
