"""Shared parity metrics for the tests (CPU emulation and GPU)."""
import torch

from oracle import raster_oracle as ro

IMG_ABS_TOL = 1e-5        # north-star: <= 1e-5 abs on RGB / depth
GRAD_REL_TOL = 1e-4       # north-star: <= 1e-4 rel on all gradients
FLIP_ABS_BOUND = 2e-2     # a flipped alpha<1/255 / T<1e-4 decision moves a pixel by at most ~alpha*T*c
FLIP_FRACTION = 2e-3      # at most 0.2 % of the pixels may sit on a flipped decision


def check_image(name, got, ref, aux, scale=1.0):
    """got/ref [C,H,W]; ref is the float64 oracle; aux from want_aux=True (or None)."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    assert torch.isfinite(got).all(), name
    err = (got - ref).abs().amax(0)
    H, W = err.shape
    tol = IMG_ABS_TOL * scale
    n_bad = int((err > tol).sum())
    assert n_bad <= FLIP_FRACTION * H * W, f"{name}: {n_bad} pixels above {tol:g} (max {err.max():.3g})"
    assert err.max().item() <= FLIP_ABS_BOUND * scale, f"{name}: max abs err {err.max():.3g}"
    if aux is not None:
        fm = ro.fragile_pixel_mask(aux, H, W)
        solid = err[~fm]
        if solid.numel():
            assert solid.max().item() <= tol, f"{name}: {solid.max():.3g} on a pixel with no near-threshold decision"
    return err.max().item(), n_bad


def rel_err(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return ((got - ref).norm() / ref.norm().clamp(min=1e-30)).item()


def fragile_mask(aux, H, W, eps_pix=1e-4, eps_gauss=2e-6):
    """Pixels whose value may legitimately differ by a flipped threshold decision (see
    oracle/raster_oracle.fragile_pixel_mask).  Gradient comparisons zero the upstream gradient on
    these pixels on BOTH sides, so that a flip (an O(1/255) discontinuity that any float32
    implementation -- the reference included -- takes at its own rounding) cannot masquerade as a
    gradient error; the image checks count and bound the flips separately."""
    return ro.fragile_pixel_mask(aux, H, W, eps_pix=eps_pix, eps_gauss=eps_gauss)


def check_grad(name, got, ref, tol=GRAD_REL_TOL):
    """Norm-wise relative error over the tensor plus an element-wise check with a floor at
    1e-3 * max|ref| (float32 atomics make tiny entries relatively noisy)."""
    got_d, ref_d = got.detach().double().cpu(), ref.detach().double().cpu()
    assert torch.isfinite(got_d).all(), name
    r = rel_err(got_d, ref_d)
    if r > tol:
        d = (got_d - ref_d).abs().reshape(-1)
        top = torch.topk(d, min(5, d.numel()))
        detail = ", ".join(f"[{int(i)}] got {got_d.reshape(-1)[i]:.6g} ref {ref_d.reshape(-1)[i]:.6g}" for i in top.indices)
        raise AssertionError(f"{name}: norm-wise rel err {r:.3g} > {tol:g}; |ref| {ref_d.norm():.4g}; worst: {detail}")
    floor = 1e-3 * ref_d.abs().max().clamp(min=1e-30)
    elem = ((got_d - ref_d).abs() / ref_d.abs().clamp(min=floor))
    frac_bad = (elem > 50 * tol).double().mean().item()
    assert frac_bad <= 1e-3, f"{name}: {frac_bad:.3g} of the entries off by more than {50 * tol:g}"
    return r
