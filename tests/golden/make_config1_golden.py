"""Generates tests/golden/config1_golden.npz: BASELINE.json configs[0] (10k Gaussians, 640x512,
m = 2, SH degree 3, seed 0) rendered + differentiated by the float64 plain-C oracle through the
reference's `render` formulation (oracle/render_oracle.py, backend "c").

    python tests/golden/make_config1_golden.py

PARITY UNPINNED for the rasteriser core (no reference implementation on disk); the Python half of
the path is pinned by ref_python_half.npz.  The image planes are stored sub-sampled (every 4th row
and column) to keep the fixture small; gradients are stored in full (float32).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "free-surgs_b200")]
from fsgs_b200.synth import make_scene  # noqa: E402
from oracle import raster_oracle as ro  # noqa: E402
from oracle import render_oracle as R  # noqa: E402


def main():
    sc = make_scene(10000, 640, 512, size_mult=2.0, seed=0)
    dt = torch.float64
    params = {k: v.to(dt).requires_grad_(True) for k, v in sc.params.items()}
    r, t = sc.pose_q.to(dt).requires_grad_(True), sc.pose_t.to(dt).requires_grad_(True)
    out = R.render(params, r, t, sc.camera, 3, sc.camera.campos, True, True, backend="c", want_aux=True)
    # pixels sitting on a threshold decision take no upstream gradient (tests/parity.py::fragile_mask)
    mask = ro.fragile_pixel_mask(out["_aux"], 512, 640, eps_pix=1e-4, eps_gauss=2e-6, eps_order=0.0)
    keep = (~mask).to(dt)
    loss = (out["render"] * sc.grads_out["G_rgb"].to(dt) * keep).sum() + \
           (out["render_dep"] * sc.grads_out["G_dep"].to(dt) * keep).sum()
    out["render_w2c"].retain_grad()
    loss.backward()
    planes = torch.cat([out["render"], out["_depth_sil"]], 0).detach()
    arrays = {"planes_sub4": planes[:, ::4, ::4].numpy(), "radii": out["radii"].numpy(),
              "num_rendered_rect": np.int64(out["_num_rendered"]), "loss": np.float64(loss.item()),
              "fragile_mask_bits": np.packbits(mask.numpy().reshape(-1)),
              "g_pose": out["render_w2c"].grad[:3].numpy(), "g_r": r.grad.numpy(), "g_t": t.grad.numpy(),
              "g_means2D": out["viewspace_points"].grad.float().numpy()}
    for k, v in params.items():
        arrays["g_" + k] = v.grad.float().numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config1_golden.npz")
    np.savez_compressed(path, **arrays)
    print("fragile pixel fraction", mask.float().mean().item())
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB", "R_rect", int(out["_num_rendered"]), "loss", loss.item())


if __name__ == "__main__":
    main()
