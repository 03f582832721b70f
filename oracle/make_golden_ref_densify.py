"""ORACLE tooling (test infrastructure): golden vectors for densification + checkpoints from the REFERENCE's own code.

Run in the build container only (needs /root/reference):

    python oracle/make_golden_ref_densify.py        # writes tests/golden/ref_densify.npz

Imports the reference's ``scene.gaussian_model.GaussianModel`` unmodified (absent third-party imports stubbed, 'cuda'
redirected to the CPU, exactly as oracle/make_golden_ref_python.py does) and records, on seeded inputs:

  scene/gaussian_model.py:678  add_densification_stats       (accumulators after two calls)
  scene/gaussian_model.py:656  densify_and_prune             (all parameters, Adam moments and statistics after it;
                                                              torch.manual_seed pins the split step's torch.normal)
  scene/gaussian_model.py:452  reset_opacity
  scene/gaussian_model.py:86   capture()                     (the checkpoint tuple, as torch.save'd bytes in a
                                                              side file -- tests load it into OUR mirror)

tests/test_densify_golden.py replays the same inputs through fsgs_b200/densify.py.
"""
from __future__ import annotations

import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_ref_python as base  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "ref_densify.npz")
SEED_SPLIT = 4242


def seeded_inputs(P=400):
    g = torch.Generator().manual_seed(20261017)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32)
    inp = {"_xyz": rn(P, 3) * 0.5 + torch.tensor([0.0, 0.0, 1.0]), "_features_dc": 0.5 * rn(P, 1, 3),
           "_features_rest": 0.1 * rn(P, 15, 3), "_opacity": 2.0 * rn(P, 1), "_scaling": rn(P, 3) * 0.8 - 4.6,
           "_rotation": rn(P, 4)}
    grads = {k: 0.01 * rn(*v.shape) for k, v in inp.items()}          # one Adam step populates the moments
    accum = rn(P, 1).abs() * 3e-4
    denom = torch.randint(0, 4, (P, 1), generator=g).float()          # zeros included: never-seen Gaussians -> NaN grads
    max_radii = torch.randint(0, 40, (P,), generator=g).float()
    vs_grad = rn(P, 3) * 1e-3
    vs_grad[:, 2] = 0
    vis = torch.rand(P, generator=g) > 0.3
    return inp, grads, accum, denom, max_radii, vs_grad, vis


OPT = dict(percent_dense=0.01, position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01,
           position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001)
SCENE_RADIUS = 0.75
MAX_GRAD, MIN_OPACITY, MAX_SCREEN = 0.0002, 0.05, 20


def main():
    base._install_stubs()
    base._cpu_redirect()
    import scene.gaussian_model as gm                         # noqa: E402

    inp, grads, accum, denom, max_radii, vs_grad, vis = seeded_inputs()
    opt = types.SimpleNamespace(**OPT)
    m = gm.GaussianModel(3, opt)
    m.params = {k: torch.nn.Parameter(v.clone().requires_grad_(True)) for k, v in inp.items()}
    m.spatial_lr_scale = 5.0
    m.active_sh_degree = 2
    m.training_setup(opt)
    for k, p in m.params.items():
        p.grad = grads[k].clone()
    m.optimizer.step()
    m.optimizer.zero_grad(set_to_none=True)
    out = {}
    # add_densification_stats, twice
    m.variables['xyz_gradient_accum'] = torch.zeros(inp["_xyz"].shape[0], 1)
    m.variables['denom'] = torch.zeros(inp["_xyz"].shape[0], 1)
    vs = types.SimpleNamespace(grad=vs_grad)
    m.add_densification_stats(vs, vis)
    m.add_densification_stats(vs, ~vis | (vs_grad[:, 0] > 0))
    out["stats_accum"] = m.variables['xyz_gradient_accum'].clone()
    out["stats_denom"] = m.variables['denom'].clone()
    # the checkpoint tuple BEFORE densification (what train.py:371-373 saves)
    m.variables['xyz_gradient_accum'] = accum.clone()
    m.variables['denom'] = denom.clone()
    m.variables['max_radii2D'] = max_radii.clone()
    m.variables['scene_radius'] = torch.tensor(SCENE_RADIUS)
    buf = io.BytesIO()
    torch.save((m.capture(), 1234), buf)
    ck = os.path.join(os.path.dirname(OUT), "ref_chkpnt_gaussians.pth")
    with open(ck, "wb") as f:
        f.write(buf.getvalue())
    for k, p in m.params.items():
        out["pre" + k] = p.detach().clone()
    # densify_and_prune
    torch.manual_seed(SEED_SPLIT)
    m.densify_and_prune(MAX_GRAD, MIN_OPACITY, MAX_SCREEN, None)
    for k, p in m.params.items():
        out["post" + k] = p.detach().clone()
        st = m.optimizer.state[p]
        out["post_exp_avg" + k] = st["exp_avg"].clone()
        out["post_exp_avg_sq" + k] = st["exp_avg_sq"].clone()
    for k in ("xyz_gradient_accum", "denom", "max_radii2D"):
        out["post_var_" + k] = m.variables[k].clone()
    # reset_opacity
    m.reset_opacity()
    out["reset_opacity"] = m.params["_opacity"].detach().clone()
    out["reset_opacity_exp_avg"] = m.optimizer.state[m.params["_opacity"]]["exp_avg"].clone()
    arrays = {k: v.detach().cpu().numpy() for k, v in out.items()}
    np.savez_compressed(OUT, **arrays)
    print("wrote", os.path.normpath(OUT), f"{os.path.getsize(OUT) / 1024:.1f} KiB;", "P", inp["_xyz"].shape[0], "->",
          out["post_xyz"].shape[0], "; checkpoint", os.path.normpath(ck), f"{os.path.getsize(ck) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
