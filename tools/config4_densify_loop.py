"""BASELINE.json configs[3] under the profiler: 2 M Gaussians, 1280x1024, a mapping loop with densification on
(render -> L1 image loss -> backward with the densification statistics folded in -> Adam step; densify_and_prune
every `every` iterations).  For ncu captures on the GPU box:

   ncu --set full --clock-control none -k regex:^k_ -c 40 -o gpurun_out/r2_config4 python tools/config4_densify_loop.py 7 3

   python tools/config4_densify_loop.py [iterations] [densify every]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "free-surgs_b200")]
import torch  # noqa: E402

from fsgs_b200 import densify  # noqa: E402
from fsgs_b200 import frame_render as render  # noqa: E402
from fsgs_b200 import model  # noqa: E402
from fsgs_b200.synth import make_scene  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 7
every = int(sys.argv[2]) if len(sys.argv) > 2 else 3
P = int(sys.argv[3]) if len(sys.argv) > 3 else 2_000_000
LR = {"_xyz": 1.6e-4 * 5, "_features_dc": 0.0025, "_features_rest": 0.0025 / 20, "_opacity": 0.05, "_scaling": 0.005,
      "_rotation": 0.001}
sc = make_scene(P, 1280, 1024, size_mult=2.0, seed=0)
poses, pc = model.scene_to_device(sc, "cuda")
with torch.no_grad():
    target = render.render(poses, 0, pc, gs_grad=False, cam_grad=False)["render"].clone()
    pc.params["_features_dc"].add_(0.2 * torch.randn_like(pc.params["_features_dc"]))
    pc.params["_xyz"].add_(2e-3 * torch.randn_like(pc.params["_xyz"]))
densify.training_setup(pc, LR)
pc.variables['scene_radius'] = torch.tensor(0.75, device="cuda")
pc.fold_densification_stats = True
for it in range(1, iters + 1):
    pc.optimizer.zero_grad(set_to_none=True)
    out = render.render(poses, 0, pc, gs_grad=True, cam_grad=False)
    loss = (out["render"] - target).abs().mean()
    loss.backward()
    pc.optimizer.step()
    print(f"it {it}: P {pc.params['_xyz'].shape[0]} instances {int(out['num_rendered'][0])} loss {float(loss):.5f}", flush=True)
    if it % every == 0:
        densify.densify_and_prune(pc, 2e-6, 0.005, None)
        print(f"   densify_and_prune -> P {pc.params['_xyz'].shape[0]}", flush=True)
torch.cuda.synchronize()
