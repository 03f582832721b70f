"""A few fused render fwd+bwd steps at the bench configuration, for ncu captures (run on the GPU box).
   python tools/one_step.py [P] [m] [steps] [full|track|track_frozen]
   full = mapping-style step (all gradients); track = the reference's tracking step (gs_grad=False, parameters
   still trainable); track_frozen = tracking against a frozen model (the library's pose-only backward)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "free-surgs_b200")]
import torch  # noqa: E402

from fsgs_b200 import frame_render as render  # noqa: E402
from fsgs_b200 import model  # noqa: E402
from fsgs_b200.synth import make_scene  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
m = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = sys.argv[4] if len(sys.argv) > 4 else "full"
sc = make_scene(P, 1280, 1024, size_mult=m, seed=0)
poses, pc = model.scene_to_device(sc, "cuda")
G = torch.cat([sc.grads_out["G_rgb"], sc.grads_out["G_dep"][None]]).cuda()
if mode == "track_frozen":
    for v in pc.params.values():
        v.requires_grad_(False)
for _ in range(steps):
    pc.zero_grad()
    poses.pose_param_net.zero_grad(set_to_none=True)
    if mode == "full":
        out = render.render(poses, 0, pc, gs_grad=True, cam_grad=True)
        loss = (out["render"] * G[:3]).sum() + (out["render_dep"] * G[3]).sum()
    else:
        out = render.render(poses, 0, pc, gs_grad=False, cam_grad=True)
        loss = (out["render"] * G[:3]).sum()
    loss.backward()
torch.cuda.synchronize()
print("R", out["num_rendered"], "loss", float(loss))
