"""Where does the un-fused drop-in path (render_two_pass: PyTorch pre-processing + two GaussianRasterizer calls)
spend its time?  Run on the GPU box.   python tools/two_pass_profile.py [P]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "free-surgs_b200")]
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from fsgs_b200 import _lib, model  # noqa: E402
from fsgs_b200 import frame_render as render  # noqa: E402
from fsgs_b200.synth import make_scene  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
sc = make_scene(P, 1280, 1024, size_mult=2.0, seed=0)
poses, pc = model.scene_to_device(sc, "cuda")
G = torch.cat([sc.grads_out["G_rgb"], sc.grads_out["G_dep"][None]]).cuda()


def step():
    pc.zero_grad()
    poses.pose_param_net.zero_grad(set_to_none=True)
    out = render.render_two_pass(poses, 0, pc, gs_grad=True, cam_grad=True)
    ((out["render"] * G[:3]).sum() + (out["render_dep"] * G[3]).sum()).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"host issue {t_host / 5 * 1e3:.2f} ms/step, wall {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms/step")
_lib.profile_enable(True)
for _ in range(5):
    step()
prof = _lib.profile_collect()
_lib.profile_enable(False)
print({k: (round(v[0] / 5, 3), v[1] // 5) for k, v in prof.items() if v[1]})
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as pr:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(pr.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
