"""Host-side profile of one fused render step (where do the CPU milliseconds go?).  Run on the GPU box."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "free-surgs_b200")]
import torch  # noqa: E402

from fsgs_b200 import frame_render as render  # noqa: E402
from fsgs_b200 import model  # noqa: E402
from fsgs_b200.synth import make_scene  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
sc = make_scene(P, 1280, 1024, size_mult=2.0, seed=0)
poses, pc = model.scene_to_device(sc, "cuda")
G = torch.cat([sc.grads_out["G_rgb"], sc.grads_out["G_dep"][None]]).cuda()


def step():
    pc.zero_grad()
    poses.pose_param_net.zero_grad(set_to_none=True)
    out = render.render(poses, 0, pc, gs_grad=True, cam_grad=True)
    loss = (out["render"] * G[:3]).sum() + (out["render_dep"] * G[3]).sum()
    loss.backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
N = 30
t0 = time.perf_counter()
for _ in range(N):
    step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host issue time {t_host / N * 1e3:.3f} ms/step, wall incl. drain {t_all / N * 1e3:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(35)
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
