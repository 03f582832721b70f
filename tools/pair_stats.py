"""How full are the compositors' (warp, entry) visits?  (run on the GPU box, with an instrumented build:
   nvcc ... -DFSGS_PAIR_STATS -o gpurun_ab/lib_stats.so;  FSGS_RASTER_LIB=gpurun_ab/lib_stats.so python tools/pair_stats.py [P] [m])
Prints, for one fused forward + backward at the bench configuration, the number of (warp, entry) visits of each
compositor, the (pixel, entry) pairs that actually contribute, and how often only one 4x4 half of a warp's 8x4 block
is touched."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "free-surgs_b200")]
import torch  # noqa: E402

from fsgs_b200 import _lib, model  # noqa: E402
from fsgs_b200 import frame_render as render  # noqa: E402
from fsgs_b200.synth import make_scene  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
m = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
sc = make_scene(P, 1280, 1024, size_mult=m, seed=0)
poses, pc = model.scene_to_device(sc, "cuda")
G = torch.cat([sc.grads_out["G_rgb"], sc.grads_out["G_dep"][None]]).cuda()
lib = ctypes.CDLL(_lib.LIB_PATH)
out8 = (ctypes.c_ulonglong * 8)()
lib.fsgs_debug_pair_stats(out8)          # clear
pc.zero_grad()
out = render.render(poses, 0, pc, gs_grad=True, cam_grad=True)
((out["render"] * G[:3]).sum() + (out["render_dep"] * G[3]).sum()).backward()
lib.fsgs_debug_pair_stats(out8)
v = list(out8)
print(f"P={P} m={m} R={out['num_rendered']}")
print(f"backward: visits {v[0]}  valid pairs {v[1]}  = {v[1] / max(v[0], 1) / 32:.3f} of the lanes;"
      f" empty visits {v[2]} ({v[2] / max(v[0], 1):.3f}); one 4x4 half only {v[3]} ({v[3] / max(v[0], 1):.3f})")
print(f"forward:  visits {v[4]}  contributing pairs {v[5]} = {v[5] / max(v[4], 1) / 32:.3f} of the lanes")
