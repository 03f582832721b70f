"""How much of each tile's depth-sorted list is ever composited?  (run on the GPU box)
   python tools/list_stats.py [P] [m]
Prints R = sum of list lengths, the sum over tiles of the deepest contributor (what the forward has to
walk before every pixel of the tile is saturated or the list ends) and the same per 8x4 warp block."""
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
W, H = 1280, 1024
sc = make_scene(P, W, H, size_mult=m, seed=0)
poses, pc = model.scene_to_device(sc, "cuda")
xyz = pc.params['_xyz']
means2D = torch.zeros_like(xyz, requires_grad=True) + 0
planes, radii, stats = render.render_planes(
    xyz, pc.params['_features_dc'], pc.params['_features_rest'], pc.params['_opacity'], pc.params['_scaling'],
    pc.params['_rotation'], poses.get_pose(0), means2D, pc.cam, poses.cam_center, pc.active_sh_degree)
img = planes[0].grad_fn.saved_tensors[-1]
import ctypes  # noqa: E402
off = (ctypes.c_size_t * 6)()
_lib.lib().fsgs_img_offsets(W, H, off)
HW, T = W * H, (W // 16) * (H // 16)
raw = img.view(torch.uint8)
final_T = raw[off[0]:off[0] + 4 * HW].view(torch.float32).view(H, W)
n_contrib = raw[off[1]:off[1] + 4 * HW].view(torch.int32).view(H, W)
tile_off = raw[off[3]:off[3] + 4 * (T + 1)].view(torch.int32).long()
n = tile_off[1:] - tile_off[:-1]
nc_t = n_contrib.view(H // 16, 16, W // 16, 16).permute(0, 2, 1, 3).reshape(T, 256)
maxlast = nc_t.max(dim=1).values.long()
nc_b = n_contrib.view(H // 4, 4, W // 8, 8).permute(0, 2, 1, 3).reshape(-1, 32)
blast = nc_b.max(dim=1).values.long()
sat = (final_T < 1e-3).float().mean().item()
print(f"P={P} m={m} R={int(n.sum())} rect={stats[1]} tiles={T}")
print(f"list length: mean {n.float().mean():.1f} max {int(n.max())}  p50 {int(n.float().median())}")
print(f"deepest contributor per tile: sum {int(maxlast.sum())} = {maxlast.sum().item() / n.sum().item():.3f} of R; mean {maxlast.float().mean():.1f}")
print(f"per 8x4 block: sum {int(blast.sum())} (x1/8 = {blast.sum().item() / 8 / n.sum().item():.3f} of R); per pixel mean n_contrib {n_contrib.float().mean():.1f}")
print(f"pixels with final_T < 1e-3: {sat:.3f}")
for q in (64, 128, 256, 512, 1024):
    print(f"tiles with maxlast <= {q}: {(maxlast <= q).float().mean():.3f}   with n <= {q}: {(n <= q).float().mean():.3f}")
