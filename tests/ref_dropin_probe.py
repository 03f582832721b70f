"""Helper process of tests/test_ref_dropin_cpu.py: import the REFERENCE's own ``gaussian_renderer.render``,
``GaussianModel`` and ``PoseModel`` (unmodified, from the checkout given as argv[1]) with free-surgs_b200/ on the
path, build a model the reference's way, call ``render`` with the reference's own arguments on the CPU, and report
how far the call got.  In a container without a GPU the expected end of the road is the library's refusal to run
without a CUDA device -- raised from inside ``fsgs_b200.rasterizer`` after travelling through the reference's render,
its ``GaussianRasterizer(raster_settings=...)(**rendervar)`` keyword call and our settings tuple: the proof that
the package names, the 12-field settings tuple and the call convention line up, and that there is no CPU fallback."""
import json
import os
import sys
import tempfile
import traceback
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = sys.argv[1]
sys.path[:0] = [ref, os.path.join(ROOT, "free-surgs_b200"), os.path.join(ROOT, "tools"), ROOT]

import ref_shims  # noqa: E402
import run_config3  # noqa: E402

ref_shims.install()
run_config3._cpu_redirect()

import torch  # noqa: E402
from make_scared_synth import write_sequence  # noqa: E402

out = {"stage": "start"}
try:
    import diff_gaussian_rasterization as dgr
    out["diff_gaussian_rasterization"] = os.path.relpath(dgr.__file__, ROOT)
    from gaussian_renderer import render                        # the reference's
    import gaussian_renderer
    out["gaussian_renderer"] = os.path.abspath(gaussian_renderer.__file__)
    from scene import GaussianModel
    from scene.pose_optimizer import PoseModel
    from simple_knn._C import distCUDA2                          # noqa: F401  (resolves to our stand-in)
    data = tempfile.mkdtemp(prefix="fsgs_dropin_")
    write_sequence(data, n_frames=2, W=96, H=80, P=600, renderer="oracle")
    args = SimpleNamespace(source_path=data, data_type=0, frame_start=0, frame_end=-1)
    poses = PoseModel(args, device="cpu")
    out["stage"] = "PoseModel loaded"
    out["frames"] = int(poses.num_cams)
    import numpy as np
    poses.record_data['pred_w2c'][0] = np.eye(4)
    poses.record_data['pred_depths'][0, ...] = poses.record_data['monodeps'][0].float()
    opt = SimpleNamespace(percent_dense=0.01, position_lr_init=1.6e-4, position_lr_final=1.6e-6, position_lr_delay_mult=0.01,
                          position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001)
    gaussians = GaussianModel(3, opt)
    gaussians.initialize_first_timestep(0, poses)
    gaussians.training_setup(opt)
    out["stage"] = "GaussianModel initialised"
    out["n_gaussians"] = int(gaussians.params['_xyz'].shape[0])
    out["settings_type"] = type(gaussians.cam).__module__ + "." + type(gaussians.cam).__name__
    out["settings_fields"] = list(gaussians.cam._fields)
    try:
        render(poses, 0, gaussians, gs_grad=True, cam_grad=True)
        out["stage"] = "render returned"            # only possible with a GPU
    except Exception as exc:  # noqa: BLE001
        tb = traceback.extract_tb(exc.__traceback__)
        out["stage"] = "render raised"
        out["error_type"] = type(exc).__module__ + "." + type(exc).__name__
        out["error"] = str(exc)
        out["raised_in"] = os.path.relpath(tb[-1].filename, ROOT)
        out["went_through"] = [os.path.abspath(f.filename) for f in tb]
except Exception as exc:  # noqa: BLE001
    out["failed"] = repr(exc)
    out["traceback"] = traceback.format_exc()[-2000:]
print("PROBE " + json.dumps(out))
