"""CPU check of the drop-in claim ("what's weak" 6 / next-round item 2): the reference's REAL
``gaussian_renderer.render`` + ``GaussianModel`` + ``PoseModel`` are imported unmodified with free-surgs_b200/ on the
path and called; the call must reach this library's C-ABI layer through the reference's own keyword arguments and
stop there with the "no CPU fallback" error (this container has no GPU).  Needs a Free-SurGS checkout
(/root/reference in the build container, or the copy __graft_entry__.build() stages under baseline/_ref/)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_reference_render_reaches_the_c_abi_and_there_is_no_cpu_fallback():
    import run_config3
    ref = run_config3.find_reference()
    if ref is None:
        pytest.skip("no Free-SurGS checkout on this machine")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_dropin_probe.py"), ref], capture_output=True,
                       text=True, timeout=600, env=env)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("PROBE ")]
    assert lines, p.stdout[-2000:] + p.stderr[-3000:]
    out = json.loads(lines[-1][6:])
    assert "failed" not in out, out
    # the reference's imports resolved to OUR packages
    assert out["diff_gaussian_rasterization"].startswith("free-surgs_b200/diff_gaussian_rasterization"), out
    assert out["gaussian_renderer"].startswith(ref), out
    # the reference built its single settings tuple from our type, with the 12 fields in the reference's order
    assert out["settings_type"].endswith("GaussianRasterizationSettings")
    assert out["settings_fields"] == ["image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier",
                                      "viewmatrix", "projmatrix", "sh_degree", "campos", "prefiltered", "debug"]
    assert out["frames"] == 2 and out["n_gaussians"] == int(0.1 * 96 * 80)        # create_random_mask(.., 0.1)
    # ... and render() travelled through the reference's code into the library, which refuses to run on the CPU
    assert out["stage"] == "render raised", out
    assert out["error_type"].endswith("FsgsError") and "no CPU fallback" in out["error"], out
    assert out["raised_in"] == os.path.join("free-surgs_b200", "fsgs_b200", "rasterizer.py"), out
    assert any(f == os.path.join(ref, "gaussian_renderer", "__init__.py") for f in out["went_through"]), out
