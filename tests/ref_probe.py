"""Probe for the REAL ``diff_gaussian_rasterization`` package (last row).

The rasteriser Free-SurGS was developed against is an un-vendored third-party CUDA extension
(requirements.txt:26 ``git+https://github.com/ingra14m/depth-diff-gaussian-rasterization.git``, no pin); it is not
in /root/reference and cannot be fetched here, which is why the oracle's rasteriser core is "parity unpinned".
If a build of it ever IS present on the machine the tests run on -- installed into the interpreter, dropped into
``baseline/_ref/`` by the driver, or pointed to by ``FSGS_REF_RASTERIZER`` -- these helpers find it and load it
under another module name, and tests/test_gpu_reference_package.py then runs the parity cases against it
(it becomes the primary oracle; the restatement is demoted to a cross-check).

A candidate only counts if it carries a COMPILED ``_C`` extension next to its ``__init__.py``: this repository's
own drop-in package of the same name (free-surgs_b200/diff_gaussian_rasterization, a ``_C.py`` shim) and the
oracle-backed one (oracle/ref_boundary) never qualify.
"""
from __future__ import annotations

import glob
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAME = "diff_gaussian_rasterization"
OURS = (os.path.join(ROOT, "free-surgs_b200"), os.path.join(ROOT, "oracle", "ref_boundary"))


def _has_compiled_C(pkg_dir: str) -> bool:
    return any(glob.glob(os.path.join(pkg_dir, pat)) for pat in ("_C*.so", "_C*.pyd"))


def candidate_dirs():
    env = os.environ.get("FSGS_REF_RASTERIZER")
    if env:
        yield env
    base = os.path.join(ROOT, "baseline", "_ref")
    yield base
    for sub in sorted(glob.glob(os.path.join(base, "*"))):
        if os.path.isdir(sub):
            yield sub
    for p in sys.path:
        if p and os.path.isdir(p) and not any(os.path.abspath(p).startswith(o) for o in OURS):
            yield p


def find_reference_rasterizer():
    """-> directory of a real ``diff_gaussian_rasterization`` package (with a compiled ``_C``), or None."""
    seen = set()
    for d in candidate_dirs():
        for pkg in (os.path.join(d, NAME), d if os.path.basename(os.path.normpath(d)) == NAME else None):
            if not pkg or pkg in seen:
                continue
            seen.add(pkg)
            if os.path.isfile(os.path.join(pkg, "__init__.py")) and _has_compiled_C(pkg):
                if not any(os.path.abspath(pkg).startswith(o) for o in OURS):
                    return os.path.abspath(pkg)
    return None


def load_reference_rasterizer(pkg_dir=None):
    """Import the real package under the module name ``ref_diff_gaussian_rasterization`` (so that it can live next
    to ours in one process).  Returns the module, or None if there is none / it does not load on this machine."""
    pkg_dir = pkg_dir or find_reference_rasterizer()
    if pkg_dir is None:
        return None
    alias = "ref_" + NAME
    if alias in sys.modules:
        return sys.modules[alias]
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == NAME or k.startswith(NAME + ".")}
    parent = os.path.dirname(pkg_dir)
    sys.path.insert(0, parent)
    try:
        importlib.invalidate_caches()
        mod = importlib.import_module(NAME)          # its own `from . import _C` needs the real name while loading
        if not os.path.abspath(getattr(mod, "__file__", "")).startswith(pkg_dir):
            return None
        sys.modules[alias] = mod
        return mod
    except Exception as exc:  # noqa: BLE001 -- a build for another torch / arch: report, do not fail the suite
        print(f"[ref_probe] {pkg_dir} is present but does not load: {exc!r}")
        return None
    finally:
        sys.path.remove(parent)
        for k in [k for k in sys.modules if k == NAME or k.startswith(NAME + ".")]:
            del sys.modules[k]
        sys.modules.update(saved)
