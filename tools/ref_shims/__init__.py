"""Dependency shims for running the UNMODIFIED Free-SurGS driver (train.py) offline -- SURVEY.md 8f N3.

The reference imports eleven third-party packages at module top that this image does not have
(gaussian_renderer/__init__.py:14, scene/pose_optimizer.py:11-23, scene/gaussian_model.py:21,
utils/general_utils.py:18-22, vis/*.py, utils/server_utils.py:1-3).  ``install()`` makes them importable:

  * functional stand-ins (``tools/ref_shims/functional/``) for the few functions train.py actually EXECUTES:
      kornia.geometry.epipolar  essential_from_Rt / fundamental_from_essential / sampson_epipolar_distance
                                (scene/pose_optimizer.py:20-22, used by compute_epipolar_loss :732-746)
      kornia.geometry.linalg / kornia.geometry.conversions   (imported by utils/geometry_utils.py:14,
                                utils/general_utils.py:18; rigid compose / inverse, rotation -> quaternion)
      imageio.imread            (scene/pose_optimizer.py:345; PIL underneath)
      skimage.metrics.structural_similarity   (utils/general_utils.py:42; uniform 7x7 window, scipy underneath)
      lpips.LPIPS               (utils/general_utils.py:31; the AlexNet weights are a download -> returns NaN,
                                loudly labelled; PSNR / SSIM are unaffected)
  * inert stubs (any attribute is a do-nothing class) for packages that are only imported, or only used by the
    viewer / plotting paths that stay switched off (``--visualize`` / ``--log`` unset):
      matplotlib, mpl_toolkits, plyfile, torchviz, viser, nerfview, open3d, splines

A package that IS installed is never shadowed: both kinds are only registered when the real import fails.
Nothing here is on the product path; it is test / measurement infrastructure for config 3.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
FUNCTIONAL = ("kornia", "imageio", "skimage", "lpips")
INERT = ("matplotlib", "mpl_toolkits", "plyfile", "torchviz", "viser", "nerfview", "open3d", "splines")


class _InertMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _inert_class(name)


class Inert(metaclass=_InertMeta):
    """Instances swallow every call / attribute; usable as a base class, a decorator, a context manager."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return Inert()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return Inert()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __iter__(self):
        return iter(())

    def __bool__(self):
        return False


def _inert_class(name):
    return _InertMeta(name, (Inert,), {})


class _InertModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _inert_class(name)


class _InertFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, roots):
        self.roots = set(roots)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _InertModule(spec.name)

    def exec_module(self, module):
        pass


def _missing(name: str) -> bool:
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def install(verbose: bool = False):
    """Register the shims for every package of the two lists that is not installed.  Returns
    {"functional": [...], "inert": [...]} -- what was actually shimmed."""
    inert = [m for m in INERT if _missing(m)]
    functional = [m for m in FUNCTIONAL if _missing(m)]
    if inert:
        sys.meta_path.append(_InertFinder(inert))
    if functional:
        fdir = os.path.join(_HERE, "functional")
        if fdir not in sys.path:
            sys.path.append(fdir)          # appended: a real package anywhere on the path wins
    if verbose:
        print(f"[ref_shims] functional stand-ins: {functional}; inert stubs: {inert}")
    return {"functional": functional, "inert": inert}
