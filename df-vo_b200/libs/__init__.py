"""``libs`` -- drop-in replacement of DF-VO's hot-path packages.

This package provides ``libs.deep_models``, ``libs.matching``, ``libs.tracker`` and ``libs.geometry`` with
the reference's public API, backed by the dfvo_b200 CUDA library.  Everything else the reference driver
imports (``libs.dfvo`` itself, ``libs.general``, ``libs.datasets``, ``libs.flowlib``) is *not* part of the
hot path and is taken unchanged from a DF-VO checkout: put this directory first on ``sys.path`` and
point ``DFVO_REFERENCE_ROOT`` at the checkout; the line below appends the
checkout's ``libs`` directory to this package's search path so those sub-packages resolve there.
"""
import os

_ref = os.environ.get("DFVO_REFERENCE_ROOT")
if _ref:
    _p = os.path.join(_ref, "libs")
    if os.path.isdir(_p) and _p not in __path__:
        __path__.append(_p)
