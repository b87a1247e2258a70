"""Import the *unmodified* reference package from /root/reference under stubs.

Container-only (``/root/reference`` does not exist on the GPU box): used by
``tools/make_golden.py`` to generate ``tests/golden/*.npz`` and by the
``not gpu`` tests (skipped when the directory is absent) to pin the oracle.

Stubs / shims, all outside the reference's arithmetic:
  * ``tensorflow``, ``numpy_indexed``, ``rospy``, ``std_msgs``, ``nav_msgs``,
    ``geometry_msgs``: empty modules (models.py:2, bev.py:6, occgrid_to_ros.py:2-8)
  * ``numpy.core.fromnumeric.swapaxes`` import in utils.py:3 works on NumPy 2 via
    the deprecated alias; ``np.Inf`` was removed in NumPy 2 (bev.py:118,186) -> shim
  * ``cv2.imshow`` is called inside the hot path (bev.py:132,213) and raises on
    headless OpenCV -> no-op
"""
import importlib
import os
import sys
import types

REFERENCE_DIR = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "bev.py"))


class _Anything(types.ModuleType):
    """Module stub whose every attribute is another stub (never called on the
    paths we exercise)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Anything(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):  # e.g. decorators
        return _Anything(self.__name__ + "()")


def load():
    """Return the reference as the package ``reference`` (modules bev, models,
    utils, image_processing_utils importable as attributes)."""
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    import numpy as np
    import cv2

    if not hasattr(np, "Inf"):
        np.Inf = np.inf  # bev.py:118,186
    cv2.imshow = lambda *a, **k: None  # bev.py:132,213
    for name in ("tensorflow", "numpy_indexed", "rospy"):
        sys.modules.setdefault(name, _Anything(name))
    for pkg, subs in (("std_msgs", ("msg",)), ("nav_msgs", ("msg",)),
                      ("geometry_msgs", ("msg",))):
        m = sys.modules.setdefault(pkg, _Anything(pkg))
        for s in subs:
            sys.modules.setdefault(pkg + "." + s, getattr(m, s))
    parent = os.path.dirname(REFERENCE_DIR)
    if parent not in sys.path:
        sys.path.insert(0, parent)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = importlib.import_module("reference")
        for sub in ("utils", "image_processing_utils", "bev", "models"):
            importlib.import_module("reference." + sub)
    return ref
