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


class _GroupBy:
    """numpy_indexed.group_by(keys).min(values) -> (unique keys, per-key minimum), the only use
    the reference makes of the package (bev.py:156, :229); numpy_indexed is not installed."""

    def __init__(self, keys):
        import numpy as np
        self.keys = np.asarray(keys)

    def min(self, values):
        import numpy as np
        values = np.asarray(values)
        uniq = np.unique(self.keys)
        return uniq, np.array([values[self.keys == k].min() for k in uniq])


def deterministic_laserscan(ref):
    """Make the reference's laserscan branch a function of its input: OR WARP_FILL_OUTLIERS into its
    two cv2.warpPolar calls (bev.py:148,160,219,235 leave outliers uninitialised) and give it the
    group_by().min it imports from the absent numpy_indexed.  Returns an undo callable."""
    import cv2
    orig = cv2.warpPolar

    def filled(src, dsize, center, maxRadius, flags):
        return orig(src, dsize, center, maxRadius, flags | cv2.WARP_FILL_OUTLIERS)

    cv2.warpPolar = filled
    old_gb = getattr(ref.bev.npi, "group_by", None)
    ref.bev.npi.group_by = _GroupBy

    def undo():
        cv2.warpPolar = orig
        if old_gb is not None:
            ref.bev.npi.group_by = old_gb
    return undo


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
