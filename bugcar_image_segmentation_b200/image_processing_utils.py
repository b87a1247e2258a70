"""Drop-in module name for the reference's ``image_processing_utils.py``.

None of these helpers is called on the hot path in the reference snapshot
(``contour_noise_removal`` is imported by models.py:6 and never used; ``create_skeleton``
calls ``create_occupancy_grid`` with a stale signature and cannot run).
Only the pure-geometry helper is provided; the OpenCV-based filters are "next" rows
 and raise until they exist as GPU kernels -- there is no CPU fallback
in this package.
"""
import numpy as np


def find_intersection_line(line1, line2):
    """image_processing_utils.py:63-91: intersection of two lines, each given by two
    points; ``None`` for parallel lines."""
    def coeffs(line):
        (x1, y1), (x2, y2) = line
        if x2 - x1 == 0:
            return 1.0, 0.0, x1
        a = (y2 - y1) / (x2 - x1)
        return a, -1.0, (x1 * y2 - x2 * y1) / (x2 - x1)

    a1, b1, c1 = coeffs(line1)
    a2, b2, c2 = coeffs(line2)
    if a1 == a2:
        return None
    return np.linalg.solve(np.array([[a1, b1], [a2, b2]]), np.array([c1, c2]))


def contour_noise_removal(segmap):
    raise NotImplementedError("contour_noise_removal (image_processing_utils.py:4-44) is not on the reference's "
                              "hot path and has no GPU kernel yet (SURVEY.md 8f-2)")


def clahe(img):
    raise NotImplementedError("clahe (image_processing_utils.py:46-61) is an optional camera pre-filter outside "
                              "the accelerated path")
