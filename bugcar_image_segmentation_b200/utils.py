"""Drop-in for the host-side helper of the reference's ``utils.py`` that the hot path's
calibration needs.  ``freeze_session`` (utils.py:49-83, TensorFlow graph freezing) and
``testDevice`` (utils.py:86-90, camera probing) are offline tooling / I/O and are not
part of this package."""
import numpy as np


def order_points_counter_clockwise(points, x_axis):
    """utils.py:10-44: order the fiducial corners counter-clockwise relative to the axis
    ``x_axis`` = [centre, point on the axis]: left-of-axis points by x, then right-of-axis
    points by x (in the axis-aligned frame).  Unlike the reference, ``x_axis`` is not
    modified in place (utils.py:15)."""
    points = np.asarray(points, dtype=np.float64)
    x_axis = np.asarray(x_axis, dtype=np.float64)
    centre = x_axis[0]
    d = x_axis[1] - centre
    ang = -np.arctan2(d[1], d[0])
    rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
    local = (rot @ (points - centre).T).T
    left = sorted((i for i in range(len(points)) if not local[i, 1] < 0), key=lambda i: local[i, 0])
    right = sorted((i for i in range(len(points)) if local[i, 1] < 0), key=lambda i: local[i, 0])
    return points[left + right]
