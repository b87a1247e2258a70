"""CPU restatement of the "laserscan-like" branch of create_occupancy_grid[_binary]
(bev.py:145-164 and bev.py:216-240).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's output on this branch is NOT a function of its input: both ``cv2.warpPolar``
calls run without ``WARP_FILL_OUTLIERS``, so every polar pixel whose ray leaves the grid, and
every grid cell that maps outside the polar image, keeps whatever the freshly allocated
destination held (two calls on one label map differed in 13 of 10 000 cells when probed).  This
module restates the branch with those pixels defined as 0 -- exactly what the reference computes
when ``WARP_FILL_OUTLIERS`` is OR-ed into both calls.

PARTIALLY PINNED: ``tests/test_oracle_laser.py`` checks (a) the two coordinate maps against
``cv2.warpPolar`` itself, and (b) the whole branch against the reference's own code run with
that one flag patched in (plus a three-line ``numpy_indexed.group_by().min`` stand-in: the
package is not installed); ``tests/golden/laser.npz`` holds those outputs.

Steps (binary variant, bev.py:145-164; the 3-way variant differs where noted):
  1. polar = warpPolar(grid, dsize, centre = (Wc/2 - 1, Hc), maxRadius = max(Wc, Hc), LINEAR),
     nearest interpolation; dsize = (Wc, Hc) [binary, :148] or (-1, -1) -> (round(R), round(R*pi))
     [3-way, :219].  Rows are angles, columns radii.
  2. per row, the first column whose pixel equals the obstacle value (100 [:154] / 3 [:227]).
  3. a filled radius-1 circle (a 5-pixel plus, clipped) of value 100 / 1 at each such point
     on a zero image (:157-158 / :232-233).
  4. warpPolar(..., WARP_INVERSE_MAP) back to (Wc, Hc) (:160 / :235).
  5. binary: int8 cast, cells that were unknown (255) become -1 (:161-163);
     3-way: cells that were not 3 keep their value, then 0 -> -1, v -> 200 - 100 v (:236, :244-245).
"""
import numpy as np

F = np.float32


def polar_dsize(max_radius):
    """cv::warpPolar with dsize <= 0: (width, height) = (round(R), round(R * pi))."""
    return int(np.rint(max_radius)), int(np.rint(max_radius * np.pi))


def forward_map(dw, dh, cx, cy, max_radius, src_w, src_h):
    """(dh, dw) int64: flat source index read by polar pixel (phi, rho), -1 outside the source."""
    cx, cy = float(F(cx)), float(F(cy))                       # Point2f
    k_angle = 2 * np.pi / dh
    k_mag = max_radius / dw
    rhos = (np.arange(dw) * k_mag).astype(F).astype(np.float64)
    phi = np.arange(dh) * k_angle
    mx = (rhos[None, :] * np.cos(phi)[:, None] + cx).astype(F)
    my = (rhos[None, :] * np.sin(phi)[:, None] + cy).astype(F)
    sx, sy = np.rint(mx).astype(np.int64), np.rint(my).astype(np.int64)      # cvRound: half to even
    ok = (sx >= 0) & (sx < src_w) & (sy >= 0) & (sy < src_h)
    return np.where(ok, sy * src_w + sx, -1)


def _fast_atan2_deg(y, x):
    """cv::fastAtan32f (degrees): the 7th-order odd polynomial of mathfuncs_core, fp32 throughout."""
    s = F(180 / np.pi)
    p1, p3 = F(0.9997878412794807) * s, F(-0.3258083974640975) * s
    p5, p7 = F(0.1555786518463281) * s, F(-0.04432655554792128) * s
    ax, ay = np.abs(x), np.abs(y)
    eps = F(np.finfo(np.float64).eps)
    wide = ax >= ay
    c = np.where(wide, ay / (ax + eps), ax / (ay + eps)).astype(F)
    c2 = (c * c).astype(F)
    a = ((((p7 * c2 + p5).astype(F) * c2 + p3).astype(F) * c2 + p1).astype(F) * c).astype(F)
    a = np.where(wide, a, F(90) - a).astype(F)
    a = np.where(x < 0, F(180) - a, a).astype(F)
    return np.where(y < 0, F(360) - a, a).astype(F)


def inverse_map(dw, dh, cx, cy, max_radius, pol_w, pol_h):
    """(dh, dw) int64: flat index into the (pol_h, pol_w) polar image read by cartesian cell
    (y, x), -1 outside.  The polar image is wrapped by one row top and bottom (BORDER_WRAP)."""
    cx, cy = F(cx), F(cy)
    k_angle = 2 * np.pi / pol_h
    k_mag = max_radius / pol_w
    bx = (np.arange(dw).astype(F) - cx).astype(F)
    by = (np.arange(dh).astype(F) - cy).astype(F)
    X, Y = np.meshgrid(bx, by)
    mag = np.sqrt((X * X + Y * Y).astype(F)).astype(F)
    ang = (_fast_atan2_deg(Y, X) * F(np.pi / 180)).astype(F)
    rho = (mag.astype(np.float64) / k_mag).astype(F)
    phi = ((ang.astype(np.float64) / k_angle).astype(F) + F(1)).astype(F)
    sx, sy = np.rint(rho).astype(np.int64), np.rint(phi).astype(np.int64)
    ok = (sx >= 0) & (sx < pol_w) & (sy >= 0) & (sy < pol_h + 2)
    return np.where(ok, ((sy - 1) % pol_h) * pol_w + sx, -1)


def _first_hits(polar_is_target):
    """per polar row the first column that holds the obstacle value, -1 if none"""
    any_ = polar_is_target.any(axis=1)
    return np.where(any_, polar_is_target.argmax(axis=1), -1)


def _plus_image(first, pol_h, pol_w):
    img = np.zeros((pol_h, pol_w), bool)
    for r in np.flatnonzero(first >= 0):
        c = int(first[r])
        for dr, dc in ((0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)):
            rr, cc = r + dr, c + dc
            if 0 <= rr < pol_h and 0 <= cc < pol_w:
                img[rr, cc] = True
    return img


def _ray_marks(values, target, pol_w, pol_h):
    hc, wc = values.shape
    centre = (wc / 2 - 1, hc)                                  # bev.py:148 / :219
    radius = max(wc, hc)
    fwd = forward_map(pol_w, pol_h, centre[0], centre[1], radius, wc, hc)
    polar = np.where(fwd >= 0, values.reshape(-1)[np.maximum(fwd, 0)], 0)
    plus = _plus_image(_first_hits(polar == target), pol_h, pol_w)
    inv = inverse_map(wc, hc, centre[0], centre[1], radius, pol_w, pol_h)
    return np.where(inv >= 0, plus.reshape(-1)[np.maximum(inv, 0)], False)


def laserscan_binary(grid_int8):
    """grid_int8: the int8 grid of the non-laserscan binary branch (-1 / 0 / 100).
    Returns (grid_int8, new_occ_grid) as bev.py:164 does."""
    g = np.asarray(grid_int8, np.int8)
    hc, wc = g.shape
    marks = _ray_marks(g.view(np.uint8), 100, wc, hc)          # dsize = (Wc, Hc)
    new = np.where(marks, 100, 0).astype(np.int8)
    new[g == -1] = -1
    return g, new


def laserscan_3way(template_u8):
    """template_u8: the resized template of bev.py:209-212 (values 0..3).  Returns the int8 grid."""
    t = np.asarray(template_u8, np.uint8)
    hc, wc = t.shape
    pol_w, pol_h = polar_dsize(max(wc, hc))                    # dsize = (-1, -1)
    marks = _ray_marks(t, 3, pol_w, pol_h)
    new = np.where(t != 3, t, marks.astype(np.uint8)).astype(np.int64)
    return np.where(new == 0, -1, 200 - new * 100).astype(np.int8)
