"""CPU restatement of ``contour_noise_removal`` (image_processing_utils.py:4-44) in terms of
connected components instead of contours.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PINNED: ``tests/test_oracle_contour.py`` compares it with the unmodified reference function
(OpenCV 4.13 ``morphologyEx`` / ``findContours`` / ``fillPoly``) on seeded masks, and
``tests/golden/contour.npz`` holds reference outputs.

What the reference computes, restated (each item cites the line it follows):

  1. ``closed = morphologyEx(segmap, MORPH_CLOSE, ones(k,k))``, ``k = int(min(h,w)/50)``
     (:6-9): grey dilate then erode with the anchor at ``k//2``; OpenCV's default border
     never wins a max (dilate) or a min (erode).
  2. ``findContours(closed, RETR_LIST, CHAIN_APPROX_SIMPLE)`` (:12): one contour per outer
     border of every 8-connected component C of ``closed != 0`` and one per hole border of every
     4-connected background region H that does not reach the image frame.  The contour is the
     closed chain of border pixels (foreground pixels with a background pixel of that region in
     their 4-neighbourhood; the frame counts as background).
  3. ``fillPoly(cnt_map, [cnt], 1)`` (:35) sets the chain itself plus every pixel strictly
     enclosed by it.  For an outer border that is C, its holes and everything nested in them;
     for a hole border it is H, everything nested in H, and the ring of C pixels 4-adjacent to H.
  4. a contour is kept when that set covers more than ``0.4`` of the bottom strip
     ``rows int(h*0.9)..h`` (:19-28, :37-39).  The ``cnt.shape[0] > 2`` filter (:13) only drops
     single pixels and straight one-pixel lines, which can never reach the threshold when
     ``h >= 30`` and ``w >= 3``.
  5. ``fillPoly(out, kept, 1)`` (:42) fills ALL kept contours in one call: chains are drawn,
     interiors follow the even-odd rule over the whole set.  A pixel is therefore 1 when it lies
     on a kept chain, or when an odd number of kept contours strictly enclose it.
"""
import numpy as np
from scipy import ndimage

LENGTH_RATIO = 0.1          # image_processing_utils.py:19
MASK_AREA_THRESH = 0.4      # image_processing_utils.py:31


def close_kxk(seg: np.ndarray) -> np.ndarray:
    """image_processing_utils.py:6-9 without OpenCV: grey close with a k x k box."""
    h, w = seg.shape
    k = int(min(h, w) / 50)
    if k < 1:
        raise ValueError("contour_noise_removal needs min(h, w) >= 50")
    a = k // 2
    lo, hi = a, k - 1 - a            # window covers offsets [-a, k-1-a]

    def sweep(img, red, fill):
        p = np.full((h + k - 1, w + k - 1), fill, dtype=np.uint8)
        p[lo:lo + h, lo:lo + w] = img
        out = p[0:h, 0:w].copy()
        for dy in range(k):
            for dx in range(k):
                out = red(out, p[dy:dy + h, dx:dx + w])
        return out

    d = sweep(seg, np.maximum, 0)
    return sweep(d, np.minimum, 255)


def contour_noise_removal(segmap: np.ndarray) -> np.ndarray:
    seg = np.ascontiguousarray(segmap, dtype=np.uint8)
    h, w = seg.shape
    fg = close_kxk(seg) != 0
    y_top = int(h * (1 - LENGTH_RATIO))
    thresh = (w * (h - y_top)) * MASK_AREA_THRESH

    # components on a frame-padded image: node ids of fg (8-connected) and bg (4-connected)
    P = np.zeros((h + 2, w + 2), bool)
    P[1:-1, 1:-1] = fg
    lf, nf = ndimage.label(P, structure=np.ones((3, 3), int))
    lb, nb = ndimage.label(~P)
    ext = lb[0, 0]
    # node numbering: fg 1..nf ; bg nf+1..nf+nb
    node = np.where(P, lf, lb + nf)
    ext = ext + nf
    N = nf + nb + 1
    # raster-first pixel of each node
    flat = node.ravel()
    first = np.full(N, -1, np.int64)
    idx = np.arange(flat.size)
    # np.minimum.at is slow but fine at oracle sizes
    order = np.argsort(flat, kind="stable")
    fs = flat[order]
    starts = np.r_[0, np.flatnonzero(np.diff(fs)) + 1]
    first[fs[starts]] = order[starts]
    W2 = w + 2
    up = np.full(N, -1, np.int64)
    for n in range(1, N):
        if n == ext or first[n] < 0:
            continue
        if n <= nf:        # component: the pixel left of its raster-first pixel is the enclosing region
            up[n] = flat[first[n] - 1]
        else:              # hole: the pixel above its raster-first pixel belongs to the enclosing component
            up[n] = flat[first[n] - W2]
    strip = np.zeros((h + 2, w + 2), bool)
    strip[1 + y_top:1 + h, 1:1 + w] = True
    own = np.bincount(flat[strip.ravel()], minlength=N).astype(np.int64)
    enc = own.copy()
    for n in range(1, N):
        if n == ext or own[n] == 0:
            continue
        a = up[n]
        while a != -1 and a != ext:
            enc[a] += own[n]
            a = up[a]
    # ring of component pixels 4-adjacent to each hole, counted inside the strip
    ring = np.zeros(N, np.int64)
    nbr = [node[0:-2, 1:-1], node[2:, 1:-1], node[1:-1, 0:-2], node[1:-1, 2:]]
    core = node[1:-1, 1:-1]
    corefg = fg
    upc = up[core]                                  # enclosing region of the pixel's component (valid where fg)
    seen = []
    for nb_ in nbr:
        is_hole = corefg & (nb_ > nf) & (nb_ != upc)
        dup = np.zeros_like(is_hole)
        for s in seen:
            dup |= (s == nb_)
        cnt = is_hole & ~dup & strip[1:-1, 1:-1]
        np.add.at(ring, nb_[cnt], 1)
        seen.append(np.where(is_hole, nb_, -1))
    keep = np.zeros(N, bool)
    keep[1:nf + 1] = enc[1:nf + 1] > thresh
    keep[nf + 1:] = (enc[nf + 1:] + ring[nf + 1:]) > thresh
    keep[ext] = False
    # parity of kept contours strictly enclosing everything inside node n, EXCLUDING n's own contour
    above = np.zeros(N, np.int64)
    for n in range(1, N):
        if n == ext or first[n] < 0:
            continue
        a, par = up[n], 0
        while a != -1 and a != ext:
            par ^= int(keep[a])
            a = up[a]
        above[n] = par
    out = np.zeros((h, w), np.uint8)
    # background pixels: own hole contour + everything above
    bgpix = ~corefg & (core != ext)
    out[bgpix] = (above[core] ^ keep[core])[bgpix]
    # foreground pixels
    on_outer = np.zeros((h, w), bool)
    on_kept_ring = np.zeros((h, w), bool)
    for nb_ in nbr:
        nb_bg = nb_ > nf
        on_outer |= corefg & nb_bg & (nb_ == upc)
        on_kept_ring |= corefg & nb_bg & (nb_ != upc) & keep[nb_]
    kc = keep[core]
    val = np.where(kc & on_outer, 1, above[core] ^ kc)
    val = np.where(on_kept_ring, 1, val)
    out[corefg] = val[corefg]
    return out


def contour_noise_removal_cv2(segmap: np.ndarray) -> np.ndarray:
    """The same function through the OpenCV calls the reference makes
    (image_processing_utils.py:9 morphologyEx, :12 findContours(mode 1 = RETR_LIST, method 2 =
    CHAIN_APPROX_SIMPLE), :35/:42 fillPoly): the CPU baseline bench.py times, and a second,
    contour-based opinion for the connected-component restatement above."""
    import cv2
    seg = np.ascontiguousarray(segmap, dtype=np.uint8)
    h, w = seg.shape
    k = int(min(h, w) / 50)
    closed = cv2.morphologyEx(seg, cv2.MORPH_CLOSE, np.ones((k, k), np.uint8))
    contours, _ = cv2.findContours(closed, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    y_top = int(h * (1 - LENGTH_RATIO))
    limit = w * (h - y_top) * MASK_AREA_THRESH
    kept = []
    for c in contours:
        if c.shape[0] <= 2:
            continue
        filled = np.zeros_like(seg)
        cv2.fillPoly(filled, [c], 1)
        if np.count_nonzero(filled[y_top:h]) > limit:
            kept.append(c)
    out = np.zeros((h, w), np.uint8)
    if kept:
        cv2.fillPoly(out, kept, 1)
    return out
