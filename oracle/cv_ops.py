"""Integer/fp64 restatements (NumPy) of the four OpenCV primitives on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  These are the executable
specification for the CUDA kernels K1 (resize) and K9 (warp + morph + nearest).
The arithmetic lives in a third-party dependency of the reference,
``opencv-python`` (requirements.txt:1, unpinned; this image has 4.13.0), reached
from the reference's call sites

    cv2.resize(bgr, (512, 256))                       models.py:87
    cv2.warpPerspective(segmap, M, (ww, wh))          bev.py:114, bev.py:182
    cv2.morphologyEx(occ, cv2.MORPH_OPEN, ones(3,3))  bev.py:130, bev.py:198
    cv2.resize(t, (Wc, Hc), interpolation=NEAREST)    bev.py:139, bev.py:209

Each function below restates OpenCV's published algorithm for exactly those
flag combinations and is pinned bit-exact against cv2 4.13.0 by
``tests/test_oracle_cv.py`` (random fuzz) and the committed golden vectors.
"""
import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS  # 32


# --------------------------------------------------------------------------- warp
def invert3x3(M):
    """OpenCV inverts the src->dst homography in fp64 (cv::invert, 3x3 closed
    form: adjugate times 1/det).  imgwarp.cpp warpPerspective()."""
    M = np.asarray(M, dtype=np.float64).reshape(3, 3)
    a = M
    t = np.empty(9, np.float64)
    d = (a[0, 0] * (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1])
         - a[0, 1] * (a[1, 0] * a[2, 2] - a[1, 2] * a[2, 0])
         + a[0, 2] * (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]))
    if d == 0.0:
        return np.zeros((3, 3), np.float64)
    d = 1.0 / d
    t[0] = (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) * d
    t[1] = (a[0, 2] * a[2, 1] - a[0, 1] * a[2, 2]) * d
    t[2] = (a[0, 1] * a[1, 2] - a[0, 2] * a[1, 1]) * d
    t[3] = (a[1, 2] * a[2, 0] - a[1, 0] * a[2, 2]) * d
    t[4] = (a[0, 0] * a[2, 2] - a[0, 2] * a[2, 0]) * d
    t[5] = (a[0, 2] * a[1, 0] - a[0, 0] * a[1, 2]) * d
    t[6] = (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]) * d
    t[7] = (a[0, 1] * a[2, 0] - a[0, 0] * a[2, 1]) * d
    t[8] = (a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]) * d
    return t.reshape(3, 3)


def warp_block_width(dst_w, dst_h):
    """Column-block width of OpenCV's WarpPerspectiveInvoker (BLOCK_SZ = 32).  The
    fp64 source coordinate of a pixel is evaluated as
    ``(Mi0*x_block + Mi1*y + Mi2) + Mi0*(x - x_block)``, so the block origin is
    part of the arithmetic."""
    bh0 = min(16, dst_h)
    bw0 = min(1024 // bh0, dst_w)
    return bw0


def warp_coords_fixed(Mi, dst_w, dst_h):
    """Fixed-point (1/32 px) source coordinates X, Y (int32 arrays (dst_h,dst_w))
    for every destination pixel.  ``Mi`` is the *inverse* (dst->src) matrix."""
    Mi = np.asarray(Mi, np.float64).reshape(9)
    bw0 = warp_block_width(dst_w, dst_h)
    xs = np.arange(dst_w, dtype=np.int64)
    xb = ((xs // bw0) * bw0).astype(np.float64)[None, :]     # block origin
    x1 = (xs % bw0).astype(np.float64)[None, :]               # offset in block
    ys = np.arange(dst_h, dtype=np.float64)[:, None]
    X0 = Mi[0] * xb + Mi[1] * ys + Mi[2]
    Y0 = Mi[3] * xb + Mi[4] * ys + Mi[5]
    W0 = Mi[6] * xb + Mi[7] * ys + Mi[8]
    W = W0 + Mi[6] * x1
    with np.errstate(divide="ignore", invalid="ignore"):
        Ws = np.where(W != 0.0, INTER_TAB_SIZE / W, 0.0)
    lo, hi = float(-2 ** 31), float(2 ** 31 - 1)
    fX = np.maximum(lo, np.minimum(hi, (X0 + Mi[0] * x1) * Ws))
    fY = np.maximum(lo, np.minimum(hi, (Y0 + Mi[3] * x1) * Ws))
    # saturate_cast<int>(double) == cvRound == round-half-to-even
    X = np.rint(fX).astype(np.int64).clip(-2 ** 31, 2 ** 31 - 1)
    Y = np.rint(fY).astype(np.int64).clip(-2 ** 31, 2 ** 31 - 1)
    return X, Y


def warp_perspective_u8(src, M, dsize):
    """``cv2.warpPerspective(src, M, dsize)`` for single-channel uint8 with the
    default flags (INTER_LINEAR, BORDER_CONSTANT value 0)."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    assert src.ndim == 2
    dst_w, dst_h = int(dsize[0]), int(dsize[1])
    sh, sw = src.shape
    X, Y = warp_coords_fixed(invert3x3(M), dst_w, dst_h)
    # saturate_cast<short>(X >> INTER_BITS)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    ax = (X & (INTER_TAB_SIZE - 1))
    ay = (Y & (INTER_TAB_SIZE - 1))
    acc = np.zeros((dst_h, dst_w), np.int64)
    for dy in (0, 1):
        wy = ay if dy else (INTER_TAB_SIZE - ay)
        yy = sy + dy
        for dx in (0, 1):
            wx = ax if dx else (INTER_TAB_SIZE - ax)
            xx = sx + dx
            inside = (xx >= 0) & (xx < sw) & (yy >= 0) & (yy < sh)
            p = np.where(inside, src[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)], 0)
            acc += p.astype(np.int64) * wy * wx
    return ((acc + 512) >> 10).astype(np.uint8)


# ------------------------------------------------------------------------- resize
def _linear_axis_tables(dst_n, src_n, clamp_frac=True):
    """Per-axis source index pair and 11-bit coefficients of OpenCV's
    resizeGeneric_ INTER_LINEAR table builder (resize.cpp).  The x axis zeroes
    the fraction when the tap pair leaves the row (``clamp_frac``); the y axis
    keeps the fraction and only clamps the two row indices (resizeGeneric_Invoker
    ``clip(sy0 - ksize2 + 1 + k, 0, ssize.height)``), which differs by one LSB
    after the fixed-point vertical pass when upscaling."""
    scale = float(src_n) / float(dst_n)
    d = np.arange(dst_n)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)   # fx = (float)((dx+0.5)*scale_x - 0.5)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_frac:
        lo = s < 0
        f[lo] = 0.0
        s[lo] = 0
        hi = s >= src_n - 1
        f[hi] = 0.0
        s[hi] = src_n - 1
        s1 = np.minimum(s + 1, src_n - 1)
    else:
        s1 = np.clip(s + 1, 0, src_n - 1)
        s = np.clip(s, 0, src_n - 1)
    # cbuf in float, then saturate_cast<short>(cbuf * INTER_RESIZE_COEF_SCALE)
    c0 = np.rint((np.float32(1.0) - f).astype(np.float32) * np.float32(2048.0)).astype(np.int64)
    c1 = np.rint(f * np.float32(2048.0)).astype(np.int64)
    return s, s1, c0, c1


def resize_bilinear_u8(src, dsize):
    """``cv2.resize(src, dsize)`` (default INTER_LINEAR) for uint8 HxWxC.

    Follows resize.cpp: identity copy when sizes match; the INTER_LINEAR ->
    INTER_AREA substitution when both scale factors are exactly 2 (2x2 box mean
    with rounding); otherwise the 11-bit fixed-point separable kernel."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    squeeze = src.ndim == 2
    if squeeze:
        src = src[:, :, None]
    sh, sw, _ = src.shape
    dw, dh = int(dsize[0]), int(dsize[1])
    if (sh, sw) == (dh, dw):
        out = src.copy()
    elif sw == 2 * dw and sh == 2 * dh:
        s = src.astype(np.int64)
        out = ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    else:
        x0, x1, a0, a1 = _linear_axis_tables(dw, sw)
        y0, y1, b0, b1 = _linear_axis_tables(dh, sh, clamp_frac=False)
        s = src.astype(np.int64)
        # horizontal pass: int32 rows scaled by 2^11
        H = s[:, x0, :] * a0[None, :, None] + s[:, x1, :] * a1[None, :, None]
        S0 = H[y0]
        S1 = H[y1]
        b0 = b0[:, None, None]
        b1 = b1[:, None, None]
        out = ((((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2)
        out = np.clip(out, 0, 255).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def nearest_index_table(dst_n, src_n):
    """Source index per destination index of OpenCV's resizeNN:
    ``min(floor(d * (1 / (dst_n / src_n))), src_n - 1)`` in fp64."""
    inv_scale = float(dst_n) / float(src_n)
    ifx = 1.0 / inv_scale
    d = np.arange(dst_n, dtype=np.float64)
    return np.minimum(np.floor(d * ifx).astype(np.int64), src_n - 1)


def resize_nearest(src, dsize):
    """``cv2.resize(src, dsize, interpolation=cv2.INTER_NEAREST)`` (2-D)."""
    sh, sw = src.shape[:2]
    dw, dh = int(dsize[0]), int(dsize[1])
    xi = nearest_index_table(dw, sw)
    yi = nearest_index_table(dh, sh)
    return src[yi][:, xi]


# -------------------------------------------------------------------------- morph
def erode3(mask):
    """3x3 erosion with OpenCV's default border (outside = +max, i.e. ignored)."""
    m = np.pad(mask.astype(np.uint8), 1, constant_values=255)
    out = np.full(mask.shape, 255, np.uint8)
    for dy in range(3):
        for dx in range(3):
            out = np.minimum(out, m[dy:dy + mask.shape[0], dx:dx + mask.shape[1]])
    return out


def dilate3(mask):
    """3x3 dilation with OpenCV's default border (outside = 0)."""
    m = np.pad(mask.astype(np.uint8), 1, constant_values=0)
    out = np.zeros(mask.shape, np.uint8)
    for dy in range(3):
        for dx in range(3):
            out = np.maximum(out, m[dy:dy + mask.shape[0], dx:dx + mask.shape[1]])
    return out


def morph_open3(mask):
    """``cv2.morphologyEx(mask, cv2.MORPH_OPEN, np.ones((3,3)))``."""
    return dilate3(erode3(mask))
