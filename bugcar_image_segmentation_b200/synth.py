"""Seeded synthetic inputs: camera frames, label maps and the
four BEV calibrations A-D.  The reference ships no data and its calibration JSON
is git-ignored (reference .gitignore:47), so every benchmark/parity input comes
from here.  NumPy only."""
import numpy as np

INPUT_H, INPUT_W = 256, 512     # models.py:19


def noise_frame(seed, h=INPUT_H, w=INPUT_W):
    """uniform uint8 BGR frame"""
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def blocky_frame(seed, h=INPUT_H, w=INPUT_W):
    """8x16 random colour blocks upsampled (nearest) + N(0,8) noise"""
    rng = np.random.default_rng(seed)
    b = rng.integers(0, 256, (8, 16, 3)).astype(np.float64)
    img = np.kron(b, np.ones(((h + 7) // 8, (w + 15) // 16, 1)))[:h, :w]
    img = img + rng.normal(0.0, 8.0, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def frames(n, seed0=1234, kind="mixed"):
    """(n,256,512,3) uint8; frame i is seeded seed0+i (config 4: rank r frame i
    uses seed 1234 + r*B + i)."""
    out = np.empty((n, INPUT_H, INPUT_W, 3), np.uint8)
    for i in range(n):
        blocky = (kind == "blocky") or (kind == "mixed" and (i % 2 == 1))
        out[i] = blocky_frame(seed0 + i) if blocky else noise_frame(seed0 + i)
    return out


def label_map(seed, classes=3, h=INPUT_H, w=INPUT_W, block=16):
    """blocky random label map in {0..classes-1}"""
    rng = np.random.default_rng(seed)
    small = rng.integers(0, classes, ((h + block - 1) // block, (w + block - 1) // block))
    return np.kron(small, np.ones((block, block), np.int64))[:h, :w].astype(np.uint8)


def perspective_transform(src, dst):
    """3x3 homography from 4 point pairs (same linear system as
    cv2.getPerspectiveTransform, solved in fp64 from float32-rounded points)."""
    src = np.asarray(src, np.float32).astype(np.float64)
    dst = np.asarray(dst, np.float32).astype(np.float64)
    A = np.zeros((8, 8))
    b = np.zeros(8)
    for i in range(4):
        x, y = src[i]
        u, v = dst[i]
        A[i] = [x, y, 1, 0, 0, 0, -x * u, -y * u]
        A[i + 4] = [0, 0, 0, x, y, 1, -x * v, -y * v]
        b[i] = u
        b[i + 4] = v
    h = np.linalg.solve(A, b)
    return np.append(h, 1.0).reshape(3, 3)


CALIBRATIONS = {          # name: (warp_w, warp_h, cm_per_px)
    "A": (500, 500, 2),   # trivial crop
    "B": (600, 400, 2),   # left_x = 50, top_y = -100
    "C": (500, 500, 3),   # fractional cell (3.33 px)
    "D": (400, 500, 2),   # warp narrower than the grid
    "E": (1280, 720, 4),  # 720p-sized warp (config 5 post-processing)
}


def calibration(name, in_rows=INPUT_H, in_cols=INPUT_W):
    """dict with the reference's JSON keys (bev.py:29-37) for a synthetic camera
    mount: road trapezoid -> central 40 % of the warped image."""
    ww, wh, cm = CALIBRATIONS[name]
    sx, sy = in_cols / 512.0, in_rows / 256.0
    src = [[180 * sx, 140 * sy], [332 * sx, 140 * sy], [512 * sx, 256 * sy], [0, 256 * sy]]
    dst = [[.3 * ww, 0], [.7 * ww, 0], [.7 * ww, wh], [.3 * ww, wh]]
    M = perspective_transform(src, dst)
    return {"output image size": [ww, wh], "input image size": [in_rows, in_cols],
            "bev matrix": M.reshape(-1).tolist(), "distance to target": [0, 100],
            "tile_length": 60, "cm_per_px": cm, "yaw": 0.0, "is_laserscan": False}


# ---- colour-region frames (tools/train_synthetic.py; parity inputs with real class structure)
PALETTE = np.array([     # 15 BGR colours standing for the class ids of note_label:1-15
    [128, 64, 128], [255, 255, 255], [232, 35, 244], [70, 70, 70], [156, 102, 102],
    [153, 153, 190], [30, 170, 250], [0, 220, 220], [153, 153, 153], [35, 142, 107],
    [180, 130, 70], [60, 20, 220], [142, 0, 0], [100, 60, 0], [32, 11, 119]], np.uint8)


def region_frame(seed, h=INPUT_H, w=INPUT_W, n_rect=24):
    """(bgr uint8 (h,w,3), class map uint8 (h,w)): random rectangles filled with palette
    colours, +-12 % brightness jitter per rectangle, N(0,8) pixel noise."""
    rng = np.random.default_rng(seed)
    lab = np.full((h, w), rng.integers(0, 15), np.uint8)
    gain = np.ones((h, w), np.float64)
    for _ in range(n_rect):
        y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
        rh, rw = int(rng.integers(h // 8, h // 2)), int(rng.integers(w // 8, w // 2))
        lab[y0:y0 + rh, x0:x0 + rw] = rng.integers(0, 15)
        gain[y0:y0 + rh, x0:x0 + rw] = rng.uniform(0.88, 1.12)
    img = PALETTE[lab].astype(np.float64) * gain[:, :, None] + rng.normal(0.0, 8.0, (h, w, 3))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8), lab


def road_mask(seed, h=INPUT_H, w=INPUT_W):
    """Binary road mask (uint8 0/1, or 0/arbitrary non-zero for seed % 7 == 6) as predict_binary
    would deliver it, for contour_noise_removal parity: seed % 7 picks white noise, coarse
    blocks, nested rings anchored in the bottom strip (kept hole contours, islands in holes),
    or a road trapezoid with elliptical holes, islands and salt-and-pepper noise."""
    rng = np.random.default_rng(seed)
    kind = seed % 7
    yy, xx = np.mgrid[0:h, 0:w]
    if kind == 0:
        return (rng.random((h, w)) < rng.uniform(0.2, 0.8)).astype(np.uint8)
    if kind == 1:
        s = int(rng.integers(2, 16))
        m = (rng.random((h // s + 1, w // s + 1)) < rng.uniform(0.3, 0.8)).astype(np.uint8)
        return np.ascontiguousarray(np.kron(m, np.ones((s, s), np.uint8))[:h, :w])
    m = np.zeros((h, w), np.uint8)
    if kind == 2:
        x0 = 0 if rng.random() < .5 else int(rng.integers(0, w // 8))
        x1 = w if rng.random() < .5 else w - int(rng.integers(0, w // 8))
        y0 = int(rng.integers(0, h // 2))
        y1 = h if rng.random() < .5 else h - int(rng.integers(0, 4))
        v = 1
        for _ in range(int(rng.integers(1, 6))):
            if x1 - x0 < 4 or y1 - y0 < 4:
                break
            if rng.random() < 0.5:
                m[y0:y1, x0:x1] = v
            else:
                cx, cy = (x0 + x1) / 2, (y0 + y1) / 2
                m[((xx - cx) / ((x1 - x0) / 2)) ** 2 + ((yy - cy) / ((y1 - y0) / 2)) ** 2 <= 1] = v
            v ^= 1
            x0 += int(rng.integers(1, max(2, (x1 - x0) // 6)))
            x1 -= int(rng.integers(1, max(2, (x1 - x0) // 6)))
            y0 += int(rng.integers(1, max(2, (y1 - y0) // 6)))
            y1 -= int(rng.integers(1, max(2, (y1 - y0) // 8)))
        m[rng.random((h, w)) < rng.uniform(0, 0.05)] ^= 1
        return m
    top = int(rng.integers(h // 4, 3 * h // 4))
    m[(yy > top) & (np.abs(xx - w / 2) < (yy - top) * rng.uniform(1, 4) + 20)] = 1
    for _ in range(int(rng.integers(0, 6))):
        cx, cy = int(rng.integers(0, w)), int(rng.integers(h // 2, h))
        rx, ry = int(rng.integers(5, max(6, w // 2))), int(rng.integers(3, max(4, h // 6)))
        e = ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2
        m[e < 1] = 0
        if rng.random() < 0.7:
            m[e < rng.uniform(0.1, 0.7)] = 1
            if rng.random() < 0.5:
                m[e < rng.uniform(0.01, 0.1)] = 0
    m[rng.random((h, w)) < rng.uniform(0, 0.08)] ^= 1
    if kind == 6:
        m = m * rng.integers(1, 255, (h, w)).astype(np.uint8)
    return m
