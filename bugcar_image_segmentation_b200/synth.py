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
