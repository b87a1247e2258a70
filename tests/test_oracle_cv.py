"""Pins oracle/cv_ops.py (integer restatements of the four OpenCV primitives on the path)
against cv2 and against the committed golden warps generated from cv2 4.13."""
import hashlib

import numpy as np
import pytest

from conftest import golden
from bugcar_image_segmentation_b200 import synth
from oracle import cv_ops

cv2 = pytest.importorskip("cv2")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("cal", ["A", "B", "C", "D"])
def test_warp_matches_golden(cal):
    g = golden(f"warp_{cal}.npz")
    c = synth.calibration(cal)
    ww, wh = c["output image size"]
    for s in g["label_seeds"]:
        lab = synth.label_map(int(s), 3)
        out = cv_ops.warp_perspective_u8(np.add(lab, 1), c["bev matrix"], (ww, wh))
        assert np.array_equal(out, g[f"warp3_{s}"])


def test_warp_fuzz_vs_cv2():
    rng = np.random.default_rng(5)
    for it in range(12):
        sh, sw = int(rng.integers(8, 200)), int(rng.integers(8, 300))
        dw, dh = int(rng.integers(5, 400)), int(rng.integers(5, 300))
        src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        quad = np.array([[0, 0], [sw, 0], [sw, sh], [0, sh]], np.float64) + rng.normal(0, 0.15 * min(sh, sw), (4, 2))
        dst = np.array([[0, 0], [dw, 0], [dw, dh], [0, dh]], np.float64) + rng.normal(0, 0.1 * min(dw, dh), (4, 2))
        M = synth.perspective_transform(quad, dst)
        ref = cv2.warpPerspective(src, M, (dw, dh))
        assert np.array_equal(cv_ops.warp_perspective_u8(src, M, (dw, dh)), ref), it


def test_warp_singular_denominator():
    # w == 0 rows/cols: OpenCV maps them to source (0, 0)
    src = np.arange(64, dtype=np.uint8).reshape(8, 8) + 7
    M = np.array([[1.0, 0, 0], [0, 1.0, 0], [0.25, 0, -1.0]])
    assert np.array_equal(cv_ops.warp_perspective_u8(src, M, (12, 9)), cv2.warpPerspective(src, M, (12, 9)))


@pytest.mark.parametrize("hw", [(256, 512), (512, 1024), (720, 1280), (375, 621), (120, 160), (1080, 1920), (255, 511)])
def test_resize_bilinear_vs_cv2(hw):
    h, w = hw
    for seed, f in ((1, synth.noise_frame), (2, synth.blocky_frame)):
        img = f(seed, h, w)
        assert np.array_equal(cv_ops.resize_bilinear_u8(img, (512, 256)), cv2.resize(img, (512, 256)))


def test_resize_golden():
    g = golden("resize.npz")
    p = golden("pre.npz")
    for k in ("native", "p720", "odd", "up", "x2"):
        h, w, seed = (int(v) for v in p[k + "_hw_seed"])
        frame = synth.blocky_frame(seed, h, w) if seed % 2 else synth.noise_frame(seed, h, w)
        r = cv_ops.resize_bilinear_u8(frame, (512, 256))
        assert _sha(r) == str(g[k + "_sha"])


def test_nearest_vs_cv2():
    rng = np.random.default_rng(9)
    for it in range(40):
        sh, sw = int(rng.integers(1, 700)), int(rng.integers(1, 700))
        dh, dw = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        src = rng.integers(0, 4, (sh, sw), dtype=np.uint8)
        assert np.array_equal(cv_ops.resize_nearest(src, (dw, dh)), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_NEAREST))


def test_morph_open_vs_cv2():
    rng = np.random.default_rng(3)
    for p in (0.3, 0.6, 0.9):
        m = (rng.random((97, 131)) < p).astype(np.uint8)
        assert np.array_equal(cv_ops.morph_open3(m), cv2.morphologyEx(m, cv2.MORPH_OPEN, kernel=np.ones((3, 3))))
    for shape in ((1, 1), (2, 5), (3, 3)):
        m = np.ones(shape, np.uint8)
        assert np.array_equal(cv_ops.morph_open3(m), cv2.morphologyEx(m, cv2.MORPH_OPEN, kernel=np.ones((3, 3))))
