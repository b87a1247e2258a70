"""Pins oracle/contour_oracle.py (connected-component restatement of contour_noise_removal,
image_processing_utils.py:4-44) against the reference: committed outputs of the reference
function (tests/golden/contour.npz, tools/make_golden.py) and, in the build container, a live
run of the unmodified function under import stubs."""
import hashlib

import numpy as np
import pytest

from conftest import golden
from bugcar_image_segmentation_b200 import synth
from oracle import contour_oracle, refstub


def _cases():
    g = golden("contour.npz")
    for s in g["seeds"]:
        h, w = (int(v) for v in g["shapes"][int(s) % 4])
        yield g, int(s), h, w


def test_matches_reference_golden():
    kept = 0
    for g, s, h, w in _cases():
        m = synth.road_mask(s, h, w)
        want = np.unpackbits(g[f"out_{s}"])[:h * w].reshape(h, w)
        got = contour_oracle.contour_noise_removal(m)
        assert got.dtype == np.uint8 and np.array_equal(got, want), f"seed {s}"
        kept += int(want.any())
    assert kept >= 10          # the vectors are not all-empty


def test_close_matches_cv2_golden():
    for g, s, h, w in _cases():
        c = contour_oracle.close_kxk(synth.road_mask(s, h, w))
        assert hashlib.sha256(c.tobytes()).hexdigest() == str(g[f"closed_sha_{s}"]), f"seed {s}"


def test_kept_hole_contour_is_xored():
    """fillPoly over ALL kept contours follows the even-odd rule (image_processing_utils.py:42):
    a hole whose own contour covers > 40 % of the bottom strip stays empty, the ring around it
    is drawn, and an island inside it that is large enough comes back."""
    m = np.zeros((256, 512), np.uint8)
    m[100:256, :] = 1
    m[150:254, 20:500] = 0           # hole: 24 strip rows x 480 > 5324.8
    m[160:246, 40:480] = 1           # island: 16 strip rows x 440 > 5324.8 (gaps wider than the 5x5 close)
    m[200:240, 150:350] = 0          # small hole in the island: its contour is not kept -> filled
    out = contour_oracle.contour_noise_removal(m)
    assert out[120, 5] == 1 and out[155, 30] == 0 and out[180, 100] == 1 and out[220, 200] == 1
    assert out[149, 100] == 1 and out[50, 50] == 0


def test_empty_and_full():
    z = np.zeros((256, 512), np.uint8)
    assert not contour_oracle.contour_noise_removal(z).any()
    assert contour_oracle.contour_noise_removal(z + 1).all()
    with pytest.raises(ValueError):
        contour_oracle.contour_noise_removal(np.zeros((40, 80), np.uint8))


@pytest.mark.skipif(not refstub.available(), reason="/root/reference only exists in the build container")
def test_live_reference():
    ref = refstub.load()
    f = ref.image_processing_utils.contour_noise_removal
    for s in range(100, 135):
        h, w = [(256, 512), (120, 200), (50, 50), (300, 300), (360, 640)][s % 5]
        m = synth.road_mask(s, h, w)
        assert np.array_equal(contour_oracle.contour_noise_removal(m), f(m)), f"seed {s} shape {(h, w)}"
    m = np.zeros((256, 512), np.uint8)
    m[100:256, :] = 1
    m[150:254, 20:500] = 0
    m[160:246, 40:480] = 1
    m[200:240, 150:350] = 0
    assert np.array_equal(contour_oracle.contour_noise_removal(m), f(m))
